"""Re-export (TEST INFRASTRUCTURE): synthetic MANO assets live in dir_b200/synth.py."""
from dir_b200.synth import KINTREE_PARENTS, make_mano_arrays, mano_buffers  # noqa: F401
