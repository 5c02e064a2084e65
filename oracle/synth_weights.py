"""Re-export (TEST INFRASTRUCTURE): the deterministic random-init weights live in dir_b200/synth.py."""
from dir_b200.synth import fingerprint, load_key_shapes, make_state_dict, make_tensor  # noqa: F401
