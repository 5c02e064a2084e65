"""N3: MANO pickles without chumpy + checkpoint variants. CPU only.

The real MANO_{LEFT,RIGHT}.pkl are licence-gated, so the fixture below writes pickles with the same structure: a
protocol-2 dict whose array members are instances of a class pickled as `chumpy.ch.Ch` (state = instance dict with
the payload under 'x', as chumpy's Ch.__getstate__ produces), J_regressor a scipy csc_matrix, no 'betas' member.
"""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from dir_b200 import assets
from oracle import synth


def _write_mano_pickles(root, same_shapedirs_x=False):
    ch_pkg, ch_mod = types.ModuleType("chumpy"), types.ModuleType("chumpy.ch")

    class Ch:  # pickled by reference as chumpy.ch.Ch
        def __init__(self, x):
            self.x = np.asarray(x, dtype=np.float64)
            self._dirty_vars = set()
            self._itr = None

        def __getstate__(self):
            return dict(self.__dict__)

    Ch.__module__, Ch.__qualname__ = "chumpy.ch", "Ch"
    ch_mod.Ch = Ch
    ch_pkg.ch = ch_mod
    saved = {k: sys.modules.get(k) for k in ("chumpy", "chumpy.ch")}
    sys.modules["chumpy"], sys.modules["chumpy.ch"] = ch_pkg, ch_mod
    try:
        right_sd = None
        for side in ("right", "left"):
            a = synth.make_mano_arrays(side)
            sd = a["shapedirs"].astype(np.float64)
            if same_shapedirs_x:
                if side == "right":
                    right_sd = sd
                else:
                    sd = sd.copy()
                    sd[:, 0, :] = right_sd[:, 0, :]
            dd = {
                "hands_components": a["hands_components"].astype(np.float64),
                "hands_mean": a["hands_mean"].astype(np.float64),
                "hands_coeffs": np.zeros((4, 45)),
                "shapedirs": Ch(sd),
                "posedirs": a["posedirs"].astype(np.float64),
                "v_template": a["v_template"].astype(np.float64),
                "J_regressor": sp.csc_matrix(a["J_regressor"].astype(np.float64)),
                "J": Ch(np.zeros((16, 3))),
                "weights": Ch(a["weights"]),
                "f": a["f"].astype(np.uint32),
                "kintree_table": a["kintree_table"],
                "bs_style": "lbs", "bs_type": "lrotmin",
            }
            with open(os.path.join(root, f"MANO_{side.upper()}.pkl"), "wb") as f:
                pickle.dump(dd, f, protocol=2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.fixture()
def mano_dir(tmp_path):
    _write_mano_pickles(str(tmp_path))
    assert "chumpy" not in sys.modules
    return str(tmp_path)


def test_pickle_reads_without_chumpy(mano_dir):
    dd = assets.read_mano_pickle(os.path.join(mano_dir, "MANO_RIGHT.pkl"))
    a = synth.make_mano_arrays("right")
    for k in ("shapedirs", "posedirs", "v_template", "weights", "hands_components", "hands_mean", "J_regressor"):
        assert isinstance(dd[k], np.ndarray), k
        np.testing.assert_allclose(dd[k], a[k].astype(np.float64), rtol=0, atol=0)
    assert dd["J_regressor"].shape == (16, 778)
    assert dd["betas"].shape == (10,) and not dd["betas"].any()          # ready_arguments :42-43
    assert dd["pose"].shape == (48,) and dd["trans"].shape == (3,)       # :36-41
    assert dd["bs_type"] == "lrotmin"
    with open(os.path.join(mano_dir, "MANO_LEFT.pkl"), "rb") as f:
        assert assets.read_mano_pickle(f.read())["f"].shape == (1538, 3)  # bytes input


def test_layer_buffers_match_manolayer_registration(mano_dir):
    for side in ("left", "right"):
        got = assets.mano_layer_buffers(assets.read_mano_pickle(os.path.join(mano_dir, f"MANO_{side.upper()}.pkl")))
        want = synth.mano_buffers(side)
        assert tuple(got) == assets.MANO_BUFFER_NAMES
        for k in assets.MANO_BUFFER_NAMES:
            assert tuple(got[k].shape) == tuple(want[k].shape), k
            assert got[k].dtype == (torch.int64 if k == "th_faces" else torch.float32)
            np.testing.assert_array_equal(got[k].numpy(), want[k])
    flat = assets.mano_layer_buffers(assets.read_mano_pickle(os.path.join(mano_dir, "MANO_LEFT.pkl")), flat_hand_mean=True)
    assert not flat["th_hands_mean"].any()


def test_fix_shape(tmp_path):
    _write_mano_pickles(str(tmp_path), same_shapedirs_x=True)
    st = assets.mano_state_from_dir(str(tmp_path))
    l, r = (st[f"init_regressor.mano_layer_{s}.th_shapedirs"] for s in ("left", "right"))
    assert torch.equal(l[:, 0, :], -r[:, 0, :])                          # models/dir.py:306-309 applied
    assert not torch.equal(l[:, 1, :], -r[:, 1, :])
    for owner in assets.MANO_OWNERS[1:]:
        assert torch.equal(st[f"{owner}.mano_layer_left.th_shapedirs"], l)


def test_state_from_dir_covers_every_mano_key(mano_dir):
    st = assets.mano_state_from_dir(mano_dir)
    shapes = synth.load_key_shapes()
    mano_keys = {k for k in shapes if "mano_layer_" in k}
    assert set(st) == mano_keys and len(st) == 60
    for k, v in st.items():
        assert list(v.shape) == shapes[k], k
    with pytest.raises(FileNotFoundError):
        assets.mano_state_from_dir(os.path.join(mano_dir, "nope"))


def test_module_reads_pickles_at_construction(mano_dir):
    from dir_b200 import DIR
    net = DIR(21, mano_dir)
    sd = net.state_dict()
    want = synth.mano_buffers("left")["th_posedirs"]
    np.testing.assert_array_equal(sd["decoder.projecter_3.regressor.mano_layer_left.th_posedirs"].numpy(), want)
    # a checkpoint that lacks the MANO buffers is complete once the pickles were read
    state = {k: v for k, v in synth.make_state_dict().items() if "mano_layer_" not in k}
    net.load_state_dict(state, strict=False)
    assert {k for k in sd if "mano_layer_" in k} <= net._loaded_keys
    np.testing.assert_array_equal(net.state_dict()["init_regressor.mano_layer_right.th_weights"].numpy(),
                                  synth.mano_buffers("right")["th_weights"])
    # and without pickles the module still constructs (buffers then come from the checkpoint)
    assert DIR(21, "./misc/mano")._asset_keys == set()


def test_checkpoint_variants(tmp_path):
    keys = synth.load_key_shapes()
    small = {k: torch.zeros(1) for k in list(keys)[:5]}
    p = str(tmp_path / "DIR.pth")
    torch.save({"net": small, "optimizer": {}, "schedule": {}, "last_epoch": 3}, p)      # train.py:139-149
    st, rep = assets.read_checkpoint(p, expected_keys=keys)
    assert set(st) == set(small) and len(rep["missing"]) == len(keys) - 5 and rep["unexpected"] == []
    st2, _ = assets.read_checkpoint({"net": {"module." + k: v for k, v in small.items()}})
    assert set(st2) == set(small)
    st3, rep3 = assets.read_checkpoint(dict(small, extra=torch.zeros(1)), expected_keys=keys)
    assert rep3["unexpected"] == ["extra"] and "extra" in st3
    with pytest.raises(ValueError):
        assets.read_checkpoint([1, 2, 3])


@pytest.mark.skipif(not os.path.isdir("/root/reference/manopth"), reason="reference tree not present")
def test_against_reference_manolayer(mano_dir):
    """The reference's own ManoLayer.__init__, fed our chumpy-free reader's output, registers the same buffers."""
    from oracle import ref_shims
    ref = ref_shims.load_reference()
    import manopth.manolayer as ref_manolayer

    class R:
        def __init__(self, a):
            self.r = np.asarray(a)

    def ready(path, *a, **k):
        dd = assets.read_mano_pickle(path)
        out = dict(dd)
        for s in ("betas", "shapedirs", "posedirs", "v_template", "weights"):
            out[s] = R(dd[s])
        out["J_regressor"] = sp.csc_matrix(dd["J_regressor"])
        return out

    old = ref_manolayer.ready_arguments
    ref_manolayer.ready_arguments = ready
    try:
        for side in ("left", "right"):
            layer = ref_manolayer.ManoLayer(root_rot_mode="6D", joint_rot_mode="axisang", use_pca=True, mano_root=mano_dir,
                                            side=side, ncomps=45, center_idx=0, flat_hand_mean=False, robust_rot=True)
            got = assets.mano_layer_buffers(assets.read_mano_pickle(os.path.join(mano_dir, f"MANO_{side.upper()}.pkl")))
            ref_sd = layer.state_dict()
            assert set(ref_sd) == set(got)
            for k in got:
                assert ref_sd[k].dtype == got[k].dtype and torch.equal(ref_sd[k], got[k]), k
    finally:
        ref_manolayer.ready_arguments = old
