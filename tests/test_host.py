"""CPU tests of the host side: the C-ABI library loads and exports every symbol of include/dirb200.h, the
drop-in module reproduces the reference's state_dict contract, record unpacking, failure behaviour without a
GPU, and the sharding helpers under a world_size-2 gloo job."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "dir_b200", "libdirb200.so")
    if not os.path.exists(so):
        sys.path.insert(0, ROOT)
        import __graft_entry__

        __graft_entry__.build()
    import dir_b200

    return dir_b200.load_library()


def header_symbols():
    with open(os.path.join(ROOT, "include", "dirb200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dirb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol(lib):
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dirb200.h but not exported"
    from dir_b200 import capi

    assert set(capi.EXPORTS) == set(names)


def test_create_fails_loudly_without_b200(lib):
    import dir_b200
    from dir_b200 import capi

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = capi.Config(0, 8, 1, 0, 2, 0)
    h = ctypes.c_void_p()
    rc = lib.dirb200_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -3 and not h.value
    assert b"CUDA" in lib.dirb200_last_error(None) or b"device" in lib.dirb200_last_error(None)
    m = dir_b200.DIR(21, "./misc/mano")
    with pytest.raises(dir_b200.DirB200Error):
        m({"img": torch.zeros(1, 3, 256, 256)}, None, None)
    bad = capi.Config(7, 8, 1, 0, 2, 0)
    assert lib.dirb200_create(ctypes.byref(bad), ctypes.byref(h)) == -1  # unknown precision


def test_module_contract(golden_dir):
    import json

    import dir_b200

    with open(os.path.join(golden_dir, "state_dict_keys.json")) as f:
        ref = json.load(f)
    m = dir_b200.DIR(21, "./misc/mano", 0)
    sd = m.state_dict()
    assert set(sd) == set(ref) and all(list(sd[k].shape) == ref[k] for k in ref)
    assert not m.training  # inference module
    with pytest.raises(ValueError):
        dir_b200.DIR(16, "./misc/mano")
    m.train()
    with pytest.raises(NotImplementedError):
        m({"img": torch.zeros(1, 3, 256, 256)}, None, None)
    # strict load of a synthetic checkpoint, like apps/eval.py:107-108
    from oracle.synth import make_state_dict

    res = m.load_state_dict(make_state_dict(0, prefix="decoder.projecter_4.interaction."), strict=False)
    assert len(res.unexpected_keys) == 0


def test_unpack_record_layout():
    import dir_b200
    from dir_b200 import capi

    B = 3
    rec = torch.arange(B * capi.RECORD_FLOATS, dtype=torch.float32).reshape(B, capi.RECORD_FLOATS)
    outs = dir_b200.DIR.unpack_record(rec, {"seg": None})
    assert len(outs) == 4 and outs[0]["pd_rel_joint"] is None
    for i in range(3):
        o = outs[i]
        assert o["pd_mesh_xyz_left"].shape == (B, 778, 3) and o["pd_joint_uv_right"].shape == (B, 21, 2)
        assert o["pd_offset"].shape == (B, 3) and o["pd_proj_left"].shape == (B, 3)
        base = i * capi.STAGE_FLOATS
        assert float(o["pd_mesh_xyz_left"][1, 0, 0]) == capi.RECORD_FLOATS + base
        assert float(o["pd_mesh_xyz_right"][0, 0, 0]) == base + 2334
        assert float(o["pd_joint_xyz_left"][0, 2, 1]) == base + 4668 + 7
        assert float(o["pd_joint_uv_right"][0, 20, 1]) == base + 4836 + 41
        assert float(o["pd_offset"][2, 2]) == 2 * capi.RECORD_FLOATS + base + 4886
    assert 4887 == 2 * 2334 + 2 * 63 + 2 * 42 + 3 + 3 + 3


def test_shard_bounds_cover():
    from dir_b200.dist import padded_shard, shard_bounds

    for n in (1, 7, 128, 1000):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == padded_shard(n, w)


def _gloo_worker(rank, world, port, n, q):
    import torch.distributed as dist

    from dir_b200.dist import assemble, broadcast_bytes, padded_shard, shard_bounds

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        uid = broadcast_bytes(bytes(range(128)) if rank == 0 else None)
        assert uid == bytes(range(128))
        # every image's "record" is its global index; shards are padded like the NCCL all-gather needs
        lo, hi = shard_bounds(n, rank, world)
        pad = padded_shard(n, world)
        local = torch.zeros(pad, 5)
        local[: hi - lo] = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1).expand(-1, 5)
        bufs = [torch.empty(pad, 5) for _ in range(world)]
        dist.all_gather(bufs, local)
        full = assemble(torch.cat(bufs, 0), n, world)
        ok = bool((full[:, 0] == torch.arange(n, dtype=torch.float32)).all()) and full.shape == (n, 5)

        # forward_sharded end to end with a stand-in for the GPU module (incl. n < world: a rank with an empty shard
        # must still enter the collective instead of raising and leaving its peers hanging)
        import dir_b200
        from dir_b200 import capi
        from dir_b200.dist import forward_sharded

        class Fake:
            unpack_record = staticmethod(dir_b200.DIR.unpack_record)

            def _device(self):
                return torch.device("cpu")

            def run_raw(self, img):
                assert img.shape[0] > 0, "run_raw must not be called for an empty shard"
                return {"record": img[:, 0, 0, :1].expand(-1, capi.RECORD_FLOATS).contiguous()}

            def allgather_records(self, rec):
                out = [torch.empty_like(rec) for _ in range(world)]
                dist.all_gather(out, rec)
                return torch.cat(out, 0)

        for m in (n, 1):
            img = torch.arange(m, dtype=torch.float32).view(m, 1, 1, 1).expand(m, 3, 4, 4).contiguous()
            outs = forward_sharded(Fake(), img)
            ok = ok and outs[2]["pd_offset"].shape == (m, 3) and bool(
                (outs[0]["pd_mesh_xyz_left"][:, 0, 0] == torch.arange(m, dtype=torch.float32)).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [8, 7])
def test_sharded_gather_world2_gloo(n):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]


def test_hrnet_key_inventory_is_the_generators():
    """The shipped inventory of the HRNet extension equals what oracle/hrnet_oracle.py derives (python -m oracle.hrnet_oracle)."""
    import dir_b200
    from dir_b200.module import reference_key_shapes
    from oracle.synth import hrnet_key_shapes

    shipped = reference_key_shapes("hrnet_w32")
    assert shipped == hrnet_key_shapes(32)
    m = dir_b200.DIR(21, "./misc/mano", backbone="hrnet_w32")
    assert set(m.state_dict()) == set(shipped)
    assert tuple(m.state_dict()["init_regressor.mano_left.weight"].shape) == (64, 256)
    assert reference_key_shapes("hrnet_w48") == hrnet_key_shapes(48)
    m48 = dir_b200.DIR(21, "./misc/mano", backbone="hrnet_w48")
    assert tuple(m48.state_dict()["init_regressor.mano_left.weight"].shape) == (64, 384)
    assert tuple(m48.state_dict()["decoder.skip_layer3.conv1.conv.weight"].shape) == (128, 96, 1, 1)
    with pytest.raises(ValueError):
        dir_b200.DIR(21, "./misc/mano", backbone="hrnet_w64")
