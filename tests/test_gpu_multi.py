"""GPU, needs >= 2 devices (skipped on a single-GPU box): batch-sharded forward + the library's NCCL all-gather
must reproduce the single-GPU records bit for bit."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import dir_b200
    from dir_b200.dist import forward_sharded, init_nccl
    from oracle.synth import make_state_dict

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        net = dir_b200.DIR(21, "./misc/mano", precision="fp32", aux_outputs=False, max_batch=8).to(f"cuda:{rank}")
        net.load_state_dict(make_state_dict(0), strict=False)
        init_nccl(net)
        img = torch.randn(7, 3, 256, 256, generator=torch.Generator().manual_seed(5)).to(f"cuda:{rank}")
        outs = forward_sharded(net, img)  # ragged: 4 + 3 images
        single = net.unpack_record(net.run_raw(img)["record"])
        ok = all(torch.equal(outs[i][k], single[i][k]) for i in range(3) for k in outs[i] if outs[i][k] is not None)
        torch.cuda.synchronize()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_forward_equals_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert res == [(0, True), (1, True)]
