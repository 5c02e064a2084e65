"""Experiment: do two concurrent forwards on two streams hide the inter-kernel drain/fill? (images/s, resident inputs)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dir_b200
from oracle.synth import make_state_dict

dev = torch.device("cuda:0")
net = dir_b200.DIR(21, "./misc/mano", precision="bf16", aux_outputs=True, max_batch=128).to(dev)
net.load_state_dict(make_state_dict(0), strict=False)
net.eval()
g = torch.Generator().manual_seed(1)
xs = [torch.randn(128, 3, 256, 256, generator=g).to(dev) for _ in range(4)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def run(nb, per, K=30):
    def step(i):
        if nb == 1:
            net.run_raw(xs[i % 4][:per])
            return
        ev = torch.cuda.Event(); ev.record()
        for s in range(nb):
            streams[s].wait_event(ev)
            with torch.cuda.stream(streams[s]):
                net._workspace = wss[s]
                net.run_raw(xs[(i + 2 * s) % 4][s * per % 128: s * per % 128 + per] if per < 128 else xs[(i + 2 * s) % 4])
        for s in range(nb):
            torch.cuda.current_stream().wait_stream(streams[s])
    wss = [dict(), dict()]
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"streams={nb} batch/stream={per}: {ms:.3f} ms/step, {nb * per / ms * 1000:.0f} img/s")

run(1, 128)
run(1, 64)
run(2, 64)
run(2, 128)
