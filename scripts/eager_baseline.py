"""PyTorch-eager-on-B200 baseline (SURVEY 2 / BASELINE.md 3(4)): the reference's own op sequence (oracle restatement,
same ATen/cuDNN/cuBLAS calls as models/dir.py) executed on the GPU, timed with CUDA events.
Prints one JSON line per (batch, tf32) combination. Test infrastructure (imports oracle/)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dir_oracle as O  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402


def eager_rate(sd_gpu, batch, allow_tf32, autocast=False, steps=10, warmup=3):
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    imgs = [torch.randn(batch, 3, 256, 256, device=dev) for _ in range(4)]
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        for i in range(warmup):
            O.dir_forward(sd_gpu, imgs[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            O.dir_forward(sd_gpu, imgs[i % 4])
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"batch": batch, "conv_tf32": allow_tf32, "bf16_autocast": autocast, "ms_per_step": ms,
            "images_per_s": batch / ms * 1000.0}


if __name__ == "__main__":
    sd = {k: v.cuda() for k, v in make_state_dict(0).items()}
    for b in (32, 128):
        for tf32, ac in ((False, False), (True, False), (True, True)):
            print(json.dumps(eager_rate(sd, b, tf32, ac)), flush=True)
