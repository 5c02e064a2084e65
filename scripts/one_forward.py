"""Runs a few whole forwards on cuda:0 (the command line ncu wraps for launch lists and --set full captures).
    python scripts/one_forward.py --precision bf16 --batch 128 --iters 4 [--no-aux]
Test infrastructure: uses the synthetic weights of oracle/synth.py."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dir_b200  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--precision", default="bf16")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--iters", type=int, default=4)
ap.add_argument("--no-aux", action="store_true")
ap.add_argument("--backbone", default="resnet50")
a = ap.parse_args()
net = dir_b200.DIR(21, "./misc/mano", precision=a.precision, aux_outputs=not a.no_aux, max_batch=a.batch,
                   backbone=a.backbone).cuda()
net.load_state_dict(make_state_dict(0, backbone=a.backbone) if a.backbone != "resnet50" else make_state_dict(0), strict=False)
img = torch.randn(a.batch, 3, 256, 256, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(a.iters):
    net.run_raw(img)
torch.cuda.synchronize()
h = net._handle
print("launches per forward:", h.lib.dirb200_forward_launches(h.h, a.batch))
