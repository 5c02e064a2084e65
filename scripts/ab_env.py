"""A/B of one engine switch on the whole forward: python scripts/ab_env.py DIRB200_NO_PREACT_FOLD [--batch 128] [--precision bf16]
Builds two modules in one process (the handle reads its switches at creation), alternates 5 x 20 timed forwards of each,
prints the median ms per forward of both. Test infrastructure (imports oracle/ for the weights)."""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dir_b200  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("env")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--backbone", default="resnet50")
args = ap.parse_args()
sd = make_state_dict(0, backbone=args.backbone) if args.backbone != "resnet50" else make_state_dict(0)


def build():
    m = dir_b200.DIR(21, "./misc/mano", precision=args.precision, max_batch=args.batch, backbone=args.backbone).cuda()
    m.load_state_dict(sd, strict=False)
    m._ensure_handle()
    return m


on = build()
name, _, value = args.env.partition("=")  # NAME or NAME=VALUE
os.environ[name] = value or "1"
off = build()
del os.environ[name]
args.env = name + "=" + (value or "1")
imgs = [torch.randn(args.batch, 3, 256, 256, generator=torch.Generator().manual_seed(i)).cuda() for i in range(4)]
res = {"default": [], args.env: []}
for m in (on, off):
    for i in range(5):
        m.run_raw(imgs[i % 4])
torch.cuda.synchronize()
for rnd in range(int(os.environ.get("AB_ROUNDS", "5"))):
    for name, m in (("default", on), (args.env, off)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            m.run_raw(imgs[i % 4])
        b.record()
        torch.cuda.synchronize()
        res[name].append(a.elapsed_time(b) / 20)
for k, v in res.items():
    print(f"{k:32s} median {statistics.median(v):.4f} ms/forward  (runs: {' '.join(f'{x:.3f}' for x in v)})  "
          f"{args.batch / statistics.median(v) * 1e3:.0f} images/s")
