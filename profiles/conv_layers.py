"""Per-conv-layer roofline table of one eager forward (CUDA events around every conv launch, warm):
   python profiles/conv_layers.py [--batch 128] [--precision bf16]  > profiles/conv_layers_rNN.txt
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dir_b200  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()
pk = {"tflops": 1364.7, "hbm": 6549.4}
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    d = json.load(open(p))
    pk = {"tflops": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"]}
# tf32 MMA issue rate measured by scripts/mma_probe.cu (profiles/mma_probe_r2.txt): 3064 flop/clk/SM = 891 TFLOP/s at
# 1965 MHz; precision fp32 spends three tf32 MMAs per algorithmic MAC (3xTF32), precision tf32 one
if args.precision == "fp32":
    pk["tflops"] = 891.0 / 3
elif args.precision == "tf32":
    pk["tflops"] = 891.0
net = dir_b200.DIR(21, "./misc/mano", precision=args.precision, max_batch=args.batch).cuda()
net.load_state_dict(make_state_dict(0), strict=False)
img = torch.randn(args.batch, 3, 256, 256, generator=torch.Generator().manual_seed(0)).cuda()
for _ in range(3):
    net.run_raw(img)
torch.cuda.synchronize()
h = net._handle
h.profile_layer("")
net.run_raw(img)
torch.cuda.synchronize()
rows = h.profile_dump()
h.profile_read()
h.profile_layer(None)
tot = sum(r["ms"] for r in rows)
floor = 0.0
print(f"# B={args.batch} {args.precision}; peaks: {pk['tflops']:.1f} algorithmic TFLOP/s "
      f"({'cuBLAS bf16 sustained' if args.precision == 'bf16' else 'tf32 MMA issue rate / MMAs per MAC'}), {pk['hbm']} GB/s (measured)")
print(f"{'layer':58s} tc {'k':>3s} {'cin':>5s} {'cout':>5s} {'ms':>7s} {'TF/s':>7s} {'GB/s':>7s} {'floor_ms':>8s} {'x floor':>7s} bound")
for r in rows:
    t_c = r["flops"] / (pk["tflops"] * 1e12) * 1e3
    t_m = r["bytes"] / (pk["hbm"] * 1e9) * 1e3
    fl = max(t_c, t_m)
    floor += fl
    print(f"{r['layer'][:58]:58s} {r['tc']:2d} {r['kernel']:>3s} {r['cin']:5d} {r['cout']:5d} {r['ms']:7.3f} "
          f"{r['flops'] / r['ms'] / 1e9:7.1f} {r['bytes'] / r['ms'] / 1e6:7.0f} {fl:8.3f} {r['ms'] / fl:7.2f} "
          f"{'tensor' if t_c > t_m else 'hbm'}")
print(f"# total conv ms {tot:.3f}; sum of per-layer roofline floors {floor:.3f} ms; ratio {tot / floor:.2f}")
