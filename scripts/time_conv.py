"""Times single conv layers through the conv_layer seam (CUDA events around the conv launch only, engine profile):
    python scripts/time_conv.py [--batch 128] key[:res] ...   e.g. backbone.layer3.1.conv3.weight:res
Used for epilogue / TMA experiments on individual layers. Test infrastructure (synthetic weights from oracle/)."""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import dir_b200  # noqa: E402
from dir_b200 import seams  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--precision", default="bf16")
ap.add_argument("keys", nargs="+")
a = ap.parse_args()
sd = make_state_dict(0)
net = dir_b200.DIR(21, "./misc/mano", precision=a.precision, max_batch=a.batch).cuda()
net.load_state_dict(sd, strict=False)
SIZE = {"layer1": 64, "layer2": 32, "layer3": 16, "layer4": 8}
for spec in a.keys:
    key, _, opt = spec.partition(":")
    cout, cin, kh, _ = sd[key].shape
    S = SIZE[key.split(".")[1]]
    x = torch.relu(torch.randn(a.batch, cin, S, S, device="cuda"))
    res = torch.randn(a.batch, cout, S, S, device="cuda") if opt == "res" else None
    h = net._ensure_handle() or net._handle
    times = []
    for i in range(8):
        h.profile_layer(key)
        seams.conv_layer(net, key, x, res)
        torch.cuda.synchronize()
        rows = h.profile_dump()
        h.profile_read()
        h.profile_layer(None)
        if i >= 2:
            times.append(rows[0]["ms"])
    r = rows[0]
    ms = statistics.median(times)
    print(f"{spec:48s} {ms * 1e3:7.1f} us  {r['flops'] / ms / 1e9:7.1f} TF/s  {r['bytes'] / ms / 1e6:6.0f} GB/s")
