"""PyTorch-eager baseline of the HRNet-W32 extension on the GPU: the self-authored oracle's op sequence (oracle/hrnet_oracle.py)
timed like scripts/eager_baseline.py. One JSON line per mode. Test infrastructure."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as H  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

sd = {k: v.cuda() for k, v in make_state_dict(0, backbone="hrnet_w32").items()}
torch.backends.cudnn.benchmark = True
for batch in (32, 128):
    imgs = [torch.randn(batch, 3, 256, 256, device="cuda") for _ in range(2)]
    for name, tf32, ac in (("fp32", False, False), ("tf32_convs_torch_default", True, False), ("bf16_autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
            for i in range(3):
                H.dir_forward(sd, imgs[i % 2], 32)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(6):
                H.dir_forward(sd, imgs[i % 2], 32)
            e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"backbone": "hrnet_w32", "batch": batch, "mode": name,
                          "images_per_s": batch * 6 / (e0.elapsed_time(e1) / 1000.0)}), flush=True)
