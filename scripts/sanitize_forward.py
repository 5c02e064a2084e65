"""Small whole-forward run for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python scripts/sanitize_forward.py bf16"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dir_b200  # noqa: E402
from oracle.synth import make_state_dict  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
net = dir_b200.DIR(21, "./misc/mano", precision=precision, max_batch=4).cuda()
net.load_state_dict(make_state_dict(0), strict=False)
img = torch.randn(3, 3, 256, 256, generator=torch.Generator().manual_seed(0)).cuda()
o = net.run_raw(img)
torch.cuda.synchronize()
print(precision, "record finite:", bool(torch.isfinite(o["record"]).all()), "checksum", float(o["record"].double().sum()))
