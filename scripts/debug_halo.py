import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import dir_b200
from dir_b200 import seams
from oracle.synth import make_state_dict
from test_gpu_conv import expected
sd = make_state_dict(0)
m = dir_b200.DIR(21, "./misc/mano", precision="bf16", max_batch=8).cuda(); m.load_state_dict(sd, strict=False)
key = "backbone.layer1.0.conv2.weight"
for (B, H, W) in ((2, 64, 64), (1, 32, 32), (3, 16, 16)):
    g = torch.Generator().manual_seed(1)
    x = torch.relu(torch.randn(B, 64, H, W, generator=g))
    want = expected(sd, key, x, None, round_bf16=True)
    got, used = seams.conv_layer(m, key, x.cuda())
    got = got.cpu()
    err = (got - want).abs()
    print(f"B{B} H{H} W{W} used={used} max err {float(err.max()):.4f} of max {float(want.abs().max()):.3f}; rel {float(err.max()/want.abs().max()):.3e}")
    print(" by x :", [f"{float(v):.2f}" for v in err.amax(dim=(0, 1, 2))[:8]], "...", [f"{float(v):.2f}" for v in err.amax(dim=(0, 1, 2))[-4:]])
    print(" by y :", [f"{float(v):.2f}" for v in err.amax(dim=(0, 1, 3))[:8]], "...", [f"{float(v):.2f}" for v in err.amax(dim=(0, 1, 3))[-4:]])
    print(" by ch:", [f"{float(v):.2f}" for v in err.amax(dim=(0, 2, 3))[:16]])
    print(" by b :", [f"{float(v):.2f}" for v in err.amax(dim=(1, 2, 3))])
    # delta-input probe: which taps are seen? single nonzero pixel at (y=5,x=7), channel 3
    xd = torch.zeros(1, 64, H, W); xd[0, 3, 5, 7] = 1.0
    wd = expected(sd, key, xd, None, round_bf16=True); gd, _ = seams.conv_layer(m, key, xd.cuda()); gd = gd.cpu()
    base = expected(sd, key, torch.zeros(1, 64, H, W), None, round_bf16=True)
    nzw = ((wd - base).abs() > 1e-6).any(dim=1)[0].nonzero().tolist()
    nzg = ((gd - base).abs() > 1e-3).any(dim=1)[0].nonzero().tolist()
    print(" delta probe: want nonzero at", nzw[:12], " got nonzero at", nzg[:20])

print("---- structured probes (H=W=64, B=1)")
H = W = 64
for name, xin in (("zeros", torch.zeros(1, 64, H, W)), ("ones", torch.ones(1, 64, H, W)),
                  ("ch0 only", torch.cat([torch.ones(1, 1, H, W), torch.zeros(1, 63, H, W)], 1)),
                  ("ch8 only", torch.cat([torch.zeros(1, 8, H, W), torch.ones(1, 1, H, W), torch.zeros(1, 55, H, W)], 1))):
    wd = expected(sd, key, xin, None, round_bf16=True)
    gd, _ = seams.conv_layer(m, key, xin.cuda()); gd = gd.cpu()
    print(name, "want[0,:4,10,10]", [f"{float(v):.3f}" for v in wd[0, :4, 10, 10]], "got", [f"{float(v):.3f}" for v in gd[0, :4, 10, 10]],
          "| got[0,:4,30,31]", [f"{float(v):.3f}" for v in gd[0, :4, 30, 31]], "max err", f"{float((gd - wd).abs().max()):.3f}")
