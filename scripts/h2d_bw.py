"""Pinned host->device copy bandwidth of one 100 MB batch (what bounds the fp32 e2e number)."""
import torch
x = torch.empty(128, 3, 256, 256).pin_memory()
d = torch.empty_like(x, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    d.copy_(x, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"H2D pinned: {x.numel() * 4 / ms / 1e6:.1f} GB/s ({ms:.3f} ms per 100.7 MB)")
