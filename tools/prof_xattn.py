"""Driver for ncu captures of the masked cross-attention kernels (Lk = 16384, B = 2, Q = 100, 8 heads, 80 % masked)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import functional as fn  # noqa: E402
Lk = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
q = torch.randn(2, 100, 256, device="cuda").requires_grad_(); k = torch.randn(2, Lk, 256, device="cuda").requires_grad_()
v = torch.randn(2, Lk, 256, device="cuda").requires_grad_()
mask = (torch.rand(2, 100, Lk, device="cuda") < 0.8).to(torch.uint8)
ra = torch.ones(200, dtype=torch.int32, device="cuda")
for _ in range(3):
    out = fn.masked_cross_attention(q, k, v, mask, ra, 8)
    torch.autograd.grad(out, (q, k, v), torch.ones_like(out))
torch.cuda.synchronize()
