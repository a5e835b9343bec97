"""Profiling driver for ncu: MSDeformAttn forward/backward at the C5(i) and C2 shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import msda_case
from partdistillation_b200 import functional as fn
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, N, shapes in (("C5i", 1, [(256, 256), (128, 128), (64, 64), (32, 32)]), ("C2", 2, [(32, 32), (64, 64), (128, 128)])):
    value, loc, attn, shapes, fb, bb = msda_case(N, shapes, 4.0, False)
    for _ in range(2):
        flush.zero_()
        out = fn.ms_deform_attn(value, shapes, None, loc, attn)
        go = torch.randn_like(out)
        flush.zero_()
        torch.autograd.grad(out, (value, loc, attn), go)
torch.cuda.synchronize()
