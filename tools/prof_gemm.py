"""Profiling driver for ncu: a few launches of the tcgen05 GEMM at the encoder-linear and mask-einsum shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gemm import gemm
M = 43008
x = torch.randn(1, M, 256, device="cuda"); w = torch.randn(1, 256, 256, device="cuda"); out = torch.empty(1, M, 256, device="cuda")
e = torch.randn(2, 100, 256, device="cuda"); f = torch.randn(2, 65536, 256, device="cuda"); o = torch.empty(2, 100, 65536, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_(); gemm(x, w, M, 256, 256, out=out)
    flush.zero_(); gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, out=o)
torch.cuda.synchronize()
