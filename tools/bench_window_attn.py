"""Timing of the fused Swin window attention at the Swin-B 1024x1024 stage shapes (B=2)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import functional as fn
def timeit(f, iters=10):
    for _ in range(3): f()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts)//2] * 1e3
tot = 0
for H, heads, blocks in ((256, 4, 2), (128, 8, 2), (64, 16, 18), (32, 32, 2)):
    C = heads * 32
    qkv = torch.randn(2, H, H, 3 * C, device="cuda"); qb = torch.randn(3 * C, device="cuda"); bias = torch.randn(heads, 144, 144, device="cuda")
    t0 = timeit(lambda: fn.swin_window_attention(qkv, qb, bias, heads, 12, 0, 32 ** -0.5))
    t1 = timeit(lambda: fn.swin_window_attention(qkv, qb, bias, heads, 12, 6, 32 ** -0.5))
    print(f"stage {H}x{H} heads {heads}: {t0:.1f} us plain, {t1:.1f} us shifted")
    tot += (t0 + t1) / 2 * blocks
print(f"total per backbone forward: {tot / 1e3:.2f} ms")
