import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from partdistillation_b200 import functional as fn
for rows, N in [(200, 256), (2048, 256), (8192, 256), (32768, 256), (131072, 256), (43008, 1024)]:
    x = torch.randn(rows, N, device="cuda")
    for f, name in ((lambda: fn.col_sum(x), "pdb"), (lambda: x.sum(0), "aten")):
        for _ in range(5):
            f()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(50):
            f()
        b.record()
        torch.cuda.synchronize()
        print(rows, N, name, f"{a.elapsed_time(b) / 50 * 1e3:.1f} us")
