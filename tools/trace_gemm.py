"""Timeline of CTA 0 of the persistent tcgen05 GEMM (debug trace): per k-block waits of every warp role."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gemm import gemm
from partdistillation_b200 import _lib
lib = _lib.load()
lib.pdb_debug_set_trace.argtypes = [ctypes.c_void_p]
M = 43008
x = torch.randn(1, M, 256, device="cuda"); w = torch.randn(1, 256, 256, device="cuda"); out = torch.empty(1, M, 256, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for flags in (0,):
    for _ in range(3):
        gemm(x, w, M, 256, 256, out=out)
    trace = torch.zeros(4 * 256 * 4, dtype=torch.int64, device="cuda")
    flush.zero_()
    lib.pdb_debug_set_trace(trace.data_ptr())
    gemm(x, w, M, 256, 256, out=out)
    torch.cuda.synchronize()
    lib.pdb_debug_set_trace(None)
    t = trace.cpu().view(4, 256, 4)
    t0 = int(t[0, 0, 0])
    r = lambda v: int(v) - t0 if int(v) else -1
    print(f"=== flags {flags}: it | TMA issue | split: start raw_full alo_empty done | MMA: start ready issued | (clk since first TMA)")
    for it in range(40):
        print(f"{it:3d} | {r(t[0,it,0]):6d} | {r(t[2,it,0]):6d} {r(t[2,it,1]):6d} {r(t[2,it,2]):6d} {r(t[2,it,3]):6d} | {r(t[1,it,0]):6d} {r(t[1,it,1]):6d} {r(t[1,it,2]):6d}  accwait@{r(t[1,it,3])}")
    print("epilogue tiles: wait_start acc_full done")
    for k in range(5):
        print(k, r(t[3, k, 0]), r(t[3, k, 1]), r(t[3, k, 2]))
