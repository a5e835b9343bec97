"""tcgen05 GEMM at the mask-einsum and step shapes with A_lo in shared memory vs tensor memory
(pdb_debug_set_gemm_alo_tmem: 0 = shared memory, 2 = TMEM with the 3-stage ring, 1 = TMEM with the deep ring), plus the CTA-0 timeline of the einsum forward.  Usage: python tools/sweep_gemm.py [--trace]  (--trace needs a library built with PDB_GEMM_TRACE=1; an optional tools/libpdb200_old.so is timed next to the current build)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gemm import gemm as gemm_new, split_lo, timeit  # noqa: E402
from partdistillation_b200 import _lib  # noqa: E402

lib = _lib.load()
lib.pdb_debug_set_trace.argtypes = [ctypes.c_void_p]
if hasattr(lib, "pdb_debug_set_gemm_alo_tmem"):
    lib.pdb_debug_set_gemm_alo_tmem.argtypes = [ctypes.c_int]
else:
    lib.pdb_debug_set_gemm_alo_tmem = lambda v: 0
HAS_LA = hasattr(lib, "pdb_debug_set_gemm_lookahead")
if HAS_LA:
    lib.pdb_debug_set_gemm_lookahead.argtypes = [ctypes.c_int]
HAS_PF = hasattr(lib, "pdb_debug_set_gemm_prefetch")       # experiment knobs: only in experimental builds of the library
HAS_DBG = hasattr(lib, "pdb_debug_set_gemm_dbg")
if HAS_PF:
    lib.pdb_debug_set_gemm_prefetch.argtypes = [ctypes.c_int]
if HAS_DBG:
    lib.pdb_debug_set_gemm_dbg.argtypes = [ctypes.c_int]


def set_pf(v):
    if HAS_PF:
        lib.pdb_debug_set_gemm_prefetch(v)


_OLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpdb200_old.so")
old_lib = None
if os.path.exists(_OLD):          # optional A/B partner: an earlier build of the library (same C ABI), timed on the same GPU
    old_lib = ctypes.CDLL(_OLD)
    old_lib.pdb_gemm_tf32x3.restype, old_lib.pdb_gemm_tf32x3.argtypes = _lib.SIGNATURES["pdb_gemm_tf32x3"]
USE_OLD = False


def gemm(A, B, M, N, K, batch=1, a_mn=0, b_mn=0, c_trans=0, bias=None, relu=0, accumulate=0, ksplit=1, out=None, B_lo=None):
    if not USE_OLD:
        return gemm_new(A, B, M, N, K, batch, a_mn, b_mn, c_trans, bias, relu, accumulate, ksplit, out, B_lo)
    lda, ldb = A.stride(-2), B.stride(-2)
    sa = A.stride(0) if batch > 1 else 0
    sb = B.stride(0) if batch > 1 else 0
    rc = old_lib.pdb_gemm_tf32x3(A.data_ptr(), B.data_ptr(), B_lo.data_ptr() if B_lo is not None else None, out.data_ptr(),
                                 bias.data_ptr() if bias is not None else None, M, N, K, batch, lda, ldb, out.stride(-2), sa, sb,
                                 out.stride(0), a_mn, b_mn, c_trans, relu, accumulate, ksplit, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    return out


def cases():
    e = torch.randn(2, 100, 256, device="cuda"); f = torch.randn(2, 65536, 256, device="cuda")
    o = torch.empty(2, 100, 65536, device="cuda"); el = split_lo(e)
    yield "einsum fwd (presplit embed)", 186.9e6, lambda: gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, out=o, B_lo=el)
    go = torch.randn(2, 100, 65536, device="cuda"); gf = torch.empty(2, 65536, 256, device="cuda")
    yield "einsum grad_feat (A MN, B MN)", 4 * (2 * 100 * 65536 + 2 * 65536 * 256), lambda: gemm(go, e, 65536, 256, 100, batch=2, a_mn=1, b_mn=1, out=gf)
    ge = torch.zeros(2, 100, 256, device="cuda")
    yield "einsum grad_embed (B MN split-K 64)", 4 * (2 * 100 * 65536 + 2 * 65536 * 256), lambda: gemm(go, f, 100, 256, 65536, batch=2, b_mn=1, accumulate=1, ksplit=64, out=ge)
    M = 43008
    for N, K in ((256, 256), (1024, 256), (256, 1024)):
        xx = torch.randn(1, M, K, device="cuda"); w = torch.randn(1, N, K, device="cuda"); bias = torch.randn(N, device="cuda")
        out = torch.empty(1, M, N, device="cuda"); wl = split_lo(w)
        yield f"linear {M}x{K}->{N} presplit", 4 * (M * K + M * N), (lambda xx=xx, w=w, N=N, K=K, bias=bias, out=out, wl=wl: gemm(xx, w, M, N, K, bias=bias, out=out, B_lo=wl))
    # dgrad: dx = dy (M x N) . W (N x K): B MN-major
    dy = torch.randn(1, M, 1024, device="cuda"); w = torch.randn(1, 1024, 256, device="cuda"); dx = torch.empty(1, M, 256, device="cuda")
    yield "dgrad 43008x1024 -> 256 (B MN)", 4 * (M * 1024 + M * 256), lambda: gemm(dy, w, M, 256, 1024, b_mn=1, out=dx)
    # wgrad: dW (N x K) = dy^T (N x M) . x (M x K): A MN, B MN, split-K
    x = torch.randn(1, M, 256, device="cuda"); dw = torch.zeros(1, 1024, 256, device="cuda")
    yield "wgrad 1024x256 over 43008 (A MN, B MN, split-K 32)", 4 * (M * 1024 + M * 256), lambda: gemm(dy, x, 1024, 256, M, a_mn=1, b_mn=1, accumulate=1, ksplit=32, out=dw)


def trace_einsum(pf, alo=1):
    lib.pdb_debug_set_gemm_alo_tmem(alo)
    set_pf(pf)
    e = torch.randn(2, 100, 256, device="cuda"); f = torch.randn(2, 65536, 256, device="cuda")
    o = torch.empty(2, 100, 65536, device="cuda"); el = split_lo(e)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    run = lambda: gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, out=o, B_lo=el)
    for _ in range(3):
        run()
    trace = torch.zeros(4 * 256 * 4, dtype=torch.int64, device="cuda")
    flush.zero_()
    lib.pdb_debug_set_trace(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.pdb_debug_set_trace(None)
    t = trace.cpu().view(4, 256, 4)
    t0 = int(t[0, 0, 0])
    r = lambda v: int(v) - t0 if int(v) else -1
    ck = t[3, 200]
    dclk, dns = int(ck[2] - ck[0]), int(ck[3] - ck[1])
    print(f"=== einsum fwd, A_lo in TMEM {alo}, L2 prefetch {pf}: kernel body {dclk} clk in {dns} ns -> SM clock {dclk / max(dns, 1) * 1e3:.0f} MHz")
    print(f"=== einsum fwd, it | TMA issue | split: start raw_full alo_empty done | MMA: start ready issued | (clk since first TMA)")
    for it in range(56):
        print(f"{it:3d} | {r(t[0,it,0]):6d} | {r(t[2,it,0]):6d} {r(t[2,it,1]):6d} {r(t[2,it,2]):6d} {r(t[2,it,3]):6d} | {r(t[1,it,0]):6d} {r(t[1,it,1]):6d} {r(t[1,it,2]):6d}  accwait@{r(t[1,it,3])}")
    print("epilogue tiles: wait_start acc_full done")
    for k in range(7):
        print(k, r(t[3, k, 0]), r(t[3, k, 1]), r(t[3, k, 2]))


def check_einsum():
    e = torch.randn(2, 100, 256, device="cuda"); f = torch.randn(2, 65536, 256, device="cuda")
    el = split_lo(e)
    ref = torch.einsum("bqc,bpc->bqp", e.double(), f.double())
    for alo in (0, 2, 1):
        lib.pdb_debug_set_gemm_alo_tmem(alo)
        for lo in (None, el):
            o = gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, B_lo=lo)
            torch.cuda.synchronize()
            err = float((o.double() - ref).abs().max() / ref.abs().max())
            print(f"einsum fwd full, A_lo in TMEM {alo}, presplit {lo is not None}: rel err {err:.2e}", "OK" if err < 1e-5 else "FAIL", flush=True)
    lib.pdb_debug_set_gemm_alo_tmem(1)


def batched_einsum():
    """8 back-to-back launches per event pair over 4 feature buffers (the way bench.py times the roofline kernel)."""
    e = torch.randn(2, 100, 256, device="cuda"); el = split_lo(e)
    fs = [torch.randn(2, 65536, 256, device="cuda") for _ in range(4)]
    os_ = [torch.empty(2, 100, 65536, device="cuda") for _ in range(4)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for alo in (0, 2, 1):
        lib.pdb_debug_set_gemm_alo_tmem(alo)
        ts = []
        for rep in range(12):
            flush.zero_(); flush.zero_()
            s, en = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record()
            for j in range(8):
                gemm(fs[j % 4], e, 65536, 100, 256, batch=2, c_trans=1, out=os_[j % 4], B_lo=el)
            en.record()
            torch.cuda.synchronize()
            if rep >= 2:
                ts.append(s.elapsed_time(en) / 8 * 1e3)
        t = sum(ts) / len(ts)
        print(f"einsum fwd, 8 launches per event pair over 4 buffer sets, A_lo in TMEM {alo}: {t:.1f} us per launch -> {186.9e6 / t / 1e3:.0f} GB/s", flush=True)
    lib.pdb_debug_set_gemm_alo_tmem(1)


def ablate_einsum():
    """Where the einsum's time goes: switch off one stage at a time (results are garbage, timing only)."""
    e = torch.randn(2, 100, 256, device="cuda"); f = torch.randn(2, 65536, 256, device="cuda")
    o = torch.empty(2, 100, 65536, device="cuda"); el = split_lo(e)
    run = lambda: gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, out=o, B_lo=el)
    for flags, what in ((0, "full kernel"), (1, "no epilogue loads / stores"), (2, "no correction MMAs"), (4, "no split work"),
                        (3, "no epilogue, no correction MMAs"), (5, "no epilogue, no split"), (8, "no MMAs"), (13, "TMA + barriers only"),
                        (0, "full kernel again")):
        lib.pdb_debug_set_gemm_dbg(flags)
        t = timeit(run, iters=15)
        print(f"ablation dbg={flags:2d} ({what:34s}): {t * 1e6:6.1f} us", flush=True)
    lib.pdb_debug_set_gemm_dbg(0)


def main():
    check_einsum()
    if HAS_DBG:
        ablate_einsum()
    batched_einsum()
    global USE_OLD
    for name, nbytes, run in cases():
        row = []
        if old_lib is not None:
            for rep in range(2):           # old / new interleaved twice: order effects show up as differing repeats
                USE_OLD = True
                t = timeit(run, iters=15)
                USE_OLD = False
                row.append(f"old lib: {t * 1e6:6.1f}")
                t = timeit(run, iters=15)
                row.append(f"new lib: {t * 1e6:6.1f}")
        for alo in (0, 1):
            lib.pdb_debug_set_gemm_alo_tmem(alo)
            for la in ((0, 1) if HAS_LA else (0,)):
                if HAS_LA:
                    lib.pdb_debug_set_gemm_lookahead(la)
                t = timeit(run, iters=15)
                row.append(f"alo{alo} la{la}: {t * 1e6:6.1f}")
        set_pf(4)
        print(f"{name:52s} {' | '.join(row)}", flush=True)
    if "--trace" in sys.argv:
        trace_einsum(0, 1)
        trace_einsum(6, 1)
    set_pf(4)
    lib.pdb_debug_set_gemm_alo_tmem(1)


if __name__ == "__main__":
    main()
