"""GPU check of pdb_gemm_tf32x3 over every operand-layout combination against an fp64 product (and a timing
of the encoder-sized problems).  Usage: python tools/test_gemm.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import _lib  # noqa: E402


def gemm(A, B, M, N, K, batch=1, a_mn=0, b_mn=0, c_trans=0, bias=None, relu=0, accumulate=0, ksplit=1, out=None, B_lo=None):
    lib = _lib.load()
    lda = A.stride(-2)
    ldb = B.stride(-2)
    sa = A.stride(0) if batch > 1 else 0
    sb = B.stride(0) if batch > 1 else 0
    shape = (batch, N, M) if c_trans else (batch, M, N)
    if out is None:
        out = torch.zeros(shape, device="cuda") if accumulate else torch.empty(shape, device="cuda")
    ldc = out.stride(-2)
    sc = out.stride(0)
    rc = lib.pdb_gemm_tf32x3(A.data_ptr(), B.data_ptr(), B_lo.data_ptr() if B_lo is not None else None, out.data_ptr(),
                             bias.data_ptr() if bias is not None else None,
                             M, N, K, batch, lda, ldb, ldc, sa, sb, sc, a_mn, b_mn, c_trans, relu, accumulate, ksplit,
                             torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "pdb_gemm_tf32x3")
    return out


def split_lo(x):
    lo = torch.empty_like(x)
    _lib.check(_lib.load().pdb_split_lo(x.data_ptr(), lo.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream), "split_lo")
    return lo


def check(name, M, N, K, batch=1, a_mn=0, b_mn=0, c_trans=0, use_bias=False, relu=0, accumulate=0, ksplit=1, presplit=False):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((batch, K, M) if a_mn else (batch, M, K), device="cuda", generator=g)
    B = torch.randn((batch, K, N) if b_mn else (batch, N, K), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
    Am = A.double().transpose(1, 2) if a_mn else A.double()
    Bm = B.double().transpose(1, 2) if b_mn else B.double()
    ref = Am @ Bm.transpose(1, 2)
    if use_bias:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    if c_trans:
        ref = ref.transpose(1, 2)
    out = gemm(A, B, M, N, K, batch, a_mn, b_mn, c_trans, bias, relu, accumulate, ksplit, B_lo=split_lo(B) if presplit else None)
    torch.cuda.synchronize()
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"{name:40s} M={M} N={N} K={K} b={batch} a_mn={a_mn} b_mn={b_mn} ct={c_trans} ks={ksplit}: rel err {err:.2e}",
          "OK" if err < 1e-5 else "FAIL", flush=True)
    return err < 1e-5


def timeit(f, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        f()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); f(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    lib = _lib.load()
    ok = check("NT plain", 256, 128, 64)
    ok &= check("NT ragged + bias + relu", 300, 200, 100, use_bias=True, relu=1)
    ok &= check("NT batch c_trans (einsum fwd)", 1024, 100, 256, batch=2, c_trans=1)
    ok &= check("NT N=32 tile", 500, 24, 256, use_bias=True)
    ok &= check("A K-major, B MN-major (dgrad)", 300, 256, 200, b_mn=1)
    ok &= check("A MN, B MN (wgrad) split-K", 256, 256, 4000, a_mn=1, b_mn=1, accumulate=1, ksplit=8)
    ok &= check("A MN, B K-major", 384, 96, 128, a_mn=1)
    ok &= check("einsum grad_feat (A MN, B MN, batch)", 4096, 256, 100, batch=2, a_mn=1, b_mn=1)
    ok &= check("einsum grad_embed (B MN, split-K)", 100, 256, 4096, batch=2, b_mn=1, accumulate=1, ksplit=4)
    ok &= check("B MN ragged N=112-ish", 300, 100, 64, b_mn=1)
    ok &= check("many tiles persistent", 43008, 288, 256, use_bias=True, relu=1)
    for ps in (True,):
        ok &= check("presplit NT plain", 256, 128, 64, presplit=ps)
        ok &= check("presplit NT ragged + bias + relu", 300, 200, 100, use_bias=True, relu=1, presplit=ps)
        ok &= check("presplit einsum fwd", 1024, 100, 256, batch=2, c_trans=1, presplit=ps)
        ok &= check("presplit N=24", 500, 24, 256, use_bias=True, presplit=ps)
        ok &= check("presplit dgrad (B MN)", 300, 256, 200, b_mn=1, presplit=ps)
        ok &= check("presplit B MN ragged", 300, 100, 64, b_mn=1, presplit=ps)
        ok &= check("presplit B MN N=288", 4096, 288, 256, b_mn=1, presplit=ps)
        ok &= check("presplit many tiles", 43008, 288, 256, use_bias=True, relu=1, presplit=ps)
    ok &= check("einsum fwd full", 65536, 100, 256, batch=2, c_trans=1)
    # timing at the encoder shapes (rows = 43008)
    M = 43008
    x = torch.randn(1, M, 256, device="cuda")
    for N, K in ((256, 256), (1024, 256), (256, 1024), (288, 256)):
        xx = torch.randn(1, M, K, device="cuda")
        w = torch.randn(1, N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        out = torch.empty(1, M, N, device="cuda")
        t = timeit(lambda: gemm(xx, w, M, N, K, bias=bias, out=out))
        wl = split_lo(w)
        tp = timeit(lambda: gemm(xx, w, M, N, K, bias=bias, out=out, B_lo=wl))
        torch.backends.cuda.matmul.allow_tf32 = False
        t2 = timeit(lambda: torch.nn.functional.linear(xx[0], w[0], bias))
        fl = 2.0 * M * N * K
        print(f"linear {M}x{K} -> {N}: tcgen05 3xTF32 {t * 1e6:.1f} us ({fl / t / 1e12:.1f} TFLOP/s fp32-equivalent), pre-split B {tp * 1e6:.1f} us, "
              f"cuBLAS fp32 {t2 * 1e6:.1f} us ({fl / t2 / 1e12:.1f} TFLOP/s)", flush=True)
    # mask einsum forward / backward shapes
    e = torch.randn(2, 100, 256, device="cuda"); f = torch.randn(2, 65536, 256, device="cuda"); o = torch.empty(2, 100, 65536, device="cuda")
    t = timeit(lambda: gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, out=o))
    print(f"einsum fwd: {t * 1e6:.1f} us -> {186.9e6 / t / 1e9:.0f} GB/s algorithmic")
    el = split_lo(e)
    t = timeit(lambda: gemm(f, e, 65536, 100, 256, batch=2, c_trans=1, out=o, B_lo=el))
    print(f"einsum fwd, pre-split embed: {t * 1e6:.1f} us -> {186.9e6 / t / 1e9:.0f} GB/s algorithmic")
    print("RESULT", "PASS" if ok else "FAIL")


if __name__ == "__main__":
    main()
