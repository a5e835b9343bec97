"""Decoder-shaped Linear products: the mma.sync short-A kernel against the persistent tcgen05 kernel (CUDA events, L2-warm as in
the step, 200 launches each).  python tools/bench_small_gemm.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from partdistillation_b200 import functional as fn


def t(f, n=50, reps=10):
    """us per call inside a replayed CUDA graph of n calls (the step runs as a graph: no host launch cost)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            f()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        for _ in range(n):
            f()
    graph.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (n * reps) * 1e3


for M in (200, 400, 64):
    for (N, K, b_mn, relu) in [(256, 256, 0, 0), (256, 256, 1, 0), (2048, 256, 0, 1), (256, 2048, 0, 0), (256, 2048, 1, 0),
                               (2048, 256, 1, 0), (768, 256, 0, 0)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        C = torch.empty(M, N, device="cuda")
        small = t(lambda: fn.gemm_small(A, B, M, N, K, lda=K, ldb=N if b_mn else K, b_mn=bool(b_mn), bias=bias, relu=relu))
        tc = t(lambda: fn.gemm_tf32x3(A, B, C, M, N, K, lda=K, ldb=N if b_mn else K, ldc=N, b_mn=bool(b_mn), bias=bias, relu=relu))
        print(f"M={M} N={N} K={K} b_mn={b_mn} relu={relu}: small {small:.1f} us  tcgen05 {tc:.1f} us")
