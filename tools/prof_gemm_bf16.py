"""Driver for ncu / timing of the bf16 tcgen05 GEMM at the C3 backbone shapes."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import functional as fn  # noqa: E402
from tools.microbench import timeit  # noqa: E402

g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rows, K, N, act, odt in ((51200, 512, 2048, 2, torch.bfloat16), (51200, 512, 2048, 0, torch.bfloat16), (51200, 512, 2048, 0, torch.float32),
                             (51200, 2048, 512, 0, torch.bfloat16), (204800, 256, 768, 0, torch.bfloat16), (3200, 256, 2048, 1, torch.bfloat16),
                             (3200, 2048, 256, 0, torch.bfloat16)):
    x = torch.randn(rows, K, generator=g).cuda().to(torch.bfloat16)
    w = torch.randn(N, K, generator=g).cuda().to(torch.bfloat16)
    b = torch.randn(N, generator=g).cuda()
    t = timeit(lambda: fn.gemm_bf16(x, w, b, act, odt), flush=flush)
    xf, wf = x.float(), w.float()
    tt = timeit(lambda: torch.nn.functional.linear(x, w, b.to(torch.bfloat16)), flush=flush)
    print(f"{rows}x{K}->{N} act={act} out={odt}: {t*1e6:.1f} us = {2.0*rows*K*N/t/1e12:.0f} TFLOP/s; torch (cuBLAS bf16, no act) {tt*1e6:.1f} us", flush=True)
