"""Differential fuzzing of the masked cross-attention kernels' own code (csrc/xattn.cu, host build through
tests/native/cuda_on_cpu.h) against a float64 masked softmax attention with the all-masked-row reset
(mask2former_transformer_decoder.py:84,102-114,405): random batch / query / key counts (tile boundaries +-1, single keys),
random mask densities incl. fully masked rows and unmasked calls; forward and all three gradients.
    python tests/fuzz/fuzz_xattn_host.py [seconds]        # round 1: 367 cases in 150 s, worst error 19 % of tolerance
No GPU needed; test tooling only."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import pathlib  # noqa: E402
import tempfile  # noqa: E402

from host_kernels import build_host_library  # noqa: E402


def main():
    from partdistillation_b200 import _lib
    from partdistillation_b200 import functional as fn
    lib = build_host_library(pathlib.Path(tempfile.mkdtemp()), "xattn.cu", "xattn_section.inc", "xattn_kernels_host.cpp",
                             ("masked_xattn_workspace_bytes", "masked_xattn_forward", "masked_xattn_backward"))
    _lib.load = lambda: lib
    fn._need_cuda = lambda *a: None
    fn._stream = lambda: None
    g = torch.Generator().manual_seed(2024)
    heads, E = 8, 256
    t0, n, worst = time.time(), 0, [0.0] * 4
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    while time.time() - t0 < budget:
        B = int(torch.randint(1, 3, (1,), generator=g))
        Q = int(torch.randint(1, 140, (1,), generator=g))
        Lk = int(torch.randint(2, 200, (1,), generator=g))
        q = (torch.randn(B, Q, E, generator=g) * 0.3).double()
        k, v = torch.randn(B, Lk, E, generator=g).double(), torch.randn(B, Lk, E, generator=g).double()
        mask = row_any = None
        if torch.rand(1, generator=g) < 0.8:
            dens = float(torch.rand(1, generator=g))
            mask = (torch.rand(B, Q, Lk, generator=g) < dens).to(torch.uint8)
            if Q > 1:
                mask[0, 0] = 1
            row_any = (~mask.bool().all(-1)).view(-1).to(torch.int32)
        qr, kr, vr = (t.clone().requires_grad_() for t in (q, k, v))
        d = E // heads
        s = qr.view(B, Q, heads, d).transpose(1, 2) @ kr.view(B, Lk, heads, d).transpose(1, 2).transpose(-1, -2)
        if mask is not None:
            m = mask.bool().clone()
            m[m.all(-1)] = False
            s = s.masked_fill(m[:, None], float("-inf"))
        ref = (s.softmax(-1) @ vr.view(B, Lk, heads, d).transpose(1, 2)).transpose(1, 2).reshape(B, Q, E)
        go = torch.randn(ref.shape, generator=g).double()
        rg = torch.autograd.grad(ref, (qr, kr, vr), go)
        qc, kc, vc = (t.float().requires_grad_() for t in (q, k, v))
        out = fn.masked_cross_attention(qc, kc, vc, mask, row_any, heads)
        out_d = out.detach()
        gg = torch.autograd.grad(out, (qc, kc, vc), go.float())
        errs = [rel(out_d.double(), ref.detach())] + [rel(a.double(), b) for a, b in zip(gg, rg)]
        tols = [5e-6, 2e-5, 2e-5, 2e-5]
        # a gradient that is exactly 0 in exact arithmetic (e.g. dK with a single attended key) has no relative scale
        bad = [e > t and float(r.abs().max()) > 1e-6 for e, t, r in zip(errs, tols, [ref.detach(), *rg])]
        if any(bad):
            print("FAIL", B, Q, Lk, mask is not None, errs)
            return 1
        worst = [max(w, e / t) if float(r.abs().max()) > 1e-6 else w for w, e, t, r in zip(worst, errs, tols, [ref.detach(), *rg])]
        n += 1
    print("cases", n, "worst/tol", [round(w, 3) for w in worst])
    return 0


if __name__ == "__main__":
    sys.exit(main())
