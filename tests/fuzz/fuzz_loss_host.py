"""Differential fuzzing of the loss-side kernels' own code (csrc/loss.cu, host build through tests/native/cuda_on_cpu.h)
against the oracle: point sampling (float and 0/1 maps, gathers, coordinates outside [0, 1]), matcher cost matrices over
ragged target counts, and the fused point BCE + dice with its gradient — random, non-square shapes and point counts.
    python tests/fuzz/fuzz_loss_host.py [seconds]        # round 1: 240 cases in 120 s, all within tolerance
No GPU needed; test tooling only (the oracle is the checker)."""
import os
import pathlib
import sys
import tempfile
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import m2f_oracle as O  # noqa: E402
from host_kernels import build_host_library  # noqa: E402

OPS = ("point_sample_forward", "point_sample_backward", "matcher_cost", "lsap_batched", "point_loss_forward",
       "point_loss_backward", "class_rows_forward", "class_rows_backward")


def main():
    from partdistillation_b200 import _lib
    from partdistillation_b200 import functional as fn
    lib = build_host_library(pathlib.Path(tempfile.mkdtemp()), "loss.cu", "loss_section.inc", "loss_kernels_host.cpp", OPS)
    _lib.load = lambda: lib
    fn._need_cuda = lambda *a: None
    fn._stream = lambda: None
    g = torch.Generator().manual_seed(31)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
    t0, n = time.time(), 0
    while time.time() - t0 < budget:
        H, W, P = ri(1, 20), ri(1, 20), ri(1, 70)
        # ---- point sampling with gradient
        R = ri(1, 6)
        img = torch.randn(R, H, W, generator=g)
        pts = torch.rand(R, P, 2, generator=g) * 1.4 - 0.2
        ir = img.clone().requires_grad_()
        ref = O.point_sample(ir[:, None], pts).squeeze(1)
        ic = img.clone().requires_grad_()
        out = fn.point_sample(ic, pts)
        go = torch.randn(out.shape, generator=g)
        ok = torch.allclose(out.detach(), ref.detach(), rtol=1e-5, atol=1e-6)
        ok &= torch.allclose(torch.autograd.grad(out, ic, go)[0], torch.autograd.grad(ref, ir, go)[0], rtol=1e-4, atol=1e-5)
        # ---- matcher cost over ragged target counts (incl. 0)
        B, Q = ri(1, 4), ri(1, 9)
        Ks = [ri(0, 5) for _ in range(B)]
        if sum(Ks):
            logits = torch.randn(B, Q, 3, generator=g)
            pm = torch.randn(B, Q, H, W, generator=g) * 3
            Hg, Wg = H * ri(1, 4), W * ri(1, 4)
            tm = [torch.rand(k, Hg, Wg, generator=g) > 0.5 for k in Ks]
            labels = [torch.randint(0, 2, (k,), generator=g) for k in Ks]
            coords = [torch.rand(1, P, 2, generator=g) for _ in range(B)]
            off = [0]
            for k in Ks:
                off.append(off[-1] + k)
            call = torch.cat(coords)
            pred_pts = fn.point_sample(pm.flatten(0, 1), call, None, torch.arange(B).repeat_interleave(Q).int())
            tgt_pts = fn.point_sample(torch.cat(tm).to(torch.uint8), call, None,
                                      torch.arange(B).repeat_interleave(torch.tensor(Ks)).int())
            cost = fn.matcher_cost(pred_pts, tgt_pts, logits.softmax(-1).flatten(0, 1), torch.cat(labels).int(), off, Q,
                                   2.0, 5.0, 5.0)
            for b in range(B):
                if Ks[b]:
                    C = O.matcher_costs(logits[b], pm[b], labels[b], tm[b], coords[b], 2.0, 5.0, 5.0)
                    ok &= torch.allclose(cost[Q * off[b]: Q * off[b + 1]].view(Q, Ks[b]), C, rtol=1e-4, atol=1e-5)
            # ---- fused point loss on random (prediction, target) pairs
            Nm = ri(1, 5)
            pidx = torch.randint(0, B * Q, (Nm,), generator=g)
            gidx = torch.randint(0, off[-1], (Nm,), generator=g)
            pc = torch.rand(Nm, P, 2, generator=g)
            pr = pm.clone().requires_grad_()
            lg = O.point_sample(pr.flatten(0, 1)[pidx][:, None], pc).squeeze(1)
            lb = O.point_sample(torch.cat(tm)[gidx][:, None].float(), pc).squeeze(1)
            bce = F.binary_cross_entropy_with_logits(lg, lb, reduction="none").mean(1)
            s = lg.sigmoid()
            dice = 1 - (2 * (s * lb).sum(-1) + 1) / (s.sum(-1) + lb.sum(-1) + 1)
            wb, wd = torch.rand(Nm, generator=g), torch.rand(Nm, generator=g)
            ((bce * wb).sum() + (dice * wd).sum()).backward()
            pk = pm.clone().requires_grad_()
            mb, md = fn.point_loss(pk.flatten(0, 1), pidx, torch.cat(tm).to(torch.uint8), gidx, pc)
            ok &= torch.allclose(mb.detach(), bce.detach(), rtol=1e-5, atol=1e-6)
            ok &= torch.allclose(md.detach(), dice.detach(), rtol=1e-5, atol=1e-6)
            ((mb * wb).sum() + (md * wd).sum()).backward()
            ok &= torch.allclose(pk.grad, pr.grad, rtol=1e-4, atol=1e-6)
        if not ok:
            print("FAIL", dict(H=H, W=W, P=P, R=R, B=B, Q=Q, Ks=Ks))
            return 1
        n += 1
    print("cases", n, "all within tolerance")
    return 0


if __name__ == "__main__":
    sys.exit(main())
