"""Differential fuzzing of the inference post-processing kernels' own code (csrc/postprocess_kernels.cuh, host build through
tests/native/cuda_on_cpu.h) against the oracle expression F.interpolate -> crop -> F.interpolate -> gate -> threshold /
score * sigmoid arg-max, popcounts and pycocotools-style IoU: random logit sizes, padded / image / output geometries (up- and
down-sampling, 1-pixel outputs, widths around word boundaries), gated and ungated, repeated queries in the selection.
    python tests/fuzz/fuzz_postprocess_host.py [seconds]        # round 1: 904 random geometries in 150 s, all checks pass, 0 flips
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
from host_kernels import POSTPROCESS_OPS, build_plain_harness  # noqa: E402


def unpack(bits, width):
    b = bits.to(torch.int64) & 0xFFFFFFFF
    return ((b[..., None] >> torch.arange(32)) & 1).bool().flatten(-2)[..., :width]


def main():
    from partdistillation_b200 import _lib
    from partdistillation_b200 import functional as fn
    lib = build_plain_harness(pathlib.Path(tempfile.mkdtemp()), "postprocess_kernels_host.cpp", POSTPROCESS_OPS)
    _lib.load = lambda: lib
    fn._need_cuda = lambda *a: None
    fn._stream = lambda: None
    g = torch.Generator().manual_seed(77)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=g))
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 180.0
    t0, n, flips_total, px_total = time.time(), 0, 0, 0
    while time.time() - t0 < budget:
        h, w = ri(1, 12), ri(1, 12)
        Hp, Wp = h * ri(1, 5), w * ri(1, 5)
        Hi, Wi = ri(max(1, Hp // 2), Hp + 1), ri(max(1, Wp // 2), Wp + 1)
        if torch.rand(1, generator=g) < 0.3:
            Ho, Wo = Hi, Wi
        else:
            Ho, Wo = ri(1, 2 * Hi + 2), ri(1, 2 * Wi + 2)
        Q, K = ri(1, 7), ri(1, 9)
        logits = torch.randn(Q, h, w, generator=g) * 2
        sel = torch.randint(0, Q, (K,), generator=g)
        scores = torch.rand(K, generator=g) + 0.05
        gate = (torch.rand(Ho, Wo, generator=g) > 0.3) if torch.rand(1, generator=g) < 0.6 else None
        bits, label, sb = fn.postprocess_masks(logits, sel, (Hp, Wp), (Hi, Wi), (Ho, Wo), gate=gate, scores=scores,
                                               want_bits=True, want_label=True, score_threshold=0.5)
        up = F.interpolate(logits[None], size=(Hp, Wp), mode="bilinear", align_corners=False)[0]
        ref = O.sem_seg_postprocess(up, (Hi, Wi), Ho, Wo)[sel]
        if gate is not None:
            ref = ref * gate
        got = unpack(bits, Wo)
        flips = got[:K] != (ref > 0)
        sm = scores[:, None, None] * ref.sigmoid()
        ok = not (flips & (ref.abs() > 1e-4)).any()
        ok &= bool(torch.equal(got[K], got[:K].any(0)))
        ok &= not unpack(bits, 32 * bits.shape[-1])[..., Wo:].any()
        if K > 1:
            top2 = sm.topk(2, dim=0)[0]
            ok &= not ((label.long() != sm.argmax(0)) & ((top2[0] - top2[1]) > 1e-5)).any()
        else:
            ok &= bool((label == 0).all())
        ok &= not ((unpack(sb, Wo) != (sm > 0.5)) & ((sm - 0.5).abs() > 1e-5)).any()
        ok &= bool(torch.equal(fn.bits_popcount(bits), got.flatten(1).sum(1)))
        gt = torch.rand(ri(1, 5), Ho, Wo, generator=g) > 0.5
        ok &= bool(torch.equal(fn.bits_iou(bits[:K], fn.pack_bits(gt)), O.mask_iou(got[:K], gt)))
        ok &= bool(torch.equal(fn.unpack_bits(fn.pack_bits(gt), Wo), gt))
        m8 = torch.zeros(2, Hp, Wp, dtype=torch.bool)
        m8[:, :Hi, :Wi] = torch.rand(2, Hi, Wi, generator=g) > 0.6
        exp = O.sem_seg_postprocess(m8.float(), (Hi, Wi), Ho, Wo)
        rz = fn.resize_bool_masks(m8, (Hi, Wi), (Ho, Wo))
        ok &= not ((rz != exp.bool()) & (exp.abs() > 1e-6)).any()      # `.bool()` of an interpolated float: exact away from 0
        if not ok:
            print("FAIL", dict(h=h, w=w, padded=(Hp, Wp), image=(Hi, Wi), out=(Ho, Wo), Q=Q, K=K, gated=gate is not None))
            return 1
        flips_total += int(flips.sum())
        px_total += flips.numel()
        n += 1
    print("cases", n, "threshold flips inside the noise band", flips_total, "of", px_total)
    return 0


if __name__ == "__main__":
    sys.exit(main())
