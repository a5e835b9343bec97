"""Eval-branch post-processing on one B200: the packed-mask path of this library against the
reference's dense PyTorch expression on the same GPU, per image at the BASELINE geometry (Q = 100 queries, 256^2 logits,
1024^2 image).  CUDA-event timing, L2 flushed between iterations.  Prints one JSON object; nothing here is a bench.py
value.      python tools/bench_postprocess.py [--queries 100] [--size 1024] [--gt 6]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import functional as fn  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def reference_dense(logits, scores, gate, gt, size):
    """proposal_model.py:225-230,243,375,258-302 + a device-side IoU in place of the pycocotools round trip (:414)."""
    up = F.interpolate(logits[None], size=size, mode="bilinear", align_corners=False)[0]
    up = F.interpolate(up[None], size=size, mode="bilinear", align_corners=False)[0]          # sem_seg_postprocess
    up = up * gate
    obj = up.topk(1, dim=0)[0] > 0
    masks = up > 0
    area = masks.flatten(1).sum(1) / obj.flatten(1).sum(1)
    a = masks.flatten(1).float()
    b = gt.flatten(1).float()
    inter = a @ b.t()
    iou = inter / (a.sum(1)[:, None] + b.sum(1)[None] - inter).clamp(min=1)
    return masks, area, iou


def packed(logits, sel, scores, gate, gt, size):
    bits, _ = fn.postprocess_masks(logits, sel, size, size, size, gate=gate, scores=scores, want_bits=True)
    K = sel.shape[0]
    counts = fn.bits_popcount(bits)
    iou = fn.bits_iou(bits[:K], fn.pack_bits(gt))
    return bits, counts[:K] / counts[K:], iou


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=100)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--gt", type=int, default=6)
    a = ap.parse_args()
    Q, S, G = a.queries, a.size, a.gt
    g = torch.Generator().manual_seed(0)
    low = torch.randn(Q, S // 16, S // 16, generator=g) * 3
    logits = (F.interpolate(low[None], size=(S // 4, S // 4), mode="bicubic")[0] + 0.3 * torch.randn(Q, S // 4, S // 4, generator=g)).cuda()
    scores = torch.rand(Q, generator=g).cuda()
    sel = torch.arange(Q, device="cuda")
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    gate = (((yy - S / 2) ** 2 + (xx - S / 2) ** 2) < (0.4 * S) ** 2).cuda()
    lab = torch.randint(0, G, (S // 16, S // 16), generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1)
    gt = torch.stack([lab == k for k in range(G)]).cuda() & gate
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    size = (S, S)
    with torch.no_grad():
        t_ref = timeit(lambda: reference_dense(logits, scores, gate, gt, size), iters=10, flush=flush)
        t_new = timeit(lambda: packed(logits, sel, scores, gate, gt, size), iters=20, flush=flush)
        t_kernel = timeit(lambda: fn.postprocess_masks(logits, sel, size, size, size, gate=gate, scores=scores), iters=20, flush=flush)
        t_label = timeit(lambda: fn.postprocess_masks(logits, sel, size, size, size, gate=gate, scores=scores, want_label=True),
                         iters=20, flush=flush)
        m_ref, area_ref, iou_ref = reference_dense(logits, scores, gate, gt, size)
        bits, area_new, iou_new = packed(logits, sel, scores, gate, gt, size)
        mism = float((fn.unpack_bits(bits[:Q], S) != m_ref).float().mean())
    alg = 4 * Q * (S // 4) ** 2 + S * S + (Q + 1) * S * S // 8
    print(json.dumps({
        "workload": f"eval post-processing, 1 image, Q={Q}, logits {S // 4}^2 -> {S}^2, {G} gt parts",
        "reference_dense_torch_ms": round(t_ref * 1e3, 3),
        "packed_path_ms": round(t_new * 1e3, 3),
        "postprocess_masks_kernel_ms": round(t_kernel * 1e3, 3),
        "postprocess_masks_kernel_with_label_ms": round(t_label * 1e3, 3),
        "kernel_algorithmic_bytes": alg,
        "kernel_achieved_gbs": round(alg / t_kernel * 1e-9, 1),
        "mask_bit_mismatch_fraction_vs_aten": mism,
        "max_iou_diff": float((iou_new - iou_ref.double()).abs().max()),
        "max_area_ratio_diff": float((area_new - area_ref).abs().max()),
    }))


if __name__ == "__main__":
    main()
