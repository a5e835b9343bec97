"""Same-box bars ("the reference PyTorch path on the B200"): the reference's own PyTorch expressions for
the three named operators, timed on the GPU next to this library's kernels at the C2 shapes (B = 2, 3 levels
32^2 / 64^2 / 128^2, 100 queries, 256^2 mask features).  The expressions are restated inline from the reference
(ops/functions/ms_deform_attn_func.py:55-75; mask2former_transformer_decoder.py:84,107-110,449) — this tool does not import
oracle/.  CUDA-event timing, L2 flushed between iterations; prints one JSON object, nothing here is a bench.py value.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import functional as fn  # noqa: E402
from tools.microbench import msda_case, timeit  # noqa: E402


def msda_core_pytorch(value, shapes, loc, attn):
    """ms_deform_attn_core_pytorch: per level grid_sample (bilinear, zeros, align_corners=False), weighted sum."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    grids = 2 * loc - 1
    outs = []
    for lid, v in enumerate(value.split([h * w for h, w in shapes], dim=1)):
        h, w = shapes[lid]
        v = v.flatten(2).transpose(1, 2).reshape(N * M, D, h, w)
        g = grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        outs.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    out = (torch.stack(outs, dim=-2).flatten(-2) * a).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()


def fwd_bwd(make_out, inputs, flush):
    with torch.no_grad():
        tf = timeit(make_out, iters=10, flush=flush)
    out = make_out()
    go = torch.randn_like(out)
    tb = timeit(lambda: torch.autograd.grad(out, inputs, go, retain_graph=True), iters=10, flush=flush)
    return tf * 1e6, tb * 1e6


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    res = {}
    shapes = [(32, 32), (64, 64), (128, 128)]
    value, loc, attn, shapes, fb, bb = msda_case(2, shapes, 4.0)
    rf, rb = fwd_bwd(lambda: msda_core_pytorch(value, shapes, loc, attn), (value, loc, attn), flush)
    nf, nb = fwd_bwd(lambda: fn.ms_deform_attn(value, shapes, None, loc, attn), (value, loc, attn), flush)
    res["msda_C2"] = dict(reference_pytorch_fwd_us=rf, reference_pytorch_bwd_us=rb, pdb_fwd_us=nf, pdb_bwd_us=nb,
                          fwd_speedup=rf / nf, bwd_speedup=rb / nb)

    B, Q, C, H, W = 2, 100, 256, 256, 256
    e = torch.randn(B, Q, C, device="cuda").requires_grad_()
    f = torch.randn(B, C, H, W, device="cuda").requires_grad_()
    fcl = f.detach().contiguous(memory_format=torch.channels_last).requires_grad_()
    rf, rb = fwd_bwd(lambda: torch.einsum("bqc,bchw->bqhw", e, f), (e, f), flush)
    nf, nb = fwd_bwd(lambda: fn.mask_einsum(e, fcl), (e, fcl), flush)
    res["mask_einsum"] = dict(reference_pytorch_fwd_us=rf, reference_pytorch_bwd_us=rb, pdb_fwd_us=nf, pdb_bwd_us=nb,
                              fwd_speedup=rf / nf, bwd_speedup=rb / nb)

    mha = torch.nn.MultiheadAttention(256, 8, dropout=0.0).cuda()
    for Lk in (1024, 4096, 16384):
        q = torch.randn(100, 2, 256, device="cuda").requires_grad_()
        k = torch.randn(Lk, 2, 256, device="cuda").requires_grad_()
        v = torch.randn(Lk, 2, 256, device="cuda").requires_grad_()
        keep = torch.rand(2, 100, Lk, device="cuda") < 0.8
        bool_mask = (~keep).unsqueeze(1).repeat(1, 8, 1, 1).flatten(0, 1)              # (B * heads, Q, Lk), True = masked
        rf, rb = fwd_bwd(lambda: mha(q, k, v, attn_mask=bool_mask)[0], (q, k, v), flush)   # in/out projections included
        qh, kh, vh = (t.detach().transpose(0, 1).contiguous().requires_grad_() for t in (q, k, v))
        mask = (~keep).to(torch.uint8).contiguous()
        ra = torch.ones(200, dtype=torch.int32, device="cuda")
        nf, nb = fwd_bwd(lambda: fn.masked_cross_attention(qh, kh, vh, mask, ra, 8), (qh, kh, vh), flush)
        res[f"masked_xattn_Lk{Lk}"] = dict(reference_mha_incl_projections_fwd_us=rf, reference_mha_incl_projections_bwd_us=rb,
                                           pdb_core_fwd_us=nf, pdb_core_bwd_us=nb)
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/reference_ops.json", "w"), indent=1)


if __name__ == "__main__":
    main()
