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


def ref_cuda_op(res, flush):
    """The reference's own CUDA kernels (ops/src/cuda/ms_deform_im2col_cuda.cuh:242-304 forward, :306-408 backward for
    D = 32), compiled unmodified for sm_100a by baseline/build_ref_msda.sh, next to this library's kernels on the same
    inputs: C2 (3 levels, N = 2) and C5(i) (4 levels, N = 1)."""
    import ctypes as C
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "libref_msda.so")
    if not os.path.exists(so):
        res["ref_cuda_op"] = {"unavailable": "baseline/_ref/libref_msda.so not built (bash baseline/build_ref_msda.sh)"}
        return
    lib = C.CDLL(so)
    from partdistillation_b200 import _lib
    pdb = _lib.load()
    p = C.c_void_p
    lib.ref_msda_forward_f32.argtypes = [p] * 6 + [C.c_int] * 8 + [p]
    lib.ref_msda_backward_f32.argtypes = [p] * 9 + [C.c_int] * 8 + [p]
    for name, N, shapes in (("C2_3lvl_N2", 2, [(32, 32), (64, 64), (128, 128)]),
                            ("C5i_4lvl_N1", 1, [(256, 256), (128, 128), (64, 64), (32, 32)])):
        value, loc, attn, shapes, fb, bb = msda_case(N, shapes, 4.0)
        S, L = value.shape[1], len(shapes)
        sh = torch.tensor(shapes, dtype=torch.int64, device="cuda")
        st = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
        out = torch.zeros(N, S, 256, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        v, lo, a = value.detach(), loc.detach(), attn.detach()

        def ref_fwd():
            assert lib.ref_msda_forward_f32(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lo.data_ptr(), a.data_ptr(), out.data_ptr(),
                                            N, S, 8, 32, L, S, 4, 128, stream) == 0
        go = torch.randn(N, S, 256, device="cuda")
        gv, gl, ga = torch.zeros_like(v), torch.zeros_like(lo), torch.zeros_like(a)

        def ref_bwd():      # zero-filling the three gradient buffers is part of the reference's backward (at::zeros_like)
            gv.zero_(); gl.zero_(); ga.zero_()
            assert lib.ref_msda_backward_f32(v.data_ptr(), sh.data_ptr(), st.data_ptr(), lo.data_ptr(), a.data_ptr(), go.data_ptr(),
                                             gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), N, S, 8, 32, L, S, 4, 128, stream) == 0
        tf = timeit(ref_fwd, iters=10, flush=flush)
        tb = timeit(ref_bwd, iters=10, flush=flush)
        with torch.no_grad():
            ours = fn.ms_deform_attn(v, shapes, None, lo, a)
            nf = timeit(lambda: fn.ms_deform_attn(v, shapes, None, lo, a), iters=10, flush=flush)
            pdb.pdb_debug_set_msda_path(4)
            nf_tma = timeit(lambda: fn.ms_deform_attn(v, shapes, None, lo, a), iters=10, flush=flush)
            pdb.pdb_debug_set_msda_path(1)
            nf_l1 = timeit(lambda: fn.ms_deform_attn(v, shapes, None, lo, a), iters=10, flush=flush)
            pdb.pdb_debug_set_msda_path(0)
        o = fn.ms_deform_attn(value, shapes, None, loc, attn)
        nb = timeit(lambda: torch.autograd.grad(o, (value, loc, attn), go, retain_graph=True), iters=10, flush=flush)
        g_ours = torch.autograd.grad(o, (value, loc, attn), go, retain_graph=True)
        ref_fwd(); ref_bwd(); torch.cuda.synchronize()
        rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
        res["ref_cuda_op_" + name] = dict(
            ref_cuda_sm100a_fwd_us=tf * 1e6, ref_cuda_sm100a_bwd_incl_zero_fill_us=tb * 1e6, pdb_fwd_us=nf * 1e6,
            pdb_fwd_tma_tiles_us=nf_tma * 1e6, pdb_fwd_l1_tiled_us=nf_l1 * 1e6, pdb_bwd_us=nb * 1e6, fwd_speedup=tf / nf,
            bwd_speedup=tb / nb, fwd_algorithmic_bytes=fb, bwd_algorithmic_bytes=bb,
            out_rel_diff=rel(ours, out.view_as(ours)), grad_value_rel_diff=rel(g_ours[0], gv), grad_loc_rel_diff=rel(g_ours[1], gl),
            grad_attn_rel_diff=rel(g_ours[2], ga))


def oracle_step_on_gpu(res):
    """The reference algorithm as plain PyTorch on THIS GPU (SURVEY.md §8c's GPU-PyTorch bar): the oracle's restatement of
    the reference modules (pinned to the unmodified reference by tests/test_oracle_golden.py) with every tensor on the
    device, fp32, TF32 off — forward + loss + backward of configs[1]'s batch of 2 images, eager ATen / cuBLAS / cuDNN kernels."""
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle"))
    sys.path.insert(0, root)
    try:
        import bench
        import m2f_oracle as O
        from partdistillation_b200 import compat, presets
        cfg = presets.make_cfg("ProposalModel", "swin_b", 100, 10, 12544, 0.0, device="cpu")
        torch.manual_seed(0)
        sd = {k: v.detach().cuda() for k, v in compat.META_ARCH_REGISTRY.get("ProposalModel")(cfg).state_dict().items()}
        for k, v in sd.items():
            if k.startswith("sem_seg_head.") and v.is_floating_point():
                v.requires_grad_(True)
        hp = dict(num_classes=1, dec_layers=10, num_points_match=12544, num_points_loss=12544, w_class=2.0, w_mask=5.0,
                  w_dice=5.0, eos_coef=0.1, oversample_ratio=3.0, importance_ratio=0.0)
        batch = []
        for i in range(2):
            img, m = bench.synth_image_and_masks(i)
            batch.append({"image": img.float().cuda(), "gt_masks": m.cuda()})
        mean, std = [123.675, 116.280, 103.530], [58.395, 57.120, 57.375]

        def one():
            for v in sd.values():
                v.grad = None
            with torch.device("cuda"):           # the oracle creates its index / position tensors with the default device
                x = O.prepare_images(batch, mean, std, 32)
                with torch.no_grad():
                    feats = O.swin_forward(sd, "backbone.", x, 128, [2, 2, 18, 2], [4, 8, 16, 32], 12)
                losses = O.head_and_loss(sd, feats, O.prepare_targets(batch, 1024, 1024), hp)
            sum(losses.values()).backward()
        for _ in range(2):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            one()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        res["reference_pytorch_gpu_step"] = dict(ms_per_step=dt * 1e3, images_per_s=2 / dt,
                                                 what="oracle/m2f_oracle.py on cuda:0, fwd + loss + bwd of 2 x 1024^2 images, fp32, "
                                                      "no optimizer step, frozen Swin-B under no_grad")
    except Exception as ex:          # the oracle is CPU test infrastructure; a device mismatch inside it is reported, not fatal
        import traceback
        res["reference_pytorch_gpu_step"] = {"failed": f"{type(ex).__name__}: {ex}"[:300], "where": traceback.format_exc()[-700:]}


def main():
    if "--oracle-step-only" in sys.argv:
        res = {}
        oracle_step_on_gpu(res)
        print(json.dumps(res, indent=1))
        return
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
    ref_cuda_op(res, flush)
    oracle_step_on_gpu(res)
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/reference_ops.json", "w"), indent=1)


if __name__ == "__main__":
    main()
