"""Host-side operators: torch.autograd Functions over the C ABI of libpdb200.so.

Every function here requires CUDA tensors and the compiled library; nothing falls back to PyTorch
or to the CPU (the reference's native op is CUDA-only as well: ops/src/ms_deform_attn.h:44).
"""
import os

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib

_DT = {torch.float32: 0, torch.float64: 1}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("partdistillation_b200 operators are CUDA-only (no CPU implementation)")


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


_table_cache = {}


def host_table(values, dtype, device):
    """Small host-side index table (python ints / floats) as a device tensor.  Tables are cached by content: the
    per-step tables of the criterion / matcher depend only on the per-image target counts, so steady-state
    steps issue no host->device copy for them (and a CUDA-graph capture of the step contains none)."""
    key = (tuple(values), dtype, str(device))
    t = _table_cache.get(key)
    if t is None:
        if len(_table_cache) > 512:
            _table_cache.clear()
        t = _table_cache[key] = torch.tensor(list(values), dtype=dtype).to(device)
    return t


# --------------------------------------------------------------------------------------------------
# MSDeformAttn  (ops/functions/ms_deform_attn_func.py:35-52 — same call signature)
# --------------------------------------------------------------------------------------------------
_shape_cache = {}


def _host_levels(spatial_shapes, level_start_index):
    """Accepts python sequences or the reference's int64 device tensors; the tensor form costs one
    device->host sync per distinct tensor (cached), so callers on the hot path pass sequences."""
    if isinstance(spatial_shapes, torch.Tensor):
        key = (spatial_shapes.data_ptr(), spatial_shapes._version, tuple(spatial_shapes.shape),
               level_start_index.data_ptr() if isinstance(level_start_index, torch.Tensor) else None)
        hit = _shape_cache.get(key)
        if hit is None:
            shapes = [tuple(int(v) for v in r) for r in spatial_shapes.tolist()]
            starts = ([int(v) for v in level_start_index.tolist()] if isinstance(level_start_index, torch.Tensor)
                      else [int(v) for v in level_start_index])
            if len(_shape_cache) > 64:
                _shape_cache.clear()
            hit = _shape_cache[key] = (shapes, starts)
        return hit
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    if level_start_index is None:
        starts, s = [], 0
        for h, w in shapes:
            starts.append(s)
            s += h * w
    elif isinstance(level_start_index, torch.Tensor):
        starts = [int(v) for v in level_start_index.tolist()]
    else:
        starts = [int(v) for v in level_start_index]
    return shapes, starts


# Opt-in reduced-precision staging of the encoder's value pyramid (csrc/msda_tile.cu): the forward gathers
# from an fp16 head-major copy of `value` (weights, products, accumulation fp32); the backward is unchanged (fp32 value).
# Off by default: the fp32 parity contract of the mask logits is stated for exact fp32 staging.
msda_value_half = bool(int(os.environ.get("PDB_MSDA_HALF", "0")))


def _msda_half_eligible(value, shapes, starts, Lq, L, P):
    N, S, M, D = value.shape
    if not (msda_value_half and value.dtype == torch.float32 and D == 32 and P == 4 and Lq == S and 1 <= L <= 4):
        return False
    s = 0
    for (h, w), st in zip(shapes, starts):
        if h < 2 or w < 2 or st != s:
            return False
        s += h * w
    return s == S


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step=128):
        _need_cuda(value, sampling_locations, attention_weights)
        # the reference op hard-asserts contiguity (ms_deform_attn_cuda.cu:34-44)
        for name, t in (("value", value), ("sampling_loc", sampling_locations), ("attn_weight", attention_weights)):
            if not t.is_contiguous():
                raise RuntimeError(f"{name} tensor has to be contiguous")
        if value.dtype not in _DT or sampling_locations.dtype != value.dtype or attention_weights.dtype != value.dtype:
            raise RuntimeError("ms_deform_attn: value / sampling_loc / attn_weight must share dtype float32 or float64")
        N, S, M, D = value.shape
        _, Lq, _, L, P, _ = sampling_locations.shape
        step = min(N, int(im2col_step))
        if N % step != 0:
            raise RuntimeError(f"batch({N}) must divide im2col_step({step})")
        shapes, starts = _host_levels(value_spatial_shapes, value_level_start_index)
        if len(shapes) != L:
            raise RuntimeError(f"ms_deform_attn: {len(shapes)} spatial shapes for L={L}")
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        hs = _lib.host_i64([v for hw in shapes for v in hw])
        st = _lib.host_i64(starts)
        if _msda_half_eligible(value, shapes, starts, Lq, L, P):
            value_h = torch.empty((N, M, S, D), dtype=torch.float16, device=value.device)
            _lib.check(_lib.load().pdb_msda_pack_value_h(value.data_ptr(), value_h.data_ptr(), N, S, M, D, _stream()),
                       "pdb_msda_pack_value_h")
            rc = _lib.load().pdb_msda_forward_h(value_h.data_ptr(), hs, st, sampling_locations.data_ptr(),
                                                attention_weights.data_ptr(), out.data_ptr(), N, S, M, D, Lq, L, P, _stream())
            _lib.check(rc, "pdb_msda_forward_h")
        else:
            rc = _lib.load().pdb_msda_forward(value.data_ptr(), hs, st, sampling_locations.data_ptr(),
                                              attention_weights.data_ptr(), out.data_ptr(), N, S, M, D, Lq, L, P,
                                              _DT[value.dtype], _stream())
            _lib.check(rc, "pdb_msda_forward")
        ctx.save_for_backward(value, sampling_locations, attention_weights)
        ctx.levels = (shapes, starts)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, loc, attn = ctx.saved_tensors
        shapes, starts = ctx.levels
        N, S, M, D = value.shape
        _, Lq, _, L, P, _ = loc.shape
        grad_output = _c(grad_output)
        gv = torch.empty_like(value)
        gl = torch.empty_like(loc)
        ga = torch.empty_like(attn)
        rc = _lib.load().pdb_msda_backward(value.data_ptr(), _lib.host_i64([v for hw in shapes for v in hw]),
                                           _lib.host_i64(starts), loc.data_ptr(), attn.data_ptr(),
                                           grad_output.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                                           N, S, M, D, Lq, L, P, _DT[value.dtype], _stream())
        _lib.check(rc, "pdb_msda_backward")
        return gv, None, None, gl, ga, None


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step=128):
    return MSDeformAttnFunction.apply(value, spatial_shapes, level_start_index, sampling_locations,
                                      attention_weights, im2col_step)


# --------------------------------------------------------------------------------------------------
# mask-head einsum  bqc,bchw->bqhw  (mask2former_transformer_decoder.py:449)
# --------------------------------------------------------------------------------------------------
def _pixel_major(t):
    """(B, C, H, W) logical tensor -> the same tensor in channels_last memory (pixel-major (B, HW, C))."""
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


class GradFanOut(Function):
    """Identity on a tensor that feeds several mask_einsum calls (mask_features: the 10 prediction heads of the decoder).  Their
    backward GEMMs accumulate into ONE gradient buffer kept in ``holder`` (first one stores, the rest red.add) and hand autograd
    None; this node runs after all of them (autograd orders a node behind every consumer of its output, whether or not the
    consumer returned a gradient) and delivers the buffer — instead of 10 full-size gradients and 9 ATen add_ passes."""

    @staticmethod
    def forward(ctx, x, holder):
        ctx.holder = holder
        ctx.set_materialize_grads(False)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        buf = ctx.holder.pop("buf", None)
        if buf is None:
            return g, None
        if g is not None:                  # consumers that went through plain autograd
            buf = buf.add_(g)
        return buf, None


def grad_fan_out(x):
    """-> (x', holder) for mask_einsum(..., grad_acc=holder); a no-op pair (x, None) when x needs no gradient."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        return x, None
    holder = {}
    return GradFanOut.apply(x, holder), holder


class MaskEinsumFunction(Function):
    @staticmethod
    def forward(ctx, mask_embed, mask_features, embed_lo=None, grad_acc=None):
        ctx.grad_acc = grad_acc
        _need_cuda(mask_embed, mask_features)
        if mask_embed.dtype != torch.float32 or mask_features.dtype != torch.float32:
            raise RuntimeError("mask_einsum: float32 only")
        mask_embed, mask_features = _c(mask_embed), _pixel_major(mask_features)
        B, Q, C = mask_embed.shape
        Bf, Cf, H, W = mask_features.shape
        if B != Bf or C != Cf:
            raise RuntimeError(f"mask_einsum: shapes {tuple(mask_embed.shape)} x {tuple(mask_features.shape)}")
        out = torch.empty((B, Q, H, W), dtype=torch.float32, device=mask_embed.device)
        # embed_lo: the tf32 low parts of mask_embed (split_lo), so that the GEMM does not re-split the same (Q, C) operand in
        # every one of its ~1000 pixel tiles (47 vs 51 us at the BASELINE shape); without it the kernel splits in place
        if embed_lo is not None and (embed_lo.shape != mask_embed.shape or not embed_lo.is_contiguous()
                                     or embed_lo.dtype != torch.float32):
            raise RuntimeError("mask_einsum: embed_lo must be a contiguous float32 tensor shaped like mask_embed")
        rc = _lib.load().pdb_mask_einsum_forward(mask_embed.data_ptr(), embed_lo.data_ptr() if embed_lo is not None else None,
                                                 mask_features.data_ptr(), out.data_ptr(), B, Q, C, H * W, _stream())
        _lib.check(rc, "pdb_mask_einsum_forward")
        ctx.save_for_backward(mask_embed, mask_features)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        mask_embed, mask_features = ctx.saved_tensors
        B, Q, C = mask_embed.shape
        H, W = mask_features.shape[-2:]
        grad_out = _c(grad_out)
        ge = torch.empty_like(mask_embed) if ctx.needs_input_grad[0] else None
        acc, holder = 0, ctx.grad_acc
        gf = None
        if ctx.needs_input_grad[1]:
            if holder is not None and "buf" in holder:
                gf, acc = holder["buf"], 1
            else:
                gf = torch.empty_like(mask_features, memory_format=torch.channels_last)
                if holder is not None:
                    holder["buf"] = gf
        rc = _lib.load().pdb_mask_einsum_backward(mask_embed.data_ptr(), mask_features.data_ptr(), grad_out.data_ptr(),
                                                  ge.data_ptr() if ge is not None else None,
                                                  gf.data_ptr() if gf is not None else None, acc,
                                                  B, Q, C, H * W, _stream())
        _lib.check(rc, "pdb_mask_einsum_backward")
        return ge, (None if holder is not None else gf), None, None


def mask_einsum(mask_embed, mask_features, embed_lo=None, presplit=True, grad_acc=None):
    """torch.einsum("bqc,bchw->bqhw") (mask2former_transformer_decoder.py:449).  The tiny embed operand is pre-split into its
    tf32 hi / lo parts by one extra launch (presplit) unless the caller hands in embed_lo = split_lo(mask_embed) itself.
    Under bf16 autocast the embed arrives as bf16 (output of the mask MLP); the contraction itself stays on the fp32-accurate
    kernel (the reference's autocast einsum rounds operands AND the mask logits to bf16; keeping fp32 here is the more precise
    side of the documented tolerance).  grad_acc: the holder of grad_fan_out(mask_features) when several calls share the
    features (mask_features must then be that call's first result)."""
    if mask_embed.dtype != torch.float32:
        mask_embed = mask_embed.float()
    if mask_features.dtype != torch.float32:
        mask_features = mask_features.float()
    if embed_lo is None and presplit and mask_embed.is_cuda and mask_embed.dtype == torch.float32 \
            and mask_embed.numel() % 4 == 0:
        embed_lo = split_lo(_c(mask_embed.detach()))
    return MaskEinsumFunction.apply(mask_embed, mask_features, embed_lo, grad_acc)


# --------------------------------------------------------------------------------------------------
# attention mask  (mask2former_transformer_decoder.py:453-457 and the reset at :405)
# --------------------------------------------------------------------------------------------------
def build_attention_mask(pred_masks, size):
    """pred_masks (B, Q, H, W) f32 logits -> (mask uint8 (B, Q, h*w) with 1 = masked, row_any int32 (B*Q,)).
    Not differentiable (the reference detaches it)."""
    _need_cuda(pred_masks)
    pm = _c(pred_masks.detach())
    if pm.dtype != torch.float32:
        pm = pm.float()
    B, Q, H, W = pm.shape
    h, w = int(size[0]), int(size[1])
    mask = torch.empty((B, Q, h * w), dtype=torch.uint8, device=pm.device)
    row_any = torch.zeros((B * Q,), dtype=torch.int32, device=pm.device)
    rc = _lib.load().pdb_attn_mask_build(pm.data_ptr(), mask.data_ptr(), row_any.data_ptr(), B, Q, H, W, h, w,
                                         _stream())
    _lib.check(rc, "pdb_attn_mask_build")
    return mask, row_any


def reset_fully_masked_rows(mask, row_any):
    """In-place ``attn_mask[rows that are all True] = False`` on the compact mask."""
    B, Q, hw = mask.shape
    rc = _lib.load().pdb_attn_mask_reset_rows(mask.data_ptr(), row_any.data_ptr(), B * Q, hw, _stream())
    _lib.check(rc, "pdb_attn_mask_reset_rows")
    return mask


# --------------------------------------------------------------------------------------------------
# masked cross-attention core  (nn.MultiheadAttention inside CrossAttentionLayer, :84,102-114)
# --------------------------------------------------------------------------------------------------
_xattn_passes_now = [3]


def _set_xattn_passes(passes):
    """3xTF32 (fp32-accurate) or a single TF32 product per MMA in the attention kernels; the library keeps the value until told
    otherwise, so the setter is only called on a change."""
    if _xattn_passes_now[0] != passes:
        f = getattr(_lib.load(), "pdb_set_xattn_passes", None)       # absent only in the CPU-tier host builds of the tests
        if f is not None:
            f(passes)
        _xattn_passes_now[0] = passes


class MaskedCrossAttentionFunction(Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, row_any, heads):
        _need_cuda(q, k, v, mask, row_any)
        # under torch.autocast(bfloat16) the reference's attention rounds q, k, p, v to bf16 (2^-9): one TF32 product (operands
        # truncated at 2^-10) stays inside that; the backward (which runs outside the autocast region) follows the forward
        ctx.passes = 1 if (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16) else 3
        _set_xattn_passes(ctx.passes)
        q, k, v = _c(q), _c(k), _c(v)
        if q.dtype != torch.float32 or k.dtype != torch.float32 or v.dtype != torch.float32:
            raise RuntimeError("masked_cross_attention: float32 only")
        B, Q, E = q.shape
        Lk = k.shape[1]
        d = E // heads
        if mask is not None:
            mask = _c(mask)
            if mask.dtype != torch.uint8 or tuple(mask.shape) != (B, Q, Lk):
                raise RuntimeError("masked_cross_attention: mask must be uint8 (B, Q, Lk)")
        lib = _lib.load()
        ws_bytes = lib.pdb_masked_xattn_workspace_bytes(B, heads, Q, Lk, d)
        if ws_bytes < 0:
            raise RuntimeError(f"masked_cross_attention: unsupported shape B={B} heads={heads} Q={Q} Lk={Lk} d={d}")
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=q.device)
        out = torch.empty_like(q)
        lse = torch.empty((B, heads, Q), dtype=torch.float32, device=q.device)
        rc = lib.pdb_masked_xattn_forward(q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                          mask.data_ptr() if mask is not None else None,
                                          row_any.data_ptr() if row_any is not None else None,
                                          out.data_ptr(), lse.data_ptr(), ws.data_ptr(), B, heads, Q, Lk, d, _stream())
        _lib.check(rc, "pdb_masked_xattn_forward")
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.mask, ctx.row_any, ctx.heads = mask, row_any, heads
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        q, k, v, out, lse = ctx.saved_tensors
        mask, row_any, heads = ctx.mask, ctx.row_any, ctx.heads
        B, Q, E = q.shape
        Lk = k.shape[1]
        grad_out = _c(grad_out)
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _set_xattn_passes(ctx.passes)
        rc = _lib.load().pdb_masked_xattn_backward(
            q.data_ptr(), k.data_ptr(), v.data_ptr(), mask.data_ptr() if mask is not None else None,
            row_any.data_ptr() if row_any is not None else None, out.data_ptr(), lse.data_ptr(), grad_out.data_ptr(),
            gq.data_ptr(), gk.data_ptr(), gv.data_ptr(), B, heads, Q, Lk, E // heads, _stream())
        _lib.check(rc, "pdb_masked_xattn_backward")
        return gq, gk, gv, None, None, None


def masked_cross_attention(q, k, v, mask, row_any, heads):
    """q (B, Q, E) pre-scaled, k/v (B, Lk, E), mask uint8 (B, Q, Lk) or None -> (B, Q, E)."""
    return MaskedCrossAttentionFunction.apply(q, k, v, mask, row_any, heads)


# --------------------------------------------------------------------------------------------------
# point sampling / matcher / point loss  (matcher.py:100-168, criterion.py:147-207)
# --------------------------------------------------------------------------------------------------
class PointSampleFunction(Function):
    @staticmethod
    def forward(ctx, src, coords, map_index, coord_index):
        _need_cuda(src, coords)
        src, coords = _c(src), _c(coords)
        R_src, H, W = src.shape
        P = coords.shape[1]
        R = map_index.shape[0] if map_index is not None else (coord_index.shape[0] if coord_index is not None else R_src)
        if src.dtype == torch.float32:
            sd = 0
        elif src.dtype in (torch.uint8, torch.bool):
            sd = 1
        elif src.dtype == torch.int32:               # bit-packed maps: (R, H, W / 32) words of 32 pixels
            sd = 2
            W *= 32
        else:
            raise RuntimeError("point_sample: maps must be float32, uint8/bool or bit-packed int32 words")
        out = torch.empty((R, P), dtype=torch.float32, device=src.device)
        if R:       # nothing to sample for an empty row set (a batch without targets): empty tensors have null pointers
            rc = _lib.load().pdb_point_sample_forward(src.data_ptr(), sd, map_index.data_ptr() if map_index is not None else None,
                                                      coords.data_ptr(), coord_index.data_ptr() if coord_index is not None else None,
                                                      out.data_ptr(), R, P, H, W, _stream())
            _lib.check(rc, "pdb_point_sample_forward")
        ctx.save_for_backward(coords)
        ctx.meta = (map_index, coord_index, tuple(src.shape), sd)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (coords,) = ctx.saved_tensors
        map_index, coord_index, shape, sd = ctx.meta
        if sd != 0 or not ctx.needs_input_grad[0]:
            return None, None, None, None
        grad_out = _c(grad_out)
        R, P = grad_out.shape
        gs = torch.zeros(shape, dtype=torch.float32, device=grad_out.device)
        if R:
            rc = _lib.load().pdb_point_sample_backward(grad_out.data_ptr(), map_index.data_ptr() if map_index is not None else None,
                                                       coords.data_ptr(), coord_index.data_ptr() if coord_index is not None else None,
                                                       gs.data_ptr(), R, P, shape[1], shape[2], _stream())
            _lib.check(rc, "pdb_point_sample_backward")
        return gs, None, None, None


def point_sample(src, coords, map_index=None, coord_index=None):
    """src (R_src, H, W) f32 / uint8, coords (Rc, P, 2) in [0,1] (x, y) -> (R, P) f32.
    ``map_index`` / ``coord_index`` (int32) select the map / point set of each output row."""
    return PointSampleFunction.apply(src, coords, map_index, coord_index)


def matcher_cost(pred_pts, tgt_pts, cls_prob, tgt_label, tgt_offset, Q, w_class, w_mask, w_dice):
    """Cost matrices of every image in one launch; returns the flat f32 buffer (sum_b Q*K_b)."""
    _need_cuda(pred_pts, tgt_pts, cls_prob, tgt_label)
    pred_pts, tgt_pts, cls_prob = _c(pred_pts), _c(tgt_pts), _c(cls_prob)
    B = len(tgt_offset) - 1
    P = pred_pts.shape[1]
    Kc = cls_prob.shape[-1]
    cost = torch.empty((Q * int(tgt_offset[-1]),), dtype=torch.float32, device=pred_pts.device)
    rc = _lib.load().pdb_matcher_cost(pred_pts.data_ptr(), tgt_pts.data_ptr(), cls_prob.data_ptr(), tgt_label.data_ptr(),
                                      _lib.host_i32(tgt_offset), cost.data_ptr(), B, Q, Kc, P,
                                      float(w_class), float(w_mask), float(w_dice), _stream())
    _lib.check(rc, "pdb_matcher_cost")
    return cost


def lsap_batched(cost, tgt_offset, Q):
    """Returns (pred_idx, tgt_idx) int64 (Ktot,), per image ordered by ascending matched cost."""
    _need_cuda(cost)
    B = len(tgt_offset) - 1
    Kt = int(tgt_offset[-1])
    pred_idx = torch.empty((Kt,), dtype=torch.int64, device=cost.device)
    tgt_idx = torch.empty((Kt,), dtype=torch.int64, device=cost.device)
    rc = _lib.load().pdb_lsap_batched(cost.data_ptr(), _lib.host_i32(tgt_offset), pred_idx.data_ptr(), tgt_idx.data_ptr(),
                                      B, Q, _stream())
    _lib.check(rc, "pdb_lsap_batched")
    return pred_idx, tgt_idx


class PointLossFunction(Function):
    @staticmethod
    def forward(ctx, pred, pred_index, gt, gt_index, coords):
        _need_cuda(pred, pred_index, gt, gt_index, coords)
        pred, gt, coords = _c(pred), _c(gt), _c(coords)
        if pred.dtype != torch.float32 or gt.dtype not in (torch.uint8, torch.bool, torch.int32):
            raise RuntimeError("point_loss: pred float32, gt uint8/bool or bit-packed int32 words")
        Rp, H, W = pred.shape
        bits = int(gt.dtype == torch.int32)          # (Rg, Hg, Wg / 32) words of 32 pixels (padded widths are multiples of 32)
        _, Hg, Wg = gt.shape
        if bits:
            Wg *= 32
        Nm, P = coords.shape[0], coords.shape[1]
        sums = torch.empty((Nm, 4), dtype=torch.float32, device=pred.device)
        # few pairs (B = 2: a dozen per decoder output): the points of a pair are split over several CTAs as well
        splits = 1 if Nm >= 148 or Nm == 0 else max(1, min((P + 255) // 256, (2 * 148 + Nm - 1) // Nm))
        ctx.splits = splits
        if Nm:      # no matched pair (a batch without targets): the losses are empty sums, as in the reference
            partial = torch.empty((Nm, splits, 4), dtype=torch.float32, device=pred.device) if splits > 1 else None
            rc = _lib.load().pdb_point_loss_forward(pred.data_ptr(), pred_index.data_ptr(), gt.data_ptr(), gt_index.data_ptr(),
                                                    coords.data_ptr(), sums.data_ptr(),
                                                    partial.data_ptr() if partial is not None else None, splits,
                                                    Nm, P, H, W, Hg, Wg, bits, _stream())
            _lib.check(rc, "pdb_point_loss_forward")
        ctx.save_for_backward(pred, pred_index, gt, gt_index, coords, sums)
        bce = sums[:, 0] / P
        dice = 1 - (2 * sums[:, 1] + 1) / (sums[:, 2] + sums[:, 3] + 1)
        return bce, dice

    @staticmethod
    @once_differentiable
    def backward(ctx, g_bce, g_dice):
        pred, pred_index, gt, gt_index, coords, sums = ctx.saved_tensors
        Rp, H, W = pred.shape
        bits = int(gt.dtype == torch.int32)
        _, Hg, Wg = gt.shape
        if bits:
            Wg *= 32
        Nm, P = coords.shape[0], coords.shape[1]
        g_bce, g_dice = _c(g_bce.float()), _c(g_dice.float())
        gp = torch.zeros_like(pred)
        if Nm:
            rc = _lib.load().pdb_point_loss_backward(pred.data_ptr(), pred_index.data_ptr(), gt.data_ptr(), gt_index.data_ptr(),
                                                     coords.data_ptr(), sums.data_ptr(), g_bce.data_ptr(), g_dice.data_ptr(),
                                                     gp.data_ptr(), ctx.splits, Nm, P, H, W, Hg, Wg, bits, _stream())
            _lib.check(rc, "pdb_point_loss_backward")
        return gp, None, None, None, None


def point_loss(pred, pred_index, gt, gt_index, coords):
    """Per matched pair: (mean BCE over the points, dice loss).  pred (Rp,H,W) f32, gt (Rg,Hg,Wg) uint8."""
    return PointLossFunction.apply(pred, pred_index, gt, gt_index, coords)


# --------------------------------------------------------------------------------------------------
# PartDistillation fp64 classifier rows  (part_distillation_transformer_decoder.py:107,215-238)
# --------------------------------------------------------------------------------------------------
class ClassRowsFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, obj, num_parts):
        _need_cuda(x, weight, bias, obj)
        x = _c(x.float())
        if weight.dtype != torch.float64 or bias.dtype != torch.float64:
            raise RuntimeError("class_rows: weight/bias must be float64 (the reference's class_embed is .double())")
        B, Q, C = x.shape
        Ncls = weight.shape[0]
        out = torch.empty((B, Q, num_parts + 1), dtype=torch.float64, device=x.device)
        rc = _lib.load().pdb_class_rows_forward(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), obj.data_ptr(),
                                                out.data_ptr(), B, Q, C, num_parts, Ncls, _stream())
        _lib.check(rc, "pdb_class_rows_forward")
        ctx.save_for_backward(x, weight, obj)
        ctx.num_parts = num_parts
        ctx.bias_ref = bias
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x, weight, obj = ctx.saved_tensors
        B, Q, C = x.shape
        Ncls = weight.shape[0]
        grad_out = _c(grad_out.double())
        gx = torch.empty_like(x)
        # The kernel ADDS the <= B*(P+1) touched rows into its gradient buffers.  When the trainer has pre-allocated the
        # (zero-filled, flat) fp64 gradients, the rows go straight into them and autograd gets None — no dense 360 MB zero
        # tensor + accumulation per decoder layer (O = 22 000 x P = 8 classes: part_distillation_transformer_decoder.py:107).
        direct = (weight.grad is not None and weight.grad.dtype == torch.float64 and weight.grad.is_contiguous()
                  and ctx.bias_ref.grad is not None and ctx.bias_ref.grad.dtype == torch.float64)
        gw = weight.grad if direct else torch.zeros_like(weight)       # dense zero rows otherwise: AdamW parity with the reference
        gb = ctx.bias_ref.grad if direct else torch.zeros((Ncls,), dtype=torch.float64, device=x.device)
        rc = _lib.load().pdb_class_rows_backward(x.data_ptr(), weight.data_ptr(), obj.data_ptr(), grad_out.data_ptr(),
                                                 gx.data_ptr(), gw.data_ptr(), gb.data_ptr(), B, Q, C, ctx.num_parts, Ncls,
                                                 _stream())
        _lib.check(rc, "pdb_class_rows_backward")
        return (gx, None, None, None, None) if direct else (gx, gw, gb, None, None)


def class_rows(x, weight, bias, obj, num_parts):
    return ClassRowsFunction.apply(x, weight, bias, obj, num_parts)


# --------------------------------------------------------------------------------------------------
# dense contractions on the tcgen05 tensor cores (3xTF32, fp32-accurate)  — csrc/gemm_tc.cu
# --------------------------------------------------------------------------------------------------
def split_lo(x):
    """lo = x - trunc_tf32(x) for a contiguous fp32 tensor whose numel is a multiple of 4 (the pre-split B operand)."""
    lo = torch.empty_like(x)
    _lib.check(_lib.load().pdb_split_lo(x.data_ptr(), lo.data_ptr(), x.numel(), _stream()), "pdb_split_lo")
    return lo


_weight_lo = {}          # frozen weights: (data_ptr, shape) -> lo tensor (computed once)
weights_epoch = 0        # bumped by the trainer after every optimizer step (trainable weights change in place)


def forget_weight(weight):
    """Drops the cached low part of a derived frozen weight that is about to be released (its address may be reused)."""
    _weight_lo.pop((weight.data_ptr(), tuple(weight.shape)), None)
    for transposed in (False, True):
        _bf16_weights.pop((weight.data_ptr(), tuple(weight.shape), transposed), None)


def weight_lo(weight, rows):
    """Pre-split low parts of a Linear weight for GEMMs with many row tiles (every 128-row tile would otherwise
    re-split the same weight tile).  Frozen weights are split once; trainable ones once per optimizer step."""
    if rows < 2048 or weight.numel() % 4 or weight.data_ptr() % 16 or not weight.is_contiguous():
        return None
    if not (weight.is_leaf or (weight._base is not None and weight._base.is_leaf)):
        # a temporary (e.g. a scaled / concatenated weight rebuilt inside autograd on every call): its address says nothing about
        # its contents, so nothing is cached
        return split_lo(weight.detach())
    key = (weight.data_ptr(), tuple(weight.shape))
    epoch = weights_epoch if weight.requires_grad or weight._base is not None and weight._base.requires_grad else -1
    hit = _weight_lo.get(key)
    if hit is None or hit[0] != epoch:
        if len(_weight_lo) > 4096:
            _weight_lo.clear()
        hit = _weight_lo[key] = (epoch, split_lo(weight.detach()))
    return hit[1]


def gemm_tf32x3(A, B, C, M, N, K, *, batch=1, lda, ldb, ldc, sa=0, sb=0, sc=0, a_mn=False, b_mn=False, c_trans=False,
                bias=None, relu=False, accumulate=False, ksplit=1, B_lo=None):
    """Raw entry point: C_b[m][n] (+)= sum_k A_b(m,k) B_b(n,k) (+bias[n]) (ReLU); see include/pdb200.h."""
    rc = _lib.load().pdb_gemm_tf32x3(A.data_ptr(), B.data_ptr(), B_lo.data_ptr() if B_lo is not None else None,
                                     C.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     M, N, K, batch, lda, ldb, ldc, sa, sb, sc, int(a_mn), int(b_mn), int(c_trans),
                                     int(relu), int(accumulate), int(ksplit), _stream())
    _lib.check(rc, "pdb_gemm_tf32x3")
    return C


class SplitColumns(Function):
    """x[..., :n], x[..., n:] as views; backward writes the two gradients into the column ranges of ONE new tensor (autograd's own
    slice backward builds a zero-filled full-size tensor per slice and adds them)."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.n, ctx.shape = n, x.shape
        ctx.set_materialize_grads(False)
        return x[..., :n], x[..., n:]

    @staticmethod
    def backward(ctx, ga, gb):
        if ga is None and gb is None:
            return None, None
        ref = ga if ga is not None else gb
        g = torch.empty(ctx.shape, dtype=ref.dtype, device=ref.device)
        for part, grad in ((g[..., :ctx.n], ga), (g[..., ctx.n:], gb)):
            if grad is None:
                part.zero_()
            else:
                part.copy_(grad)
        return g, None


def split_columns(x, n):
    return SplitColumns.apply(x, n) if torch.is_grad_enabled() and x.requires_grad else (x[..., :n], x[..., n:])


small_gemm_rows = 512      # row count up to which Linear forward / input-gradient products take the mma.sync kernel (0 disables)


def _small_gemm(M, N, K, relu):
    """Measured on B200 inside a replayed graph (tools/bench_small_gemm.py, profiles/r02_small_gemm.txt): the mma.sync kernel wins
    while its 32 x 64 tiles do not fill the SMs (5.1 vs 8.0 us at 200 x 256 x 256) and for long contractions, which it splits
    over K (10.4 vs 35.4 us at K = 2048); the tcgen05 kernel wins from ~150 tiles on (200 x 2048 x 256: 9.6 vs 9.5 us)."""
    if not (0 < M <= small_gemm_rows and relu in (0, 1, False, True)):
        return False
    if getattr(_lib.load(), "pdb_gemm_small_tf32x3", None) is None:       # absent only in the CPU-tier host builds of the tests
        return False
    return ((N + 63) // 64) * ((M + 31) // 32) < 148 or (K >= 1024 and not relu)


def gemm_small(A, B, M, N, K, *, lda, ldb, b_mn=False, bias=None, relu=False):
    """C[m][n] = sum_k A[m][k] B(n, k) (+ bias) (ReLU) for short A (csrc/gemm_small.cu); returns a new (M, N) fp32 tensor.
    Long contractions are split over K (red.add into a zero-filled C) until ~2 CTAs per SM are in flight."""
    ctas = ((N + 63) // 64) * ((M + 31) // 32)
    chunks = (K + 31) // 32
    ksplit = 1
    if not relu and chunks >= 16 and ctas < 148:
        ksplit = max(1, min(chunks // 4, (296 + ctas - 1) // ctas))
    C = (torch.zeros if ksplit > 1 else torch.empty)((M, N), dtype=torch.float32, device=A.device)
    rc = _lib.load().pdb_gemm_small_tf32x3(A.data_ptr(), B.data_ptr(), C.data_ptr(), bias.data_ptr() if bias is not None else None,
                                           M, N, K, lda, ldb, N, int(b_mn), int(bool(relu)), ksplit, _stream())
    _lib.check(rc, "pdb_gemm_small_tf32x3")
    return C


def _split_k(M, N, K):
    """K slices so that a weight-gradient / reduction-shaped GEMM fills the 148 SMs."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    return max(1, min((K + 31) // 32, (2 * 148 + tiles - 1) // tiles))


def col_sum(x2, into=None):
    """Sum over the rows of a contiguous fp32 (rows, N) matrix: the bias gradient of a Linear layer.  Short matrices (the
    decoder's 200 rows) take one CTA per 32 columns, tall ones row blocks that meet through red.add.  ``into``: a
    preallocated fp32 (N,) gradient to ADD the sums to (returns None then)."""
    if x2.is_cuda and x2.dtype == torch.bfloat16 and x2.shape[0] > 0 and x2.shape[1] % 2 == 0:
        fb = getattr(_lib.load(), "pdb_col_sum_bf16", None)
        if fb is not None:
            x2 = _c(x2)
            out = into if into is not None else torch.empty((x2.shape[1],), dtype=torch.float32, device=x2.device)
            _lib.check(fb(x2.data_ptr(), out.data_ptr(), x2.shape[0], x2.shape[1], 1 if into is not None else 0, _stream()),
                       "pdb_col_sum_bf16")
            return None if into is not None else out
        x2 = x2.float()
    f = getattr(_lib.load(), "pdb_col_sum", None)             # absent only in the CPU-tier host builds of the tests
    if f is None or not x2.is_cuda or x2.dtype != torch.float32 or x2.shape[0] == 0:
        if into is not None:
            into.add_(x2.sum(0))
            return None
        return x2.sum(0)
    x2 = _c(x2)
    out = into if into is not None else torch.empty((x2.shape[1],), dtype=torch.float32, device=x2.device)
    _lib.check(f(x2.data_ptr(), out.data_ptr(), x2.shape[0], x2.shape[1], 1 if into is not None else 0, _stream()), "pdb_col_sum")
    return None if into is not None else out


# Weight / bias gradients of LinearFunction go STRAIGHT into the parameter's preallocated gradient when there is one (the trainer
# keeps every trainable parameter's .grad as a view of its zero-filled flat buffer): the weight-gradient GEMM accumulates into it
# (it is a split-K red.add product anyway) and autograd receives None — no zero-filled (N, K) temporary, no AccumulateGrad add,
# and for row slices of a packed parameter (the q / k / v blocks of in_proj_weight) no SliceBackward (zeros + copy of the whole
# parameter per use).  Without a preallocated gradient (plain autograd use, the parity tests) nothing changes.
direct_param_grads = True


def _direct_grad(t, shape):
    if not direct_param_grads or t is None or not t.requires_grad or not t.is_contiguous():
        return None
    p = t if t.is_leaf else t._base
    if p is None or not p.is_leaf or p.grad is None or p.grad.dtype != torch.float32 or not p.grad.is_contiguous() \
            or p.grad.shape != p.shape:
        return None
    off = t.storage_offset() - p.storage_offset()
    if off < 0 or off + t.numel() > p.numel():
        return None
    g = p.grad.view(-1)[off:off + t.numel()].view(shape)
    return g if g.data_ptr() % 16 == 0 else None


def linear_supported(x, weight):
    return (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32
            and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0 and weight.is_contiguous())


class LinearFunction(Function):
    """y = x W^T + b (optionally ReLU) with forward, input-gradient and weight-gradient GEMMs on tcgen05."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        _need_cuda(x, weight, bias)
        N, K = weight.shape
        x2 = _c(x).view(-1, K)
        if x2.data_ptr() % 16:
            x2 = x2.clone()
        M = x2.shape[0]
        w_lo = weight_lo(weight, M)
        if _small_gemm(M, N, K, relu) and weight.data_ptr() % 16 == 0:
            out = gemm_small(x2, weight, M, N, K, lda=K, ldb=K, bias=bias, relu=relu)
        else:
            out = torch.empty((M, N), dtype=torch.float32, device=x.device)
            if M > 0:
                gemm_tf32x3(x2, weight, out, M, N, K, lda=K, ldb=K, ldc=N, bias=bias, relu=relu, B_lo=w_lo)
        ctx.w_lo = w_lo
        ctx.relu = int(relu)            # epilogue activation: 0 none, 1 ReLU, 2 GELU (forward-only: see linear())
        ctx.has_bias = bias is not None
        ctx.wref, ctx.bref = weight, bias       # the caller's tensors (leaf parameters or row slices of one): see _direct_grad
        ctx.save_for_backward(x2, weight, out if ctx.relu == 1 else None)
        return out.view(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, weight, out = ctx.saved_tensors
        N, K = weight.shape
        M = x2.shape[0]
        gy2 = _c(gy).view(M, N)
        if ctx.relu == 2:
            raise RuntimeError("LinearFunction: the fused GELU epilogue is forward-only")
        if ctx.relu:
            gy2 = torch.ops.aten.threshold_backward(gy2, out, 0.0)      # gy * (out > 0) in one pass
        if gy2.data_ptr() % 16:
            gy2 = gy2.clone()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # dx[m,i] = sum_o gy[m,o] W[o,i]:  A = gy (K-major), B(n=i,k=o) = W[o*K+i] (MN-major)
            if _small_gemm(M, K, N, 0) and weight.data_ptr() % 16 == 0:
                gx = gemm_small(gy2, weight, M, K, N, lda=N, ldb=K, b_mn=True)
            else:
                gx = torch.empty((M, K), dtype=torch.float32, device=gy.device)
                gemm_tf32x3(gy2, weight, gx, M, K, N, lda=N, ldb=K, ldc=K, b_mn=True, B_lo=ctx.w_lo)
            gx = gx.view(*gy.shape[:-1], K)
        if ctx.needs_input_grad[1]:
            tgt = _direct_grad(ctx.wref, (N, K))
            gw = tgt if tgt is not None else torch.zeros((N, K), dtype=torch.float32, device=gy.device)
            # dW[o,i] = sum_m gy[m,o] x[m,i]:  A(m'=o,k=m) = gy[m*N+o], B(n'=i,k=m) = x[m*K+i]  (both MN-major)
            gemm_tf32x3(gy2, x2, gw, N, K, M, lda=N, ldb=K, ldc=K, a_mn=True, b_mn=True, accumulate=True,
                        ksplit=_split_k(N, K, M))
            if tgt is not None:
                gw = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(gy2, into=_direct_grad(ctx.bref, (N,)))
        return gx, gw, gb, None


class FFNFunction(Function):
    """y = relu(x W1^T + b1) W2^T + b2 as ONE autograd node (the FFN of the encoder and decoder layers, msdeformattn.py:120-124,
    mask2former_transformer_decoder.py:167-171).  Forward = two LinearFunction products; what the single node buys is the
    backward: dh = (dy W2) * (h > 0) leaves the input-gradient GEMM already gated (pdb_gemm_tf32x3_gated, gate = h) instead of
    a GEMM pass followed by a threshold_backward pass over the (rows, d_ffn) tensor.  Weight / bias gradients as in
    LinearFunction (straight into preallocated parameter gradients when there are any)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        _need_cuda(x, w1, w2)
        F1, K = w1.shape
        N = w2.shape[0]
        x2 = _c(x).view(-1, K)
        if x2.data_ptr() % 16:
            x2 = x2.clone()
        M = x2.shape[0]
        lo1, lo2 = weight_lo(w1, M), weight_lo(w2, M)
        h = torch.empty((M, F1), dtype=torch.float32, device=x.device)
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        if M > 0:
            gemm_tf32x3(x2, w1, h, M, F1, K, lda=K, ldb=K, ldc=F1, bias=b1, relu=True, B_lo=lo1)
            gemm_tf32x3(h, w2, y, M, N, F1, lda=F1, ldb=F1, ldc=N, bias=b2, B_lo=lo2)
        ctx.lo = (lo1, lo2)
        ctx.refs = (w1, b1, w2, b2)
        ctx.save_for_backward(x2, h, w1, w2)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, h, w1, w2 = ctx.saved_tensors
        r1, rb1, r2, rb2 = ctx.refs
        lo1, lo2 = ctx.lo
        F1, K = w1.shape
        N = w2.shape[0]
        M = x2.shape[0]
        gy2 = _c(gy).view(M, N)
        if gy2.data_ptr() % 16:
            gy2 = gy2.clone()
        gx = gw1 = gb1 = gw2 = gb2 = None
        need_h = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (rb1 is not None and ctx.needs_input_grad[2])
        if need_h:
            # dh[m, f] = (sum_o gy[m, o] W2[o, f]) * (h[m, f] > 0): A = gy (K-major), B(n = f, k = o) = W2[o*F1 + f] (MN-major)
            gh = torch.empty((M, F1), dtype=torch.float32, device=gy.device)
            rc = _lib.load().pdb_gemm_tf32x3_gated(gy2.data_ptr(), w2.data_ptr(), lo2.data_ptr() if lo2 is not None else None,
                                                   gh.data_ptr(), h.data_ptr(), M, F1, N, 1, N, F1, F1, 0, 0, 0, 0, 1, _stream())
            _lib.check(rc, "pdb_gemm_tf32x3_gated")
            if ctx.needs_input_grad[0]:
                gx = torch.empty((M, K), dtype=torch.float32, device=gy.device)
                gemm_tf32x3(gh, w1, gx, M, K, F1, lda=F1, ldb=K, ldc=K, b_mn=True, B_lo=lo1)
                gx = gx.view(*gy.shape[:-1], K)
            if ctx.needs_input_grad[1]:
                tgt = _direct_grad(r1, (F1, K))
                gw1 = tgt if tgt is not None else torch.zeros((F1, K), dtype=torch.float32, device=gy.device)
                gemm_tf32x3(gh, x2, gw1, F1, K, M, lda=F1, ldb=K, ldc=K, a_mn=True, b_mn=True, accumulate=True,
                            ksplit=_split_k(F1, K, M))
                gw1 = None if tgt is not None else gw1
            if rb1 is not None and ctx.needs_input_grad[2]:
                gb1 = col_sum(gh, into=_direct_grad(rb1, (F1,)))
        if ctx.needs_input_grad[3]:
            tgt = _direct_grad(r2, (N, F1))
            gw2 = tgt if tgt is not None else torch.zeros((N, F1), dtype=torch.float32, device=gy.device)
            gemm_tf32x3(gy2, h, gw2, N, F1, M, lda=N, ldb=F1, ldc=F1, a_mn=True, b_mn=True, accumulate=True,
                        ksplit=_split_k(N, F1, M))
            gw2 = None if tgt is not None else gw2
        if rb2 is not None and ctx.needs_input_grad[4]:
            gb2 = col_sum(gy2, into=_direct_grad(rb2, (N,)))
        return gx, gw1, gb1, gw2, gb2


ffn_fused_rows = 2048       # from this many rows on the two-Linear FFN runs as one node with the gated input-gradient GEMM


def ffn(x, w1, b1, w2, b2):
    """relu(x W1^T + b1) W2^T + b2.  Tall fp32 CUDA inputs (the encoder's 43 008 rows) take FFNFunction; short ones (the decoder's
    200 rows, where a threshold_backward pass costs 2 us and the short-A GEMM kernels matter more), autocast and everything the
    GEMM cannot take go through two ``linear`` calls."""
    rows = x.numel() // max(1, x.shape[-1])
    if (rows >= ffn_fused_rows and w1.shape[0] >= 256 and not _autocast_bf16(x, w1) and linear_supported(x, w1) and linear_supported(x, w2)
            and b1 is not None and b2 is not None and w1.data_ptr() % 16 == 0 and w2.data_ptr() % 16 == 0
            and getattr(_lib.load(), "pdb_gemm_tf32x3_gated", None) is not None):
        return FFNFunction.apply(x, w1, b1, w2, b2)
    return linear(linear(x, w1, b1, relu=True), w2, b2)


class Conv1x1Function(Function):
    """1x1 convolution as a GEMM over pixels.  An NCHW-contiguous input is read in place as an MN-major operand
    (pixels contiguous), a channels-last one as a K-major operand; the result is the logical (B, O, H, W) tensor in
    channels-last memory (pixel-major (B, HW, O)), which is what the deformable encoder flattens to anyway."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, weight, bias)
        B, C, H, W = x.shape
        O = weight.shape[0]
        w2 = weight.view(O, C)
        HW = H * W
        cl = x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous()
        if not cl:
            x = _c(x)
        out = torch.empty((B, HW, O), dtype=torch.float32, device=x.device)
        w_lo = weight_lo(w2, B * HW)
        if cl:      # (B*HW, C) rows
            gemm_tf32x3(x, w2, out, B * HW, O, C, lda=C, ldb=C, ldc=O, bias=bias, B_lo=w_lo)
        else:       # A_b(m = pixel, k = channel) = x[b, k, m]
            gemm_tf32x3(x, w2, out, HW, O, C, batch=B, lda=HW, ldb=C, ldc=O, sa=C * HW, sb=0, sc=HW * O, a_mn=True,
                        bias=bias, B_lo=w_lo)
        ctx.save_for_backward(x, weight)
        ctx.cl, ctx.has_bias, ctx.w_lo = cl, bias is not None, w_lo
        return out.view(B, H, W, O).permute(0, 3, 1, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        B, C, H, W = x.shape
        O = weight.shape[0]
        HW = H * W
        w2 = weight.view(O, C)
        gy_pm = _c(gy.permute(0, 2, 3, 1)).view(B, HW, O)           # no copy when gy is channels-last
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gxp = torch.empty((B * HW, C), dtype=torch.float32, device=gy.device)
            gemm_tf32x3(gy_pm, w2, gxp, B * HW, C, O, lda=O, ldb=C, ldc=C, b_mn=True, B_lo=ctx.w_lo)
            gx = gxp.view(B, H, W, C).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            gw = torch.zeros((O, C), dtype=torch.float32, device=gy.device)
            ks = _split_k(O, C, HW)
            if ctx.cl:      # B_b(n = channel, k = pixel) = x[b, k, n]  (MN-major)
                gemm_tf32x3(gy_pm, x, gw, O, C, HW, batch=B, lda=O, ldb=C, ldc=C, sa=HW * O, sb=HW * C, sc=0,
                            a_mn=True, b_mn=True, accumulate=True, ksplit=ks)
            else:           # x NCHW: B_b(n = channel, k = pixel) = x[b, n, k]  (K-major)
                gemm_tf32x3(gy_pm, x, gw, O, C, HW, batch=B, lda=O, ldb=HW, ldc=C, sa=HW * O, sb=C * HW, sc=0,
                            a_mn=True, accumulate=True, ksplit=ks)
            gw = gw.view_as(weight)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(gy_pm.reshape(-1, gy_pm.shape[-1]))
        return gx, gw, gb


def conv1x1(x, weight, bias=None):
    """F.conv2d with a 1x1 kernel on the tensor cores (fp32 CUDA, channel counts / pixel count multiples of 4);
    returns the logical NCHW result in channels-last memory."""
    B, C, H, W = x.shape
    ok = (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and tuple(weight.shape[2:]) == (1, 1)
          and C % 4 == 0 and weight.shape[0] % 4 == 0 and (H * W) % 4 == 0 and weight.is_contiguous())
    if not ok:
        return torch.nn.functional.conv2d(x, weight, bias)
    return Conv1x1Function.apply(x, weight, bias)


def _gemm_taps(A, Bw, C, M, N, Ck, batch, a_rows, tap_off, bias=None, relu=False, B_lo=None):
    rc = _lib.load().pdb_gemm_taps_tf32x3(A.data_ptr(), Bw.data_ptr(), B_lo.data_ptr() if B_lo is not None else None,
                                          C.data_ptr(), bias.data_ptr() if bias is not None else None, M, N, Ck, batch, a_rows,
                                          Ck, N, a_rows * Ck, M * N, len(tap_off), _lib.host_i32(tap_off), int(relu), _stream())
    _lib.check(rc, "pdb_gemm_taps_tf32x3")
    return C


def _conv_taps(xp, w9, B, H, W, Cin, Cout, bias=None):
    """The 9-tap GEMM of a 3x3 convolution on the zero-padded (B, H + 3, W + 2, Cin) image -> (B, H, W, Cout).  With the
    cropping store (pdb_gemm_taps_cropped_tf32x3) the result is written at the true width; otherwise on the padded-width grid
    and cropped by a copy."""
    Wp = W + 2
    taps = [ky * Wp + kx for ky in range(3) for kx in range(3)]
    lo = split_lo(w9)
    f = getattr(_lib.load(), "pdb_gemm_taps_cropped_tf32x3", None)        # absent only in the CPU-tier host builds of the tests
    if f is not None and Cout > 112 and Cout % 4 == 0:
        out = torch.empty((B, H, W, Cout), dtype=torch.float32, device=xp.device)
        rc = f(xp.data_ptr(), w9.data_ptr(), lo.data_ptr(), out.data_ptr(), bias.data_ptr() if bias is not None else None,
               H * Wp, Cout, Cin, B, (H + 3) * Wp, Cin, Cout, (H + 3) * Wp * Cin, H * W * Cout, 9, _lib.host_i32(taps), 0, Wp, W,
               _stream())
        _lib.check(rc, "pdb_gemm_taps_cropped_tf32x3")
        return out
    full = torch.empty((B, H * Wp, Cout), dtype=torch.float32, device=xp.device)
    _gemm_taps(xp, w9, full, H * Wp, Cout, Cin, B, (H + 3) * Wp, taps, bias=bias, B_lo=lo)
    return full.view(B, H, Wp, Cout)[:, :, :W].contiguous()


def _pad_nhwc(x_nhwc):
    """(B, H, W, C) -> zero-padded (B, H + 3, W + 2, C): one row above, one column left / right, two rows below (the
    second one only keeps the shifted reads of the padded-width grid inside the image's own buffer)."""
    B, H, W, C = x_nhwc.shape
    f = getattr(_lib.load(), "pdb_pad_nhwc", None)          # absent only in the CPU-tier host builds of the tests
    if f is not None and x_nhwc.is_cuda and x_nhwc.dtype == torch.float32 and C % 4 == 0:
        x_nhwc = _c(x_nhwc)
        xp = torch.empty((B, H + 3, W + 2, C), dtype=torch.float32, device=x_nhwc.device)
        _lib.check(f(x_nhwc.data_ptr(), xp.data_ptr(), B, H, W, C, 1, 2, 1, 1, _stream()), "pdb_pad_nhwc")
        return xp
    xp = x_nhwc.new_zeros((B, H + 3, W + 2, C))
    xp[:, 1:H + 1, 1:W + 1] = x_nhwc
    return xp


class Conv3x3Function(Function):
    """3x3 / stride 1 / pad 1 convolution as ONE tensor-core GEMM: the zero-padded NHWC image is the A operand, the
    9 taps are 9 K segments whose rows are shifted by ky * (W + 2) + kx (pdb_gemm_taps_tf32x3); outputs are computed
    on the padded-width pixel grid and the two garbage columns per row dropped.  The input gradient is the same
    routine with the flipped, transposed kernel; the weight gradient is 9 split-K GEMMs over the pixels."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _need_cuda(x, weight, bias)
        B, C, H, W = x.shape
        O = weight.shape[0]
        xp = _pad_nhwc(x.permute(0, 2, 3, 1))
        w9 = weight.permute(0, 2, 3, 1).reshape(O, 9 * C).contiguous()
        out = _conv_taps(xp, w9, B, H, W, C, O, bias)
        ctx.save_for_backward(xp, weight)
        ctx.dims = (B, C, H, W, O)
        ctx.has_bias = bias is not None
        return out.permute(0, 3, 1, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xp, weight = ctx.saved_tensors
        B, C, H, W, O = ctx.dims
        Wp = W + 2
        taps = [ky * Wp + kx for ky in range(3) for kx in range(3)]
        gy_nhwc = _c(gy.permute(0, 2, 3, 1))
        gx = gw = gb = None
        gyp = None
        if ctx.needs_input_grad[0]:
            gyp = _pad_nhwc(gy_nhwc)
            w9t = weight.flip(2, 3).permute(1, 2, 3, 0).reshape(C, 9 * O).contiguous()
            gx = _conv_taps(gyp, w9t, B, H, W, O, C).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            rows, K = (H + 3) * Wp, H * Wp
            if gyp is not None:
                # gy on the padded-width grid with zero garbage columns is the padded image read Wp + 1 pixels in:
                # gyp[(y + 1) * Wp + x + 1] = gy[y][x], and x = W, W + 1 land on its zero right / left border columns
                gyf = gyp.view(B, rows * O)[:, (Wp + 1) * O:]
                sa = rows * O
            else:
                gyf = torch.nn.functional.pad(gy_nhwc, (0, 0, 0, 2))          # zero garbage columns of the padded-width grid
                sa = K * O
            gw9 = torch.zeros((O, 9 * C), dtype=torch.float32, device=gy.device)
            xflat = xp.view(B, rows * C)
            ks = _split_k(O, C, K)
            for t, off in enumerate(taps):
                # gw9[o, t*C + i] = sum_{b,q} gyf[b, q, o] * xp[b, q + off, i]
                gemm_tf32x3(gyf, xflat[:, off * C:], gw9[:, t * C:], O, C, K, batch=B, lda=O, ldb=C, ldc=9 * C,
                            sa=sa, sb=rows * C, sc=0, a_mn=True, b_mn=True, accumulate=True, ksplit=ks)
            gw = gw9.view(O, 3, 3, C).permute(0, 3, 1, 2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(gy_nhwc.reshape(-1, gy_nhwc.shape[-1]))
        return gx, gw, gb


# --------------------------------------------------------------------------------------------------
# bf16 autocast path: nn.Linear on tcgen05 kind::f16 (csrc/gemm_bf16.cu)
# --------------------------------------------------------------------------------------------------
def gemm_bf16(a, b, bias=None, act=0, out_dtype=torch.bfloat16, ksplit=1, into=None, transposed=False):
    """a (M, K), b (N, K) bf16 with contiguous rows -> a @ b^T (+ bias) (act) as (M, N) bf16 or fp32.  ksplit > 1 (fp32 result
    only): K is cut into slices that run on different SMs and meet through red.add.  ``into``: an existing contiguous fp32
    (M, N) tensor to ADD the product to (returns it).  ``transposed``: a is (K, M) and b is (K, N) — a^T @ b with both read in
    place as MN-major operands (the weight gradient dy^T x)."""
    _need_cuda(a, b)
    if transposed:
        K, M = a.shape
        N = b.shape[1]
    else:
        M, K = a.shape
        N = b.shape[0]
    if into is not None:
        out = into
    elif ksplit > 1:
        out = torch.zeros((M, N), dtype=out_dtype, device=a.device)
    else:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    if M and N:
        rc = _lib.load().pdb_gemm_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), None if bias is None else bias.data_ptr(), M, N, K,
                                       a.stride(0), b.stride(0), N, int(act), 1 if out_dtype == torch.bfloat16 else 0,
                                       int(ksplit), int(into is not None), 3 if transposed else 0, _stream())
        _lib.check(rc, "pdb_gemm_bf16")
    return out


bf16_wgrad_in_place = True     # A/B switch: False materialises dy^T and x^T in front of the weight-gradient GEMM (first version)


def _split_k_bf16(M, N, K):
    """K slices of a bf16 weight-gradient product (128 x 128 output tiles, 64-wide k-blocks): ~2 tiles per SM, >= 8 k-blocks each."""
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    return max(1, min((K + 511) // 512, (2 * 148 + tiles - 1) // tiles))


_bf16_weights = {}


def weight_bf16(weight, transposed=False):
    """bf16 copy of a weight (optionally transposed), cached until the weight changes (optimizer step / load)."""
    if not (weight.is_leaf or (weight._base is not None and weight._base.is_leaf)):     # a temporary: see weight_lo
        w = weight.detach()
        return (w.t() if transposed else w).to(torch.bfloat16).contiguous()
    key = (weight.data_ptr(), tuple(weight.shape), transposed)
    tag = (weight._version, weights_epoch)
    hit = _bf16_weights.get(key)
    if hit is None or hit[0] != tag:
        if len(_bf16_weights) > 4096:
            _bf16_weights.clear()
        w = weight.detach()
        hit = _bf16_weights[key] = (tag, (w.t() if transposed else w).to(torch.bfloat16).contiguous())
    return hit[1]


def _rows8(t):
    """(R, M) bf16 with contiguous rows and M padded to a multiple of 8 with zeros (the contraction length of a GEMM)."""
    R, M = t.shape
    if M % 8 == 0:
        return _c(t)
    out = torch.zeros((R, (M + 7) // 8 * 8), dtype=t.dtype, device=t.device)
    out[:, :M] = t
    return out


class LinearBF16Function(Function):
    """nn.Linear under bf16 autocast: y = x W^T + b (optionally ReLU / GELU in the epilogue), operands bf16, accumulation fp32,
    output bf16 — what torch.autocast makes of F.linear; the bias stays fp32 inside the epilogue.  Backward: dx = dy W and
    dW = dy^T x on the same kernel (bf16 operands, dW accumulated and returned in fp32), db = column sums in fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, out_fp32=False):
        _need_cuda(x, weight)
        N, K = weight.shape
        x2 = _c(x.reshape(-1, K).to(torch.bfloat16))
        if x2.data_ptr() % 16:
            x2 = x2.clone()
        # torch.autocast casts the bias to bf16 as well; the epilogue adds its fp32 image to the fp32 accumulator.
        # out_fp32: the consumer is one of this library's fp32 kernels (attention cores, mask einsum): the epilogue writes the
        # fp32 accumulator instead of rounding to bf16 and converting back in a separate pass.
        y = gemm_bf16(x2, weight_bf16(weight), None if bias is None else _c(bias.to(torch.bfloat16).float()), act,
                      torch.float32 if out_fp32 else torch.bfloat16)
        ctx.save_for_backward(x2, weight, y if act == 1 else None)
        ctx.wref, ctx.bref = weight, bias       # the caller's tensors (leaves or row slices of one): see _direct_grad
        ctx.meta = (tuple(x.shape), x.dtype, bias is not None, act)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x2, weight, y = ctx.saved_tensors
        shape, xdtype, has_bias, act = ctx.meta
        N, K = weight.shape
        gy = gy.reshape(-1, N).to(torch.bfloat16)
        if act == 1:
            gy = gy * (y > 0)
        gy = _c(gy)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_bf16(gy, weight_bf16(weight, transposed=True)).view(shape).to(xdtype)
        if ctx.needs_input_grad[1]:
            # dW = dy^T x: a handful of output tiles over a contraction as long as the token count -> split over K; straight
            # into the parameter's preallocated fp32 gradient when there is one (see _direct_grad)
            tgt = _direct_grad(ctx.wref, (N, K)) if weight.dtype == torch.float32 else None
            ks = _split_k_bf16(N, K, gy.shape[0])
            if bf16_wgrad_in_place and N % 8 == 0 and K % 8 == 0 and gy.data_ptr() % 16 == 0 and x2.data_ptr() % 16 == 0:
                ops = dict(a=gy, b=x2, transposed=True)           # dy (rows, N) and x (rows, K) read in place, MN-major
            else:
                ops = dict(a=_rows8(gy.t()), b=_rows8(x2.t()))
            if tgt is not None:
                gemm_bf16(out_dtype=torch.float32, ksplit=ks, into=tgt, **ops)
            else:
                gw = gemm_bf16(out_dtype=torch.float32, ksplit=ks, **ops).to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            gb = col_sum(gy, into=_direct_grad(ctx.bref, (N,)))
        return gx, gw, gb, None, None


def _autocast_bf16(x, weight):
    return (x.is_cuda and torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16
            and weight.dtype in (torch.float32, torch.bfloat16) and weight.shape[1] % 8 == 0)


def conv3x3(x, weight, bias=None):
    """F.conv2d(kernel 3, stride 1, padding 1) on the tensor cores for fp32 CUDA tensors with channel counts that are
    multiples of 32; returns the logical NCHW result in channels-last memory."""
    ok = (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and tuple(weight.shape[2:]) == (3, 3)
          and x.shape[1] % 32 == 0 and weight.shape[0] % 32 == 0 and weight.shape[1] == x.shape[1])
    if not ok:
        return torch.nn.functional.conv2d(x, weight, bias, padding=1)
    return Conv3x3Function.apply(x, weight, bias)


def linear(x, weight, bias=None, relu=False, gelu=False, out_fp32=False):
    """nn.Linear (optionally fused ReLU) on the tensor cores for fp32 CUDA tensors whose feature sizes are
    multiples of 4; other dtypes (autocast halves, the fp64 classifier) and tiny ragged heads go through
    the library GEMM.  gelu=True applies nn.GELU() (erf form); it is folded into the GEMM epilogue when nothing
    on the path needs a gradient (the frozen backbone), otherwise it runs as a separate differentiable op.
    Under torch.autocast(bfloat16) (and for bf16 inputs) the product runs on the bf16 tensor-core kernel and returns bf16
    (``out_fp32``: fp32, for consumers that are fp32 kernels of this library; no effect outside autocast, where the result is
    fp32 anyway)."""
    if _autocast_bf16(x, weight) or (x.is_cuda and x.dtype == torch.bfloat16 and weight.shape[1] % 8 == 0
                                     and weight.dtype in (torch.float32, torch.bfloat16)):
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or
                                                  (bias is not None and bias.requires_grad))
        if gelu and needs_grad:
            return torch.nn.functional.gelu(LinearBF16Function.apply(x, weight, bias, 0))
        return LinearBF16Function.apply(x, weight, bias, 2 if gelu else (1 if relu else 0), out_fp32)
    if linear_supported(x, weight):
        if gelu:
            needs_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad or
                                                      (bias is not None and bias.requires_grad))
            if needs_grad:
                return torch.nn.functional.gelu(LinearFunction.apply(x, weight, bias, False))
            return LinearFunction.apply(x, weight, bias, 2)
        return LinearFunction.apply(x, weight, bias, relu)
    if (x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and weight.shape[1] % 4 == 0
            and weight.shape[0] % 4 != 0):
        # ragged head (the 2-column class_embed of the proposal model): zero rows up to a multiple of 4, same tensor-core GEMM
        n, pad = weight.shape[0], (-weight.shape[0]) % 4
        w = torch.nn.functional.pad(weight, (0, 0, 0, pad))
        b = None if bias is None else torch.nn.functional.pad(bias, (0, pad))
        return linear(x, w, b, relu=relu, gelu=gelu)[..., :n]
    y = torch.nn.functional.linear(x, weight.to(x.dtype), None if bias is None else bias.to(x.dtype))
    if gelu:
        return torch.nn.functional.gelu(y)
    return torch.relu(y) if relu else y


# --------------------------------------------------------------------------------------------------
# pixel grouping affinity  (pixel_grouping_model.py:139-144,197-211)
# --------------------------------------------------------------------------------------------------
def group_affinity(feat, centroids, mask, metric="dot", geometry=None, two_stage=True):
    """feat (C, h, w) f32, centroids (Kc, C) f32, mask (H, W) bool/uint8 -> labels (H, W) int32: 0 outside the mask,
    1 + argmax_k affinity(bilinear(feat) at the pixel, centroid k) inside.  ``geometry`` = (padded size, image size,
    output size) when the features are up-sampled to the padded size, cropped and resized again (sem_seg_postprocess);
    ``mask`` is then at the output size.  Default: one bilinear pass to the size of ``mask``.  ``two_stage`` (default):
    the C-channel contraction runs once at feature resolution (linearity of the interpolation; include/pdb200.h)."""
    _need_cuda(feat, centroids, mask)
    if metric not in ("dot", "l2"):
        raise ValueError(f"distance metric {metric!r} (dot / l2)")
    feat, centroids = _c(feat.float()), _c(centroids.float())
    mask = _c(mask).view(torch.uint8) if mask.dtype == torch.bool else _c(mask.to(torch.uint8))
    C, h, w = feat.shape
    Kc = centroids.shape[0]
    H, W = mask.shape
    if two_stage and C > Kc:
        # contraction with the centroids at feature resolution (pdb_group_scores), then the per-pixel kernels on Kc score maps
        scores = torch.empty((Kc, h, w), dtype=torch.float32, device=feat.device)
        _lib.check(_lib.load().pdb_group_scores(feat.data_ptr(), centroids.data_ptr(), scores.data_ptr(), C, Kc, h, w,
                                                0 if metric == "dot" else 1, _stream()), "pdb_group_scores")
        feat, centroids, metric, C = scores, host_table(torch.eye(Kc).flatten().tolist(), torch.float32, feat.device).view(Kc, Kc), "dot", Kc
    labels = torch.empty((H, W), dtype=torch.int32, device=feat.device)
    if geometry is not None:
        (Hp, Wp), (Hi, Wi), (Ho, Wo) = [(int(a), int(b)) for a, b in geometry]
        if (Ho, Wo) != (H, W):
            raise RuntimeError(f"group_affinity: mask shape {(H, W)} != output size {(Ho, Wo)}")
        if not ((Hp, Wp) == (Hi, Wi) == (Ho, Wo)):
            rc = _lib.load().pdb_group_affinity_resized(feat.data_ptr(), centroids.data_ptr(), mask.data_ptr(),
                                                        labels.data_ptr(), C, Kc, h, w, Hp, Wp, Hi, Wi, Ho, Wo,
                                                        0 if metric == "dot" else 1, _stream())
            _lib.check(rc, "pdb_group_affinity_resized")
            return labels
    rc = _lib.load().pdb_group_affinity(feat.data_ptr(), centroids.data_ptr(), mask.data_ptr(), labels.data_ptr(), C, Kc, h, w,
                                        H, W, 0 if metric == "dot" else 1, _stream())
    _lib.check(rc, "pdb_group_affinity")
    return labels


def group_affinity_batched(feats, centroids, masks, metric="dot"):
    """B images of one geometry in two launches: feats (B, C, h, w) f32, centroids (B, Kc, C), masks (B, H, W) bool/uint8 ->
    labels (B, H, W) int32, image by image what ``group_affinity`` returns."""
    _need_cuda(feats, centroids, masks)
    if metric not in ("dot", "l2"):
        raise ValueError(f"distance metric {metric!r} (dot / l2)")
    feats, centroids = _c(feats.float()), _c(centroids.float())
    masks = _c(masks).view(torch.uint8) if masks.dtype == torch.bool else _c(masks.to(torch.uint8))
    B, C, h, w = feats.shape
    Kc = centroids.shape[1]
    H, W = masks.shape[1:]
    labels = torch.empty((B, H, W), dtype=torch.int32, device=feats.device)
    scores = torch.empty((B, Kc, h, w), dtype=torch.float32, device=feats.device)
    eye = host_table(torch.eye(Kc).flatten().tolist(), torch.float32, feats.device)
    rc = _lib.load().pdb_group_affinity_batched(feats.data_ptr(), centroids.data_ptr(), masks.data_ptr(), labels.data_ptr(),
                                                scores.data_ptr(), eye.data_ptr(), B, C, Kc, h, w, H, W,
                                                0 if metric == "dot" else 1, _stream())
    _lib.check(rc, "pdb_group_affinity_batched")
    return labels


# --------------------------------------------------------------------------------------------------
# Inference post-processing on bit-packed masks  (proposal_model.py:220-302,369-432; utils/utils.py:35-42)
# --------------------------------------------------------------------------------------------------
def _u8(mask):
    return _c(mask).view(torch.uint8) if mask.dtype == torch.bool else _c(mask.to(torch.uint8))


def postprocess_masks(logits, sel, padded_size, image_size, out_size, gate=None, scores=None, want_bits=True,
                      want_label=False, score_threshold=None):
    """logits (Q, h, w) f32 of ONE image, sel (K) query indices -> (bits (K + 1, Ho, ceil(Wo / 32)) int32 words or
    None, label (Ho, Wo) int32 or None).  Row k of ``bits`` is ``resize(logits[sel[k]]) * gate > 0`` with ``resize`` =
    bilinear to ``padded_size``, crop to ``image_size``, bilinear to ``out_size`` (both align_corners=False); row K is
    the OR over k.  ``label`` = argmax_k scores[k] * sigmoid(resize * gate).  With ``score_threshold`` a third result is
    returned: words (K, Ho, Ww) of ``scores[k] * sigmoid(resize * gate) > score_threshold``."""
    _need_cuda(logits, sel, gate, scores)
    if logits.dtype != torch.float32 or logits.dim() != 3:
        raise RuntimeError("postprocess_masks: logits must be float32 (Q, h, w)")
    logits = _c(logits)
    sel = _c(sel.to(torch.int32))
    Q, h, w = logits.shape
    K = sel.shape[0]
    (Hp, Wp), (Hi, Wi), (Ho, Wo) = [(int(a), int(b)) for a, b in (padded_size, image_size, out_size)]
    if gate is not None:
        gate = _u8(gate)
        if tuple(gate.shape[-2:]) != (Ho, Wo) or gate.numel() != Ho * Wo:
            raise RuntimeError(f"postprocess_masks: gate shape {tuple(gate.shape)} != output size {(Ho, Wo)}")
    need_scores = want_label or score_threshold is not None
    if need_scores:
        if scores is None:
            raise RuntimeError("postprocess_masks: the label map / score bits need the scores")
        scores = _c(scores.float())
    Ww = (Wo + 31) // 32
    bits = torch.empty((K + 1, Ho, Ww), dtype=torch.int32, device=logits.device) if want_bits else None
    label = torch.empty((Ho, Wo), dtype=torch.int32, device=logits.device) if want_label else None
    sbits = torch.empty((K, Ho, Ww), dtype=torch.int32, device=logits.device) if score_threshold is not None else None
    rc = _lib.load().pdb_postprocess_masks(logits.data_ptr(), sel.data_ptr(),
                                           scores.data_ptr() if need_scores else None,
                                           gate.data_ptr() if gate is not None else None,
                                           bits.data_ptr() if want_bits else None,
                                           label.data_ptr() if want_label else None,
                                           sbits.data_ptr() if sbits is not None else None,
                                           float(score_threshold) if score_threshold is not None else 0.0,
                                           Q, K, h, w, Hp, Wp, Hi, Wi, Ho, Wo, _stream())
    _lib.check(rc, "pdb_postprocess_masks")
    if score_threshold is not None:
        return bits, label, sbits
    return bits, label


def resize_bool_masks(masks, image_size, out_size):
    """Zero-padded bool / uint8 (G, Hp, Wp) -> bool (G, Ho, Wo): ``sem_seg_postprocess(masks.float(), image_size, Ho,
    Wo).bool()`` (crop, bilinear, non-zero)."""
    _need_cuda(masks)
    masks = _u8(masks)
    G, Hp, Wp = masks.shape
    (Hi, Wi), (Ho, Wo) = [(int(a), int(b)) for a, b in (image_size, out_size)]
    out = torch.empty((G, Ho, Wo), dtype=torch.uint8, device=masks.device)
    if G:
        rc = _lib.load().pdb_resize_masks_u8(masks.data_ptr(), out.data_ptr(), G, Hp, Wp, Hi, Wi, Ho, Wo, _stream())
        _lib.check(rc, "pdb_resize_masks_u8")
    return out.view(torch.bool)


def pack_bits(masks):
    """bool / uint8 (R, Ho, Wo) -> int32 words (R, Ho, ceil(Wo / 32)); bit i of a word = pixel 32 * word + i."""
    _need_cuda(masks)
    masks = _u8(masks)
    R, Ho, Wo = masks.shape
    bits = torch.empty((R, Ho, (Wo + 31) // 32), dtype=torch.int32, device=masks.device)
    if R:
        rc = _lib.load().pdb_pack_bits(masks.data_ptr(), bits.data_ptr(), R, Ho, Wo, _stream())
        _lib.check(rc, "pdb_pack_bits")
    return bits


def unpack_bits(bits, width, rows=None):
    """int32 words (R0, Ho, Ww) -> bool (R, Ho, width), gathering ``rows`` (int tensor) if given."""
    _need_cuda(bits, rows)
    bits = _c(bits)
    _, Ho, Ww = bits.shape
    if (int(width) + 31) // 32 != Ww:
        raise RuntimeError(f"unpack_bits: width {width} does not match {Ww} words per row")
    if rows is not None:
        rows = _c(rows.to(torch.int32))
    R = rows.shape[0] if rows is not None else bits.shape[0]
    out = torch.empty((R, Ho, int(width)), dtype=torch.uint8, device=bits.device)
    if R:
        rc = _lib.load().pdb_unpack_bits(bits.data_ptr(), rows.data_ptr() if rows is not None else None, out.data_ptr(),
                                         R, Ho, int(width), _stream())
        _lib.check(rc, "pdb_unpack_bits")
    return out.view(torch.bool)


def bits_popcount(bits):
    """int32 words (R, ...) -> int64 (R,) number of set bits per row."""
    _need_cuda(bits)
    bits = _c(bits)
    R = bits.shape[0]
    counts = torch.zeros((R,), dtype=torch.int64, device=bits.device)
    if R and bits[0].numel():
        rc = _lib.load().pdb_bits_popcount(bits.data_ptr(), counts.data_ptr(), R, bits[0].numel(), _stream())
        _lib.check(rc, "pdb_bits_popcount")
    return counts


def bits_iou(a, b):
    """Pairwise mask IoU of packed masks a (Ka, Ho, Ww), b (Kb, Ho, Ww) -> float64 (Ka, Kb) with pycocotools' rleIou
    semantics for iscrowd = 0: |a & b| / |a | b|, exactly 0 where the intersection is empty."""
    _need_cuda(a, b)
    a, b = _c(a), _c(b)
    if a.shape[1:] != b.shape[1:]:
        raise RuntimeError(f"bits_iou: packed shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    Ka, Kb = a.shape[0], b.shape[0]
    inter = torch.zeros((Ka, Kb), dtype=torch.int64, device=a.device)
    if Ka and Kb:
        rc = _lib.load().pdb_bits_intersect(a.data_ptr(), b.data_ptr(), inter.data_ptr(), Ka, Kb, a[0].numel(), _stream())
        _lib.check(rc, "pdb_bits_intersect")
    union = bits_popcount(a)[:, None] + bits_popcount(b)[None] - inter
    return torch.where(inter > 0, inter.double() / union.clamp(min=1).double(), torch.zeros((), dtype=torch.float64, device=a.device))


# --------------------------------------------------------------------------------------------------
# Swin window attention, forward (frozen backbone)  (modeling/backbone/swin.py:78-176)
# --------------------------------------------------------------------------------------------------
def window_attention(qkv, bias, mask, heads, scale):
    """qkv (Bw, N, 3*heads*32) f32 from the qkv Linear, bias (heads, N, N), mask (nW, N, N) additive or None ->
    (Bw, N, heads*32).  No autograd: used when the backbone is frozen."""
    _need_cuda(qkv, bias, mask)
    qkv, bias = _c(qkv), _c(bias.float())
    Bw, N, C3 = qkv.shape
    d = C3 // (3 * heads)
    if mask is not None:
        mask = _c(mask.float())
    out = torch.empty((Bw, N, heads * d), dtype=torch.float32, device=qkv.device)
    rc = _lib.load().pdb_window_attention_forward(qkv.data_ptr(), bias.data_ptr(), mask.data_ptr() if mask is not None else None,
                                                  out.data_ptr(), Bw, N, heads, d, mask.shape[0] if mask is not None else 1,
                                                  float(scale), _stream())
    _lib.check(rc, "pdb_window_attention_forward")
    return out


swin_attention_tensor_cores = os.environ.get("PDB_SWIN_ATTN", "mma") != "ffma"     # A/B switch: the FFMA kernel of round 1


def swin_window_attention(qkv, qkv_bias, bias, heads, window_size, shift, scale, out_dtype=torch.float32):
    """qkv (B, H, W, 3*heads*32) f32 in token order -> (B, H, W, heads*32): the whole shifted-window attention of a Swin
    block (pad, roll, partition, mask, attention, reverse, roll back, crop) in one kernel.  No autograd.
    out_dtype = torch.bfloat16 (autocast: the consumer is the bf16 projection GEMM) is written by the tensor-core kernel itself."""
    _need_cuda(qkv, bias)
    qkv, bias = _c(qkv), _c(bias.float())
    B, H, W, C3 = qkv.shape
    d = C3 // (3 * heads)
    qb = _c(qkv_bias.float()) if qkv_bias is not None else None
    lib = _lib.load()
    tc = getattr(lib, "pdb_swin_window_attention_forward_tc", None)     # absent only in the CPU-tier host builds of the tests
    if swin_attention_tensor_cores and tc is not None and d == 32 and int(window_size) in (12, 8, 4):
        # tensor-core kernel (csrc/window_attn_mma.cu): 3xTF32 = fp32-accurate; a single TF32 pass under bf16 autocast
        passes = 1 if (torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16) else 3
        half = out_dtype == torch.bfloat16
        out = torch.empty((B, H, W, heads * d), dtype=torch.bfloat16 if half else torch.float32, device=qkv.device)
        rc = tc(qkv.data_ptr(), qb.data_ptr() if qb is not None else None, bias.data_ptr(), out.data_ptr(), B, H, W, heads, d,
                int(window_size), int(shift), float(scale), passes, int(half), _stream())
        _lib.check(rc, "pdb_swin_window_attention_forward_tc")
        return out
    out = torch.empty((B, H, W, heads * d), dtype=torch.float32, device=qkv.device)
    rc = lib.pdb_swin_window_attention_forward(qkv.data_ptr(), qb.data_ptr() if qb is not None else None, bias.data_ptr(),
                                               out.data_ptr(), B, H, W, heads, d, int(window_size), int(shift),
                                               float(scale), _stream())
    _lib.check(rc, "pdb_swin_window_attention_forward")
    return out if out_dtype == torch.float32 else out.to(out_dtype)


# --------------------------------------------------------------------------------------------------
# LayerNorm (+ fused residual add)
# --------------------------------------------------------------------------------------------------
class LayerNormFunction(Function):
    """y = LayerNorm(x (+ residual)); optionally also returns the sum.  Forward is one pass of the warp-per-row kernel;
    backward is ATen's native_layer_norm_backward on the saved mean / rstd and the saved sum."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, want_sum):
        _need_cuda(x, residual, weight, bias)
        C = x.shape[-1]
        x2 = _c(x).view(-1, C)
        r2 = _c(residual).view(-1, C) if residual is not None else None
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        # with a residual the kernel also writes the sum whenever a backward will follow: the LayerNorm gradient needs it, and one
        # extra write here is cheaper than re-adding x + residual there (a 3-pass ATen add per LayerNorm of the encoder)
        z = torch.empty_like(x2) if (r2 is not None and (want_sum or any(ctx.needs_input_grad[:4]))) else None
        mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
        rc = _lib.load().pdb_layer_norm_forward(x2.data_ptr(), r2.data_ptr() if r2 is not None else None, weight.data_ptr(),
                                                bias.data_ptr(), y.data_ptr(), z.data_ptr() if z is not None else None,
                                                mean.data_ptr(), rstd.data_ptr(), rows, C, float(eps), _stream())
        _lib.check(rc, "pdb_layer_norm_forward")
        ctx.save_for_backward(x2 if z is None else None, weight, bias, mean, rstd, z)
        ctx.shape = x.shape
        ctx.want_sum = want_sum
        ctx.has_residual = r2 is not None
        if want_sum:
            return y.view(x.shape), (z if z is not None else x2).view(x.shape)
        return y.view(x.shape), None

    @staticmethod
    @once_differentiable
    def backward(ctx, gy, gz):
        x2, weight, bias, mean, rstd, z = ctx.saved_tensors
        zin = z if z is not None else x2                # no residual: the input itself
        C = zin.shape[-1]
        has_res = ctx.has_residual
        mask = [ctx.needs_input_grad[0] or (has_res and ctx.needs_input_grad[1]), ctx.needs_input_grad[2],
                ctx.needs_input_grad[3]]
        gin, gw, gb = torch.ops.aten.native_layer_norm_backward(_c(gy).view(-1, C), zin, [C], mean.view(-1, 1), rstd.view(-1, 1),
                                                                weight, bias, mask)
        if gin is not None:
            if gz is not None:
                gin = gin + _c(gz).view(-1, C)
            gin = gin.view(ctx.shape)
        elif gz is not None:
            gin = gz
        return (gin if ctx.needs_input_grad[0] else None, gin if (has_res and ctx.needs_input_grad[1]) else None,
                gw, gb, None, None)


def _layer_norm_nograd(x, weight, bias, eps, residual, return_sum, residual_scale, out_dtype):
    """The forward-only variants of the LayerNorm kernel (per-sample residual scale, bf16 output); None when they do not apply."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, residual, weight, bias)):
        return None
    f = getattr(_lib.load(), "pdb_layer_norm_forward_scaled", None)
    if f is None or out_dtype not in (torch.float32, torch.bfloat16):
        return None
    if x.is_cuda and residual is not None and residual.dtype == torch.bfloat16:
        residual = residual.float()
    ok = (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 2 and weight is not None and bias is not None
          and weight.dtype == torch.float32 and bias.dtype == torch.float32 and x.shape[-1] % 4 == 0 and x.shape[-1] <= 2048
          and (residual is None or (residual.dtype == torch.float32 and residual.shape == x.shape))
          and (residual_scale is None or (residual is not None and residual_scale.numel() == x.shape[0])))
    if not ok:
        return None
    C = x.shape[-1]
    x2 = _c(x).view(-1, C)
    r2 = _c(residual).view(-1, C) if residual is not None else None
    sc = _c(residual_scale.reshape(-1).float()) if residual_scale is not None else None
    rows = x2.shape[0]
    y = torch.empty((rows, C), dtype=out_dtype, device=x.device)
    z = torch.empty_like(x2) if (return_sum and r2 is not None) else None
    mean = torch.empty((rows,), dtype=torch.float32, device=x.device)
    rstd = torch.empty((rows,), dtype=torch.float32, device=x.device)
    rc = f(x2.data_ptr(), r2.data_ptr() if r2 is not None else None, sc.data_ptr() if sc is not None else None,
           rows // max(1, x.shape[0]), _c(weight).data_ptr(), _c(bias).data_ptr(), y.data_ptr(), z.data_ptr() if z is not None else None,
           mean.data_ptr(), rstd.data_ptr(), rows, C, float(eps), int(out_dtype == torch.bfloat16), _stream())
    _lib.check(rc, "pdb_layer_norm_forward_scaled")
    if return_sum:
        return y.view(x.shape), (z if z is not None else x2).view(x.shape)
    return y.view(x.shape)


def layer_norm(x, weight, bias, eps=1e-5, residual=None, return_sum=False, residual_scale=None, out_dtype=None):
    """F.layer_norm over the last dimension of ``x + residual`` (``residual`` optional).  With ``return_sum`` also
    returns the sum (pre-norm residual streams).  fp32 CUDA tensors with C % 4 == 0 use the kernel; anything else goes
    through torch.  ``residual_scale``: one factor per sample (numel == x.shape[0]) applied to ``residual`` — stochastic
    depth's keep / (1 - p) — folded into the same pass when nothing needs a gradient (the frozen backbone).
    ``out_dtype`` = torch.bfloat16: the normalised output is written as bf16 by the kernel (autocast: it feeds a bf16 GEMM; the
    sum stays fp32) — also only on the no-gradient path; otherwise the result is converted afterwards."""
    if out_dtype is not None and out_dtype != torch.float32:
        r = _layer_norm_nograd(x, weight, bias, eps, residual, return_sum, residual_scale, out_dtype)
        if r is not None:
            return r
        r = layer_norm(x, weight, bias, eps, residual, return_sum, residual_scale)
        return (r[0].to(out_dtype), r[1]) if return_sum else r.to(out_dtype)
    if x.is_cuda and (x.dtype == torch.bfloat16 or (residual is not None and residual.dtype == torch.bfloat16)):
        # autocast: LayerNorm runs (and returns) fp32, the residual stream stays fp32
        x = x.float()
        residual = None if residual is None else residual.float()
    if residual_scale is not None:
        r = _layer_norm_nograd(x, weight, bias, eps, residual, return_sum, residual_scale, torch.float32)
        if r is not None:
            return r
        residual = residual * residual_scale.view((x.shape[0],) + (1,) * (x.dim() - 1)).to(residual.dtype)
    ok = (x.is_cuda and x.dtype == torch.float32 and weight is not None and bias is not None and weight.dtype == torch.float32
          and x.shape[-1] % 4 == 0 and x.shape[-1] <= 2048 and (residual is None or residual.shape == x.shape))
    if not ok:
        z = x if residual is None else x + residual
        y = torch.nn.functional.layer_norm(z, (x.shape[-1],), weight, bias, eps)
        return (y, z) if return_sum else y
    y, z = LayerNormFunction.apply(x, residual, weight, bias, eps, return_sum)
    return (y, z) if return_sum else y


# --------------------------------------------------------------------------------------------------
# FPN top-down step: lateral + bilinear up-sampling  (msdeformattn.py:352-356)  — csrc/upsample.cu
# --------------------------------------------------------------------------------------------------
class UpsampleAddFunction(Function):
    """lateral + F.interpolate(x, size=lateral.shape[-2:], mode="bilinear", align_corners=False) on channels-last maps: one
    forward pass, gather backward (deterministic); the gradient of ``lateral`` is the incoming gradient itself."""

    @staticmethod
    def forward(ctx, x, lateral):
        _need_cuda(x, lateral)
        B, C, h, w = x.shape
        H, W = lateral.shape[-2:]
        xp = x.permute(0, 2, 3, 1)                        # pixel-major view; rows of a batch item must be dense
        if not (xp.stride(3) == 1 and xp.stride(2) == C and xp.stride(1) == w * C and xp.stride(0) % 4 == 0
                and xp.data_ptr() % 16 == 0):
            xp = xp.contiguous()
        lp = _c(lateral.permute(0, 2, 3, 1))
        out = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
        rc = _lib.load().pdb_upsample_add_forward(xp.data_ptr(), lp.data_ptr(), out.data_ptr(), B, h, w, H, W, C, xp.stride(0),
                                                  _stream())
        _lib.check(rc, "pdb_upsample_add_forward")
        ctx.geom = (B, C, h, w, H, W)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        B, C, h, w, H, W = ctx.geom
        gp = _c(gy.permute(0, 2, 3, 1))                   # no copy when gy is channels-last
        gx = None
        if ctx.needs_input_grad[0]:
            gxp = torch.empty((B, h, w, C), dtype=torch.float32, device=gy.device)
            rc = _lib.load().pdb_upsample_backward(gp.data_ptr(), gxp.data_ptr(), B, h, w, H, W, C, _stream())
            _lib.check(rc, "pdb_upsample_backward")
            gx = gxp.permute(0, 3, 1, 2)
        return gx, (gy if ctx.needs_input_grad[1] else None)


def upsample_add(x, lateral):
    """lateral + bilinear up-sampling of x to lateral's size (align_corners=False), fp32 CUDA maps with C % 4 == 0 on the
    library kernel; anything else through F.interpolate."""
    if (x.is_cuda and x.dtype == torch.float32 and lateral.dtype == torch.float32 and x.dim() == 4 and x.shape[1] % 4 == 0
            and x.shape[:2] == lateral.shape[:2] and getattr(_lib.load(), "pdb_upsample_add_forward", None) is not None):
        return UpsampleAddFunction.apply(x, lateral)
    return lateral + torch.nn.functional.interpolate(x, size=lateral.shape[-2:], mode="bilinear", align_corners=False)


# --------------------------------------------------------------------------------------------------
# GroupNorm (+ ReLU) on channels-last maps  (msdeformattn.py:249-287: Conv2d(norm=get_norm("GN", C), activation=F.relu))
# --------------------------------------------------------------------------------------------------
class GroupNormFunction(Function):
    """x: logical (B, C, H, W) in channels-last memory.  Returns the same logical shape, channels-last."""

    @staticmethod
    def forward(ctx, x, weight, bias, num_groups, eps, relu):
        _need_cuda(x, weight, bias)
        B, C, H, W = x.shape
        xp = x.permute(0, 2, 3, 1)
        if not xp.is_contiguous():
            xp = xp.contiguous()
        w, b = _c(weight), _c(bias)
        y = torch.empty_like(xp)
        stats = torch.zeros((B, num_groups, 2), dtype=torch.float64, device=x.device)
        mean = torch.empty((B, num_groups), dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        rc = _lib.load().pdb_group_norm_forward(xp.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), stats.data_ptr(),
                                                mean.data_ptr(), rstd.data_ptr(), B, H * W, C, num_groups, float(eps), int(relu),
                                                _stream())
        _lib.check(rc, "pdb_group_norm_forward")
        ctx.save_for_backward(xp, w, b, mean, rstd)
        ctx.cfg = (B, C, H, W, num_groups, bool(relu))
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        xp, w, b, mean, rstd = ctx.saved_tensors
        B, C, H, W, G, relu = ctx.cfg
        gyp = gy.permute(0, 2, 3, 1)
        if not gyp.is_contiguous() or gyp.data_ptr() % 16:
            gyp = gyp.contiguous()
        sums = torch.zeros((B, C, 2), dtype=torch.float64, device=gy.device)
        coef = torch.empty((B, G, 2), dtype=torch.float32, device=gy.device)
        gx = torch.empty_like(xp) if ctx.needs_input_grad[0] else None
        rc = _lib.load().pdb_group_norm_backward(gyp.data_ptr(), xp.data_ptr(), w.data_ptr(), b.data_ptr(), mean.data_ptr(),
                                                 rstd.data_ptr(), sums.data_ptr(), coef.data_ptr(),
                                                 gx.data_ptr() if gx is not None else None, B, H * W, C, G, int(relu), _stream())
        _lib.check(rc, "pdb_group_norm_backward")
        gw = gb = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            tot = sums.sum(0).float()
            gw, gb = tot[:, 0].contiguous(), tot[:, 1].contiguous()
        return (gx.permute(0, 3, 1, 2) if gx is not None else None), gw, gb, None, None, None


def group_norm_supported(x, num_groups, weight, bias):
    if not (torch.is_tensor(x) and x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and weight is not None
            and bias is not None and weight.dtype == torch.float32 and bias.dtype == torch.float32):
        return False
    C = x.shape[1]
    if C % num_groups or (C // num_groups) % 4 or C // 4 > 256 or 256 % (C // 4) or num_groups > 256 or x.numel() == 0:
        return False
    # channels-last memory (what conv1x1 / conv3x3 of this module return); an NCHW input would need a transposing copy first
    return x.permute(0, 2, 3, 1).is_contiguous() and x.data_ptr() % 16 == 0


def group_norm(x, num_groups, weight, bias, eps=1e-5, relu=False):
    """F.group_norm (optionally followed by ReLU) for a logical (B, C, H, W) map.  fp32 CUDA maps in channels-last memory
    use the pixel-major kernels; everything else goes through torch."""
    if group_norm_supported(x, num_groups, weight, bias) and not torch.is_autocast_enabled():
        return GroupNormFunction.apply(x, weight, bias, int(num_groups), float(eps), bool(relu))
    y = torch.nn.functional.group_norm(x, num_groups, weight, bias, eps)
    return torch.relu(y) if relu else y
