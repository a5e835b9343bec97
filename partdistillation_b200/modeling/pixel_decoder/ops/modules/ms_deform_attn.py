"""MSDeformAttn module — same constructor, parameters and forward signature as the reference's
ops/modules/ms_deform_attn.py:38-131; the sampling core is the sm_100a kernel behind
``functional.ms_deform_attn`` (the reference silently falls back to grid_sample: §0.1 of SURVEY.md).
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from ..... import functional as PF


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        self.im2col_step = 128
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """Offsets start as a ring of n_heads directions, point p at (p+1) pixels; uniform weights."""
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            ang = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
            ring = torch.stack([ang.cos(), ang.sin()], -1)
            ring = ring / ring.abs().max(-1, keepdim=True)[0]
            steps = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, -1, 1)
            bias = ring.view(self.n_heads, 1, 1, 2) * steps.expand(self.n_heads, self.n_levels, self.n_points, 1)
            self.sampling_offsets.bias.copy_(bias.reshape(-1))
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            nn.init.xavier_uniform_(self.value_proj.weight)
            self.value_proj.bias.zero_()
            nn.init.xavier_uniform_(self.output_proj.weight)
            self.output_proj.bias.zero_()

    fold_normalizer = True      # False: the reference's expression, two projections and a division pass

    def _merged_projection(self, spatial_shapes, normalizer, query):
        """[sampling_offsets / (W_l, H_l) ; attention_weights] as one (M*L*P*3, C) weight and bias.  Frozen weights (the recipe
        freezes the encoder) and python shapes: built once per (shapes, weight versions) and kept; otherwise rebuilt inside
        autograd on every call (four element-wise launches on 288 x 256 numbers)."""
        M, L, P = self.n_heads, self.n_levels, self.n_points
        so, aw = self.sampling_offsets, self.attention_weights
        frozen = not (torch.is_grad_enabled() and any(t.requires_grad for t in (so.weight, so.bias, aw.weight, aw.bias)))
        key = None
        if frozen and not isinstance(spatial_shapes, torch.Tensor):
            key = (tuple((int(h), int(w)) for h, w in spatial_shapes), query.device, query.dtype,
                   so.weight._version, so.bias._version, aw.weight._version, aw.bias._version,
                   so.weight.data_ptr(), aw.weight.data_ptr())
            hit = self.__dict__.get("_merged")
            if hit is not None and hit[0] == key:
                return hit[1], hit[2]
            normalizer = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=torch.float32, device=query.device)
        elif normalizer is None:
            normalizer = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=torch.float32, device=query.device)
        inv = (1.0 / normalizer.to(torch.float32)).view(1, L, 1, 2).expand(M, L, P, 2).reshape(-1)
        W = torch.cat([so.weight * inv[:, None], aw.weight])
        b = torch.cat([so.bias * inv, aw.bias])
        if key is not None:
            old = self.__dict__.get("_merged")
            if old is not None:
                PF.forget_weight(old[1])        # its address may be handed to the next tensor of this shape
            W, b = W.detach(), b.detach()
            self.__dict__["_merged"] = (key, W, b)
        return W, b

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, offset_normalizer=None):
        """query (N, Lq, C); reference_points (N, Lq, L, 2|4) in [0, 1]; input_flatten (N, S, C);
        input_spatial_shapes: python [(H, W)] (no sync) or the reference's (L, 2) int64 tensor."""
        N, Lq, _ = query.shape
        S = input_flatten.shape[1]
        M, L, P = self.n_heads, self.n_levels, self.n_points
        value = PF.linear(input_flatten, self.value_proj.weight, self.value_proj.bias)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], 0.0)
        value = value.view(N, S, M, self.d_model // M)
        if reference_points.shape[-1] == 2 and self.fold_normalizer and query.is_cuda:
            # ONE projection for offsets and attention logits (the query is read once), with the division by (W_l, H_l) of
            # `offsets / offset_normalizer` (ms_deform_attn.py:107-109) folded into the offset rows of the weight and bias
            if offset_normalizer is None and isinstance(input_spatial_shapes, torch.Tensor):
                offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            W, b = self._merged_projection(input_spatial_shapes, offset_normalizer, query)
            off, logits = PF.split_columns(PF.linear(query, W, b), M * L * P * 2)
            offsets = off.view(N, Lq, M, L, P, 2)
            weights = F.softmax(logits.reshape(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
            loc = reference_points[:, :, None, :, None, :] + offsets
            out = PF.ms_deform_attn(value, input_spatial_shapes, input_level_start_index, loc.contiguous(),
                                    weights, self.im2col_step)
            return PF.linear(out, self.output_proj.weight, self.output_proj.bias)
        offsets = PF.linear(query, self.sampling_offsets.weight, self.sampling_offsets.bias).view(N, Lq, M, L, P, 2)
        weights = F.softmax(PF.linear(query, self.attention_weights.weight, self.attention_weights.bias).view(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
        if reference_points.shape[-1] == 2:
            if offset_normalizer is None:
                if isinstance(input_spatial_shapes, torch.Tensor):
                    offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
                else:
                    offset_normalizer = torch.tensor([[w, h] for h, w in input_spatial_shapes],
                                                     dtype=query.dtype, device=query.device)
            loc = reference_points[:, :, None, :, None, :] + offsets / offset_normalizer[None, None, None, :, None, :]
        elif reference_points.shape[-1] == 4:
            loc = reference_points[:, :, None, :, None, :2] \
                + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
        else:
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but get {reference_points.shape[-1]} instead.")
        out = PF.ms_deform_attn(value, input_spatial_shapes, input_level_start_index, loc.contiguous(),
                                weights, self.im2col_step)
        return PF.linear(out, self.output_proj.weight, self.output_proj.bias)
