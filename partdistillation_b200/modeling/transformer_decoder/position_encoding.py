"""Sine position embedding (position_encoding.py:16-56 of the reference), cached per shape.

The embedding is a pure function of (H, W, num_pos_feats): the reference recomputes it 6x per step;
here it is computed once per shape and device and expanded over the batch without a copy."""
import math

import torch
from torch import nn


class PositionEmbeddingSine(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        self.num_pos_feats = num_pos_feats
        self.temperature = temperature
        self.normalize = normalize
        self.scale = 2 * math.pi if scale is None else scale
        self._cache = {}

    def _compute(self, H, W, device):
        ones = torch.ones((1, H, W), dtype=torch.float32, device=device)
        y_embed = ones.cumsum(1)
        x_embed = ones.cumsum(2)
        if self.normalize:
            eps = 1e-6
            y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
            x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        dim_t = torch.arange(self.num_pos_feats, dtype=torch.float32, device=device)
        dim_t = self.temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_t
        pos_y = y_embed[:, :, :, None] / dim_t
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2).contiguous()      # (1, 2F, H, W)

    def forward(self, x, mask=None):
        if mask is not None:
            raise NotImplementedError("padding masks are not produced on the training path")
        B, _, H, W = x.shape
        key = (H, W, x.device)
        pos = self._cache.get(key)
        if pos is None:
            with torch.no_grad():
                pos = self._cache[key] = self._compute(H, W, x.device)
        return pos.expand(B, -1, -1, -1)
