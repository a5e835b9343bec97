"""MSDeformAttn pixel decoder (reference: modeling/pixel_decoder/msdeformattn.py:27-362).

Same registry name, constructor / ``from_config`` keys, parameter names and ``forward_features``
contract.  Differences are structural only: spatial shapes stay on the host (no int64 shape tensors
read back per step), reference points / position embeddings are cached per shape, and the sampling
core is the sm_100a kernel.
"""
import copy
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from ... import functional as PF
from ...compat import SEM_SEG_HEADS_REGISTRY, Conv2d, ShapeSpec, c2_xavier_fill, configurable, get_norm
from ..transformer_decoder.position_encoding import PositionEmbeddingSine
from .ops.modules import MSDeformAttn


def _clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


def _activation(name):
    try:
        return {"relu": F.relu, "gelu": F.gelu, "glu": F.glu}[name]
    except KeyError:
        raise RuntimeError(f"activation should be relu/gelu, not {name}.")


class MSDeformAttnTransformerEncoderLayer(nn.Module):
    """deformable self-attention -> +res -> LN -> FFN -> +res -> LN  (msdeformattn.py:96-135)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.activation = _activation(activation)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None,
                offset_normalizer=None):
        q = src if pos is None else src + pos
        src2 = self.self_attn(q, reference_points, src, spatial_shapes, level_start_index, padding_mask,
                              offset_normalizer=offset_normalizer)
        src = PF.layer_norm(src, self.norm1.weight, self.norm1.bias, self.norm1.eps, residual=self.dropout1(src2))
        if self.activation is F.relu and not (self.training and self.dropout2.p > 0):
            src2 = PF.ffn(src, self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias)
        else:
            hidden = self.activation(PF.linear(src, self.linear1.weight, self.linear1.bias))
            src2 = PF.linear(self.dropout2(hidden), self.linear2.weight, self.linear2.bias)
        return PF.layer_norm(src, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=self.dropout3(src2))


class MSDeformAttnTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers):
        super().__init__()
        self.layers = _clones(encoder_layer, num_layers)
        self.num_layers = num_layers
        self._ref_cache = {}

    @staticmethod
    def get_reference_points(spatial_shapes, valid_ratios, device):
        """Pixel centres of every level, normalised, times the valid ratios (msdeformattn.py:144-157)."""
        pts = []
        for lvl, (H_, W_) in enumerate(spatial_shapes):
            ref_y, ref_x = torch.meshgrid(torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device),
                                          torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device),
                                          indexing="ij")
            ref_y = ref_y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H_)
            ref_x = ref_x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W_)
            pts.append(torch.stack((ref_x, ref_y), -1))
        ref = torch.cat(pts, 1)
        return ref[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios=None, pos=None, padding_mask=None):
        shapes = [(int(h), int(w)) for h, w in spatial_shapes]
        B = src.shape[0]
        if valid_ratios is None:       # no padding masks on the training path -> ratios are exactly 1: cacheable
            key = (tuple(shapes), B, src.device)
            hit = self._ref_cache.get(key)
            if hit is None:
                with torch.no_grad():
                    ones = torch.ones((B, len(shapes), 2), dtype=torch.float32, device=src.device)
                    ref = self.get_reference_points(shapes, ones, src.device).contiguous()
                    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32, device=src.device)
                hit = self._ref_cache[key] = (ref, norm)
            reference_points, normalizer = hit
        else:
            reference_points = self.get_reference_points(shapes, valid_ratios, src.device)
            normalizer = None
        out = src
        for layer in self.layers:
            out = layer(out, pos, reference_points, shapes, level_start_index, padding_mask,
                        offset_normalizer=normalizer)
        return out


class MSDeformAttnTransformerEncoderOnly(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_encoder_layers=6, dim_feedforward=1024, dropout=0.1,
                 activation="relu", num_feature_levels=4, enc_n_points=4):
        super().__init__()
        self.d_model = d_model
        self.nhead = nhead
        layer = MSDeformAttnTransformerEncoderLayer(d_model, dim_feedforward, dropout, activation,
                                                    num_feature_levels, nhead, enc_n_points)
        self.encoder = MSDeformAttnTransformerEncoder(layer, num_encoder_layers)
        self.level_embed = nn.Parameter(torch.Tensor(num_feature_levels, d_model))
        self._reset_parameters()

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        nn.init.normal_(self.level_embed)

    def forward(self, srcs, pos_embeds):
        """srcs / pos_embeds: per level (B, C, H, W), low resolution first.  Returns
        (memory (B, S, C), spatial_shapes [(H, W)], level_start_index [int])."""
        shapes, starts, flat, posf, s = [], [], [], [], 0
        for lvl, (src, pos) in enumerate(zip(srcs, pos_embeds)):
            h, w = src.shape[-2:]
            shapes.append((h, w))
            starts.append(s)
            s += h * w
            flat.append(src.flatten(2).transpose(1, 2))
            posf.append(pos.flatten(2).transpose(1, 2) + self.level_embed[lvl].view(1, 1, -1))
        src_flatten = torch.cat(flat, 1)
        pos_flatten = torch.cat(posf, 1)
        memory = self.encoder(src_flatten, shapes, starts, None, pos_flatten, None)
        return memory, shapes, starts


@SEM_SEG_HEADS_REGISTRY.register()
class MSDeformAttnPixelDecoder(nn.Module):
    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, transformer_dropout: float, transformer_nheads: int,
                 transformer_dim_feedforward: int, transformer_enc_layers: int, conv_dim: int, mask_dim: int,
                 norm: Optional[Union[str, Callable]] = None, transformer_in_features: List[str],
                 common_stride: int):
        super().__init__()
        by_stride = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in by_stride]                    # res2 .. res5
        self.feature_strides = [v.stride for _, v in by_stride]
        self.feature_channels = [v.channels for _, v in by_stride]
        tr = sorted(((k, v) for k, v in input_shape.items() if k in transformer_in_features),
                    key=lambda kv: kv[1].stride)
        self.transformer_in_features = [k for k, _ in tr]
        tr_channels = [v.channels for _, v in tr]
        self.transformer_feature_strides = [v.stride for _, v in tr]
        self.transformer_num_feature_levels = len(tr)

        # 1x1 conv + GN per encoder level, low resolution (res5) first
        chans = tr_channels[::-1] if self.transformer_num_feature_levels > 1 else [tr_channels[-1]]
        self.input_proj = nn.ModuleList(
            [nn.Sequential(nn.Conv2d(c, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim)) for c in chans])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.constant_(proj[0].bias, 0)

        self.transformer = MSDeformAttnTransformerEncoderOnly(
            d_model=conv_dim, dropout=transformer_dropout, nhead=transformer_nheads,
            dim_feedforward=transformer_dim_feedforward, num_encoder_layers=transformer_enc_layers,
            num_feature_levels=self.transformer_num_feature_levels)
        self.pe_layer = PositionEmbeddingSine(conv_dim // 2, normalize=True)

        self.mask_dim = mask_dim
        self.mask_features = Conv2d(conv_dim, mask_dim, kernel_size=1, stride=1, padding=0)
        c2_xavier_fill(self.mask_features)
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride

        # extra FPN levels between the finest encoder level and the common stride
        self.num_fpn_levels = int(np.log2(min(self.transformer_feature_strides)) - np.log2(self.common_stride))
        lateral, output = [], []
        use_bias = norm == ""
        for idx, in_ch in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lat = Conv2d(in_ch, conv_dim, kernel_size=1, bias=use_bias, norm=get_norm(norm, conv_dim))
            out = Conv2d(conv_dim, conv_dim, kernel_size=3, stride=1, padding=1, bias=use_bias,
                         norm=get_norm(norm, conv_dim), activation=F.relu)
            c2_xavier_fill(lat)
            c2_xavier_fill(out)
            self.add_module(f"adapter_{idx + 1}", lat)
            self.add_module(f"layer_{idx + 1}", out)
            lateral.append(lat)
            output.append(out)
        self.lateral_convs = lateral[::-1]      # top-down order
        self.allow_tf32_conv = False
        self.output_convs = output[::-1]

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        h = cfg.MODEL.SEM_SEG_HEAD
        return dict(
            input_shape={k: v for k, v in input_shape.items() if k in h.IN_FEATURES},
            conv_dim=h.CONVS_DIM, mask_dim=h.MASK_DIM, norm=h.NORM,
            transformer_dropout=cfg.MODEL.MASK_FORMER.DROPOUT, transformer_nheads=cfg.MODEL.MASK_FORMER.NHEADS,
            transformer_dim_feedforward=1024,          # fixed by the reference (msdeformattn.py:310)
            transformer_enc_layers=h.TRANSFORMER_ENC_LAYERS,
            transformer_in_features=h.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES,
            common_stride=h.COMMON_STRIDE)

    def forward_features(self, features):
        """-> (mask_features (B, mask_dim, H/4, W/4), coarsest encoder map, [3 multi-scale maps])."""
        # the deformable encoder is fp32-only (:318,324); cuDNN's TF32 convolution default is switched off
        # here so that the fp32 contract (<= 1e-3 on the mask logits) holds against the fp32 reference
        with torch.autocast("cuda", enabled=False), torch.backends.cudnn.flags(allow_tf32=self.allow_tf32_conv):
            srcs, pos = [], []
            for idx, f in enumerate(self.transformer_in_features[::-1]):
                x = features[f].float()
                conv, gn = self.input_proj[idx][0], self.input_proj[idx][1]
                srcs.append(self._norm_act(gn, None, PF.conv1x1(x, conv.weight, conv.bias)))
                pos.append(self.pe_layer(x))
            y, shapes, starts = self.transformer(srcs, pos)
            bs = y.shape[0]
            out = []
            for i, (h, w) in enumerate(shapes):
                end = starts[i + 1] if i + 1 < len(starts) else y.shape[1]
                out.append(y[:, starts[i]:end].transpose(1, 2).reshape(bs, -1, h, w))
            for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
                cur = self._lateral(self.lateral_convs[idx], features[f].float())
                out.append(self._output_conv(self.output_convs[idx], PF.upsample_add(out[-1], cur)))
            multi_scale = out[:self.maskformer_num_feature_levels]
            return self._mask_features_pixel_major(out[-1]), out[0], multi_scale

    @staticmethod
    def _norm_act(norm, activation, y):
        """norm (+ activation) behind a convolution: GroupNorm (+ ReLU) runs as one pixel-major kernel pair on the
        channels-last map the convolution GEMMs produce; any other norm / activation goes through its module."""
        if isinstance(norm, nn.GroupNorm) and norm.affine and (activation is None or activation is F.relu):
            return PF.group_norm(y, norm.num_groups, norm.weight, norm.bias, norm.eps, relu=activation is F.relu)
        if norm is not None:
            y = norm(y)
        if activation is not None:
            y = activation(y)
        return y

    @classmethod
    def _lateral(cls, conv, x):
        """1x1 lateral convolution (+ its norm / activation) with the contraction on the tensor cores."""
        if tuple(conv.kernel_size) != (1, 1):
            return conv(x)
        y = PF.conv1x1(x, conv.weight, conv.bias)
        return cls._norm_act(getattr(conv, "norm", None), getattr(conv, "activation", None), y)

    @classmethod
    def _output_conv(cls, conv, x):
        """3x3 output convolution (+ norm / activation) with the contraction on the tensor cores."""
        if tuple(conv.kernel_size) != (3, 3) or tuple(conv.stride) != (1, 1) or tuple(conv.padding) != (1, 1) or \
                tuple(conv.dilation) != (1, 1) or conv.groups != 1:
            return conv(x)
        y = PF.conv3x3(x, conv.weight, conv.bias)
        return cls._norm_act(getattr(conv, "norm", None), getattr(conv, "activation", None), y)

    def _mask_features_pixel_major(self, y):
        """The 1x1 ``mask_features`` convolution as a GEMM over pixels: the result is the logical
        (B, mask_dim, H, W) tensor stored channels-last, i.e. the pixel-major (B, HW, C) operand the
        tensor-core mask einsum consumes (both of its operands are then K-major)."""
        conv = self.mask_features
        w = conv.weight.view(conv.out_channels, conv.in_channels)
        mf = PF.linear(y.permute(0, 2, 3, 1).contiguous(), w, conv.bias)   # (B, H, W, mask_dim), contiguous
        if getattr(conv, "norm", None) is not None:
            mf = conv.norm(mf.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        if getattr(conv, "activation", None) is not None:
            mf = conv.activation(mf)
        return mf.permute(0, 3, 1, 2)
