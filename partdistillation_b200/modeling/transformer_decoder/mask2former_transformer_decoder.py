"""Masked-attention transformer decoder (reference: mask2former_transformer_decoder.py:21-472).

Same registry name, ``from_config`` keys, parameter names and output dict.  Internally the decoder is
batch-first (B, Q, C) / (B, HW, C); the three hot operators are sm_100a kernels:
  * masked cross-attention core   -> functional.masked_cross_attention   (:102-114)
  * mask-head einsum              -> functional.mask_einsum              (:449)
  * attention-mask build + reset  -> functional.build_attention_mask     (:453-457, :405)
The attention mask is kept compact, (B, Q, hw) uint8 shared by all heads plus a per-row "has an attended
key" flag, instead of the reference's (B*heads, Q, hw) bool tensor; ``AttnMask.as_reference`` expands it.
"""
import math
from collections import namedtuple

import torch
from torch import nn
from torch.nn import functional as F

from ... import functional as PF
from ...compat import TRANSFORMER_DECODER_REGISTRY, Conv2d, c2_xavier_fill, configurable
from .position_encoding import PositionEmbeddingSine


class AttnMask(namedtuple("AttnMask", ["mask", "row_any"])):
    """mask (B, Q, hw) uint8, 1 = key not attended; row_any (B*Q,) int32, 0 = row fully masked."""

    def as_reference(self, num_heads, reset_full_rows=False):
        m = self.mask.bool()
        if reset_full_rows:
            m = m & (self.row_any.view(m.shape[0], m.shape[1], 1) != 0)
        return m.unsqueeze(1).repeat(1, num_heads, 1, 1).flatten(0, 1)


class _AttentionParams(nn.Module):
    """Parameter container with nn.MultiheadAttention's names: in_proj_weight/bias, out_proj.*"""

    def __init__(self, embed_dim, num_heads):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def project(self, x, lo, hi):
        """Rows lo*E .. hi*E of the packed q / k / v projection; fp32 out (the attention cores are fp32 kernels)."""
        E = self.embed_dim
        return PF.linear(x, self.in_proj_weight[lo * E:hi * E], self.in_proj_bias[lo * E:hi * E], out_fp32=True)


def _xavier(module):
    for p in module.parameters():
        if p.dim() > 1:
            nn.init.xavier_uniform_(p)


class SelfAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.self_attn = _AttentionParams(d_model, nhead)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        a = self.self_attn
        x = self.norm(tgt) if self.normalize_before else tgt
        qk_in = x if query_pos is None else x + query_pos
        B, Q, E = x.shape
        qk = a.project(qk_in, 0, 2)                                    # (B, Q, 2E): one GEMM for the query and key projections
        v = a.project(x, 2, 3)
        # the attention core is the library's own kernel (csrc/xattn.cu) with no mask: softmax(q k^T / sqrt(d)) v per head
        o = PF.masked_cross_attention((qk[..., :E] * (1.0 / math.sqrt(a.head_dim))).float(), qk[..., E:].float(), v.float(),
                                      None, None, a.num_heads)
        o = PF.linear(o.to(x.dtype), a.out_proj.weight, a.out_proj.bias)
        return tgt + o if self.normalize_before else PF.layer_norm(tgt, self.norm.weight, self.norm.bias, self.norm.eps, residual=o)


class CrossAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.multihead_attn = _AttentionParams(d_model, nhead)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None,
                memory_with_pos=None):
        """tgt (B, Q, C); memory (B, HW, C); memory_mask: AttnMask or None."""
        a = self.multihead_attn
        x = self.norm(tgt) if self.normalize_before else tgt
        q_in = x if query_pos is None else x + query_pos
        if memory_with_pos is None:
            memory_with_pos = memory if pos is None else memory + pos
        q = a.project(q_in, 0, 1) * (1.0 / math.sqrt(a.head_dim))
        k = a.project(memory_with_pos, 1, 2)
        v = a.project(memory, 2, 3)
        mask, row_any = (memory_mask.mask, memory_mask.row_any) if memory_mask is not None else (None, None)
        o = PF.masked_cross_attention(q.float(), k.float(), v.float(), mask, row_any, a.num_heads)
        o = PF.linear(o.to(x.dtype), a.out_proj.weight, a.out_proj.bias)
        return tgt + o if self.normalize_before else PF.layer_norm(tgt, self.norm.weight, self.norm.bias, self.norm.eps, residual=o)


class FFNLayer(nn.Module):
    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)
        self.normalize_before = normalize_before
        _xavier(self)

    def forward(self, tgt):
        x = self.norm(tgt) if self.normalize_before else tgt
        y = PF.linear(PF.linear(x, self.linear1.weight, self.linear1.bias, relu=True), self.linear2.weight, self.linear2.bias)
        return tgt + y if self.normalize_before else PF.layer_norm(tgt, self.norm.weight, self.norm.bias, self.norm.eps, residual=y)


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):          # the last layer feeds the fp32 mask einsum
            x = PF.linear(x, layer.weight, layer.bias, relu=i < self.num_layers - 1, out_fp32=i == self.num_layers - 1)
        return x


@TRANSFORMER_DECODER_REGISTRY.register()
class MultiScaleMaskedTransformerDecoder(nn.Module):
    _version = 2

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # checkpoints older than version 2 call query_feat "static_query" (reference :216-237)
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            for k in list(state_dict.keys()):
                if k.startswith(prefix + "static_query"):
                    state_dict[k.replace("static_query", "query_feat")] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    @configurable
    def __init__(self, in_channels, mask_classification=True, *, num_classes: int, hidden_dim: int, num_queries: int,
                 nheads: int, dim_feedforward: int, dec_layers: int, pre_norm: bool, mask_dim: int,
                 enforce_input_project: bool, query_feature_normalize: bool):
        super().__init__()
        assert mask_classification, "Only support mask classification model"
        self.mask_classification = mask_classification
        self.pe_layer = PositionEmbeddingSine(hidden_dim // 2, normalize=True)
        self.num_heads = nheads
        self.num_layers = dec_layers
        self.hidden_dim = hidden_dim
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        for _ in range(dec_layers):
            self.transformer_self_attention_layers.append(SelfAttentionLayer(hidden_dim, nheads, 0.0, normalize_before=pre_norm))
            self.transformer_cross_attention_layers.append(CrossAttentionLayer(hidden_dim, nheads, 0.0, normalize_before=pre_norm))
            self.transformer_ffn_layers.append(FFNLayer(hidden_dim, dim_feedforward, 0.0, normalize_before=pre_norm))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.num_queries = num_queries
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.num_feature_levels = 3
        self.level_embed = nn.Embedding(self.num_feature_levels, hidden_dim)
        self.input_proj = nn.ModuleList()
        for _ in range(self.num_feature_levels):
            if in_channels != hidden_dim or enforce_input_project:
                self.input_proj.append(Conv2d(in_channels, hidden_dim, kernel_size=1))
                c2_xavier_fill(self.input_proj[-1])
            else:
                self.input_proj.append(nn.Sequential())
        self._build_class_embed(hidden_dim, num_classes)
        self.mask_embed = MLP(hidden_dim, hidden_dim, mask_dim, 3)
        self.query_feature_normalize = query_feature_normalize

    def _build_class_embed(self, hidden_dim, num_classes):
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        m = cfg.MODEL.MASK_FORMER
        assert m.DEC_LAYERS >= 1
        return dict(in_channels=in_channels, mask_classification=mask_classification,
                    num_classes=cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES, hidden_dim=m.HIDDEN_DIM,
                    num_queries=m.NUM_OBJECT_QUERIES, nheads=m.NHEADS, dim_feedforward=m.DIM_FEEDFORWARD,
                    dec_layers=m.DEC_LAYERS - 1,     # the learnable queries are supervised too (:357-362)
                    pre_norm=m.PRE_NORM, enforce_input_project=m.ENFORCE_INPUT_PROJ,
                    query_feature_normalize=m.QUERY_FEATURE_NORMALIZE, mask_dim=cfg.MODEL.SEM_SEG_HEAD.MASK_DIM)

    # ------------------------------------------------------------------------------------------
    def _prepare_memory(self, x):
        """Per level: memory (B, HW, C) with the level embedding added, memory + position, size."""
        mem, mem_pos, sizes = [], [], []
        for i in range(self.num_feature_levels):
            sizes.append(tuple(x[i].shape[-2:]))
            pos = self.pe_layer(x[i], None).flatten(2).transpose(1, 2)
            src = self.input_proj[i](x[i]).flatten(2).transpose(1, 2) + self.level_embed.weight[i][None, None, :]
            mem.append(src)
            mem_pos.append(src + pos)
        return mem, mem_pos, sizes

    def _classify(self, decoder_output, targets):
        return PF.linear(decoder_output, self.class_embed.weight, self.class_embed.bias)

    def forward(self, x, mask_features, mask=None):
        assert len(x) == self.num_feature_levels
        targets = mask          # only the PartDistillation subclass reads it
        mem, mem_pos, sizes = self._prepare_memory(x)
        bs = mem[0].shape[0]
        query_embed = self.query_embed.weight.unsqueeze(0).expand(bs, -1, -1)
        output = self.query_feat.weight.unsqueeze(0).expand(bs, -1, -1)
        # the num_layers + 1 prediction heads share mask_features: their einsum backwards accumulate into one gradient buffer
        mask_features, self._mask_grad_acc = PF.grad_fan_out(mask_features.float())
        classes, masks = [], []
        c, m, attn_mask, decoder_output = self.forward_prediction_heads(output, mask_features, sizes[0], targets)
        classes.append(c)
        masks.append(m)
        for i in range(self.num_layers):
            lvl = i % self.num_feature_levels
            # rows of attn_mask that are entirely True are reset to False (:405): the kernel does this
            # through attn_mask.row_any
            output = self.transformer_cross_attention_layers[i](
                output, mem[lvl], memory_mask=attn_mask, query_pos=query_embed, memory_with_pos=mem_pos[lvl])
            output = self.transformer_self_attention_layers[i](output, query_pos=query_embed)
            output = self.transformer_ffn_layers[i](output)
            c, m, attn_mask, decoder_output = self.forward_prediction_heads(
                output, mask_features, sizes[(i + 1) % self.num_feature_levels], targets)
            classes.append(c)
            masks.append(m)
        self._mask_grad_acc = None
        out = {"pred_logits": classes[-1], "pred_masks": masks[-1], "decoder_output": decoder_output,
               "aux_outputs": self._set_aux_loss(classes if self.mask_classification else None, masks)}
        self._extra_outputs(out, output)
        return out

    def _extra_outputs(self, out, output):
        pass

    def forward_prediction_heads(self, output, mask_features, attn_mask_target_size, targets=None):
        """output (B, Q, C) -> (class logits, mask logits (B, Q, H, W), AttnMask for the next layer,
        normalised decoder output (B, Q, C))."""
        decoder_output = PF.layer_norm(output, self.decoder_norm.weight, self.decoder_norm.bias, self.decoder_norm.eps)
        outputs_class = self._classify(decoder_output, targets)
        mask_embed = self.mask_embed(decoder_output)
        if self.query_feature_normalize:
            mask_embed = F.normalize(mask_embed, p=2, dim=-1)
        # grad_acc only inside forward() (a direct caller hands in its own mask_features, not the fan-out's)
        outputs_mask = PF.mask_einsum(mask_embed.float(), mask_features, grad_acc=getattr(self, "_mask_grad_acc", None))
        mask, row_any = PF.build_attention_mask(outputs_mask, attn_mask_target_size)
        return outputs_class, outputs_mask, AttnMask(mask, row_any), decoder_output

    def _set_aux_loss(self, outputs_class, outputs_seg_masks):
        if self.mask_classification:
            return [{"pred_logits": a, "pred_masks": b} for a, b in zip(outputs_class[:-1], outputs_seg_masks[:-1])]
        return [{"pred_masks": b} for b in outputs_seg_masks[:-1]]
