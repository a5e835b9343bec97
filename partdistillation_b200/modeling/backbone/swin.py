"""Swin Transformer backbone, registered as ``D2SwinTransformer`` (reference: modeling/backbone/swin.py).

Not a kernel target of this round (frozen in the shipped recipes): plain PyTorch with
the reference's parameter names so checkpoints load unchanged; window attention goes through
``F.scaled_dot_product_attention`` with the relative-position bias and shift mask as an additive bias.
"""
import torch
import torch.nn.functional as F
from torch import nn

from ... import functional as PF
from ...compat import BACKBONE_REGISTRY, Backbone, ShapeSpec


def _autocast_half():
    """torch.bfloat16 under torch.autocast("cuda", bfloat16), else None."""
    if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        return torch.bfloat16
    return None


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm semantics)."""

    def __init__(self, p=0.0):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        return x * self.sample_scale(x)

    def sample_scale(self, x):
        """The per-sample keep / (1 - p) factors (one bernoulli draw per sample, as timm's drop_path)."""
        keep = 1.0 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return mask.div_(keep)

    def add_to(self, residual, x):
        """residual + drop_path(x) in one pass (addcmul) instead of a scale pass and an add pass."""
        if self.p == 0.0 or not self.training:
            return residual + x
        return torch.addcmul(residual, x, self.sample_scale(x))


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std, a=-2.0, b=2.0)


class Mlp(nn.Module):
    def __init__(self, dim, hidden, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop = nn.Dropout(drop)

    def forward(self, x, out_fp32=False):
        """out_fp32: under bf16 autocast the second Linear writes fp32 (the residual stream's dtype) from its accumulator."""
        if isinstance(self.act, nn.GELU) and getattr(self.act, "approximate", "none") == "none":
            h = PF.linear(x, self.fc1.weight, self.fc1.bias, gelu=True)       # GELU in the GEMM epilogue when frozen
        else:
            h = self.act(PF.linear(x, self.fc1.weight, self.fc1.bias))
        return self.drop(PF.linear(self.drop(h), self.fc2.weight, self.fc2.bias, out_fp32=out_fp32))


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def window_reverse(win, ws, H, W):
    B = win.shape[0] // ((H // ws) * (W // ws))
    x = win.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        ws = window_size[0]
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * ws - 1) * (2 * ws - 1), num_heads))
        coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0) + (ws - 1)
        self.register_buffer("relative_position_index", rel[:, :, 0] * (2 * ws - 1) + rel[:, :, 1])
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        _trunc_normal_(self.relative_position_bias_table)

    def forward(self, x, mask=None):
        """x (nW*B, N, C); mask (nW, N, N) additive (0 / -100) or None."""
        Bw, N, C = x.shape
        qkv = PF.linear(x, self.qkv.weight, self.qkv.bias)
        bias = self.relative_position_bias_table[self.relative_position_index.view(-1)].view(N, N, -1).permute(2, 0, 1)
        if (x.is_cuda and qkv.dtype == torch.float32 and C // self.num_heads == 32 and N <= 256
                and not (torch.is_grad_enabled() and (qkv.requires_grad or bias.requires_grad))
                and not (self.training and self.attn_drop.p > 0)):
            # frozen backbone: fused window attention kernel (scores, bias, shift mask, softmax, PV, head transpose)
            o = PF.window_attention(qkv, bias, mask, self.num_heads, self.scale)
            return self.proj_drop(PF.linear(o, self.proj.weight, self.proj.bias))
        qkv = qkv.reshape(Bw, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        bias = bias.unsqueeze(0)                                                # (1, heads, N, N)
        if mask is not None:
            nW = mask.shape[0]
            bias = (bias + mask[:, None]).repeat(Bw // nW, 1, 1, 1)              # (Bw, heads, N, N)
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], attn_mask=bias.to(qkv.dtype),
                                           dropout_p=self.attn_drop.p if self.training else 0.0, scale=self.scale)
        return self.proj_drop(PF.linear(o.transpose(1, 2).reshape(Bw, N, C), self.proj.weight, self.proj.bias))


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size=7, shift_size=0, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop=0.0, attn_drop=0.0, drop_path=0.0):
        super().__init__()
        self.dim, self.window_size, self.shift_size = dim, window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, (window_size, window_size), num_heads, qkv_bias, qk_scale, attn_drop, drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), drop)

    def _fused_attention(self, x, H, W, normed=None):
        """Frozen-backbone path: qkv Linear on the token grid, then ONE kernel for pad / roll / window partition /
        shift mask / attention / window reverse / roll back / crop (functional.swin_window_attention).  ``normed``: norm1(x)
        when the caller already has it."""
        a = self.attn
        B, L, C = x.shape
        half = _autocast_half()        # bf16 autocast: every hand-over between kernels already has the consumer's dtype
        if normed is None:
            normed = PF.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, out_dtype=half)
        qkv = PF.linear(normed, a.qkv.weight, a.qkv.bias, out_fp32=True)
        bias = self._relative_bias()
        # the attention core (scores, softmax, PV) reads fp32 q / k / v (written by the qkv GEMM's epilogue) either way
        o = PF.swin_window_attention(qkv.float().view(B, H, W, 3 * C), a.qkv.bias, bias, a.num_heads, self.window_size,
                                     self.shift_size, a.scale, out_dtype=half or torch.float32)
        return PF.linear(o.view(B, L, C), a.proj.weight, a.proj.bias, out_fp32=True)

    def _relative_bias(self):
        """(heads, N, N) relative position bias of this block (swin_transformer.py:91-96: a gather from the (2ws-1)^2 table),
        contiguous fp32.  A frozen table is gathered once per table version instead of once per forward."""
        a = self.attn
        t = a.relative_position_bias_table
        N = self.window_size * self.window_size
        if torch.is_grad_enabled() and t.requires_grad:
            return t[a.relative_position_index.view(-1)].view(N, N, -1).permute(2, 0, 1)
        key = (t._version, t.data_ptr(), t.device, t.dtype)
        hit = self.__dict__.get("_rel_bias")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                b = t[a.relative_position_index.view(-1)].view(N, N, -1).permute(2, 0, 1).float().contiguous()
            hit = self.__dict__["_rel_bias"] = (key, b)
        return hit[1]

    def _can_fuse(self, x):
        a = self.attn
        frozen = not (torch.is_grad_enabled() and (x.requires_grad or a.qkv.weight.requires_grad
                                                   or a.relative_position_bias_table.requires_grad))
        return (x.is_cuda and x.dtype == torch.float32 and frozen and self.dim // a.num_heads == 32
                and self.window_size * self.window_size <= 256 and a.attn_drop.p == 0 and a.proj_drop.p == 0)

    def _branch_scale(self, x):
        dp = self.drop_path
        return dp.sample_scale(x) if isinstance(dp, DropPath) and dp.p > 0.0 and dp.training else None

    def forward_fused(self, x, H, W, pending=None):
        """The block on the frozen-backbone path with every residual add (and its stochastic-depth factor) folded into the
        LayerNorm behind it.  ``pending`` = (branch, scale) of the previous block's MLP, still to be added to ``x``; returns
        (x, pending) in the same form — BasicLayer adds the last one.  Draw order of the stochastic-depth masks = the
        reference's (attention branch, then MLP branch)."""
        half = _autocast_half()
        if pending is not None:
            h, x = PF.layer_norm(x, self.norm1.weight, self.norm1.bias, self.norm1.eps, residual=pending[0],
                                 residual_scale=pending[1], return_sum=True, out_dtype=half)
        else:
            h = None
        attn = self._fused_attention(x, H, W, normed=h)
        h, x = PF.layer_norm(x, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=attn,
                             residual_scale=self._branch_scale(x), return_sum=True, out_dtype=half)
        return x, (self.mlp(h, out_fp32=True), self._branch_scale(x))

    def forward(self, x, H, W, mask_matrix):
        B, L, C = x.shape
        ws = self.window_size
        if self._can_fuse(x):
            x, (m, scale) = self.forward_fused(x, H, W)
            return x + m if scale is None else torch.addcmul(x, m, scale)
        h = self.norm1(x).view(B, H, W, C)
        pr, pb = (ws - W % ws) % ws, (ws - H % ws) % ws
        h = F.pad(h, (0, 0, 0, pr, 0, pb))
        Hp, Wp = H + pb, W + pr
        mask = None
        if self.shift_size > 0:
            h = torch.roll(h, shifts=(-self.shift_size, -self.shift_size), dims=(1, 2))
            mask = mask_matrix
        win = self.attn(window_partition(h, ws), mask)
        h = window_reverse(win, ws, Hp, Wp)
        if self.shift_size > 0:
            h = torch.roll(h, shifts=(self.shift_size, self.shift_size), dims=(1, 2))
        h = h[:, :H, :W, :].reshape(B, H * W, C)
        x = self._add_drop_path(x, h)
        return self._add_drop_path(x, self.mlp(self.norm2(x)))

    def _add_drop_path(self, residual, x):
        dp = self.drop_path
        return dp.add_to(residual, x) if isinstance(dp, DropPath) else residual + dp(x)


def _add_branch(x, pending):
    m, scale = pending
    return x + m if scale is None else torch.addcmul(x, m, scale)


class PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)

    def forward(self, x, H, W):
        B, L, C = x.shape
        x = x.view(B, H, W, C)
        if H % 2 or W % 2:
            x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
        return PF.linear(PF.layer_norm(x.view(B, -1, 4 * C), self.norm.weight, self.norm.bias, self.norm.eps), self.reduction.weight, None)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop, drop_path,
                 downsample):
        super().__init__()
        self.window_size, self.shift_size, self.depth = window_size, window_size // 2, depth
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio,
                                 qkv_bias, qk_scale, drop, attn_drop, drop_path[i]) for i in range(depth)])
        self.downsample = PatchMerging(dim) if downsample else None
        self._mask_cache = {}

    def _shift_mask(self, H, W, device):
        ws, sh = self.window_size, self.shift_size
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        key = (Hp, Wp, device)
        m = self._mask_cache.get(key)
        if m is None:
            img = torch.zeros((1, Hp, Wp, 1), device=device)
            cnt = 0
            for hs in (slice(0, -ws), slice(-ws, -sh), slice(-sh, None)):
                for wsl in (slice(0, -ws), slice(-ws, -sh), slice(-sh, None)):
                    img[:, hs, wsl, :] = cnt
                    cnt += 1
            mw = window_partition(img, ws).squeeze(-1)
            diff = mw.unsqueeze(1) - mw.unsqueeze(2)
            m = self._mask_cache[key] = diff.masked_fill(diff != 0, -100.0).masked_fill(diff == 0, 0.0)
        return m

    def forward(self, x, H, W):
        mask = self._shift_mask(H, W, x.device)
        pending = None
        for blk in self.blocks:
            if blk._can_fuse(x):
                x, pending = blk.forward_fused(x, H, W, pending)
                continue
            if pending is not None:
                x, pending = _add_branch(x, pending), None
            x = blk(x, H, W, mask)
        if pending is not None:
            x = _add_branch(x, pending)
        if self.downsample is not None:
            down = self.downsample(x, H, W)
            if down.dtype != x.dtype and x.dtype == torch.float32:
                # bf16 autocast: the reduction Linear returns bf16; the residual stream of the next stage stays fp32 (the fused
                # LayerNorm / window-attention kernels read fp32; the more precise side of the autocast tolerance)
                down = down.float()
            return x, H, W, down, (H + 1) // 2, (W + 1) // 2
        return x, H, W, x, H, W


class PatchEmbed(nn.Module):
    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm=True):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim) if norm else None

    def forward(self, x):
        p = self.patch_size
        _, _, H, W = x.shape
        if W % p or H % p:
            x = F.pad(x, (0, (p - W % p) % p, 0, (p - H % p) % p))
        x = self.proj(x)
        if self.norm is not None:
            Wh, Ww = x.shape[2:]
            x = self.norm(x.flatten(2).transpose(1, 2)).transpose(1, 2).reshape(x.shape[0], -1, Wh, Ww)
        return x

    def forward_tokens(self, x):
        """Same result as ``forward`` in token order, (B, Wh * Ww, C) + (Wh, Ww): the stride-p p x p convolution is a GEMM over
        the non-overlapping patches (one im2col copy of the image, then the tensor-core Linear + the library's LayerNorm), so the
        NCHW <-> token transposes around the reference's conv + norm (swin.py:418-447) disappear."""
        p = self.patch_size
        _, _, H, W = x.shape
        if W % p or H % p:
            x = F.pad(x, (0, (p - W % p) % p, 0, (p - H % p) % p))
        B, C, H, W = x.shape
        Wh, Ww = H // p, W // p
        patches = x.view(B, C, Wh, p, Ww, p).permute(0, 2, 4, 1, 3, 5).reshape(B, Wh * Ww, C * p * p)
        t = PF.linear(patches, self.proj.weight.view(self.proj.weight.shape[0], -1), self.proj.bias)
        if self.norm is not None:
            t = PF.layer_norm(t, self.norm.weight, self.norm.bias, self.norm.eps)
        return t, Wh, Ww


class SwinTransformer(nn.Module):
    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.2, ape=False, patch_norm=True, out_indices=(0, 1, 2, 3),
                 frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if ape:
            raise NotImplementedError("absolute position embedding is not used by any shipped config")
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.out_indices = out_indices
        self.patch_embed = PatchEmbed(patch_size, in_chans, embed_dim, patch_norm)
        self.pos_drop = nn.Dropout(drop_rate)
        dpr = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        self.layers = nn.ModuleList()
        for i in range(self.num_layers):
            self.layers.append(BasicLayer(int(embed_dim * 2 ** i), depths[i], num_heads[i], window_size, mlp_ratio,
                                          qkv_bias, qk_scale, drop_rate, attn_drop_rate,
                                          dpr[sum(depths[:i]):sum(depths[:i + 1])], i < self.num_layers - 1))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        for i in out_indices:
            self.add_module(f"norm{i}", nn.LayerNorm(self.num_features[i]))
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x):
        x, Wh, Ww = self.patch_embed.forward_tokens(x)
        x = self.pos_drop(x)
        outs = {}
        for i, layer in enumerate(self.layers):
            x_out, H, W, x, Wh, Ww = layer(x, Wh, Ww)
            if i in self.out_indices:
                n = getattr(self, f"norm{i}")
                o = PF.layer_norm(x_out, n.weight, n.bias, n.eps)
                outs[f"res{i + 2}"] = o.view(-1, H, W, self.num_features[i]).permute(0, 3, 1, 2).contiguous()
        return outs


@BACKBONE_REGISTRY.register()
class D2SwinTransformer(SwinTransformer, Backbone):
    def __init__(self, cfg, input_shape):
        s = cfg.MODEL.SWIN
        super().__init__(s.PRETRAIN_IMG_SIZE, s.PATCH_SIZE, 3, s.EMBED_DIM, s.DEPTHS, s.NUM_HEADS, s.WINDOW_SIZE,
                         s.MLP_RATIO, s.QKV_BIAS, s.QK_SCALE, s.DROP_RATE, s.ATTN_DROP_RATE, s.DROP_PATH_RATE, s.APE,
                         s.PATCH_NORM, use_checkpoint=s.USE_CHECKPOINT)
        self._out_features = s.OUT_FEATURES
        self._out_feature_strides = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}
        self._out_feature_channels = {f"res{i + 2}": self.num_features[i] for i in range(4)}

    def forward(self, x):
        assert x.dim() == 4, f"SwinTransformer takes an input of shape (N, C, H, W). Got {x.shape} instead!"
        y = super().forward(x)
        return {k: v for k, v in y.items() if k in self._out_features}

    def output_shape(self):
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n])
                for n in self._out_features}

    @property
    def size_divisibility(self):
        return 32
