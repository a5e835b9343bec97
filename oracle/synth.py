"""TEST INFRASTRUCTURE ONLY — deterministic synthetic weights / inputs shared by the golden
generator (oracle/make_golden.py) and the parity tests.

Weights are a pure function of (parameter name, shape, seed) so a golden fixture only has to
store the ``{name: (shape, dtype)}`` table, not tens of MB of tensors.
"""
import zlib

import torch


def synth_tensor(name, shape, dtype, seed=0):
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if dtype in (torch.int64, torch.int32, torch.bool):
        raise ValueError(f"integer buffers must be copied, not synthesised: {name}")
    if leaf == "empty_weight":
        raise ValueError(name)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    is_norm = (".norm" in name or "decoder_norm" in name or name.endswith(".1.weight") or name.endswith(".1.bias"))
    if "sampling_offsets.bias" in name:
        t = 1.5 * r                                   # learned offsets of a few pixels
    elif "sampling_offsets.weight" in name:
        t = 0.02 * r
    elif "attention_weights" in name:
        t = 0.05 * r
    elif "relative_position_bias_table" in name:
        t = 0.02 * r
    elif leaf == "bias":
        t = 0.02 * r
    elif is_norm and leaf == "weight" and len(shape) == 1:
        t = 1.0 + 0.1 * r
    elif "level_embed" in name or "query_feat" in name or "query_embed" in name:
        t = r
    elif len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = r * (1.0 / fan_in) ** 0.5
    else:
        t = 0.1 * r
    return t.to(dtype)


def synth_state_dict(table, seed=0):
    """table: {name: (shape, dtype_str)}"""
    return {k: synth_tensor(k, shp, getattr(torch, dt), seed) for k, (shp, dt) in table.items()}


def table_of(state_dict, skip=("relative_position_index", "empty_weight")):
    return {k: (tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in state_dict.items()
            if not any(s in k for s in skip)}


def synth_features(B, H, W, channels, seed=0):
    """Backbone outputs res2..res5 for an HxW image (strides 4..32)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return {f"res{i + 2}": torch.randn(B, c, H // (4 << i), W // (4 << i), generator=g)
            for i, c in enumerate(channels)}


def synth_masks(H, W, K, seed, block=16):
    """Block label map -> K disjoint bool masks; empty masks dropped."""
    g = torch.Generator().manual_seed(2000 + seed)
    lab = torch.randint(0, K, (H // block, W // block), generator=g)
    lab = lab.repeat_interleave(block, 0).repeat_interleave(block, 1)
    m = torch.stack([lab == k for k in range(K)])
    return m[m.flatten(1).any(1)]


def synth_batch(B, H, W, K, part_distillation=False, num_object_classes=50, seed=0):
    out = []
    for i in range(B):
        g = torch.Generator().manual_seed(3000 + seed * 131 + i)
        d = {"image": torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8),
             "gt_masks": synth_masks(H, W, K, seed * 131 + i), "height": H, "width": W}
        n = d["gt_masks"].shape[0]
        d["gt_classes"] = (torch.arange(n) % 8) if part_distillation else torch.zeros(n, dtype=torch.long)
        if part_distillation:
            d["gt_object_class"] = int(torch.randint(0, num_object_classes, (1,), generator=g))
        out.append(d)
    return out


class RecordRand:
    """Monkey-patch for torch.rand that records every draw (reference side)."""

    def __init__(self):
        self.draws = []
        self._orig = torch.rand

    def __call__(self, *size, **kw):
        kw.pop("device", None)
        t = self._orig(*size, **kw)
        self.draws.append(t.clone())
        return t

    def __enter__(self):
        torch.rand = self
        return self

    def __exit__(self, *a):
        torch.rand = self._orig


class ReplayRand:
    """rand-provider that replays recorded draws in order (oracle / product side)."""

    def __init__(self, draws, device=None):
        self.draws = list(draws)
        self.i = 0
        self.device = device

    def __call__(self, *size, **kw):
        t = self.draws[self.i]
        self.i += 1
        size = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list)) else tuple(size)
        assert tuple(t.shape) == size, (tuple(t.shape), size, self.i)
        dev = kw.get("device", self.device)
        return t.to(dev) if dev is not None else t.clone()
