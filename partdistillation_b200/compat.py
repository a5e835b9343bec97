"""Detectron2 boundary.

The drop-in contract is string-keyed registries + ``@configurable``/``from_config``
+ yacs-style ``CfgNode``.  When a real detectron2 is importable its registries, ``configurable``,
``CfgNode`` and structures are used, so the reference's train scripts resolve
``cfg.MODEL.META_ARCHITECTURE = "ProposalModel"`` etc. to the classes of this package.  Otherwise
(this image has no detectron2) the minimal equivalents below provide the same interface.
"""
import copy
import functools
from collections import namedtuple

import torch
from torch import nn
from torch.nn import functional as F


def _real_detectron2():
    try:
        import detectron2
    except Exception:
        return None
    if getattr(detectron2, "__pdb_oracle_shim__", False):     # the oracle's test-only stand-in
        return None
    return detectron2


_d2 = _real_detectron2()
HAVE_DETECTRON2 = _d2 is not None

if HAVE_DETECTRON2:   # pragma: no cover - not present in this image
    from detectron2.config import CfgNode, configurable, get_cfg
    from detectron2.layers import Conv2d, ShapeSpec, get_norm
    from detectron2.modeling import BACKBONE_REGISTRY, META_ARCH_REGISTRY, SEM_SEG_HEADS_REGISTRY, Backbone
    from detectron2.structures import BitMasks, ImageList, Instances
    from detectron2.utils.registry import Registry
else:
    class Registry:
        """name -> class; ``register`` works as a decorator or a call, keyed by ``__name__``."""

        def __init__(self, name):
            self._name = name
            self._obj_map = {}

        def register(self, obj=None):
            if obj is None:
                def deco(o):
                    self._obj_map[o.__name__] = o
                    return o
                return deco
            self._obj_map[obj.__name__] = obj
            return obj

        def get(self, name):
            if name not in self._obj_map:
                raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
            return self._obj_map[name]

        def __contains__(self, name):
            return name in self._obj_map

    class CfgNode(dict):
        """Attribute-access nested dict (the subset of yacs the hot path reads)."""

        def __init__(self, init_dict=None):
            super().__init__()
            for k, v in (init_dict or {}).items():
                self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

        def __getattr__(self, name):
            try:
                return self[name]
            except KeyError:
                raise AttributeError(name)

        def __setattr__(self, name, value):
            self[name] = value

        def clone(self):
            return copy.deepcopy(self)

    def get_cfg():
        """detectron2 defaults touched by the hot path (values as in detectron2 0.6 defaults.py)."""
        c = CfgNode()
        c.VERSION = 2
        c.INPUT = CfgNode(); c.INPUT.CROP = CfgNode(); c.INPUT.FORMAT = "RGB"
        c.SOLVER = CfgNode(); c.DATASETS = CfgNode(); c.TEST = CfgNode()
        c.DATASETS.TRAIN = ("synthetic",); c.DATASETS.TEST = ()
        c.TEST.DETECTIONS_PER_IMAGE = 100
        c.MODEL = CfgNode()
        c.MODEL.DEVICE = "cuda"
        c.MODEL.META_ARCHITECTURE = "ProposalModel"
        c.MODEL.PIXEL_MEAN = [123.675, 116.280, 103.530]
        c.MODEL.PIXEL_STD = [58.395, 57.120, 57.375]
        c.MODEL.BACKBONE = CfgNode(); c.MODEL.BACKBONE.NAME = "D2SwinTransformer"; c.MODEL.BACKBONE.FREEZE_AT = 0
        s = c.MODEL.SEM_SEG_HEAD = CfgNode()
        s.NAME = "MaskFormerHead"; s.IN_FEATURES = ["res2", "res3", "res4", "res5"]; s.IGNORE_VALUE = 255
        s.NUM_CLASSES = 1; s.CONVS_DIM = 256; s.COMMON_STRIDE = 4; s.NORM = "GN"; s.LOSS_WEIGHT = 1.0
        return c

    def configurable(init_func=None, *, from_config=None):
        assert init_func is not None and init_func.__name__ == "__init__"

        @functools.wraps(init_func)
        def wrapped(self, *args, **kwargs):
            first = args[0] if args else kwargs.get("cfg")
            if isinstance(first, CfgNode):
                init_func(self, **type(self).from_config(*args, **kwargs))
            else:
                init_func(self, *args, **kwargs)
        return wrapped

    class ShapeSpec(namedtuple("_ShapeSpec", ["channels", "height", "width", "stride"])):
        def __new__(cls, channels=None, height=None, width=None, stride=None):
            return super().__new__(cls, channels, height, width, stride)

    class Conv2d(nn.Conv2d):
        """nn.Conv2d with an optional ``norm`` sub-module (state-dict key ``<name>.norm.*``) and activation."""

        def __init__(self, *args, **kwargs):
            norm = kwargs.pop("norm", None)
            activation = kwargs.pop("activation", None)
            super().__init__(*args, **kwargs)
            self.norm = norm
            self.activation = activation

        def forward(self, x):
            x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
            if self.norm is not None:
                x = self.norm(x)
            if self.activation is not None:
                x = self.activation(x)
            return x

    def get_norm(norm, out_channels):
        if norm is None or (isinstance(norm, str) and len(norm) == 0):
            return None
        if isinstance(norm, str):
            if norm != "GN":
                raise ValueError(f"norm {norm!r} not supported without detectron2")
            return nn.GroupNorm(32, out_channels)
        return norm(out_channels)

    class Backbone(nn.Module):
        @property
        def size_divisibility(self):
            return 0

    class ImageList:
        def __init__(self, tensor, image_sizes):
            self.tensor = tensor
            self.image_sizes = image_sizes

        def __len__(self):
            return len(self.image_sizes)

        @staticmethod
        def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
            sizes = [(t.shape[-2], t.shape[-1]) for t in tensors]
            mh = max(s[0] for s in sizes); mw = max(s[1] for s in sizes)
            if size_divisibility > 1:
                d = size_divisibility
                mh = (mh + d - 1) // d * d; mw = (mw + d - 1) // d * d
            if all(s == (mh, mw) for s in sizes):
                return ImageList(torch.stack(tensors), sizes)
            out = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (mh, mw), pad_value)
            for i, t in enumerate(tensors):
                out[i, ..., : t.shape[-2], : t.shape[-1]].copy_(t)
            return ImageList(out, sizes)

    class BitMasks:
        def __init__(self, tensor):
            self.tensor = torch.as_tensor(tensor).to(torch.bool)

        def to(self, *a, **k):
            return BitMasks(self.tensor.to(*a, **k))

        def __len__(self):
            return self.tensor.shape[0]

    class Instances:
        def __init__(self, image_size, **kwargs):
            object.__setattr__(self, "_image_size", image_size)
            object.__setattr__(self, "_fields", {})
            for k, v in kwargs.items():
                self.set(k, v)

        @property
        def image_size(self):
            return self._image_size

        def __setattr__(self, name, val):
            if name.startswith("_"):
                object.__setattr__(self, name, val)
            else:
                self.set(name, val)

        def __getattr__(self, name):
            if name == "_fields" or name not in self._fields:
                raise AttributeError(name)
            return self._fields[name]

        def set(self, name, value):
            self._fields[name] = value

        def has(self, name):
            return name in self._fields

        def get(self, name):
            return self._fields[name]

        def to(self, *a, **k):
            ret = Instances(self._image_size)
            for n, v in self._fields.items():
                ret.set(n, v.to(*a, **k) if hasattr(v, "to") else v)
            return ret

        def __len__(self):
            for v in self._fields.values():
                return len(v)
            return 0

    META_ARCH_REGISTRY = Registry("META_ARCH")
    SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")
    BACKBONE_REGISTRY = Registry("BACKBONE")

class PackedBitMasks:
    """Instance masks at one bit per pixel: ``tensor`` holds int32 words (K, H, ceil(W / 32)), bit i of
    a word = pixel 32 * word + i of that row — the layout of ``functional.pack_bits`` / ``unpack_bits``.  A dataset mapper
    that decodes the RLE pseudo labels (proposal_dataset_mapper.py:113-139,201-235) can emit this instead of the
    ``BitMasks`` bool tensor: the host->device copy of the targets shrinks 8x and the meta-architectures expand the words
    on the device straight into their padded uint8 target buffer (``meta_base._prepare_pseudo_targets``).  Quacks like
    ``BitMasks`` where the training step needs it (``tensor``, ``to``, ``len``)."""

    def __init__(self, tensor, width=None):
        tensor = torch.as_tensor(tensor)
        if tensor.dtype != torch.int32 or tensor.dim() != 3:
            raise ValueError("PackedBitMasks: int32 words of shape (K, H, ceil(W / 32))")
        self.tensor = tensor
        self.width = int(width) if width is not None else 32 * tensor.shape[-1]
        if (self.width + 31) // 32 != tensor.shape[-1]:
            raise ValueError(f"PackedBitMasks: width {self.width} does not match {tensor.shape[-1]} words per row")

    @property
    def image_size(self):
        return (int(self.tensor.shape[1]), self.width)

    @classmethod
    def from_bool(cls, masks):
        """Host-side packing of a bool (K, H, W) tensor (what a mapper would do once per sample)."""
        masks = torch.as_tensor(masks).to(torch.bool)
        K, H, W = masks.shape
        Ww = (W + 31) // 32
        padded = torch.zeros((K, H, Ww * 32), dtype=torch.int64)
        padded[..., :W] = masks
        words = (padded.view(K, H, Ww, 32) << torch.arange(32)).sum(-1)
        words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
        return cls(words.to(torch.int32), W)

    def like(self, tensor):
        return PackedBitMasks(tensor, self.width)

    def to(self, *a, **k):
        return PackedBitMasks(self.tensor.to(*a, **k), self.width)

    def __len__(self):
        return self.tensor.shape[0]


# Mask2Former's own registry (maskformer_transformer_decoder.py:19) lives in the reference package,
# not in detectron2, so it is always ours.
TRANSFORMER_DECODER_REGISTRY = Registry("TRANSFORMER_MODULE")


def build_backbone(cfg, input_shape=None):
    if input_shape is None:
        input_shape = ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN))
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)


def build_sem_seg_head(cfg, input_shape):
    return SEM_SEG_HEADS_REGISTRY.get(cfg.MODEL.SEM_SEG_HEAD.NAME)(cfg, input_shape)


def build_pixel_decoder(cfg, input_shape):
    """fpn.py:25-37."""
    name = cfg.MODEL.SEM_SEG_HEAD.PIXEL_DECODER_NAME
    model = SEM_SEG_HEADS_REGISTRY.get(name)(cfg, input_shape)
    if not callable(getattr(model, "forward_features", None)):
        raise ValueError(f"pixel decoder {name} must implement forward_features(features)")
    return model


def build_transformer_decoder(cfg, in_channels, mask_classification=True):
    """maskformer_transformer_decoder.py:25-30."""
    name = cfg.MODEL.MASK_FORMER.TRANSFORMER_DECODER_NAME
    return TRANSFORMER_DECODER_REGISTRY.get(name)(cfg, in_channels, mask_classification)


def build_model(cfg):
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    return model.to(torch.device(cfg.MODEL.DEVICE))


def c2_xavier_fill(module):
    """fvcore.nn.weight_init.c2_xavier_fill: kaiming_uniform(a=1) weight, zero bias."""
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)
