# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

import torch
from torch.nn import functional as F

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
        out = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (mh, mw), pad_value)
        for i, t in enumerate(tensors):
            out[i, ..., : t.shape[-2], : t.shape[-1]].copy_(t)
        return ImageList(out.contiguous(), sizes)

class BitMasks:
    def __init__(self, tensor):
        self.tensor = torch.as_tensor(tensor).to(torch.bool)
    def to(self, *a, **k):
        return BitMasks(self.tensor.to(*a, **k))
    def __len__(self):
        return self.tensor.shape[0]

class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor

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
