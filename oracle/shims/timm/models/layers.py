# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

import torch
from torch import nn
def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)
def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)
class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob
    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask
