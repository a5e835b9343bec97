"""Driver for ncu captures of the MSDeformAttn kernels (C2 shapes by default): runs each variant a few times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import _lib, functional as fn  # noqa: E402
from tools.microbench import msda_case  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
c5 = "--c5" in sys.argv
if c5:
    value, loc, attn, shapes, fb, bb = msda_case(1, [(256, 256), (128, 128), (64, 64), (32, 32)], 4.0)
else:
    value, loc, attn, shapes, fb, bb = msda_case(2, [(32, 32), (64, 64), (128, 128)], 4.0)
lib = _lib.load()
for _ in range(3):
    if which in ("all", "f32"):
        lib.pdb_debug_set_msda_path(4)
        with torch.no_grad():
            fn.ms_deform_attn(value, shapes, None, loc, attn)
        lib.pdb_debug_set_msda_path(0)
    if which in ("all", "half"):
        fn.msda_value_half = True
        with torch.no_grad():
            fn.ms_deform_attn(value, shapes, None, loc, attn)
        fn.msda_value_half = False
    if which in ("all", "l1"):
        lib.pdb_debug_set_msda_path(1)
        with torch.no_grad():
            fn.ms_deform_attn(value, shapes, None, loc, attn)
        lib.pdb_debug_set_msda_path(0)
    if which in ("all", "bwd"):
        out = fn.ms_deform_attn(value, shapes, None, loc, attn)
        torch.autograd.grad(out, (value, loc, attn), torch.ones_like(out))
if which in ("all", "einsum"):
    B, Q, C, H, W = 2, 100, 256, 256, 256
    e = torch.randn(B, Q, C, device="cuda")
    e_lo = fn.split_lo(e)
    fs = [torch.randn(B, C, H, W, device="cuda").contiguous(memory_format=torch.channels_last) for _ in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        for f in fs:
            flush.zero_()
            fn.mask_einsum(e, f, embed_lo=e_lo)
torch.cuda.synchronize()
