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
torch.cuda.synchronize()
