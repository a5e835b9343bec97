"""CPU: the attention-mask kernel's own code (csrc/attn_mask.cu: bilinear resize of the mask logits, sigmoid() < 0.5, the
per-row "attends somewhere" flag and the all-masked-row reset of mask2former_transformer_decoder.py:405,453-457) compiled
for the host through tests/native/cuda_on_cpu.h and compared BIT FOR BIT with the attention masks the unmodified reference
recorded for its own mask logits (tests/golden/head_*.pt: pred_masks / attn_mask_bits of every decoder layer).  The GPU
twins are tests/test_ops_gpu.py::test_attn_mask_bits and tests/test_head_gpu.py; the bit-exact contract is BASELINE.json's."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def attn_mask(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("attn_mask_host")
    src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", "attn_mask.cu")).read()
    m = re.search(r"(namespace pdb \{.*?\}  // namespace pdb\n)", src, re.S)
    assert m and "attn_mask_kernel" in m.group(1) and "<<<" not in m.group(1)
    (tmp / "attn_mask_section.inc").write_text(m.group(1))
    so = str(tmp / "libattn_mask_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-I", str(tmp),
                           os.path.join(HERE, "native", "attn_mask_kernel_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.host_attn_mask_build.argtypes = [P, P, P, I, I, I, I, I, I]
    lib.host_attn_mask_reset_rows.argtypes = [P, P, I, ctypes.c_int64]

    def build(logits, size):
        B, Q, H, W = logits.shape
        h, w = size
        x = logits.contiguous().float()
        mask = torch.empty(B, Q, h * w, dtype=torch.uint8)
        row_any = torch.zeros(B * Q, dtype=torch.int32)
        lib.host_attn_mask_build(x.data_ptr(), mask.data_ptr(), row_any.data_ptr(), B, Q, H, W, h, w)
        return mask, row_any

    def reset(mask, row_any):
        B, Q, hw = mask.shape
        lib.host_attn_mask_reset_rows(mask.data_ptr(), row_any.data_ptr(), B * Q, hw)
        return mask
    return build, reset


@pytest.mark.parametrize("name", ["proposal_micro", "proposal_micro_uniform", "pd_micro"])
def test_attn_mask_kernel_reproduces_reference_bits(attn_mask, golden_dir, name):
    build, reset = attn_mask
    g = torch.load(os.path.join(golden_dir, f"head_{name}.pt"), weights_only=False)
    H, W = g["case"]["H"], g["case"]["W"]
    sizes = [(H // 32, W // 32), (H // 16, W // 16), (H // 8, W // 8)]
    assert len(g["pred_masks"]) == len(g["attn_mask_bits"]) == g["case"]["dec_layers"]      # DEC_LAYERS - 1 layers + the initial heads
    checked = 0
    for i, (logits, packed, shape) in enumerate(zip(g["pred_masks"], g["attn_mask_bits"], g["attn_mask_shapes"])):
        h, w = sizes[i % 3]
        assert shape[-1] == h * w
        ref = torch.from_numpy(np.unpackbits(packed, axis=-1)[..., :h * w]).bool()          # (B, Q, hw): head 0 of each image
        mask, row_any = build(logits, (h, w))
        assert torch.equal(mask.bool(), ref), (name, i)
        assert torch.equal(row_any.view(ref.shape[:2]) != 0, ~ref.all(-1))
        # next layer's `attn_mask[rows that are all True] = False` (:405)
        exp = ref.clone()
        exp[exp.all(-1)] = False
        assert torch.equal(reset(mask, row_any).bool(), exp)
        checked += ref.numel()
    assert checked > 5000


def test_attn_mask_kernel_threshold_dead_zone(attn_mask):
    """sigmoid(x) < 0.5 is NOT x < 0 in fp32: false for -1.79e-7 < x < 0, false at exact zeros; rows that attend nowhere
    are flagged for the reset.  Same checks as the GPU test, against torch's CPU expression."""
    import torch.nn.functional as F
    build, reset = attn_mask
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 9, 32, 32, generator=g)
    x[0, 3] = -5.0
    x[1, 2, :4] = 0.0
    x[1, 4] = x[1, 4] * 1e-7
    for size in ((16, 16), (8, 8), (4, 4)):
        mask, row_any = build(x, size)
        ref = F.interpolate(x, size=size, mode="bilinear", align_corners=False).sigmoid().flatten(2) < 0.5
        assert torch.equal(mask.bool(), ref)
        assert torch.equal(row_any.view(2, 9) != 0, ~ref.all(-1))
        assert not reset(mask, row_any)[0, 3].any()
