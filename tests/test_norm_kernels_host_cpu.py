"""CPU: the LayerNorm (+ residual) and GroupNorm (+ ReLU) kernels' own code (csrc/layernorm.cu, csrc/groupnorm.cu) compiled
for the host and driven by the GPU parity tests of tests/test_ops_gpu.py themselves (float64 torch references, same
tolerances).  ``Tensor.is_cuda`` reads True inside these tests so that functional.layer_norm / group_norm take the kernel
path they take on the B200."""
import re

import pytest
import torch

import test_ops_gpu as gpu_tests
from host_kernels import build_host_library, patch_functional


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("norm_host")
    from host_kernels import NAMESPACE_BLOCK, ROOT
    import os
    src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", "layernorm.cu")).read()
    (tmp / "layernorm_section.inc").write_text(re.search(NAMESPACE_BLOCK, src, re.S).group(1))
    return build_host_library(tmp, "groupnorm.cu", "groupnorm_section.inc", "norm_kernels_host.cpp",
                              ("layer_norm_forward", "layer_norm_forward_scaled", "group_norm_forward", "group_norm_backward"))


@pytest.fixture
def fn(monkeypatch, host_lib):
    f = patch_functional(monkeypatch, host_lib)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    return f


@pytest.mark.parametrize("rows,C,res,want_sum", [((2, 30), 256, True, False), ((5, 7), 128, False, False),
                                                  ((3, 5), 512, True, True), ((9,), 2048, True, True), ((4, 11), 36, False, True)])
def test_layer_norm_fused(fn, rows, C, res, want_sum):
    gpu_tests.test_layer_norm_fused(fn, rows, C, res, want_sum)


@pytest.mark.parametrize("B,L,C", [(2, 30, 128), (3, 5, 512)])
def test_layer_norm_with_stochastic_depth_scale(fn, B, L, C):
    gpu_tests.test_layer_norm_with_stochastic_depth_scale(fn, B, L, C)


@pytest.mark.parametrize("B,C,H,W,G,relu", [(2, 256, 8, 8, 32, True), (1, 128, 9, 7, 32, True), (3, 64, 6, 6, 8, False),
                                            (2, 256, 7, 5, 32, False), (1, 512, 5, 5, 32, True)])
def test_group_norm_channels_last(fn, B, C, H, W, G, relu):
    gpu_tests.test_group_norm_channels_last(fn, B, C, H, W, G, relu)
