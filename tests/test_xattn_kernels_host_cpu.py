"""CPU: the masked cross-attention kernels' own code (csrc/xattn.cu: split-K forward + combine, dQ and dK/dV backward;
mask2former_transformer_decoder.py:84,102-114 with the all-masked-row reset of :405) compiled for the host and driven by
the GPU parity test of tests/test_ops_gpu.py itself (fp64 reference attention, same tolerances)."""
import pytest

import test_ops_gpu as gpu_tests
from host_kernels import build_host_library, patch_functional


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return build_host_library(tmp_path_factory.mktemp("xattn_host"), "xattn.cu", "xattn_section.inc", "xattn_kernels_host.cpp",
                              ("masked_xattn_workspace_bytes", "masked_xattn_forward", "masked_xattn_backward"))


@pytest.fixture
def fn(monkeypatch, host_lib):
    return patch_functional(monkeypatch, host_lib)


@pytest.mark.parametrize("B,Q,Lk,masked", [(1, 37, 200, True), (2, 130, 77, True), (1, 20, 300, False),
                                           (1, 5, 2, False), (1, 129, 65, True), (1, 3, 8, True)])
def test_masked_cross_attention(fn, B, Q, Lk, masked):
    gpu_tests.test_masked_cross_attention(fn, B, Q, Lk, masked)
