"""CPU: the whole hot path — pixel decoder, masked-attention decoder, criterion; forward and backward — of the PRODUCT modules
on CPU tensors, against the golden tensors of the unmodified reference: the body of tests/test_head_gpu.py::
test_head_and_loss_vs_reference_golden (mask / class logits, attention-mask bits, Hungarian indices, losses, gradients).

Every SIMT operator runs its own kernel source built for the host (tests/host_kernels.py: MSDeformAttn gather / scatter,
masked cross-attention, attention-mask build, point sampling, matcher cost, LSAP, point loss, classifier rows, Layer / Group
norms); the tcgen05 GEMM interface alone is a restatement (tests/host_gemm_abi.py).  This checks the host orchestration and
the kernels together where there is no GPU; it is not a product path (the product raises without CUDA)."""
import pytest
import torch

import test_head_gpu as head_tests
from host_kernels import full_host_library, patch_functional


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return full_host_library(tmp_path_factory)


@pytest.fixture
def on_host(monkeypatch, host_lib):
    patch_functional(monkeypatch, host_lib)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))       # the wrappers take their kernel paths
    monkeypatch.setattr(head_tests, "DEV", "cpu")


@pytest.mark.parametrize("name", ["proposal_micro", "pd_micro"])
def test_head_and_loss_vs_reference_golden(on_host, golden_dir, name):
    head_tests.test_head_and_loss_vs_reference_golden(golden_dir, name)
