"""CPU: the loss-side kernels' own code (csrc/loss.cu: point sampling, matcher cost, batched Hungarian assignment, fused
point BCE + dice, PartDistillation's float64 classifier rows) compiled for the host through tests/native/cuda_on_cpu.h and
driven through partdistillation_b200/functional.py's UNMODIFIED wrappers — by running the GPU parity tests of
tests/test_ops_gpu.py themselves with ``Tensor.cuda()`` turned into the identity and the library handle swapped for the
host build.  Same inputs, same oracle comparisons, same tolerances as on the B200."""
import pytest

import test_ops_gpu as gpu_tests
from host_kernels import build_host_library, patch_functional

OPS = ("point_sample_forward", "point_sample_backward", "matcher_cost", "lsap_batched", "point_loss_forward",
       "point_loss_backward", "class_rows_forward", "class_rows_backward")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return build_host_library(tmp_path_factory.mktemp("loss_host"), "loss.cu", "loss_section.inc", "loss_kernels_host.cpp", OPS)


@pytest.fixture
def fn(monkeypatch, host_lib):
    return patch_functional(monkeypatch, host_lib)


def test_point_sample(fn):
    gpu_tests.test_point_sample(fn)


def test_matcher_cost_and_lsap_vs_oracle(fn):
    gpu_tests.test_matcher_cost_and_lsap_vs_oracle(fn)


def test_lsap_exact_vs_scipy(fn):
    gpu_tests.test_lsap_exact_vs_scipy(fn)


def test_point_loss_vs_oracle(fn):
    gpu_tests.test_point_loss_vs_oracle(fn)


def test_class_rows(fn):
    gpu_tests.test_class_rows(fn)
