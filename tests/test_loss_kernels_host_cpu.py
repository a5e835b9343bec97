"""CPU: the loss-side kernels' own code (csrc/loss.cu: point sampling, matcher cost, batched Hungarian assignment, fused
point BCE + dice, PartDistillation's float64 classifier rows) compiled for the host through tests/native/cuda_on_cpu.h and
driven through partdistillation_b200/functional.py's UNMODIFIED wrappers — by running the GPU parity tests of
tests/test_ops_gpu.py themselves with ``Tensor.cuda()`` turned into the identity and the library handle swapped for the
host build.  Same inputs, same oracle comparisons, same tolerances as on the B200."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import test_ops_gpu as gpu_tests

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OPS = ("point_sample_forward", "point_sample_backward", "matcher_cost", "lsap_batched", "point_loss_forward",
       "point_loss_backward", "class_rows_forward", "class_rows_backward")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    from partdistillation_b200 import _lib
    tmp = tmp_path_factory.mktemp("loss_host")
    src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", "loss.cu")).read()
    m = re.search(r"(namespace pdb \{.*?\n\}  // namespace pdb\n)", src, re.S)
    assert m and "<<<" not in m.group(1) and "class_rows_bwd_w" in m.group(1)
    (tmp / "loss_section.inc").write_text(m.group(1))
    so = str(tmp / "libloss_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-I", str(tmp),
                           os.path.join(HERE, "native", "loss_kernels_host.cpp"), "-o", so])
    cdll = ctypes.CDLL(so)

    class HostLib:
        def pdb_last_error(self):
            return b"host build"
    lib = HostLib()
    for name in OPS:
        f = getattr(cdll, "host_" + name)
        res, args = _lib.SIGNATURES["pdb_" + name]
        f.restype, f.argtypes = res, args[:-1]
        setattr(lib, "pdb_" + name, (lambda f: lambda *a: f(*a[:-1]))(f))
    return lib


@pytest.fixture
def fn(monkeypatch, host_lib):
    from partdistillation_b200 import _lib
    from partdistillation_b200 import functional
    monkeypatch.setattr(_lib, "load", lambda: host_lib)
    monkeypatch.setattr(functional, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(functional, "_stream", lambda: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return functional


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
