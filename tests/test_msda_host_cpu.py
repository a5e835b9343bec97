"""CPU: the MSDeformAttn kernels' own code (csrc/msda.cu — generic f32 / f64, the D = 32 fast path, the tiled encoder path and
the slot-ordered decoder path; forward and backward) compiled for the host through tests/native/cuda_on_cpu.h and checked
against (1) the golden vectors the unmodified reference produced under its own test protocol (ops/test.py: shapes :27-31,
seed :34, fp64 allclose, fp32 rtol 1e-2 / atol 1e-3 — held to 1e-5 here; gradient cases D in {30, 32, 64, 71}) and (2) the
oracle on encoder- / decoder-style inputs with locations partly outside the maps.  GPU twins: tests/test_ops_gpu.py::test_msda_*."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import m2f_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PATH = {1: "generic", 2: "d32", 3: "tiled", 4: "slot-ordered"}


@pytest.fixture(scope="module")
def msda(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("msda_host")
    src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", "msda.cu")).read()
    a = src.index("namespace pdb {") + len("namespace pdb {")
    b = src.index("template <typename T>\nstatic int fwd_generic")
    section, n = re.subn(r"extern __shared__ (?:__align__\(\d+\) )?(\w+) (\w+)\[\];",
                         r"\1* \2 = reinterpret_cast<\1*>(cpu_cuda::g_dyn_smem);", src[a:b])
    assert n == 4 and "<<<" not in section and "msda_bwd_tiled" in section
    (tmp / "msda_section.inc").write_text(section)
    so = str(tmp / "libmsda_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-I", str(tmp),
                           os.path.join(HERE, "native", "msda_kernel_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.host_msda_forward.argtypes = [P] * 6 + [I] * 8
    lib.host_msda_backward.argtypes = [P] * 9 + [I] * 8

    def tables(shapes):
        hw = torch.tensor([v for s in shapes for v in s], dtype=torch.int64)
        starts, o = [], 0
        for h, w in shapes:
            starts.append(o)
            o += h * w
        return hw, torch.tensor(starts, dtype=torch.int64)

    def forward(value, shapes, loc, attn):
        value, loc, attn = value.contiguous(), loc.contiguous(), attn.contiguous()
        N, S, M, D = value.shape
        _, Lq, _, L, P_, _ = loc.shape
        hw, st = tables(shapes)
        out = torch.empty(N, Lq, M * D, dtype=value.dtype)
        path = lib.host_msda_forward(value.data_ptr(), hw.data_ptr(), st.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                     out.data_ptr(), N, S, M, D, Lq, L, P_, 1 if value.dtype == torch.float64 else 0)
        assert path > 0
        return out, PATH[path]

    def backward(value, shapes, loc, attn, grad_out):
        value, loc, attn, grad_out = value.contiguous(), loc.contiguous(), attn.contiguous(), grad_out.contiguous()
        N, S, M, D = value.shape
        _, Lq, _, L, P_, _ = loc.shape
        hw, st = tables(shapes)
        gv = torch.zeros_like(value)                         # pdb_msda_backward zero-fills grad_value itself
        gl = torch.full_like(loc, float("nan"))              # fully overwritten by contract
        ga = torch.full_like(attn, float("nan"))
        path = lib.host_msda_backward(value.data_ptr(), hw.data_ptr(), st.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                      grad_out.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), N, S, M, D, Lq, L, P_,
                                      1 if value.dtype == torch.float64 else 0)
        assert path > 0
        return gv, gl, ga, PATH[path]
    return forward, backward


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_msda_kernel_known_answers(msda, golden_dir):
    forward, _ = msda
    g = torch.load(os.path.join(golden_dir, "msda.pt"), weights_only=False)
    c = g["kat_double"]
    out, path = forward(c["value"].double(), c["shapes"], c["loc"].double(), c["attn"].double())
    assert path == "generic" and torch.allclose(out, c["out"])
    c = g["kat_float"]
    out, path = forward(c["value"], c["shapes"], c["loc"], c["attn"])
    assert path == "generic" and torch.allclose(out, c["out"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("case", ["grad_D30", "grad_D32", "grad_D64", "grad_D71", "cfg_like"])
def test_msda_kernel_gradients(msda, golden_dir, case):
    forward, backward = msda
    c = torch.load(os.path.join(golden_dir, "msda.pt"), weights_only=False)[case]
    out, path = forward(c["value"], c["shapes"], c["loc"], c["attn"])
    assert path == ("tiled" if case == "cfg_like" else "generic")
    tol = dict(rtol=1e-9, atol=1e-12) if out.dtype == torch.float64 else dict(rtol=1e-4, atol=2e-5)
    assert torch.allclose(out, c["out"], **tol)
    gv, gl, ga, bpath = backward(c["value"], c["shapes"], c["loc"], c["attn"], c["grad_out"])
    assert bpath == path
    assert torch.allclose(gv, c["grad_value"], **tol)
    assert torch.allclose(ga, c["grad_attn"], **tol)
    assert _rel(gl, c["grad_loc"]) < (1e-9 if out.dtype == torch.float64 else 1e-4)


def _inputs(N, shapes, Lq, M=8, D=32, P=4, seed=0, spread=6.0, encoder=True):
    g = torch.Generator().manual_seed(seed)
    S = sum(h * w for h, w in shapes)
    L = len(shapes)
    value = torch.randn(N, S, M, D, generator=g)
    ref = O.encoder_reference_points(shapes, N)[:, :Lq] if encoder else torch.rand(N, Lq, 1, 2, generator=g).expand(N, Lq, L, 2)
    off = (torch.rand(N, Lq, M, L, P, 2, generator=g) * 2 - 1) * spread
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=torch.float32)
    loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    return value, loc.contiguous(), attn.contiguous()


@pytest.mark.parametrize("shapes,N,Lq,expect", [
    ([(6, 10), (3, 5), (2, 3)], 2, None, "tiled"),          # encoder self-attention: ragged 8 x 4 query patches
    ([(8, 8), (4, 4)], 1, 10, "slot-ordered"),              # decoder-style queries
    ([(8, 8), (1, 5)], 1, None, "d32"),                     # a level thinner than 2 pixels: per-corner clamping path
])
def test_msda_kernel_fast_paths_vs_oracle(msda, shapes, N, Lq, expect):
    forward, backward = msda
    S = sum(h * w for h, w in shapes)
    value, loc, attn = _inputs(N, shapes, S if Lq is None else Lq, seed=3, encoder=Lq is None)
    v, l, a = (t.clone().requires_grad_() for t in (value, loc, attn))
    ref = O.ms_deform_attn_core(v, shapes, l, a)
    go = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    rgv, rgl, rga = torch.autograd.grad(ref, (v, l, a), go)
    out, path = forward(value, shapes, loc, attn)
    assert path == expect
    gv, gl, ga, bpath = backward(value, shapes, loc, attn, go)
    assert bpath == expect
    assert _rel(out, ref.detach()) < 1e-5
    assert _rel(gv, rgv) < 1e-5 and _rel(ga, rga) < 1e-5 and _rel(gl, rgl) < 1e-4


@pytest.mark.parametrize("shapes,N,Lq,expect,loc_kind", [
    ([(12, 9), (6, 5), (3, 3), (2, 2)], 1, None, "tiled", "wild"),       # 4 levels (L * P = 16), offsets far outside the maps
    ([(5, 4), (3, 2)], 2, 3, "slot-ordered", "uniform"),                 # fewer slots than one CTA holds
    ([(7, 3)], 1, None, "tiled", "edges"),                               # one level; locations exactly on the borders
])
def test_msda_kernel_edge_cases_vs_oracle(msda, shapes, N, Lq, expect, loc_kind):
    forward, backward = msda
    g = torch.Generator().manual_seed(17)
    S = sum(h * w for h, w in shapes)
    L, M, D, P = len(shapes), 8, 32, 4
    Lq = S if Lq is None else Lq
    value = torch.randn(N, S, M, D, generator=g)
    if loc_kind == "wild":
        loc = torch.randn(N, Lq, M, L, P, 2, generator=g) * 3.0            # many samples entirely outside [0, 1]
    elif loc_kind == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    else:
        loc = torch.randint(0, 3, (N, Lq, M, L, P, 2), generator=g).float() / 2      # 0, 0.5, 1
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    v, l, a = (t.clone().requires_grad_() for t in (value, loc, attn))
    ref = O.ms_deform_attn_core(v, shapes, l, a)
    go = torch.randn(ref.shape, generator=g)
    rgv, rgl, rga = torch.autograd.grad(ref, (v, l, a), go)
    out, path = forward(value, shapes, loc, attn)
    gv, gl, ga, bpath = backward(value, shapes, loc, attn, go)
    assert path == expect and bpath == expect
    assert _rel(out, ref.detach()) < 1e-5
    assert _rel(gv, rgv) < 1e-5 and _rel(ga, rga) < 1e-5
    if loc_kind != "edges":          # on exact pixel borders the one-sided derivative is a convention (floor side), not a value
        assert _rel(gl, rgl) < 1e-4


# ------------------------------------------------------------------ the GPU tests' own bodies through functional.py's wrappers
@pytest.fixture(scope="module")
def msda_abi_lib(tmp_path_factory):
    from host_kernels import build_host_library, msda_section
    return build_host_library(tmp_path_factory.mktemp("msda_abi"), "msda.cu", "msda_section.inc", "msda_abi_host.cpp",
                              ("msda_forward", "msda_backward"), section_regex=r"(?s)(.*)", rewrite=msda_section)


@pytest.fixture
def fn(monkeypatch, msda_abi_lib):
    from host_kernels import patch_functional
    return patch_functional(monkeypatch, msda_abi_lib)


def test_gpu_body_msda_known_answers(fn, golden_dir):
    import test_ops_gpu as gpu_tests
    gpu_tests.test_msda_known_answers(fn, golden_dir)


@pytest.mark.parametrize("case", ["grad_D32", "grad_D71", "cfg_like"])
def test_gpu_body_msda_gradients(fn, golden_dir, case):
    import test_ops_gpu as gpu_tests
    gpu_tests.test_msda_gradients(fn, golden_dir, case)


def test_gpu_body_msda_decoder_style_queries(fn):
    import test_ops_gpu as gpu_tests
    gpu_tests.test_msda_decoder_style_queries(fn)


@pytest.mark.parametrize("M,Lq", [(1, 29), (3, 7), (2, 33)])
def test_msda_kernel_d32_partial_warps(msda, M, Lq):
    """Head counts that are not multiples of 4 leave warps that mix valid and out-of-range slots in the last CTA of the
    D = 32 fallback path (levels thinner than 2 pixels).  Its backward kernel used to run its butterfly shuffles with a full
    member mask while the out-of-range lanes waited at the next barrier — a deadlock by CUDA's *_sync rules, found by
    tests/fuzz/fuzz_msda_host.py on this host build (the shim deadlocks exactly where the GPU would)."""
    forward, backward = msda
    shapes = [(8, 1)]
    g = torch.Generator().manual_seed(3)
    value = torch.randn(2, 8, M, 32, generator=g)
    loc = torch.rand(2, Lq, M, 1, 4, 2, generator=g) * 1.4 - 0.2
    attn = torch.softmax(torch.randn(2, Lq, M, 4, generator=g), -1).view(2, Lq, M, 1, 4)
    v, l, a = (t.clone().requires_grad_() for t in (value, loc, attn))
    ref = O.ms_deform_attn_core(v, shapes, l, a)
    go = torch.randn(ref.shape, generator=g)
    rgv, rgl, rga = torch.autograd.grad(ref, (v, l, a), go)
    out, path = forward(value, shapes, loc, attn)
    gv, gl, ga, bpath = backward(value, shapes, loc, attn, go)
    assert path == "d32" and bpath == "d32"
    assert _rel(out, ref.detach()) < 1e-5 and _rel(gv, rgv) < 1e-5 and _rel(ga, rga) < 1e-5 and _rel(gl, rgl) < 1e-4
