"""CPU: the Swin window-attention kernels' own code (csrc/window_attn.cu: the per-window kernel and the one-kernel
shifted-window attention on the token grid, swin.py:78-176,239-300) compiled for the host and driven by the GPU parity
tests of tests/test_ops_gpu.py themselves (float64 references, same tolerances)."""
import pytest

import test_ops_gpu as gpu_tests
from host_kernels import build_host_library, dynamic_smem, patch_functional


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return build_host_library(tmp_path_factory.mktemp("win_host"), "window_attn.cu", "window_attn_section.inc",
                              "window_attn_kernels_host.cpp", ("window_attention_forward", "swin_window_attention_forward"),
                              rewrite=dynamic_smem)


@pytest.fixture
def fn(monkeypatch, host_lib):
    return patch_functional(monkeypatch, host_lib)


@pytest.mark.parametrize("Bw,N,heads,nW", [(4, 144, 2, 2), (6, 16, 2, 0), (3, 49, 3, 3)])
def test_window_attention(fn, Bw, N, heads, nW):
    gpu_tests.test_window_attention(fn, Bw, N, heads, nW)


@pytest.mark.parametrize("H,W,ws,shift,heads", [(24, 24, 12, 6, 1), (20, 30, 12, 6, 1), (16, 16, 4, 0, 2), (9, 7, 4, 2, 1)])
def test_swin_window_attention_block(fn, H, W, ws, shift, heads):
    gpu_tests.test_swin_window_attention_block(fn, H, W, ws, shift, heads)
