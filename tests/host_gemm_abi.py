"""TEST INFRASTRUCTURE ONLY.  The tcgen05 / TMA GEMM of csrc/gemm_tc.cu cannot be compiled for the host, so for whole-model
runs on the CPU tier its five C-ABI entry points are RESTATED here from their contract in include/pdb200.h (float64
accumulation, rounded to fp32) — this is a stand-in for the interface, not the kernel's code; the kernel itself is checked on
the GPU tier (tests/test_ops_gpu.py::test_gemm_tf32x3_layouts etc.).  Every other operator of a whole-model CPU run executes
its own kernel source (tests/host_kernels.py)."""
import ctypes
import math

import numpy as np
import torch


def _view(ptr, shape, strides, dtype=ctypes.c_float):
    """Strided float32 view of raw memory at ``ptr`` (element strides)."""
    n = 1 + sum((s - 1) * st for s, st in zip(shape, strides))
    arr = np.ctypeslib.as_array((dtype * n).from_address(ptr))
    return torch.from_numpy(arr).as_strided(tuple(shape), tuple(strides))


def _activate(r, mode):
    if mode == 1:
        return r.relu()
    if mode == 2:
        return torch.nn.functional.gelu(r)
    return r


def pdb_gemm_tf32x3(A, B, B_lo, C, bias, M, N, K, batch, lda, ldb, ldc, sa, sb, sc, a_mn, b_mn, c_trans, relu, accumulate,
                    ksplit, stream):
    a = _view(A, (batch, M, K), (sa, 1, lda) if a_mn else (sa, lda, 1))
    b = _view(B, (batch, N, K), (sb, 1, ldb) if b_mn else (sb, ldb, 1))
    c = _view(C, (batch, M, N), (sc, 1, ldc) if c_trans else (sc, ldc, 1))
    r = a.double() @ b.double().transpose(1, 2)
    if bias:
        r = r + _view(bias, (N,), (1,)).double()
    r = _activate(r, relu).float()
    for i in range(batch):          # sc == 0 with accumulate: every batch item reduces into the same C (weight gradients)
        if accumulate:
            c[i].add_(r[i])
        else:
            c[i].copy_(r[i])
    return 0


def pdb_gemm_taps_tf32x3(A, B, B_lo, C, bias, M, N, Ck, batch, a_rows, lda, ldc, sa, sc, taps, tap_off, relu, stream):
    offs = [int(v) for v in tap_off]
    a = _view(A, (batch, a_rows, Ck), (sa, lda, 1)).double()
    b = _view(B, (N, taps * Ck), (taps * Ck, 1)).double()
    c = _view(C, (batch, M, N), (sc, ldc, 1))
    r = torch.zeros((batch, M, N), dtype=torch.float64)
    for t, off in enumerate(offs):
        rows = max(0, min(M, a_rows - off))                     # rows beyond a_rows read as zero
        if rows:
            r[:, :rows] += a[:, off:off + rows] @ b[:, t * Ck:(t + 1) * Ck].t()
    if bias:
        r = r + _view(bias, (N,), (1,)).double()
    c.copy_(_activate(r, relu).float())
    return 0


def pdb_split_lo(x, lo, n, stream):
    xi = _view(x, (n,), (1,), ctypes.c_int32)
    hi = (xi & ~0x1FFF).view(torch.float32)                      # low 13 mantissa bits cleared
    _view(lo, (n,), (1,)).copy_(xi.view(torch.float32) - hi)
    return 0


def pdb_mask_einsum_forward(embed, embed_lo, feat, out, B, Q, C, HW, stream):
    e = _view(embed, (B, Q, C), (Q * C, C, 1)).double()
    f = _view(feat, (B, HW, C), (HW * C, C, 1)).double()
    _view(out, (B, Q, HW), (Q * HW, HW, 1)).copy_((e @ f.transpose(1, 2)).float())
    return 0


def pdb_mask_einsum_backward(embed, feat, grad_out, grad_embed, grad_feat, accumulate, B, Q, C, HW, stream):
    go = _view(grad_out, (B, Q, HW), (Q * HW, HW, 1)).double()
    if grad_embed:
        f = _view(feat, (B, HW, C), (HW * C, C, 1)).double()
        _view(grad_embed, (B, Q, C), (Q * C, C, 1)).copy_((go @ f).float())
    if grad_feat:
        e = _view(embed, (B, Q, C), (Q * C, C, 1)).double()
        gf = _view(grad_feat, (B, HW, C), (HW * C, C, 1))
        r = (go.transpose(1, 2) @ e).float()
        if accumulate:
            gf.add_(r)
        else:
            gf.copy_(r)
    return 0


ENTRY_POINTS = dict(pdb_gemm_tf32x3=pdb_gemm_tf32x3, pdb_gemm_taps_tf32x3=pdb_gemm_taps_tf32x3, pdb_split_lo=pdb_split_lo,
                    pdb_mask_einsum_forward=pdb_mask_einsum_forward, pdb_mask_einsum_backward=pdb_mask_einsum_backward)
