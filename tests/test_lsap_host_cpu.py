"""CPU: the Hungarian-assignment kernel's own code (csrc/loss.cu, lsap_kernel: float64 shortest augmenting paths with
SciPy's tie-breaking, then the reference's ascending-cost order, matcher.py:159-163) compiled for the host through
tests/native/cuda_on_cpu.h and checked bit for bit against scipy.optimize.linear_sum_assignment on randomised cost
matrices — both orientations (Q > K and K > Q), heavy ties, duplicate rows / columns, several images per launch.
The GPU twin is tests/test_ops_gpu.py::test_lsap_exact_vs_scipy; the bit-exact contract is BASELINE.json's."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

import m2f_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lsap(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("lsap_host")
    src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", "loss.cu")).read()
    m = re.search(r"// batched rectangular LSAP.*?\n// -+\n(.*?)// -+\n// fused point-sampled BCE \+ dice", src, re.S)
    assert m, "LSAP section banners not found in loss.cu"
    section = m.group(1)
    assert "lsap_kernel" in section and "cand_merge" in section and "point_loss_fwd" not in section
    (tmp / "lsap_section.inc").write_text(section)
    so = str(tmp / "liblsap_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-I", str(tmp),
                           os.path.join(HERE, "native", "lsap_kernel_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.host_lsap_batched.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_int]

    def run(costs):
        """costs: list of (Q, K_b) float32 arrays (same Q) -> list of (pred_idx, tgt_idx) int64 arrays."""
        Q = costs[0].shape[0]
        offs = np.concatenate([[0], np.cumsum([c.shape[1] for c in costs])]).astype(np.int32)
        flat = np.concatenate([np.ascontiguousarray(c, dtype=np.float32).ravel() for c in costs]) if offs[-1] else np.zeros(1, np.float32)
        pi = np.full(max(int(offs[-1]), 1), -7, dtype=np.int64)
        ti = np.full(max(int(offs[-1]), 1), -7, dtype=np.int64)
        lib.host_lsap_batched(flat.ctypes.data, offs.ctypes.data, pi.ctypes.data, ti.ctypes.data, len(costs), Q)
        out = []
        for b, c in enumerate(costs):
            n = min(Q, c.shape[1])
            s = offs[b]
            assert (pi[s + n:offs[b + 1]] == -1).all() and (ti[s + n:offs[b + 1]] == -1).all()     # unmatched tail
            out.append((pi[s:s + n].copy(), ti[s:s + n].copy()))
        return out
    return run


def reference_pairs(cost):
    """matcher.py:159-163 on one (Q, K) matrix: SciPy's assignment, re-ordered by ascending matched cost (stable)."""
    i, j = linear_sum_assignment(cost.astype(np.float64))
    order = np.argsort(cost[i, j], kind="stable")
    return i[order].astype(np.int64), j[order].astype(np.int64)


def random_costs(rng, Q, K, kind):
    if kind == "float":
        return rng.standard_normal((Q, K)).astype(np.float32) * 3
    if kind == "ties":
        return rng.integers(0, 4, (Q, K)).astype(np.float32)
    if kind == "dup":
        c = rng.standard_normal((Q, K)).astype(np.float32)
        c[Q // 2:] = c[: Q - Q // 2]                    # duplicate rows
        if K > 1:
            c[:, -1] = c[:, 0]                          # duplicate column
        return c
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["float", "ties", "dup"])
def test_lsap_kernel_matches_scipy(lsap, kind):
    rng = np.random.default_rng({"float": 1, "ties": 2, "dup": 3}[kind])
    shapes = [(10, 3), (10, 1), (25, 6), (25, 8), (7, 7), (6, 15), (12, 40), (100, 5), (100, 8), (40, 33), (33, 40)]
    for Q, K in shapes:
        for rep in range(12 if Q * K < 600 else 3):
            costs = [random_costs(rng, Q, K, kind), random_costs(rng, Q, max(1, K - 1), kind)]
            got = lsap(costs)
            for c, (pi, ti) in zip(costs, got):
                ri, rj = reference_pairs(c)
                assert c[pi, ti].sum() == pytest.approx(c[ri, rj].sum(), rel=1e-6)       # optimal in any case
                assert np.array_equal(pi, ri) and np.array_equal(ti, rj), (kind, Q, K, rep)


def test_lsap_kernel_matches_oracle_and_handles_empty_images(lsap):
    """Against the oracle's restatement (m2f_oracle.lsap_jv) and with an image that has no targets in the launch."""
    rng = np.random.default_rng(9)
    costs = [rng.standard_normal((20, 4)).astype(np.float32), np.zeros((20, 0), np.float32),
             rng.integers(0, 3, (20, 5)).astype(np.float32)]
    got = lsap(costs)
    assert got[1][0].size == 0 and got[1][1].size == 0
    for c, (pi, ti) in ((costs[0], got[0]), (costs[2], got[2])):
        oi, oj = O.lsap_jv(torch.from_numpy(c).double())
        oi, oj = np.asarray(oi), np.asarray(oj)
        order = np.argsort(c[oi, oj], kind="stable")
        assert np.array_equal(pi, oi[order]) and np.array_equal(ti, oj[order])
