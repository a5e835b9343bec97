"""Differential fuzzing of the Hungarian-assignment kernel's own code (csrc/loss.cu, host build through
tests/native/cuda_on_cpu.h) against scipy.optimize.linear_sum_assignment + the reference's ascending-cost order
(matcher.py:159-163): random batch sizes, Q in [1, 40), K in [0, 50), float / heavily tied / rounded / duplicated costs.
    python tests/fuzz/fuzz_lsap_host.py [seconds]        # round 1: 11 821 cases in 240 s, 0 mismatches
No GPU needed; nothing here is part of the product."""
import sys, os, re, subprocess, ctypes, numpy as np, pathlib, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scipy.optimize import linear_sum_assignment
tmp = pathlib.Path(tempfile.mkdtemp())
src = open(ROOT+'/partdistillation_b200/csrc/loss.cu').read()
m = re.search(r"// batched rectangular LSAP.*?\n// -+\n(.*?)// -+\n// fused point-sampled BCE \+ dice", src, re.S)
(tmp/'lsap_section.inc').write_text(m.group(1))
so = str(tmp/'l.so')
subprocess.check_call(["g++","-O1","-std=c++20","-pthread","-shared","-fPIC","-I",str(tmp),ROOT+"/tests/native/lsap_kernel_host.cpp","-o",so])
lib = ctypes.CDLL(so)
lib.host_lsap_batched.argtypes=[ctypes.c_void_p]*4+[ctypes.c_int]*2
rng = np.random.default_rng(12345)
bad = 0; n = 0
import time; t0=time.time()
while time.time() - t0 < float(sys.argv[1] if len(sys.argv) > 1 else 240):
    B = int(rng.integers(1, 9))
    Q = int(rng.integers(1, 40))
    costs=[]
    for b in range(B):
        K = int(rng.integers(0, 50))
        kind = rng.integers(0, 4)
        if kind == 0: c = rng.standard_normal((Q,K)).astype(np.float32)
        elif kind == 1: c = rng.integers(0, 3, (Q,K)).astype(np.float32)
        elif kind == 2: c = np.round(rng.standard_normal((Q,K))*2).astype(np.float32)
        else:
            c = rng.standard_normal((Q,K)).astype(np.float32)
            if K>1: c[:, rng.integers(0,K)] = c[:, 0]
            if Q>1: c[rng.integers(0,Q)] = c[0]
        costs.append(c)
    offs = np.concatenate([[0], np.cumsum([c.shape[1] for c in costs])]).astype(np.int32)
    if offs[-1]==0: continue
    flat = np.concatenate([c.ravel() for c in costs]).astype(np.float32)
    pi = np.full(offs[-1], -7, np.int64); ti = np.full(offs[-1], -7, np.int64)
    lib.host_lsap_batched(flat.ctypes.data, offs.ctypes.data, pi.ctypes.data, ti.ctypes.data, B, Q)
    for b,c in enumerate(costs):
        K=c.shape[1]; nn=min(Q,K); s=offs[b]
        if K==0: continue
        i,j = linear_sum_assignment(c.astype(np.float64))
        order = np.argsort(c[i,j], kind='stable')
        n+=1
        if not (np.array_equal(pi[s:s+nn], i[order]) and np.array_equal(ti[s:s+nn], j[order])):
            bad+=1
            if bad<=3: print("MISMATCH", Q, K, c.tolist() if Q*K<40 else '', pi[s:s+nn], i[order], ti[s:s+nn], j[order])
print("cases", n, "mismatches", bad)
