"""Shared by the CPU-tier kernel tests: builds one csrc/*.cu file's kernels for the host (tests/native/cuda_on_cpu.h) and
wraps the result as a stand-in for the loaded libpdb200.so, so that partdistillation_b200/functional.py's wrappers — and the
GPU parity tests themselves — can run on CPU tensors.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NAMESPACE_BLOCK = r"(namespace pdb \{.*?\n\}  // namespace pdb\n)"


def build_host_library(tmp, cu_file, inc_name, harness, ops, section_regex=NAMESPACE_BLOCK, rewrite=None):
    """Cuts the kernel section out of csrc/<cu_file> into <inc_name>, compiles tests/native/<harness> against it and
    returns an object exposing pdb_<op>(..., stream) -> host_<op>(...) with the ctypes signatures of _lib.SIGNATURES."""
    from partdistillation_b200 import _lib
    src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", cu_file)).read()
    blocks = re.findall(section_regex, src, re.S)          # a file may interleave several namespace blocks with launchers
    assert blocks, f"kernel section not found in {cu_file}"
    section = "\n".join(blocks)
    if rewrite is not None:
        section = rewrite(section)
    assert "<<<" not in section
    (tmp / inc_name).write_text(section)
    so = str(tmp / ("lib" + os.path.splitext(harness)[0] + ".so"))
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off", "-I", str(tmp),
                           os.path.join(HERE, "native", harness), "-o", so])
    cdll = ctypes.CDLL(so)

    class HostLib:
        launches = 0

        def pdb_last_error(self):
            return b"host build"

        def pdb_launch_count(self):
            return self.launches
    lib = HostLib()
    for name in ops:
        f = getattr(cdll, "host_" + name)
        res, args = _lib.SIGNATURES["pdb_" + name]
        stream_arg = bool(args) and name != "masked_xattn_workspace_bytes"
        f.restype, f.argtypes = res, (args[:-1] if stream_arg else args)
        setattr(lib, "pdb_" + name, (lambda f, s: lambda *a: f(*(a[:-1] if s else a)))(f, stream_arg))
    return lib


def dynamic_smem(section):
    """`extern __shared__ [__align__(n)] T name[];` -> the shim's dynamic shared memory pointer."""
    return re.sub(r"extern __shared__ (?:__align__\(\d+\) )?(\w+) (\w+)\[\];",
                  r"\1* \2 = reinterpret_cast<\1*>(cpu_cuda::g_dyn_smem);", section)


def patch_functional(monkeypatch, host_lib):
    """functional.py's wrappers on CPU tensors: library handle = host build, CUDA-only guard and stream lookup disabled,
    Tensor.cuda() = identity (so the GPU tests' own code runs as is)."""
    import torch
    from partdistillation_b200 import _lib
    from partdistillation_b200 import functional
    monkeypatch.setattr(_lib, "load", lambda: host_lib)
    monkeypatch.setattr(functional, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(functional, "_stream", lambda: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    return functional
