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


POSTPROCESS_OPS = ("postprocess_masks", "resize_masks_u8", "pack_bits", "unpack_bits", "bits_popcount", "bits_intersect",
                   "group_affinity_resized", "group_scores")


def build_plain_harness(tmp, harness, ops):
    """A harness that includes header-only kernels directly (csrc/postprocess_kernels.cuh, grouping_resized.cuh)."""
    from partdistillation_b200 import _lib
    so = str(tmp / ("lib" + os.path.splitext(harness)[0] + ".so"))
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(HERE, "native", harness), "-o", so])
    cdll = ctypes.CDLL(so)

    class HostLib:
        def pdb_last_error(self):
            return b"host build"
    lib = HostLib()
    for name in ops:
        f = getattr(cdll, "host_" + name)
        res, args = _lib.SIGNATURES["pdb_" + name]
        f.restype, f.argtypes = res, args[:-1]
        setattr(lib, "pdb_" + name, (lambda f: lambda *a: f(*a[:-1]))(f))
    return lib


HARNESSES = [
    # (csrc file, section file, harness, entry points, extra sections [(csrc file, section file)], rewrite dynamic smem)
    ("msda.cu", "msda_section.inc", "msda_abi_host.cpp", ("msda_forward", "msda_backward"), [], True),
    ("loss.cu", "loss_section.inc", "loss_kernels_host.cpp",
     ("point_sample_forward", "point_sample_backward", "matcher_cost", "lsap_batched", "point_loss_forward",
      "point_loss_backward", "class_rows_forward", "class_rows_backward"), [], False),
    ("xattn.cu", "xattn_section.inc", "xattn_kernels_host.cpp",
     ("masked_xattn_workspace_bytes", "masked_xattn_forward", "masked_xattn_backward"), [], False),
    ("groupnorm.cu", "groupnorm_section.inc", "norm_kernels_host.cpp",
     ("layer_norm_forward", "layer_norm_forward_scaled", "group_norm_forward", "group_norm_backward"), [("layernorm.cu", "layernorm_section.inc")], False),
    ("window_attn.cu", "window_attn_section.inc", "window_attn_kernels_host.cpp",
     ("window_attention_forward", "swin_window_attention_forward"), [], True),
    ("optim.cu", "optim_section.inc", "misc_kernels_host.cpp",
     ("grad_sumsq", "adamw_flat", "group_affinity", "group_scores", "attn_mask_build", "attn_mask_reset_rows"),
     [("grouping.cu", "grouping_section.inc"), ("attn_mask.cu", "attn_mask_section.inc")], True),
]


def msda_section(src):
    """msda.cu's kernels: everything inside `namespace pdb` ahead of the launch helpers (which use <<< >>>)."""
    a = src.index("namespace pdb {") + len("namespace pdb {")
    b = src.index("template <typename T>\nstatic int fwd_generic")
    return dynamic_smem(src[a:b])


def full_host_library(tmp_path_factory):
    """Every SIMT kernel file's own code built for the host + the restated GEMM interface (host_gemm_abi): enough to run
    the whole head, loss and optimizer step on CPU tensors."""
    import host_gemm_abi

    class Composite:
        launches = 0

        def pdb_last_error(self):
            return b"host build"

        def pdb_launch_count(self):
            return self.launches
    lib = Composite()
    for cu, inc, harness, ops, extra, dyn in HARNESSES:
        tmp = tmp_path_factory.mktemp(os.path.splitext(harness)[0])
        for cu2, inc2 in extra:
            src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", cu2)).read()
            (tmp / inc2).write_text(dynamic_smem("\n".join(re.findall(NAMESPACE_BLOCK, src, re.S))))
        if cu == "msda.cu":
            part = build_host_library(tmp, cu, inc, harness, ops, section_regex=r"(?s)(.*)", rewrite=msda_section)
        else:
            part = build_host_library(tmp, cu, inc, harness, ops, rewrite=dynamic_smem if dyn else None)
        for name in ops:
            setattr(lib, "pdb_" + name, getattr(part, "pdb_" + name))
    pp = build_plain_harness(tmp_path_factory.mktemp("pp_kernels"), "postprocess_kernels_host.cpp", POSTPROCESS_OPS)
    for name in POSTPROCESS_OPS:
        setattr(lib, "pdb_" + name, getattr(pp, "pdb_" + name))
    for name, f in host_gemm_abi.ENTRY_POINTS.items():
        setattr(lib, name, f)
    return lib


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
