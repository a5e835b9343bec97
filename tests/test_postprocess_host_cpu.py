"""CPU: host logic of the ProposalModel eval branch (partdistillation_b200/postprocess.py) with the six C-ABI operators
replaced by torch restatements of their contracts (include/pdb200.h, "Inference post-processing").  This checks the
selection / filtering / labelling code around the kernels against the reference's recorded outputs; the kernels
themselves are checked on the GPU (tests/test_postprocess_gpu.py).  The product has no CPU path: without the patch the
same call raises."""
import os

import pytest
import torch

import m2f_oracle as O
from postprocess_cases import oracle_resize, run_case, run_pd_case


def _pack(m):
    R, H, W = m.shape
    Ww = (W + 31) // 32
    p = torch.zeros((R, H, Ww * 32), dtype=torch.int64)
    p[..., :W] = m.to(torch.int64)
    words = (p.view(R, H, Ww, 32) << torch.arange(32)).sum(-1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


def _unpack(bits, width, rows=None):
    b = bits.to(torch.int64) & 0xFFFFFFFF
    out = ((b[..., None] >> torch.arange(32)) & 1).bool().flatten(-2)[..., :width]
    return out if rows is None else out[rows.long()]


def _postprocess_masks(logits, sel, padded, image_size, out_size, gate=None, scores=None, want_bits=True, want_label=False,
                       score_threshold=None):
    v = oracle_resize(logits, padded, image_size, out_size)[sel.long()]
    if gate is not None:
        v = v * gate
    on = v > 0
    bits = _pack(torch.cat([on, on.any(0, keepdim=True)])) if want_bits else None
    label = (scores[:, None, None] * v.sigmoid()).argmax(0).to(torch.int32) if want_label else None
    if score_threshold is not None:
        return bits, label, _pack(scores[:, None, None] * v.sigmoid() > score_threshold)
    return bits, label


def _popcount(bits):
    return _unpack(bits, 32 * bits.shape[-1]).flatten(1).sum(1)


def _iou(a, b):
    return O.mask_iou(_unpack(a, 32 * a.shape[-1]), _unpack(b, 32 * b.shape[-1]))


@pytest.fixture
def torch_ops(monkeypatch):
    from partdistillation_b200 import functional as fn
    monkeypatch.setattr(fn, "postprocess_masks", _postprocess_masks)
    monkeypatch.setattr(fn, "resize_bool_masks", lambda m, i, o: O.sem_seg_postprocess(m.float(), i, *o).bool())
    monkeypatch.setattr(fn, "pack_bits", _pack)
    monkeypatch.setattr(fn, "unpack_bits", _unpack)
    monkeypatch.setattr(fn, "bits_popcount", _popcount)
    monkeypatch.setattr(fn, "bits_iou", _iou)


@pytest.mark.parametrize("case", ["prop", "prop_filtered", "prop_nomask", "semseg", "semseg_filtered"])
def test_eval_branch_host_logic(torch_ops, golden_dir, case):
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    run_case(g, case, "cpu")


PD_CASES = ["prop", "prop_filtered", "prop_none_valid", "semseg", "semseg_filtered", "semseg_oracle_cls"]


@pytest.mark.parametrize("case", PD_CASES)
def test_pd_eval_branch_host_logic(torch_ops, golden_dir, case):
    g = torch.load(os.path.join(golden_dir, "pd_inference.pt"), weights_only=False)
    run_pd_case(g, case, "cpu")


def test_pack_helpers_round_trip():
    m = torch.rand(3, 5, 70, generator=torch.Generator().manual_seed(0)) > 0.5
    assert torch.equal(_unpack(_pack(m), 70), m)
    assert torch.equal(_popcount(_pack(m)), m.flatten(1).sum(1))


def test_eval_branch_has_no_cpu_path(golden_dir):
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        run_case(g, "prop", "cpu")


# ------------------------------------------------------------------ the kernels' per-pixel arithmetic, compiled for the host
GEOMETRIES = [
    # (h, w), padded, image size, output size
    ((32, 32), (128, 128), (96, 128), (144, 192)),
    ((32, 32), (128, 128), (128, 112), (128, 112)),
    ((40, 56), (160, 224), (150, 200), (75, 101)),
    ((40, 56), (160, 224), (150, 200), (333, 517)),
    ((8, 8), (32, 32), (32, 32), (32, 32)),
    ((256, 256), (1024, 1024), (1024, 1024), (1024, 1024)),      # BASELINE configs[1] geometry
]


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    """tests/native/postprocess_host.cpp + csrc/postprocess_math.cuh built with g++ into a temporary directory."""
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path_factory.mktemp("pp_host") / "libpp_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                           os.path.join(here, "native", "postprocess_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    ints = [ctypes.c_int] * 9
    lib.pp_host_resize.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + ints
    lib.pp_host_resize.restype = None
    lib.pp_host_resize_masks.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + ints[:7]
    lib.pp_host_resize_masks.restype = None
    return lib


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_kernel_pixel_math_matches_oracle(host_math, geom):
    """The composed two-pass bilinear value of postprocess_math.cuh (host build) against F.interpolate o crop o
    F.interpolate: within fp32 interpolation noise everywhere, identical threshold bits outside |v| < 1e-4."""
    (h, w), padded, image_size, out_size = geom
    g = torch.Generator().manual_seed(5)
    K = 3 if h < 100 else 1
    logits = (torch.randn(K, h, w, generator=g) * 2.0).contiguous()
    logits[0, : h // 2] = 0.0
    out = torch.empty(K, *out_size)
    host_math.pp_host_resize(logits.data_ptr(), out.data_ptr(), K, h, w, *padded, *image_size, *out_size)
    ref = oracle_resize(logits, padded, image_size, out_size)
    assert (out - ref).abs().max() < 5e-5
    flips = (out > 0) != (ref > 0)
    assert not (flips & (ref.abs() > 1e-4)).any()
    assert flips.float().mean() < 1e-3
    assert torch.equal(out[0, : out_size[0] // 4] == 0, ref[0, : out_size[0] // 4] == 0)    # exact zeros stay exact


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_kernel_mask_resize_math_matches_oracle(host_math, geom):
    _, padded, image_size, out_size = geom
    g = torch.Generator().manual_seed(2)
    G = 3
    m = torch.zeros(G, *padded, dtype=torch.uint8)
    lab = torch.randint(0, 4, (image_size[0] // 8 + 1, image_size[1] // 8 + 1), generator=g)
    lab = lab.repeat_interleave(8, 0).repeat_interleave(8, 1)[:image_size[0], :image_size[1]]
    for k in range(G - 1):
        m[k, :image_size[0], :image_size[1]] = (lab == k).to(torch.uint8)
    out = torch.empty(G, *out_size, dtype=torch.uint8)
    host_math.pp_host_resize_masks(m.data_ptr(), out.data_ptr(), G, *padded, *image_size, *out_size)
    exp = O.sem_seg_postprocess(m.float(), image_size, *out_size).bool()
    assert torch.equal(out.bool(), exp)
