"""CPU: host logic of the ProposalModel eval branch (partdistillation_b200/postprocess.py) with the six C-ABI operators
replaced by torch restatements of their contracts (include/pdb200.h, "Inference post-processing").  This checks the
selection / filtering / labelling code around the kernels against the reference's recorded outputs; the kernels
themselves are checked on the GPU (tests/test_postprocess_gpu.py).  The product has no CPU path: without the patch the
same call raises."""
import os

import pytest
import torch

import m2f_oracle as O
from postprocess_cases import oracle_resize, run_case


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


def _postprocess_masks(logits, sel, padded, image_size, out_size, gate=None, scores=None, want_bits=True, want_label=False):
    v = oracle_resize(logits, padded, image_size, out_size)[sel.long()]
    if gate is not None:
        v = v * gate
    on = v > 0
    bits = _pack(torch.cat([on, on.any(0, keepdim=True)])) if want_bits else None
    label = (scores[:, None, None] * v.sigmoid()).argmax(0).to(torch.int32) if want_label else None
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


def test_pack_helpers_round_trip():
    m = torch.rand(3, 5, 70, generator=torch.Generator().manual_seed(0)) > 0.5
    assert torch.equal(_unpack(_pack(m), 70), m)
    assert torch.equal(_popcount(_pack(m)), m.flatten(1).sum(1))


def test_eval_branch_has_no_cpu_path(golden_dir):
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        run_case(g, "prop", "cpu")
