"""CPU: host builds of three small kernel files — optim.cu (flat gradient norm + clip + AdamW, base_trainer.py:118-147),
grouping.cu (one-pass pixel-group affinity, pixel_grouping_model.py:139-144,197-211) and attn_mask.cu (attention-mask build,
mask2former_transformer_decoder.py:453-457) — driven by the GPU parity tests themselves (tests/test_ops_gpu.py,
tests/test_engine.py) with the library handle swapped for the host build."""
import os
import re

import pytest
import torch

import test_engine as engine_tests
import test_ops_gpu as gpu_tests
from host_kernels import NAMESPACE_BLOCK, ROOT, build_host_library, dynamic_smem, patch_functional

OPS = ("grad_sumsq", "adamw_flat", "group_affinity", "group_scores", "attn_mask_build", "attn_mask_reset_rows")


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("misc_host")
    for cu, inc in (("grouping.cu", "grouping_section.inc"), ("attn_mask.cu", "attn_mask_section.inc")):
        src = open(os.path.join(ROOT, "partdistillation_b200", "csrc", cu)).read()
        (tmp / inc).write_text(dynamic_smem("\n".join(re.findall(NAMESPACE_BLOCK, src, re.S))))
    return build_host_library(tmp, "optim.cu", "optim_section.inc", "misc_kernels_host.cpp", OPS)


@pytest.fixture
def fn(monkeypatch, host_lib):
    return patch_functional(monkeypatch, host_lib)


@pytest.mark.parametrize("H,W,h,w", [(64, 64, 32, 32), (64, 64, 16, 16), (64, 64, 8, 8), (40, 56, 20, 28), (33, 47, 9, 13)])
def test_attn_mask_bits(fn, H, W, h, w):
    gpu_tests.test_attn_mask_bits(fn, H, W, h, w)


@pytest.mark.parametrize("metric", ["dot", "l2"])
def test_group_affinity_vs_reference_golden(fn, golden_dir, metric):
    gpu_tests.test_group_affinity_vs_reference_golden(fn, golden_dir, metric)


def test_flat_adamw_matches_torch(fn, monkeypatch):
    """The trainer's flat path (one gradient-norm kernel + one AdamW kernel over all fp32 parameters) against
    torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW with the reference's per-parameter groups — the body of
    tests/test_engine.py::test_flat_adamw_matches_torch on CPU tensors, plus a learning-rate schedule."""
    from partdistillation_b200.engine import DataParallelTrainer, WarmupMultiStepLR, build_param_groups
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))       # the trainer picks the flat kernels
    sched = WarmupMultiStepLR([3], gamma=0.1, warmup_factor=0.1, warmup_iters=2)
    torch.manual_seed(0)
    a, b = engine_tests._Tiny(), engine_tests._Tiny()
    b.load_state_dict(a.state_dict())
    tr = DataParallelTrainer(a, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())
    assert tr.flat_param is not None and tr.optimizer is None          # everything on the flat kernels
    tr.set_lr_schedule(sched)
    groups = build_param_groups(b, 1e-2, 0.05)
    params = [g["params"][0] for g in groups]
    opt = torch.optim.AdamW(groups, lr=1e-2)
    lam = torch.optim.lr_scheduler.LambdaLR(opt, sched.factor)
    for step in range(5):
        batch = engine_tests._data(step)
        tr.step(batch)
        opt.zero_grad()
        sum(b(batch).values()).backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
        lam.step()
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.allclose(v, w, rtol=2e-5, atol=2e-6), k
