"""GPU: the learning-rate schedule on the flat clip + AdamW kernels (csrc/optim.cu reads the per-segment rates from device
memory) against torch.optim.AdamW driven by LambdaLR with the same factors.  Sorted last on purpose: it was written after
round 1's GPU budget was spent (CPU twin: tests/test_engine.py::test_trainer_applies_lr_schedule_cpu)."""
import pytest
import torch

from test_engine import _Tiny, _data

pytestmark = pytest.mark.gpu


def test_flat_adamw_follows_lr_schedule():
    from partdistillation_b200.engine import DataParallelTrainer, WarmupMultiStepLR, build_param_groups
    sched = WarmupMultiStepLR([4], gamma=0.1, warmup_factor=0.1, warmup_iters=3)
    torch.manual_seed(0)
    a, b = _Tiny().cuda(), _Tiny().cuda()
    b.load_state_dict(a.state_dict())
    tr = DataParallelTrainer(a, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())
    assert tr.flat_param is not None and tr.optimizer is None          # everything on the flat kernels
    tr.set_lr_schedule(sched)
    groups = build_param_groups(b, 1e-2, 0.05)
    params = [g["params"][0] for g in groups]
    opt = torch.optim.AdamW(groups, lr=1e-2)
    lam = torch.optim.lr_scheduler.LambdaLR(opt, sched.factor)
    for step in range(7):
        x, y = _data(step)
        batch = (x.cuda(), y.cuda())
        tr.step(batch)
        assert torch.allclose(tr.seg_lr, tr.seg_lr_base * sched.factor(step))
        opt.zero_grad()
        sum(b(batch).values()).backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
        lam.step()
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.allclose(v, w, rtol=2e-5, atol=2e-6), k
