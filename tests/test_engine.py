"""Training-step engine: (1) world_size-2 gloo run on CPU of the data-parallel logic (flat gradient buffer, ONE
all-reduce, 1/world scaling, clip, AdamW) against a single process fed both shards; (2) GPU: the flat
clip + AdamW kernels (csrc/optim.cu) against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW with the
reference's per-parameter groups (base_trainer.py:65-147)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.backbone = torch.nn.Linear(6, 8)
        self.head = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.LayerNorm(8), torch.nn.Linear(8, 3))
        self.embed = torch.nn.Embedding(4, 8)

    def forward(self, batch):
        x, y = batch
        h = self.head(torch.relu(self.backbone(x)) + self.embed.weight.sum(0))
        return {"loss_a": (h - y).pow(2).mean(), "loss_b": h.abs().mean() * 0.1}


def _data(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(5, 6, generator=g), torch.randn(5, 3, generator=g)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from partdistillation_b200.engine import DataParallelTrainer
    torch.manual_seed(rank)                 # every rank initialises differently (the reference seeds with seed + rank) ...
    model = _Tiny()
    tr = DataParallelTrainer(model, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())   # ... rank 0 is broadcast
    for step in range(3):
        tr.step(_data(100 * step + rank))
    torch.save({k: v.clone() for k, v in model.state_dict().items()}, f"{out_path}.{rank}")
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_data_parallel_step_gloo_world2(tmp_path):
    out = str(tmp_path / "sd")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    sd0, sd1 = torch.load(out + ".0"), torch.load(out + ".1")
    for k in sd0:                                       # replicas stay identical
        assert torch.equal(sd0[k], sd1[k]), k
    # single process, gradients averaged over the two shards by hand, torch's own clip + AdamW
    sys.path.insert(0, ROOT)
    from partdistillation_b200.engine import build_param_groups
    torch.manual_seed(0)
    model = _Tiny()
    groups = build_param_groups(model, 1e-2, 0.05)
    params = [g["params"][0] for g in groups]
    opt = torch.optim.AdamW(groups, lr=1e-2)
    for step in range(3):
        opt.zero_grad()
        for r in range(2):
            (sum(model(_data(100 * step + r)).values()) / 2).backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
    for k, v in model.state_dict().items():
        assert torch.allclose(sd0[k], v, rtol=1e-5, atol=1e-6), k


@pytest.mark.gpu
def test_flat_adamw_matches_torch():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    sys.path.insert(0, ROOT)
    from partdistillation_b200.engine import DataParallelTrainer, build_param_groups
    torch.manual_seed(0)
    a, b = _Tiny().cuda(), _Tiny().cuda()
    b.load_state_dict(a.state_dict())
    tr = DataParallelTrainer(a, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())
    assert tr.flat_param is not None and tr.optimizer is None          # everything on the flat kernels
    groups = build_param_groups(b, 1e-2, 0.05)
    params = [g["params"][0] for g in groups]
    opt = torch.optim.AdamW(groups, lr=1e-2)
    for step in range(5):
        x, y = _data(step)
        batch = (x.cuda(), y.cuda())
        tr.step(batch)
        opt.zero_grad()
        sum(b(batch).values()).backward()
        torch.nn.utils.clip_grad_norm_(params, 0.5)
        opt.step()
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.allclose(v, w, rtol=2e-5, atol=2e-6), k


def test_lr_schedules_follow_detectron2_formulas():
    """WarmupMultiStepLR / WarmupPolyLR factors against the published detectron2==0.6 expressions (un-vendored third
    party; restated independently here) and against torch's MultiStepLR after the warm-up."""
    import bisect
    from partdistillation_b200.engine import WarmupMultiStepLR, WarmupPolyLR, build_lr_schedule
    from partdistillation_b200.compat import CfgNode
    s = WarmupMultiStepLR([40, 45], gamma=0.1, warmup_factor=0.001, warmup_iters=10, warmup_method="linear")
    for it in (0, 1, 5, 9, 10, 39, 40, 44, 45, 60):
        alpha = it / 10
        warm = 1.0 if it >= 10 else 0.001 * (1 - alpha) + alpha
        assert s.factor(it) == pytest.approx(warm * 0.1 ** bisect.bisect_right([40, 45], it), rel=1e-12)
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    ref = torch.optim.lr_scheduler.MultiStepLR(opt, [40, 45], 0.1)
    for it in range(60):
        if it >= 10:
            assert s.factor(it) == pytest.approx(ref.get_last_lr()[0], rel=1e-9)
        opt.step()
        ref.step()
    c = WarmupMultiStepLR([3], warmup_factor=0.5, warmup_iters=4, warmup_method="constant")
    assert [c.factor(i) for i in range(6)] == [0.5, 0.5, 0.5, pytest.approx(0.05), pytest.approx(0.1), pytest.approx(0.1)]
    p = WarmupPolyLR(100, power=0.9, constant_ending=0.05, warmup_factor=1.0, warmup_iters=10)
    assert p.factor(0) == 1.0 and p.factor(50) == pytest.approx(0.5 ** 0.9)
    assert p.factor(99) == 0.05                                     # (1 - 0.99) ** 0.9 = 0.0158 < constant ending
    with pytest.raises(ValueError):
        WarmupMultiStepLR([5, 3])
    cfg = CfgNode({"SOLVER": {"STEPS": (40000, 45000, 70000), "MAX_ITER": 50000, "WARMUP_FACTOR": 1.0, "WARMUP_ITERS": 10}})
    b = build_lr_schedule(cfg)                                      # the recipe of sh_files/proposal_learning/train_multi.sh
    assert b.milestones == [40000, 45000] and b.factor(0) == 1.0 and b.factor(45000) == pytest.approx(0.01)


def test_trainer_applies_lr_schedule_cpu():
    """DataParallelTrainer.set_lr_schedule on a CPU model (torch.optim.AdamW groups): same trajectory as torch AdamW
    driven by LambdaLR with the same factors and the same clipping."""
    from partdistillation_b200.engine import DataParallelTrainer, WarmupMultiStepLR, build_param_groups
    sched = WarmupMultiStepLR([2], gamma=0.1, warmup_factor=0.1, warmup_iters=2)
    torch.manual_seed(0)
    a, b = _Tiny(), _Tiny()
    b.load_state_dict(a.state_dict())
    tr = DataParallelTrainer(a, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())
    tr.set_lr_schedule(sched)
    groups = build_param_groups(b, 1e-2, 0.05, freeze_keys=())
    opt = torch.optim.AdamW(groups, lr=1e-2)
    lam = torch.optim.lr_scheduler.LambdaLR(opt, sched.factor)
    for step in range(4):
        assert tr.current_lr_factor() == pytest.approx(sched.factor(step))
        tr.step(_data(step))
        opt.zero_grad()
        sum(b(_data(step)).values()).backward()
        torch.nn.utils.clip_grad_norm_([p for g in groups for p in g["params"]], 0.5)
        opt.step()
        lam.step()
    assert tr.iteration == 4
    for (k, va), vb in zip(a.state_dict().items(), b.state_dict().values()):
        assert torch.allclose(va, vb, rtol=1e-5, atol=1e-7), k


def test_trainer_state_dict_round_trip_cpu():
    """Checkpoint / resume (the reference's DefaultTrainer saves optimizer, scheduler and iteration): a trainer restored from
    model + trainer state continues exactly like the one that kept running, and the state loads into a plain
    torch.optim.AdamW with the reference's one-group-per-parameter layout."""
    from partdistillation_b200.engine import DataParallelTrainer, WarmupMultiStepLR, build_param_groups
    sched = WarmupMultiStepLR([3], gamma=0.1, warmup_factor=0.1, warmup_iters=2)
    torch.manual_seed(0)
    a = _Tiny()
    ta = DataParallelTrainer(a, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())
    ta.set_lr_schedule(sched)
    for step in range(2):
        ta.step(_data(step))
    ckpt = {"model": {k: v.clone() for k, v in a.state_dict().items()}, "trainer": ta.state_dict()}
    assert ckpt["trainer"]["iteration"] == 2 and len(ckpt["trainer"]["param_groups"]) == len(ta.params)
    b = _Tiny()
    b.load_state_dict(ckpt["model"])
    tb = DataParallelTrainer(b, base_lr=1e-2, weight_decay=0.05, clip_norm=0.5, freeze_keys=())
    tb.set_lr_schedule(sched)
    tb.load_state_dict(ckpt["trainer"])
    assert tb.iteration == 2
    # the same state in a plain torch optimizer
    c = _Tiny()
    c.load_state_dict(ckpt["model"])
    groups = build_param_groups(c, 1e-2, 0.05, freeze_keys=())
    opt = torch.optim.AdamW(groups, lr=1e-2)
    tsd = {"state": ckpt["trainer"]["state"], "param_groups": [dict(g, **{k: v for k, v in og.items() if k not in g})
                                                               for g, og in zip(ckpt["trainer"]["param_groups"],
                                                                                opt.state_dict()["param_groups"])]}
    opt.load_state_dict(tsd)
    base = [g["lr"] for g in opt.param_groups]              # the state dict carries the un-scheduled per-group rates
    for step in range(2, 5):
        ta.step(_data(step))
        tb.step(_data(step))
        for g, lr in zip(opt.param_groups, base):
            g["lr"] = lr * sched.factor(step)
        opt.zero_grad()
        sum(c(_data(step)).values()).backward()
        torch.nn.utils.clip_grad_norm_([p for g in groups for p in g["params"]], 0.5)
        opt.step()
    for (k, va), vb, vc in zip(a.state_dict().items(), b.state_dict().values(), c.state_dict().values()):
        assert torch.equal(va, vb), k
        assert torch.allclose(va, vc, rtol=1e-5, atol=1e-7), k


def test_padded_counts_rules():
    from partdistillation_b200.engine import DataParallelTrainer

    class Arch(torch.nn.Module):
        target_padding = False
        num_queries = 10

        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(4))
    tr = DataParallelTrainer(Arch(), freeze_keys=(), target_bucket=4)
    assert tr.model.target_padding is True

    class I:
        def __init__(self, k):
            self.gt_masks = type("M", (), {"tensor": torch.zeros(k, 2, 2)})()
    counts = tr._padded_counts([{"instances": I(k)} for k in (0, 1, 4, 5, 9, 10, 13)])
    assert counts == [0, 4, 4, 8, 10, 10, 13]                     # 0 stays, capped at the query count, K >= Q untouched


def test_padded_batch_allocates_then_refills_in_place():
    """engine._padded_batch: K real rows + empty slots marked class -1; a second call refills the same tensors (the static inputs
    of a captured step) and clears what the previous, larger batch left in the tail; the object class travels as a device
    scalar next to the host value."""
    from partdistillation_b200.compat import BitMasks, Instances
    from partdistillation_b200.engine import _padded_batch

    def item(k, obj, fill):
        inst = Instances((4, 4))
        inst.gt_masks = BitMasks(torch.full((k, 4, 4), fill, dtype=torch.bool))
        inst.gt_classes = torch.arange(k)
        return {"image": torch.full((3, 4, 4), k, dtype=torch.uint8), "instances": inst, "gt_object_class": obj}
    dev = torch.device("cpu")
    static = _padded_batch([item(3, 7, True)], dev, [4])
    s = static[0]
    assert s["instances"].gt_masks.tensor.shape == (4, 4, 4) and s["instances"].gt_masks.tensor[:3].all()
    assert not s["instances"].gt_masks.tensor[3].any() and s["instances"].gt_classes.tolist() == [0, 1, 2, -1]
    assert int(s["gt_object_class_dev"]) == 7 and s["gt_object_class"] == 7
    ptr = s["instances"].gt_masks.tensor.data_ptr()
    again = _padded_batch([item(1, 2, True)], dev, [4], out=static)
    assert again is static and s["instances"].gt_masks.tensor.data_ptr() == ptr
    assert s["instances"].gt_masks.tensor[:1].all() and not s["instances"].gt_masks.tensor[1:].any()
    assert s["instances"].gt_classes.tolist() == [0, -1, -1, -1] and int(s["gt_object_class_dev"]) == 2
    assert int(s["image"][0, 0, 0]) == 1


def _num_masks_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from partdistillation_b200.engine import DataParallelTrainer
    from partdistillation_b200.modeling.criterion import SetCriterion
    from partdistillation_b200.modeling.targets import TargetList
    crit = SetCriterion.__new__(SetCriterion)            # only the normaliser logic is exercised
    crit.external_num_masks = None
    t = TargetList()
    # rank 0: 3 real + 1 padding slot, rank 1: no real target at all (2 padding slots)
    t.packed_labels = torch.tensor([0, 0, 0, -1] if rank == 0 else [-1, -1], dtype=torch.int32)
    t.offsets = [0, len(t.packed_labels)]
    t.has_dummies = True
    padded = float(crit._num_masks(t, torch.device("cpu")))
    t.has_dummies = False                                # shape-based count (the unpadded contract): (4 + 2) / 2
    plain = float(crit._num_masks(t, torch.device("cpu")))

    class Arch(torch.nn.Module):                         # the trainer's pre-step all-reduce uses the RAW (unpadded) counts
        target_padding = False
        num_queries = 10

        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(4))
            self.criterion = type("Crit", (), {"external_num_masks": None})()
    arch = Arch()
    tr = DataParallelTrainer(arch, freeze_keys=(), target_bucket=4)

    class I:
        def __init__(self, k):
            self.gt_masks = type("M", (), {"tensor": torch.zeros(k, 2, 2)})()
    tr._global_num_masks([{"instances": I(3 if rank == 0 else 0)}])
    torch.save({"padded": padded, "plain": plain, "external": float(arch.criterion.external_num_masks)}, f"{out_path}.{rank}")
    dist.destroy_process_group()


def test_num_masks_counts_real_targets_gloo_world2(tmp_path):
    """criterion.py:248-254 under target bucketing: the normaliser is the mean over ranks of the REAL target counts (padding slots
    carry label -1), clamped to >= 1, whether the criterion computes it (eager) or the trainer hands it over (captured steps)."""
    out = str(tmp_path / "nm")
    mp.spawn(_num_masks_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for rank in range(2):
        r = torch.load(f"{out}.{rank}")
        assert r["padded"] == 1.5 and r["plain"] == 3.0 and r["external"] == 1.5
