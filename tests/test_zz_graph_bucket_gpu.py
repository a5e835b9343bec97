"""GPU: DataParallelTrainer(cuda_graph=True, target_bucket=m) — batches whose images carry different numbers of pseudo masks
(and, for PartDistillation, different object classes) share one captured step, and that step computes what the eager,
unpadded trainer computes (variable-K batches must reuse one graph)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class _SharedRows:
    """Capturable point provider: every row of a draw shares one device-resident pattern, so padding slots (extra rows) do not
    shift the coordinates of the real pairs and the padded and unpadded runs sample the same points."""

    def __init__(self):
        self.base = {}

    def __call__(self, *size, device=None, dtype=None, **kw):
        P = size[-2]
        if P not in self.base:          # first use is an eager warm-up step, never the capture
            self.base[P] = torch.rand(1, P, 2, generator=torch.Generator().manual_seed(P)).to(device)
        return self.base[P].expand(size[0], -1, -1).contiguous()


def _batch(seed, counts, pd, parts=4, objects=5, size=64):
    from partdistillation_b200.compat import BitMasks, Instances
    g = torch.Generator().manual_seed(seed)
    out = []
    for k in counts:
        lab = torch.randint(0, max(k, 1), (4, 4), generator=g).repeat_interleave(size // 4, 0).repeat_interleave(size // 4, 1)
        m = torch.stack([lab == i for i in range(k)]) if k else torch.zeros(0, size, size, dtype=torch.bool)
        inst = Instances((size, size))
        inst.gt_masks = BitMasks(m)
        inst.gt_classes = torch.randint(0, parts, (k,), generator=g) if pd else torch.zeros(k, dtype=torch.long)
        d = {"image": torch.randint(0, 256, (3, size, size), generator=g, dtype=torch.uint8), "instances": inst,
             "height": size, "width": size}
        if pd:
            d["gt_object_class"] = int(torch.randint(0, objects, (1,), generator=g))
        out.append(d)
    return out


@pytest.mark.parametrize("arch", ["ProposalModel", "PartDistillationModel"])
def test_bucketed_graph_replays_variable_target_counts(arch):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from partdistillation_b200 import compat, presets
    from partdistillation_b200.engine import DataParallelTrainer
    pd = arch == "PartDistillationModel"
    kw = dict(num_object_classes=5, num_part_classes=4) if pd else {}
    trainers = []
    for graph in (True, False):
        cfg = presets.make_cfg(arch, "swin_micro", num_queries=12, dec_layers=2, num_points=64, importance_sample_ratio=0.0,
                               device="cuda", **kw)
        cfg.MODEL.SWIN.DROP_PATH_RATE = 0.0       # stochastic depth draws from the global RNG: the two trainers must not diverge
        torch.manual_seed(5)
        model = compat.build_model(cfg).train()
        rand = _SharedRows()
        model.criterion.rand = model.criterion.matcher.rand = rand
        trainers.append(DataParallelTrainer(model, base_lr=1e-4, freeze_keys=("backbone", "encoder"), cuda_graph=graph,
                                            target_bucket=4 if graph else 0))
    ta, tb = trainers
    for (ka, pa), (kb, pb) in zip(ta.model.named_parameters(), tb.model.named_parameters()):
        assert ka == kb and torch.equal(pa, pb)
    plan = [(3, 2), (4, 1), (2, 4), (1, 3), (4, 4), (2, 2)]       # every image pads to 4 slots: one signature
    for step, counts in enumerate(plan):
        batch = _batch(100 + step, counts, pd)
        total_a, losses_a = ta.step(batch)
        total_b, losses_b = tb.step(batch)
        assert set(losses_a) == set(losses_b)
        for k in losses_a:
            a, b = float(losses_a[k].detach()), float(losses_b[k].detach())
            assert abs(a - b) <= 1e-3 * max(1.0, abs(b)), (step, k, a, b)
    assert len(ta._graphs) == 1
    entry = next(iter(ta._graphs.values()))
    assert "graph" in entry and ta.iteration == len(plan)         # two eager warm-ups, then capture + replays
    # the two models walked the same trajectory.  Adam moves every weight by ~lr per step whatever the gradient's size, so entries
    # whose gradient is rounding noise (atomics order) may differ by up to 2 * lr * steps; everything else agrees far tighter
    for (k, pa), pb in zip(ta.model.named_parameters(), tb.model.parameters()):
        d = (pa.detach() - pb.detach()).abs()
        assert float(d.max()) <= 2 * 1e-4 * len(plan) + 1e-6, k
        assert float((d > 1e-5).float().mean()) < 0.05, k
    # a count beyond the bucket opens a second signature
    ta.step(_batch(200, (5, 1), pd))
    assert len(ta._graphs) == 2
