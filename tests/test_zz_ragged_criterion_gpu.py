"""GPU: SetCriterion on batches whose images have very different numbers of targets — images with NONE, more targets than
queries (K > Q: SciPy's transposed orientation) — against the oracle's restatement of the reference on the same random point
draws (criterion.py:235-270, matcher.py:100-168).  Sorted last on purpose: written after round 1's GPU budget was spent; its
body already runs on the kernels' host builds in tests/test_head_host_cpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("counts", [(3, 0, 2), (20, 1)])
def test_criterion_ragged_batches_vs_oracle(counts):
    """SetCriterion on batches whose images have very different numbers of targets — including images with NONE, and
    more targets than queries (K > Q: SciPy's transposed orientation) — through the kernels' host builds, against the
    oracle's restatement of the reference on the same random point draws (criterion.py:235-270, matcher.py:100-168)."""
    import m2f_oracle as O
    import synth
    from partdistillation_b200.modeling.criterion import SetCriterion
    from partdistillation_b200.modeling.matcher import HungarianMatcher
    g = torch.Generator().manual_seed(11)
    B, Q, H, W, P = len(counts), 12, 16, 16, 48
    logits = torch.randn(B, Q, 2, generator=g)
    masks = torch.randn(B, Q, H, W, generator=g) * 2
    aux_l, aux_m = logits * 0.7 + 0.1, masks * 0.5 - 0.2
    targets = []
    for k in counts:
        m = torch.rand(k, 4 * H, 4 * W, generator=g) > 0.5
        targets.append({"labels": torch.zeros(k, dtype=torch.long), "masks": m})
    outputs = {"pred_logits": logits, "pred_masks": masks, "aux_outputs": [{"pred_logits": aux_l, "pred_masks": aux_m}]}
    rr = synth.RecordRand()
    ref = O.set_criterion(outputs, [{"labels": t["labels"], "masks": t["masks"].float()} for t in targets], 1, P, P,
                          importance_ratio=0.75, rand=rr)
    matcher = HungarianMatcher(cost_class=2.0, cost_mask=5.0, cost_dice=5.0, num_points=P)
    crit = SetCriterion(1, matcher=matcher, weight_dict={}, eos_coef=0.1, losses=["labels", "masks"], num_points=P,
                        oversample_ratio=3.0, importance_sample_ratio=0.75)
    replay = synth.ReplayRand(rr.draws, device=DEV)
    crit.rand = replay
    matcher.rand = replay
    lg, mk = logits.to(DEV).requires_grad_(), masks.to(DEV).requires_grad_()
    crit.to(DEV)
    dev_targets = [{k: v.to(DEV) for k, v in t.items()} for t in targets]
    got = crit({"pred_logits": lg, "pred_masks": mk,
                "aux_outputs": [{"pred_logits": lg * 0.7 + 0.1, "pred_masks": mk * 0.5 - 0.2}]}, dev_targets)
    assert replay.i == len(rr.draws)
    assert set(got) == set(ref)
    for k, v in ref.items():
        assert abs(float(got[k].detach()) - float(v)) <= 1e-4 * max(1.0, abs(float(v))), (k, float(got[k]), float(v))
    sum(got.values()).backward()
    assert torch.isfinite(lg.grad).all() and torch.isfinite(mk.grad).all()
    matcher.rand = torch.rand
    pairs = matcher({"pred_logits": logits.to(DEV), "pred_masks": masks.to(DEV)}, dev_targets)
    assert [len(i) for i, _ in pairs] == [min(Q, k) for k in counts]
