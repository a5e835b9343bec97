"""Diagnostic: per-parameter gradient error of the head vs the reference goldens, with the tcgen05 linear and with
the library (cuBLAS fp32) linear, to separate GEMM rounding from the inherent sensitivity of some gradients."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import synth
import test_head_gpu as T
from partdistillation_b200 import functional as PF

def run(name, use_tc):
    orig = PF.linear_supported
    if not use_tc:
        PF.linear_supported = lambda x, w: False
    g = torch.load(os.path.join(ROOT, "tests", "golden", f"head_{name}.pt"), weights_only=False)
    model, c = T._build(g)
    feats, targets = T._inputs(model, c)
    replay = synth.ReplayRand(g["rand_draws"], device="cuda")
    model.criterion.rand = replay
    model.criterion.matcher.rand = replay
    outputs = model.run_head(feats, targets)
    losses = model.criterion(outputs, targets)
    losses = {k: v * model.criterion.weight_dict[k] for k, v in losses.items()}
    pm = [o["pred_masks"] for o in outputs["aux_outputs"]] + [outputs["pred_masks"]]
    e_logit = max(float((a.cpu() - b).abs().max() / b.abs().max()) for a, b in zip(pm, g["pred_masks"]))
    e_loss = max(abs(float(losses[k]) - float(v)) / max(1.0, abs(float(v))) for k, v in g["losses"].items())
    sum(losses.values()).backward()
    named = dict(model.named_parameters())
    rows = sorted(((float((named[k].grad.cpu() - v).abs().max()) / float(v.abs().max()), k) for k, v in g["grads"].items()), reverse=True)
    print(f"{name} tc={use_tc}: logits {e_logit:.2e} losses {e_loss:.2e}; worst grads:", [(f"{e:.2e}", k.split('sem_seg_head.')[-1]) for e, k in rows[:4]])
    PF.linear_supported = orig

for name in ("proposal_micro", "proposal_micro_uniform", "pd_micro"):
    for tc in (False, True):
        run(name, tc)
