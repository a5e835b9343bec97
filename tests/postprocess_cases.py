"""Shared by the CPU host-logic test and the GPU parity test of the ProposalModel eval branch: builds the batched inputs
of tests/golden/proposal_inference.pt and compares results with the reference's recorded outputs."""
import numpy as np
import torch
import torch.nn.functional as F

import m2f_oracle as O

NEAR = 1e-4     # |interpolated logit| below which a threshold bit may differ between two fp32 bilinear implementations


def oracle_resize(logits, padded, image_size, out_size):
    up = F.interpolate(logits[None], size=padded, mode="bilinear", align_corners=False)[0]
    return O.sem_seg_postprocess(up, image_size, *out_size)


def unpack_golden(packed, shape):
    return torch.from_numpy(np.unpackbits(packed, axis=-1)[..., :shape[-1]]).bool()


def canonical_order(scores, classes, masks):
    """The reference ranks candidates with ``topk(..., sorted=False)`` (proposal_model.py:389,
    part_distillation_model.py:466): the order of the surviving candidates is implementation-defined and differs
    between ATen's CPU kernel (where the goldens were recorded) and its CUDA kernel.  Both sides are therefore compared
    as sets: sorted by (score descending, class, mask area)."""
    s = scores.detach().cpu().double().numpy()
    c = classes.detach().cpu().numpy()
    a = masks.detach().cpu().flatten(1).sum(1).numpy()
    return torch.from_numpy(np.lexsort((a, c, -s)).copy())


def proposal_stub(overrides, topk, device):
    """The eval-branch methods of ProposalModel bound to a bare object (no backbone / head needed: the golden inputs are
    head outputs)."""
    from partdistillation_b200.postprocess import ProposalInferenceMixin

    class Stub(ProposalInferenceMixin):
        pass
    s = Stub()
    s.device = torch.device(device)
    s.test_topk_per_image = topk
    s.wandb_vis_topk = topk
    for k, v in overrides.items():
        setattr(s, k, v)
    return s


def run_case(golden, case, device):
    from partdistillation_b200.compat import BitMasks, ImageList, Instances
    inp, c = golden["inputs"], golden["cases"][case]
    Q = inp["pred_logits"].shape[1]
    model = proposal_stub(c["overrides"], Q, device)
    padded = inp["padded"]
    bi = []
    for it in inp["items"]:
        H, W = it["size"]
        inst = Instances((H, W))
        inst.gt_masks = BitMasks(it["object_mask"])
        inst.gt_classes = torch.zeros(1, dtype=torch.long)
        pinst = Instances((H, W))
        pinst.gt_masks = BitMasks(it["part_masks"])
        pinst.gt_classes = it["part_classes"]
        bi.append({"instances": inst, "part_instances": pinst, "height": it["out"][0], "width": it["out"][1]})
    images = ImageList(torch.zeros(len(bi), 3, *padded, device=device), [it["size"] for it in inp["items"]])
    outputs = {"pred_logits": inp["pred_logits"].to(device), "pred_masks": inp["pred_masks"].to(device)}
    targets = model._prepare_gt_targets(bi, images)
    res = model.inference(bi, targets, images, outputs)
    assert len(res) == len(c["results"])
    for b, (r, ref, it) in enumerate(zip(res, c["results"], inp["items"])):
        dense = oracle_resize(inp["pred_masks"][b], padded, it["size"], it["out"])
        prop = r["proposals"]
        assert tuple(prop.image_size) == ref["image_size"]
        pm = prop.pred_masks.cpu()
        ref_masks = unpack_golden(ref["pred_masks"], ref["pred_shape"])
        assert pm.dtype == torch.bool
        assert tuple(pm.shape) == tuple(ref_masks.shape)
        o, ro = canonical_order(prop.scores, prop.pred_classes, pm), canonical_order(ref["scores"], ref["pred_classes"], ref_masks)
        pm, ref_masks = pm[o], ref_masks[ro]
        assert torch.allclose(prop.scores.cpu()[o], ref["scores"][ro], rtol=1e-6, atol=1e-7)
        assert torch.equal(prop.pred_classes.cpu()[o], ref["pred_classes"][ro])
        near = (dense.abs() < NEAR).any(0)
        if c["overrides"]["use_unique_per_pixel_label"]:
            # per-pixel labels: a pixel may also change owner where the two best score * sigmoid values nearly tie
            scores = inp["pred_logits"][b].softmax(-1)[:, :-1].topk(1, dim=1)[0].flatten()
            gated = dense
            if c["overrides"]["apply_masking_with_object_mask"]:
                tom = O.sem_seg_postprocess(O.pad_masks(it["object_mask"], padded).float(), it["size"], *it["out"]).bool()
                gated = dense * tom.sum(0, keepdim=True).bool()
            top2 = (scores[:, None, None] * gated.sigmoid()).topk(2, dim=0)[0]
            near = near | ((top2[0] - top2[1]) < 1e-5)
        assert not ((pm != ref_masks) & ~near[None]).any()
        assert torch.equal(r["gt_masks"].gt_masks.cpu(), unpack_golden(ref["gt_masks"], ref["gt_shape"]))
        assert torch.equal(r["gt_masks"].gt_classes.cpu(), ref["gt_classes"])
    return res


PD_FLAGS = dict(use_unique_per_pixel_label="per_pixel", min_pseudo_mask_score="min_score", min_pseudo_mask_ratio="min_ratio",
                apply_masking_with_object_mask="gate", use_oracle_classifier="oracle_classifier")


def pd_stub(golden, case, device):
    """The eval-branch methods of PartDistillationModel bound to a bare object."""
    from partdistillation_b200.postprocess import PartDistillationInferenceMixin

    class Stub(PartDistillationInferenceMixin):
        pass
    inp, c = golden["inputs"], golden["cases"][case]
    s = Stub()
    s.device = torch.device(device)
    s.num_classes = inp["pred_logits"].shape[-1] - 1
    s.test_topk_per_image = inp["topk"]
    s.wandb_vis_topk = inp["topk"]
    s.fg_score_threshold = inp["fg_score_threshold"]
    s.mode = c["mode"]
    s.majority_vote_mapping = {k: v.to(device) for k, v in inp["majority_vote_mapping"].items()}
    for k, v in c["overrides"].items():
        setattr(s, k, v)
    return s


def run_pd_case(golden, case, device):
    """PartDistillationInferenceMixin.inference on the golden inputs against the reference's recorded outputs.  Masks
    may differ only at pixels within NEAR of a decision boundary of the oracle (threshold or owner near-tie)."""
    from partdistillation_b200.compat import BitMasks, ImageList, Instances
    inp, c = golden["inputs"], golden["cases"][case]
    model = pd_stub(golden, case, device)
    padded = inp["padded"]
    bi = []
    for it, oc in zip(inp["items"], inp["object_classes"]):
        H, W = it["size"]
        inst = Instances((H, W))
        inst.gt_masks = BitMasks(it["object_mask"])
        inst.gt_classes = torch.tensor([oc])
        pinst = Instances((H, W))
        pinst.gt_masks = BitMasks(it["part_masks"])
        pinst.gt_classes = it["part_classes"]
        bi.append({"instances": inst, "part_instances": pinst, "height": it["out"][0], "width": it["out"][1]})
    images = ImageList(torch.zeros(len(bi), 3, *padded, device=device), [it["size"] for it in inp["items"]])
    outputs = {"pred_logits": inp["pred_logits"].to(device), "pred_masks": inp["pred_masks"].to(device)}
    targets = model._prepare_gt_targets(bi, images)
    res = model.inference(bi, targets, images, outputs)
    assert len(res) == len(c["results"])
    for b, (r, ref, it) in enumerate(zip(res, c["results"], inp["items"])):
        dense = oracle_resize(inp["pred_masks"][b], padded, it["size"], it["out"])
        pred = r["predictions"]
        assert tuple(pred.image_size) == ref["image_size"]
        pm = pred.pred_masks.cpu()
        ref_masks = unpack_golden(ref["pred_masks"], ref["pred_shape"])
        assert pm.dtype == torch.bool and tuple(pm.shape) == tuple(ref_masks.shape)
        o, ro = canonical_order(pred.scores, pred.pred_classes, pm), canonical_order(ref["scores"], ref["pred_classes"], ref_masks)
        pm, ref_masks = pm[o], ref_masks[ro]
        assert torch.allclose(pred.scores.cpu()[o], ref["scores"][ro], rtol=1e-6, atol=1e-7)
        assert torch.equal(pred.pred_classes.cpu()[o], ref["pred_classes"][ro])
        near = (dense.abs() < NEAR).any(0)
        if c["overrides"]["use_unique_per_pixel_label"]:
            Q, P = inp["pred_logits"].shape[1], inp["pred_logits"].shape[2] - 1
            scores, idx = inp["pred_logits"][b].softmax(-1)[:, :-1].flatten().topk(inp["topk"], sorted=False)
            gated = dense[torch.div(idx, P, rounding_mode="floor")]
            if c["overrides"]["apply_masking_with_object_mask"]:
                tom = O.sem_seg_postprocess(O.pad_masks(it["object_mask"], padded).float(), it["size"], *it["out"]).bool()
                gated = gated * tom.sum(0, keepdim=True).bool()
            top2 = (scores[:, None, None] * gated.sigmoid()).topk(2, dim=0)[0]
            near = near | ((top2[0] - top2[1]) < 1e-5)
        assert not ((pm != ref_masks) & ~near[None]).any()
        assert torch.equal(r["gt_instances"].gt_masks.cpu(), unpack_golden(ref["gt_masks"], ref["gt_shape"]))
        assert torch.equal(r["gt_instances"].gt_classes.cpu(), ref["gt_classes"])
        assert torch.equal(torch.as_tensor(r["gt_object_label"]).cpu().flatten(), ref["gt_object_label"].flatten())
    return res
