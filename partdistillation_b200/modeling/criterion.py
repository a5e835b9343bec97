"""Set criterion (reference: modeling/criterion.py:94-286): Hungarian matching per decoder output,
weighted cross-entropy on the classes, point-sampled BCE + dice on the masks.

Same constructor, ``forward(outputs, targets) -> dict`` keys and arithmetic.  Structural changes:
targets are packed once per step (targets.py), matching stays on the device (no ``.cpu()``, no SciPy),
``num_masks`` stays a device scalar (no ``.item()``), and the two point_sample calls + BCE + dice of
``loss_masks`` are one fused kernel pair.  ``torch.rand`` is called with the reference's shapes in
the reference's order (detectron2 get_uncertain_point_coords_with_randomness: two draws).
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from .. import functional as PF
from .targets import pack_targets


def calculate_uncertainty(logits):
    """-(|logit|) of the foreground class (criterion.py:77-91)."""
    assert logits.shape[1] == 1
    return -(torch.abs(logits.clone()))


class SetCriterion(nn.Module):
    def __init__(self, num_classes, matcher, weight_dict, eos_coef, losses, num_points, oversample_ratio,
                 importance_sample_ratio):
        super().__init__()
        self.num_classes = num_classes
        self.matcher = matcher
        self.weight_dict = weight_dict
        self.eos_coef = eos_coef
        self.losses = losses
        empty_weight = torch.ones(self.num_classes + 1)
        empty_weight[-1] = self.eos_coef
        self.register_buffer("empty_weight", empty_weight)
        self.num_points = num_points
        self.oversample_ratio = oversample_ratio
        self.importance_sample_ratio = importance_sample_ratio
        self.rand = torch.rand
        self.external_num_masks = None     # device float tensor set by the trainer (see forward)
        self.batched_matching = True       # the Hungarian assignments of all decoder outputs in one cost + one LSAP launch

    # ------------------------------------------------------------------------------------------
    def _flat_indices(self, targets, match, B, Q):
        """Flat (b*Q + q) prediction index and (offset_b + k) target index of every VALID pair, in
        (image, ascending cost) order — the order the reference concatenates them in (:209-225)."""
        pi, ti = match
        offs = targets.offsets
        dev = pi.device
        cache = getattr(targets, "_crit_cache", None)
        if cache is None or cache["Q"] != Q:
            base_p, base_t, keep = [], [], []
            need_filter = False
            for b in range(B):
                k = offs[b + 1] - offs[b]
                n = min(Q, k)
                need_filter |= n < k
                base_p += [b * Q] * k
                base_t += [offs[b]] * k
                keep += list(range(offs[b], offs[b] + n))
            cache = dict(Q=Q, base_p=PF.host_table(base_p, torch.int64, dev),
                         base_t=PF.host_table(base_t, torch.int64, dev),
                         keep=PF.host_table(keep, torch.int64, dev) if need_filter else None)
            targets._crit_cache = cache
        src = pi + cache["base_p"]
        tgt = ti + cache["base_t"]
        if cache["keep"] is not None:
            src, tgt = src[cache["keep"]], tgt[cache["keep"]]
        return src, tgt

    def loss_labels(self, outputs, targets, match, num_masks):
        """Weighted cross-entropy; unmatched queries are the no-object class (:126-145)."""
        logits = outputs["pred_logits"].float()
        B, Q, _ = logits.shape
        src, tgt = self._flat_indices(targets, match, B, Q)
        target_classes = torch.full((B * Q,), self.num_classes, dtype=torch.int64, device=logits.device)
        lab = targets.packed_labels[tgt].long()
        if getattr(targets, "has_dummies", False):          # padding slots (label -1) are "no object", like unmatched queries
            lab = torch.where(lab < 0, self.num_classes, lab)
        target_classes[src] = lab
        loss_ce = F.cross_entropy(logits.view(B * Q, -1), target_classes, self.empty_weight)
        return {"loss_ce": loss_ce}

    def loss_masks(self, outputs, targets, match, num_masks, loss_points=None):
        """Point-sampled sigmoid-CE + dice on the matched masks (:147-207).  ``loss_points``: the random draws of this call made
        ahead of time by ``forward`` (_draw_loss_points)."""
        pred = outputs["pred_masks"]
        B, Q, H, W = pred.shape
        src, tgt = self._flat_indices(targets, match, B, Q)
        Nm = src.shape[0]
        flat = pred.float().flatten(0, 1)
        with torch.no_grad():
            coords = self._uncertain_point_coords(flat, src, Nm, loss_points)
        bce, dice = PF.point_loss(flat, src, targets.packed_masks, tgt, coords)
        if getattr(targets, "has_dummies", False):          # pairs matched to padding slots carry no loss
            w = (targets.packed_labels[tgt] >= 0).to(bce.dtype)
            bce, dice = bce * w, dice * w
        return {"loss_mask": bce.sum() / num_masks, "loss_dice": dice.sum() / num_masks}

    def _draw_loss_points(self, Nm, dev):
        """The random numbers of one get_uncertain_point_coords_with_randomness call, in its order: the over-sampled candidates,
        then the uniformly random remainder (None when importance sampling takes every point)."""
        n_over = int(self.num_points * self.oversample_ratio)
        n_rand = self.num_points - int(self.importance_sample_ratio * self.num_points)
        over = self.rand(Nm, n_over, 2, device=dev, dtype=torch.float32)
        return over, (self.rand(Nm, n_rand, 2, device=dev) if n_rand > 0 else None)

    def _uncertain_point_coords(self, flat, src, Nm, drawn=None):
        """detectron2 get_uncertain_point_coords_with_randomness with uncertainty = -|logit|."""
        n_unc = int(self.importance_sample_ratio * self.num_points)
        over, rnd = drawn if drawn is not None else self._draw_loss_points(Nm, flat.device)
        if n_unc > 0:
            unc = -PF.point_sample(flat.detach(), over, src.to(torch.int32), None).abs()
            idx = torch.topk(unc, k=n_unc, dim=1)[1]
            coords = torch.gather(over, 1, idx[:, :, None].expand(-1, -1, 2))
            if rnd is not None:
                coords = torch.cat([coords, rnd], dim=1)
        else:
            coords = rnd            # every point is a uniformly random one: nothing to concatenate
        return coords.contiguous()

    def get_loss(self, loss, outputs, targets, indices, num_masks, **kw):
        loss_map = {"labels": self.loss_labels, "masks": self.loss_masks}
        assert loss in loss_map, f"do you really want to compute {loss} loss?"
        return loss_map[loss](outputs, targets, indices, num_masks, **kw)

    def _num_masks(self, targets, dev):
        """Average number of target masks across ranks, >= 1 (:248-254) — kept on the device.  It depends only on the target
        counts, so a trainer may all-reduce it before the step and hand it over (external_num_masks): the forward / backward then
        contains no collective and can be replayed as a CUDA graph on every rank.  With padded targets (the trainer's target
        bucketing) the real count is data, not shape: it is counted on the device from the labels (-1 = padding slot)."""
        if self.external_num_masks is not None:
            return self.external_num_masks.reshape(-1)[0]
        if getattr(targets, "has_dummies", False):
            num_masks = (targets.packed_labels >= 0).sum().float().reshape(1)
        else:
            num_masks = torch.full((1,), float(targets.total), dtype=torch.float, device=dev)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(num_masks)
            num_masks = num_masks / dist.get_world_size()
        return torch.clamp(num_masks, min=1)[0]

    def forward(self, outputs, targets):
        targets = pack_targets(targets)
        num_masks = self._num_masks(targets, outputs["pred_masks"].device)

        losses = {}
        main = {k: v for k, v in outputs.items() if k != "aux_outputs"}
        outs = [main] + list(outputs.get("aux_outputs", []))
        points = [None] * len(outs)
        if self.batched_matching and len(outs) > 1 and hasattr(self.matcher, "match_all"):
            # Every random number of the step first, in the reference's order (per output: the matcher's coordinates, then the
            # loss points; none of the shapes depends on a matching result), then the assignments of ALL outputs in one cost
            # launch and one LSAP launch (matcher.match_all), then the losses.
            B, Q = main["pred_logits"].shape[:2]
            dev = main["pred_masks"].device
            Nm = sum(min(Q, targets.offsets[b + 1] - targets.offsets[b]) for b in range(B))
            coords = []
            for i in range(len(outs)):
                coords.append(self.matcher.draw_coords(B, dev))
                if "masks" in self.losses:
                    points[i] = self._draw_loss_points(Nm, dev)
            matches = self.matcher.match_all(outs, targets, coords)
        else:
            matches = None
        for i, out in enumerate(outs):
            match = matches[i] if matches is not None else self.matcher.match_packed(out, targets)
            suffix = "" if i == 0 else f"_{i - 1}"
            for loss in self.losses:
                kw = {"loss_points": points[i]} if (loss == "masks" and points[i] is not None) else {}
                losses.update({k + suffix: v for k, v in self.get_loss(loss, out, targets, match, num_masks, **kw).items()})
        return losses

    def __repr__(self):
        body = [f"matcher: {self.matcher.__repr__(_repr_indent=8)}", f"losses: {self.losses}",
                f"weight_dict: {self.weight_dict}", f"num_classes: {self.num_classes}", f"eos_coef: {self.eos_coef}",
                f"num_points: {self.num_points}", f"oversample_ratio: {self.oversample_ratio}",
                f"importance_sample_ratio: {self.importance_sample_ratio}"]
        return "\n".join(["Criterion " + self.__class__.__name__] + [" " * 4 + line for line in body])
