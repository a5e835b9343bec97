"""Packed training targets.

The reference hands the criterion a list of per-image dicts whose "masks" are zero-padded float/bool
(K, H, W) tensors and re-casts / re-pads / re-samples them for every decoder layer
(criterion.py:165-169, matcher.py:120-134).  ``TargetList`` is that same list plus one packed copy
made once per step: all masks as uint8 (Ktot, H, W), labels as int32, host-side per-image offsets.
"""
import torch

from ..functional import host_table


class TargetList(list):
    """list of {"labels", "masks"[, "gt_object_class"]} dicts (reference format) + packed views."""

    packed_masks = None      # (Ktot, H, W) uint8
    packed_labels = None     # (Ktot,) int32
    offsets = None           # python list, len B+1
    object_classes = None    # (B,) int32 or None
    has_dummies = False      # packed_labels may hold -1: padding slots added by the trainer's target bucketing (engine.py)

    @property
    def total(self):
        return self.offsets[-1]


def pack_targets(targets):
    """Accepts a TargetList (returned as is) or the reference's plain list of dicts."""
    if isinstance(targets, TargetList) and targets.packed_masks is not None:
        return targets
    out = TargetList(targets)
    counts = [int(t["masks"].shape[0]) for t in targets]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    out.offsets = offs
    masks = [t["masks"] for t in targets]
    H = max(m.shape[-2] for m in masks)
    W = max(m.shape[-1] for m in masks)
    dev = masks[0].device
    packed = torch.zeros((offs[-1], H, W), dtype=torch.uint8, device=dev)
    for m, o in zip(masks, offs):
        if m.shape[0]:
            packed[o:o + m.shape[0], :m.shape[-2], :m.shape[-1]] = (m != 0)
    out.packed_masks = packed
    out.packed_labels = torch.cat([t["labels"] for t in targets]).to(torch.int32) if offs[-1] else \
        torch.zeros((0,), dtype=torch.int32, device=dev)
    if len(targets) and "gt_object_class" in targets[0]:
        out.object_classes = host_table([int(t["gt_object_class"]) for t in targets], torch.int32, dev)
    return out
