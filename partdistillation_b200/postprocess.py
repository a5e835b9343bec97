"""Eval branch of ``ProposalModel`` on bit-packed masks (reference
part_distillation/proposal_model.py:205-302 inference / _unique_assignment, :341-366 _prepare_gt_targets, :369-432
masking_with_object_mask / instance_inference / match_gt_labels).

The reference up-samples every query's logits to the padded image size, resizes them again to the evaluation size,
gates, thresholds and ranks them as dense fp32 maps, then run-length encodes every bool mask on the host for the IoU
against the ground-truth parts.  Here ``functional.postprocess_masks`` composes both bilinear passes, the object-mask
gate, the ``> 0`` threshold, the top-1 object map and the ``score * sigmoid`` argmax in one kernel that writes one bit per
(query, pixel); areas and pairwise intersections are popcounts on those words (``functional.bits_popcount`` /
``bits_iou``), and only the masks that survive the filters are expanded to the ``bool (R, H, W)`` tensor the evaluators
read.  The small per-query tensors (softmax, top-k, filters, label lookup) stay torch ops on the device.
"""
import torch

from . import functional as fn
from .compat import Instances


class ProposalInferenceMixin:
    """Methods of the reference's eval branch; mixed into ``ProposalModel`` (attributes: test_topk_per_image,
    wandb_vis_topk, use_unique_per_pixel_label, minimum_pseudo_mask_score, minimum_pseudo_mask_ratio,
    apply_masking_with_object_mask)."""

    iou_foreground_threshold = 0.001          # match_gt_labels (:416)

    def _prepare_gt_targets(self, inputs, images):
        """Ground-truth parts (``part_instances``) and object masks (``instances``) zero-padded to the padded batch
        size (:341-366)."""
        h_pad, w_pad = images.tensor.shape[-2:]
        dev = self.device
        new_targets = []
        for x in inputs:
            parts, objects = x["part_instances"], x["instances"]
            gt = parts.gt_masks.tensor.to(dev, non_blocking=True)
            padded = torch.zeros((gt.shape[0], h_pad, w_pad), dtype=gt.dtype, device=dev)
            padded[:, :gt.shape[1], :gt.shape[2]] = gt
            go = objects.gt_masks.tensor.to(dev, non_blocking=True)
            padded_obj = torch.zeros((go.shape[0], h_pad, w_pad), dtype=go.dtype, device=dev)
            padded_obj[:, :go.shape[1], :go.shape[2]] = go
            new_targets.append({"labels": parts.gt_classes.to(dev), "masks": padded, "object_masks": padded_obj})
        return new_targets

    def inference(self, batched_inputs, targets, images, outputs, vis=False):
        """-> list of {"proposals": Instances(pred_masks bool, pred_classes, scores), "gt_masks": Instances(gt_masks,
        gt_classes, pred_masks, pred_classes)} at each image's evaluation size (:220-254)."""
        mask_cls_results = outputs["pred_logits"]
        mask_pred_results = outputs["pred_masks"]                       # (B, Q, h, w) logits, never up-sampled densely
        padded = tuple(int(v) for v in images.tensor.shape[-2:])
        processed_results = []
        for mask_cls, logits, target, inp, image_size in zip(mask_cls_results, mask_pred_results, targets,
                                                             batched_inputs, images.image_sizes):
            image_size = (int(image_size[0]), int(image_size[1]))
            out_size = (int(inp.get("height", image_size[0])), int(inp.get("width", image_size[1])))
            target_masks = fn.resize_bool_masks(target["masks"], image_size, out_size)
            target_object_masks = fn.resize_bool_masks(target["object_masks"], image_size, out_size)
            geometry = (padded, image_size, out_size)
            instance_r = self.instance_inference(mask_cls.float(), logits.float(), target_masks, target_object_masks,
                                                 target["labels"], vis=vis, geometry=geometry)
            target_inst = Instances(out_size)
            target_inst.gt_masks = target_masks
            target_inst.gt_classes = target["labels"]
            target_inst.pred_masks = target_masks                       # for visualisation, as in the reference
            target_inst.pred_classes = target["labels"]
            processed_results.append({"proposals": instance_r, "gt_masks": target_inst})
        return processed_results

    def instance_inference(self, mask_cls, mask_pred, target_masks, target_object_masks, target_labels, vis=False,
                           geometry=None):
        """``mask_pred``: this image's (Q, h, w) mask logits; ``geometry`` = (padded size, image size, output size)
        (default: one bilinear pass to the size of ``target_masks``).  Scores, filters and labels follow :381-411."""
        out_size = tuple(int(v) for v in target_masks.shape[-2:])
        if geometry is None:
            geometry = (out_size, out_size, out_size)
        padded, image_size, out_size = geometry
        per_pixel = self.use_unique_per_pixel_label
        topk = self.wandb_vis_topk if vis and not per_pixel else self.test_topk_per_image
        scores = mask_cls.softmax(-1)[:, :-1]
        scores = scores.topk(1, dim=1)[0].flatten()
        scores, topk_indices = scores.topk(topk, sorted=False)
        gate = target_object_masks.any(dim=0) if self.apply_masking_with_object_mask else None      # (:369-376)
        bits, label = fn.postprocess_masks(mask_pred, topk_indices, padded, image_size, out_size, gate=gate,
                                           scores=scores, want_bits=True, want_label=per_pixel)
        rows, scores, cand_bits = self._unique_assignment(bits, label, scores, out_size[1])
        rows, scores, labels = self.match_gt_labels(cand_bits, rows, scores, target_masks, target_labels)
        if rows.numel() == 0:                                           # contributes nothing to the evaluation (:402-406)
            masks = torch.zeros((1, *out_size), dtype=torch.bool, device=mask_pred.device)
            scores = scores.new_zeros(1)
            labels = labels.new_zeros(1)
        else:
            masks = fn.unpack_bits(cand_bits, out_size[1], rows)
        result = Instances(out_size)
        result.pred_masks = masks
        result.pred_classes = labels
        result.scores = scores
        return result

    def _unique_assignment(self, bits, label, scores, width):
        """_unique_assignment (:258-302) on packed masks.  ``bits`` (K + 1, Ho, Ww): the K thresholded candidates and
        their OR (the object map).  Returns (row indices into the candidate words, their scores, candidate words)."""
        K = bits.shape[0] - 1
        if self.use_unique_per_pixel_label:
            ids = label.unique()                                        # queries that own at least one pixel
            obj = fn.unpack_bits(bits[K:], width)                       # (1, Ho, Wo)
            cand_bits = fn.pack_bits((label[None] == ids[:, None, None]) & obj)
            scores = scores[ids.long()]
            area = fn.bits_popcount(cand_bits)
            obj_area = fn.bits_popcount(bits[K:])
        else:
            counts = fn.bits_popcount(bits)
            cand_bits, area, obj_area = bits[:K], counts[:K], counts[K:]
        rows = torch.arange(area.shape[0], device=bits.device)
        valid = area / obj_area > self.minimum_pseudo_mask_ratio
        if valid.any():
            rows, scores = rows[valid], scores[valid]
        valid = scores > self.minimum_pseudo_mask_score
        if valid.any():
            rows, scores = rows[valid], scores[valid]
        return rows, scores, cand_bits

    def match_gt_labels(self, cand_bits, rows, scores, target_masks, target_labels):
        """Label every candidate with its best-IoU ground-truth part and drop those below the foreground threshold
        (:414-427; IoU = pycocotools rleIou semantics, utils/utils.py:35-42)."""
        ious = fn.bits_iou(cand_bits[rows], fn.pack_bits(target_masks))
        top1_ious, top1_idx = ious.topk(1, dim=1)
        top1_idx = top1_idx.flatten()
        fg = (top1_ious > self.iou_foreground_threshold).flatten()
        return rows[fg], scores[fg], target_labels[top1_idx[fg]]
