"""Eval branches of ``ProposalModel`` and ``PartDistillationModel`` on bit-packed masks (reference
part_distillation/proposal_model.py:205-302 inference / _unique_assignment, :341-366 _prepare_gt_targets, :369-432
masking_with_object_mask / instance_inference / match_gt_labels; part_distillation_model.py:239-283 inference,
:319-394 masking / match_gt_labels / _unique_assignment_with_classes, :431-501 _prepare_gt_targets /
instance_inference_with_classification).

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


def _zero_pad(masks, h_pad, w_pad, device):
    masks = masks.to(device, non_blocking=True)
    padded = torch.zeros((masks.shape[0], h_pad, w_pad), dtype=masks.dtype, device=device)
    padded[:, :masks.shape[1], :masks.shape[2]] = masks
    return padded


def _sizes(images, inp, image_size):
    image_size = (int(image_size[0]), int(image_size[1]))
    out_size = (int(inp.get("height", image_size[0])), int(inp.get("width", image_size[1])))
    return tuple(int(v) for v in images.tensor.shape[-2:]), image_size, out_size


def _apply_filters(rows, fields, area, obj_area, min_ratio, min_score):
    """The two `if loc_valid_idxs.any()` filters shared by both meta-architectures: area ratio, then score (fields[0])."""
    valid = area / obj_area > min_ratio
    if valid.any():
        rows, fields = rows[valid], [f[valid] for f in fields]
    valid = fields[0] > min_score
    if valid.any():
        rows, fields = rows[valid], [f[valid] for f in fields]
    return rows, fields


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
            new_targets.append({"labels": parts.gt_classes.to(dev),
                                "masks": _zero_pad(parts.gt_masks.tensor, h_pad, w_pad, dev),
                                "object_masks": _zero_pad(objects.gt_masks.tensor, h_pad, w_pad, dev)})
        return new_targets

    def inference(self, batched_inputs, targets, images, outputs, vis=False):
        """-> list of {"proposals": Instances(pred_masks bool, pred_classes, scores), "gt_masks": Instances(gt_masks,
        gt_classes, pred_masks, pred_classes)} at each image's evaluation size (:220-254)."""
        mask_cls_results = outputs["pred_logits"]
        mask_pred_results = outputs["pred_masks"]                       # (B, Q, h, w) logits, never up-sampled densely
        processed_results = []
        for mask_cls, logits, target, inp, image_size in zip(mask_cls_results, mask_pred_results, targets,
                                                             batched_inputs, images.image_sizes):
            geometry = _sizes(images, inp, image_size)
            _, image_size, out_size = geometry
            target_masks = fn.resize_bool_masks(target["masks"], image_size, out_size)
            target_object_masks = fn.resize_bool_masks(target["object_masks"], image_size, out_size)
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
        rows, (scores,) = _apply_filters(rows, [scores], area, obj_area, self.minimum_pseudo_mask_ratio,
                                         self.minimum_pseudo_mask_score)
        return rows, scores, cand_bits

    def match_gt_labels(self, cand_bits, rows, scores, target_masks, target_labels):
        """Label every candidate with its best-IoU ground-truth part and drop those below the foreground threshold
        (:414-427; IoU = pycocotools rleIou semantics, utils/utils.py:35-42)."""
        ious = fn.bits_iou(cand_bits[rows], fn.pack_bits(target_masks))
        top1_ious, top1_idx = ious.topk(1, dim=1)
        top1_idx = top1_idx.flatten()
        fg = (top1_ious > self.iou_foreground_threshold).flatten()
        return rows[fg], scores[fg], target_labels[top1_idx[fg]]


class PartDistillationInferenceMixin:
    """Eval branch of ``PartDistillationModel`` (attributes: num_classes, test_topk_per_image, wandb_vis_topk,
    use_unique_per_pixel_label, min_pseudo_mask_score, min_pseudo_mask_ratio, fg_score_threshold, use_oracle_classifier,
    apply_masking_with_object_mask, mode, majority_vote_mapping)."""

    def _prepare_gt_targets(self, inputs, images):
        """(:431-453)"""
        h_pad, w_pad = images.tensor.shape[-2:]
        dev = self.device
        new_targets = []
        for x in inputs:
            parts, objects = x["part_instances"], x["instances"]
            new_targets.append({"labels": parts.gt_classes.to(dev),
                                "masks": _zero_pad(parts.gt_masks.tensor, h_pad, w_pad, dev),
                                "object_mask": _zero_pad(objects.gt_masks.tensor, h_pad, w_pad, dev),
                                "gt_object_class": objects.gt_classes.to(dev)})
        return new_targets

    def inference(self, batched_inputs, targets, images, outputs, vis=False):
        """-> list of {"predictions": Instances(pred_masks, scores, pred_classes), "gt_instances": Instances,
        "gt_object_label"} (:239-283).  With ``mode == "save"`` every image's prediction is also written to disk in the
        reference's record format (``save_part_segmentation``, :285-306)."""
        processed_results = []
        for mask_cls, logits, target, inp, image_size in zip(outputs["pred_logits"], outputs["pred_masks"], targets,
                                                             batched_inputs, images.image_sizes):
            geometry = _sizes(images, inp, image_size)
            _, image_size, out_size = geometry
            target_mask = fn.resize_bool_masks(target["masks"], image_size, out_size)
            target_object_mask = fn.resize_bool_masks(target["object_mask"], image_size, out_size)
            instance_r = self.instance_inference_with_classification(
                mask_cls.float(), logits.float(), target_mask, target_object_mask, target["labels"],
                target["gt_object_class"], vis=vis, geometry=geometry)
            if self.mode == "save" and not vis:
                self.save_part_segmentation(inp, instance_r)
            target_inst = Instances(out_size)
            target_inst.gt_masks = target_mask
            target_inst.gt_classes = target["labels"]
            target_inst.pred_masks = target_mask
            target_inst.pred_classes = target["labels"]
            processed_results.append({"predictions": instance_r, "gt_instances": target_inst,
                                      "gt_object_label": target["gt_object_class"]})
        return processed_results

    def _prepare_save_targets(self, inputs, images):
        """mode == "save": the pseudo labels of the training set stand in for the ground truth (prepare_targets, :397-402;
        _prepare_pseudo_targets, :405-428), with the object mask = union of the image's part masks."""
        h_pad, w_pad = images.tensor.shape[-2:]
        dev = self.device
        new_targets = []
        for x in inputs:
            inst = x["instances"]
            masks = _zero_pad(inst.gt_masks.tensor, h_pad, w_pad, dev)
            new_targets.append({"labels": inst.gt_classes.to(dev), "masks": masks,
                                "object_mask": masks.sum(dim=0, keepdim=True), "gt_object_class": x["gt_object_class"]})
        return new_targets

    def save_part_segmentation(self, input_per_image, instance):
        """The reference's on-disk record of one image's part prediction (:285-306); RLE strings from pycocotools, as in
        the reference (utils/utils.py:15-33)."""
        import os
        import numpy as np
        if instance is None:
            return
        root = getattr(self, "root_save_path", None)
        if root is None:
            raise RuntimeError("PartDistillationModel.save_part_segmentation: no root_save_path")
        from pycocotools import mask as mask_util
        masks = instance.pred_masks
        H, W = masks.shape[1:]
        object_area = int(masks.sum())
        part_areas = masks.flatten(1).sum(-1).long().cpu()
        rles = [mask_util.encode(np.asfortranarray(m.numpy()[:, :, None].astype(np.uint8)))[0] for m in masks.cpu()]
        for rle in rles:
            rle["counts"] = rle["counts"].decode("utf-8")
        res = {"file_name": input_per_image["file_name"], "image_id": input_per_image["image_id"],
               "class_code": input_per_image["class_code"], "height": H, "width": W,
               "part_masks": [{"segmentation": rle} for rle in rles], "part_labels": instance.pred_classes.cpu(),
               "part_area_ratios": part_areas / object_area, "object_ratio": object_area / (H * W),
               "part_scores": instance.scores.cpu().numpy()}
        folder = os.path.join(root, input_per_image["class_code"])
        os.makedirs(folder, exist_ok=True)
        torch.save(res, os.path.join(folder, input_per_image["image_id"]))

    def instance_inference_with_classification(self, mask_cls, mask_pred, target_mask, target_object_mask, target_labels,
                                               target_object_label, vis=False, geometry=None):
        """``mask_pred``: this image's (Q, h, w) logits; (query, class) pairs ranked jointly (:456-501)."""
        out_size = tuple(int(v) for v in target_mask.shape[-2:])
        if geometry is None:
            geometry = (out_size, out_size, out_size)
        padded, image_size, out_size = geometry
        per_pixel = self.use_unique_per_pixel_label
        topk = self.wandb_vis_topk if vis and not per_pixel else self.test_topk_per_image
        dev = mask_pred.device
        scores = mask_cls.softmax(-1)[:, :-1]
        labels = torch.arange(self.num_classes, device=dev).unsqueeze(0).repeat(mask_cls.shape[0], 1).flatten(0, 1)
        scores, topk_indices = scores.flatten(0, 1).topk(topk, sorted=False)
        labels = labels[topk_indices]
        if self.mode == "eval":
            labels = self.majority_vote_mapping[int(target_object_label)][labels]
        queries = torch.div(topk_indices, self.num_classes, rounding_mode="floor")
        gate = target_object_mask.any(dim=0) if self.apply_masking_with_object_mask else None      # (:319-326)
        sample = lambda **kw: fn.postprocess_masks(mask_pred, queries, padded, image_size, out_size, gate=gate,
                                                   scores=scores, **kw)
        rows, scores, labels, cand_bits = self._unique_assignment_with_classes(sample, scores, labels, out_size[1])
        rows, scores, labels, gt_part_labels = self.match_gt_labels(cand_bits, rows, scores, labels, target_mask,
                                                                    target_labels)
        if rows.numel() == 0:                                           # (:481-486)
            masks = torch.zeros((1, *out_size), dtype=torch.bool, device=dev)
            scores = scores.new_zeros(1)
            labels = scores.new_ones(1).long() * self.num_classes
            gt_part_labels = scores.new_ones(1).long() * self.num_classes
        else:
            masks = fn.unpack_bits(cand_bits, out_size[1], rows)
        result = Instances(out_size)
        result.pred_masks = masks
        result.scores = scores
        result.pred_classes = gt_part_labels if self.use_oracle_classifier else labels
        return result

    def _unique_assignment_with_classes(self, sample, scores, class_labels, width):
        """_unique_assignment_with_classes (:346-394) on packed masks; ``sample(**outputs)`` runs the post-processing
        kernel on this image's candidates.  Returns (rows into the candidate words, scores, class labels, words)."""
        if self.use_unique_per_pixel_label:
            bits, label = sample(want_bits=True, want_label=True)
            K = bits.shape[0] - 1
            obj = fn.unpack_bits(bits[K:], width)                       # (1, Ho, Wo)
            ids = label.unique().long()                                 # candidates that own at least one pixel
            new_labels, inverse = class_labels[ids].unique(return_inverse=True)
            # merge the segments of one class: a pixel belongs to class c iff its owner is labelled c (:363-367)
            pixel_class = class_labels[label.long()]
            cand_bits = fn.pack_bits((pixel_class[None] == new_labels[:, None, None]) & obj)
            new_scores = scores.new_zeros(new_labels.shape[0]).scatter_reduce(0, inverse, scores[ids], "amax",
                                                                              include_self=False)
            area, obj_area = fn.bits_popcount(cand_bits), fn.bits_popcount(bits[K:])
            rows = torch.arange(new_labels.shape[0], device=bits.device)
            rows, (new_scores, new_labels) = _apply_filters(rows, [new_scores, new_labels], area, obj_area,
                                                            self.min_pseudo_mask_ratio, self.min_pseudo_mask_score)
            return rows, new_scores, new_labels, cand_bits
        bits, _, above_half = sample(want_bits=True, score_threshold=0.5)
        K = bits.shape[0] - 1
        area, obj_area = fn.bits_popcount(above_half), fn.bits_popcount(bits[K:])
        rows = torch.arange(K, device=bits.device)
        cand_bits = bits[:K]
        valid = area / obj_area > self.min_pseudo_mask_ratio
        if valid.any():
            # the reference continues with `score * sigmoid(logit)` in place of the logits here (:385-387), so the
            # masks it returns are `score * sigmoid > 0`
            _, _, cand_bits = sample(want_bits=False, score_threshold=0.0)
            rows, scores, class_labels = rows[valid], scores[valid], class_labels[valid]
        valid = scores > self.min_pseudo_mask_score
        if valid.any():
            rows, scores, class_labels = rows[valid], scores[valid], class_labels[valid]
        return rows, scores, class_labels, cand_bits

    def match_gt_labels(self, cand_bits, rows, scores, class_labels, target_mask, target_labels):
        """(:329-343): keep the candidates whose best IoU with a ground-truth part exceeds ``fg_score_threshold``."""
        ious = fn.bits_iou(cand_bits[rows], fn.pack_bits(target_mask))
        top1_ious, top1_idx = ious.topk(1, dim=1)
        top1_idx = top1_idx.flatten()
        fg = (top1_ious > self.fg_score_threshold).flatten()
        return rows[fg], scores[fg], class_labels[fg], target_labels[top1_idx[fg]]
