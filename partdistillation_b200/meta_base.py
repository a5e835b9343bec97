"""Shared training branch of the two meta-architectures
(proposal_model.py:177-204 / part_distillation_model.py:197-226 of the reference)."""
import logging
from typing import Tuple

import torch
from torch import nn

from .compat import ImageList, PackedBitMasks
from .functional import host_table
from .modeling.targets import TargetList


class Mask2FormerTrainingArch(nn.Module):
    """normalise + pad images -> backbone -> targets -> head -> criterion -> weight_dict scaling."""

    part_distillation = False

    target_padding = False          # set by the trainer's target bucketing: gt_classes == -1 marks a padding slot (engine.py)
    keep_packed_targets = True      # f3: bit-packed targets stay packed and are sampled from the words (False: expand to bytes)

    def _init_common(self, backbone, sem_seg_head, criterion, num_queries, num_classes, size_divisibility,
                     pixel_mean: Tuple[float], pixel_std: Tuple[float], test_topk_per_image, use_wandb):
        self.backbone = backbone
        self.sem_seg_head = sem_seg_head
        self.criterion = criterion
        self.num_queries = num_queries
        self.num_classes = num_classes
        if size_divisibility < 0:
            size_divisibility = self.backbone.size_divisibility
        self.size_divisibility = size_divisibility
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)
        self.test_topk_per_image = test_topk_per_image
        self.cpu_device = torch.device("cpu")
        self.logger = logging.getLogger("part_distillation")
        self.use_wandb = use_wandb
        self.metadata = None

    @property
    def device(self):
        return self.pixel_mean.device

    # pass-throughs the reference trainers call (part_proposal_train_net.py:99-109, part_distillation_train_net.py:107-118)
    def register_metadata(self, dataset_name):
        self.logger.info("%s is registered for evaluation.", dataset_name)
        self.metadata = dataset_name

    def preprocess_images(self, batched_inputs):
        images = [x["image"].to(self.device, non_blocking=True) for x in batched_inputs]
        images = [(x - self.pixel_mean) / self.pixel_std for x in images]
        return ImageList.from_tensors(images, self.size_divisibility)

    def prepare_targets(self, inputs, images):
        """Pseudo labels while training, ground-truth parts + object masks for evaluation (proposal_model.py:306-310,
        part_distillation_model.py:397-402; the eval variants live in postprocess.py)."""
        if not self.training:
            if getattr(self, "mode", "") == "save":
                return self._prepare_save_targets(inputs, images)
            return self._prepare_gt_targets(inputs, images)
        return self._prepare_pseudo_targets(inputs, images)

    def _prepare_pseudo_targets(self, inputs, images):
        """Zero-pad every image's pseudo masks to the padded batch size (proposal_model.py:313-338,
        part_distillation_model.py:405-428).  All images' masks live in ONE uint8 (Ktot, H, W) buffer;
        the per-image dicts of the reference format are views into it."""
        h_pad, w_pad = images.tensor.shape[-2:]
        dev = self.device
        inst = [x["instances"] for x in inputs]
        for i in inst:
            if not i.has("gt_masks"):
                raise ValueError("pseudo label without masks.")
        counts = [int(i.gt_masks.tensor.shape[0]) for i in inst]
        offs = [0]
        for c in counts:
            offs.append(offs[-1] + c)
        # f3: when every image's masks arrive bit-packed (and the padded width is a whole number of 32-pixel words, which
        # size_divisibility = 32 guarantees) the targets STAY packed — (Ktot, h_pad, w_pad / 32) int32 words, 1/8 of the bytes on
        # the wire and in HBM — and the matcher / criterion kernels sample the ground truth straight from the words.
        keep_bits = self.keep_packed_targets and w_pad % 32 == 0 and all(isinstance(i.gt_masks, PackedBitMasks) for i in inst) \
            and not self.use_wandb
        if keep_bits:
            packed = torch.zeros((offs[-1], h_pad, w_pad // 32), dtype=torch.int32, device=dev)
        else:
            packed = torch.zeros((offs[-1], h_pad, w_pad), dtype=torch.uint8, device=dev)
        out = TargetList()
        labels = []
        for b, (x, i) in enumerate(zip(inputs, inst)):
            view = packed[offs[b]:offs[b + 1]]
            if keep_bits:
                if counts[b]:
                    words = i.gt_masks.tensor.to(dev, non_blocking=True)      # bits beyond the image width are zero by construction
                    view[:, :words.shape[1], :words.shape[2]] = words
            elif isinstance(i.gt_masks, PackedBitMasks):    # 1 bit / pixel over the wire, expanded on the device
                if counts[b]:
                    from .functional import unpack_bits
                    m = unpack_bits(i.gt_masks.tensor.to(dev, non_blocking=True), i.gt_masks.width)
                    view[:, :m.shape[1], :m.shape[2]] = m
            else:
                m = i.gt_masks.tensor.to(dev, non_blocking=True)
                view[:, :m.shape[1], :m.shape[2]] = m
            if self.part_distillation:
                lab = i.gt_classes.to(dev, non_blocking=True).long()
                t = {"labels": lab, "masks": PackedBitMasks(view, w_pad) if keep_bits else view.view(torch.bool),
                     "gt_object_class": x["gt_object_class"]}
                if self.use_wandb:
                    t["object_mask"] = view.sum(dim=0, keepdim=True)
            else:
                lab = torch.zeros(counts[b], dtype=torch.long, device=dev)
                if self.target_padding:         # class-agnostic labels, but the padding slots (gt_classes == -1) keep their mark
                    lab = torch.where(i.gt_classes.to(dev, non_blocking=True) < 0, -1, lab)
                t = {"labels": lab, "masks": PackedBitMasks(view, w_pad) if keep_bits else view.view(torch.bool)}
                if self.use_wandb:
                    t["object_masks"] = view.sum(0, keepdim=True)
            labels.append(lab)
            out.append(t)
        out.offsets = offs
        out.packed_masks = packed
        out.packed_labels = torch.cat(labels).to(torch.int32) if offs[-1] else torch.zeros((0,), dtype=torch.int32, device=dev)
        out.has_dummies = bool(self.target_padding)
        if self.part_distillation:
            if all("gt_object_class_dev" in x for x in inputs):     # device scalars of a static (graph) batch: data, not constants
                out.object_classes = torch.stack([x["gt_object_class_dev"].reshape(()) for x in inputs]).to(torch.int32)
            else:
                out.object_classes = host_table([int(x["gt_object_class"]) for x in inputs], torch.int32, dev)
        return out

    def run_head(self, features, targets):
        return self.sem_seg_head(features, mask=targets) if self.part_distillation else self.sem_seg_head(features)

    def losses_from_features(self, features, targets):
        """The accelerated hot path proper: head + criterion + weight_dict scaling."""
        outputs = self.run_head(features, targets)
        losses = self.criterion(outputs, targets)
        for k in list(losses.keys()):
            if k in self.criterion.weight_dict:
                losses[k] = losses[k] * self.criterion.weight_dict[k]
            else:
                losses.pop(k)
        return losses

    def forward(self, batched_inputs):
        """Training branch; the registered subclasses route ``eval()`` calls to their inference mixins."""
        if not self.training:
            raise RuntimeError("Mask2FormerTrainingArch.forward is the training branch; call the registered meta-architecture")
        images = self.preprocess_images(batched_inputs)
        features = self.backbone(images.tensor)
        targets = self.prepare_targets(batched_inputs, images)
        return self.losses_from_features(features, targets)


def build_criterion(cfg, num_classes, match_points, loss_points):
    from .modeling.criterion import SetCriterion
    from .modeling.matcher import HungarianMatcher
    m = cfg.MODEL.MASK_FORMER
    matcher = HungarianMatcher(cost_class=m.CLASS_WEIGHT, cost_mask=m.MASK_WEIGHT, cost_dice=m.DICE_WEIGHT,
                               num_points=match_points)
    weight_dict = {"loss_ce": m.CLASS_WEIGHT, "loss_mask": m.MASK_WEIGHT, "loss_dice": m.DICE_WEIGHT}
    if m.DEEP_SUPERVISION:
        base = dict(weight_dict)
        for i in range(m.DEC_LAYERS - 1):
            weight_dict.update({f"{k}_{i}": v for k, v in base.items()})
    return SetCriterion(num_classes, matcher=matcher, weight_dict=weight_dict, eos_coef=m.NO_OBJECT_WEIGHT,
                        losses=["labels", "masks"], num_points=loss_points, oversample_ratio=m.OVERSAMPLE_RATIO,
                        importance_sample_ratio=m.IMPORTANCE_SAMPLE_RATIO)
