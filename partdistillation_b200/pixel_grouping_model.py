"""``PixelGroupingModel`` — evaluation-only pixel grouping baseline (reference:
part_distillation/pixel_grouping_model.py:28-218; BASELINE.json configs[3]).

Same registry name, ``from_config`` keys and output format (list of {"proposals", "gt_masks"}).  The k-means step
stays scikit-learn on the host, as in the reference (:183-193).  What changes is ``generate_part_segments``: the
reference up-samples the (C, h, w) feature map to image size, copies the masked pixels to the host and runs
``measure_distance`` + ``topk`` there (:205-218); here one sm_100a kernel interpolates the backbone-resolution
features on the fly and writes the label map (``functional.group_affinity``), so the full-resolution feature map
is never materialised and nothing but the k-means sample leaves the device."""
from typing import List, Tuple

import torch
from torch import nn
from torch.nn import functional as F

from . import functional as PF
from .compat import META_ARCH_REGISTRY, ImageList, Instances, build_backbone, configurable


@META_ARCH_REGISTRY.register()
class PixelGroupingModel(nn.Module):
    @configurable
    def __init__(self, backbone, size_divisibility: int, pixel_mean: Tuple[float], pixel_std: Tuple[float],
                 distance_metric: str = "l2", backbone_feature_key_list: List[str] = ("res4",),
                 num_superpixel_clusters: int = 4, feature_normalize: bool = False, debug: bool = False,
                 object_mask_type: str = "detic_based", wandb_vis_period: int = 100):
        super().__init__()
        self.backbone = backbone
        if size_divisibility < 0:
            size_divisibility = self.backbone.size_divisibility
        self.size_divisibility = size_divisibility
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)
        self.distance_metric = distance_metric
        self.backbone_feature_key_list = list(backbone_feature_key_list)
        self.num_superpixel_clusters = num_superpixel_clusters
        self.feature_normalize = feature_normalize
        self.wandb_vis_period = wandb_vis_period
        self.num_test_iterations = 0
        self._kmeans = None

    @classmethod
    def from_config(cls, cfg):
        g = cfg.PIXEL_GROUPING
        return dict(backbone=build_backbone(cfg), size_divisibility=cfg.MODEL.MASK_FORMER.SIZE_DIVISIBILITY,
                    pixel_mean=cfg.MODEL.PIXEL_MEAN, pixel_std=cfg.MODEL.PIXEL_STD,
                    distance_metric=g.DISTANCE_METRIC, backbone_feature_key_list=g.BACKBONE_FEATURE_KEY_LIST,
                    num_superpixel_clusters=g.NUM_SUPERPIXEL_CLUSTERS, feature_normalize=g.FEATURE_NORMALIZE,
                    wandb_vis_period=cfg.WANDB.VIS_PERIOD_TEST, debug=g.DEBUG)

    @property
    def device(self):
        return self.pixel_mean.device

    def _prepare_features(self, features):
        """Listed backbone maps, bilinearly aligned to the first one and concatenated (:114-125)."""
        maps = [features[k] for k in self.backbone_feature_key_list]
        size = maps[0].shape[-2:]
        out = torch.cat([F.interpolate(v, size=size, mode="bilinear", align_corners=False) for v in maps], 1)
        return F.normalize(out, dim=1, p=2) if self.feature_normalize else out

    def get_pixel_grouping(self, feature_per_image, pred_mask):
        """k-means centroids of the object's backbone-resolution pixels (scikit-learn, host; :183-193)."""
        data = feature_per_image[:, pred_mask].transpose(0, 1).contiguous().cpu()
        if len(data) > self.num_superpixel_clusters:
            if self._kmeans is None:
                from sklearn.cluster import KMeans
                self._kmeans = KMeans(n_clusters=self.num_superpixel_clusters, random_state=0)
            return torch.tensor(self._kmeans.fit(data).cluster_centers_).float()
        return data.new_zeros(1, feature_per_image.shape[0])

    def generate_part_segments(self, feature_per_image, object_mask, object_mask_resized, centroids=None, geometry=None):
        """feature (C, h, w); object mask at feature and at image resolution -> bool (P, H, W), one mask per
        label present (ascending label order, as :213-216).  ``geometry`` = (padded, image, output) sizes when the
        evaluation size differs from the padded batch size (:146-158)."""
        if centroids is None:
            centroids = self.get_pixel_grouping(feature_per_image, object_mask)
        labels = PF.group_affinity(feature_per_image, centroids.to(feature_per_image.device), object_mask_resized,
                                   self.distance_metric, geometry=geometry)
        present = torch.unique(labels[labels > 0])
        return labels.unsqueeze(0) == present.view(-1, 1, 1)

    def forward(self, batched_inputs):
        assert not self.training, "pixel grouping is eval only."
        return [r for r, _ in self._group(batched_inputs)]

    def _group(self, batched_inputs):
        """-> per image (result dict or None, object mask at the evaluation size)."""
        with torch.no_grad():
            images = [(x["image"].to(self.device) - self.pixel_mean) / self.pixel_std for x in batched_inputs]
            images = ImageList.from_tensors(images, self.size_divisibility)
            h_pad, w_pad = images.tensor.shape[-2:]
            features = self._prepare_features(self.backbone(images.tensor))
            out = []
            for x, feat, image_size in zip(batched_inputs, features, images.image_sizes):
                obj = x["instances"].gt_masks.tensor.to(self.device)
                masks = torch.zeros((obj.shape[0], h_pad, w_pad), dtype=obj.dtype, device=self.device)
                masks[:, :obj.shape[1], :obj.shape[2]] = obj
                image_size = (int(image_size[0]), int(image_size[1]))
                out_size = (int(x.get("height", image_size[0])), int(x.get("width", image_size[1])))
                geometry = ((int(h_pad), int(w_pad)), image_size, out_size)
                same = geometry[0] == image_size == out_size       # nothing to crop or resize
                # sem_seg_postprocess(masks, image_size, height, width)[0].bool()  (:158)
                mask_resized = masks[0].bool() if same else PF.resize_bool_masks(masks[:1], image_size, out_size)[0]
                mask_feat = F.interpolate(masks[None].float(), size=feat.shape[-2:], mode="nearest")[0, 0].bool()
                pseudo = self.generate_part_segments(feat, mask_feat, mask_resized, geometry=None if same else geometry)
                if pseudo is None:              # ProposalGenerationModel: object too small to cluster
                    out.append((None, mask_resized))
                    continue
                inst = Instances(tuple(pseudo.shape[-2:]))
                inst.pred_masks = pseudo
                inst.scores = pseudo.new_ones(pseudo.shape[0])
                res = {"proposals": inst}
                if "part_instances" in x:
                    part = x["part_instances"].gt_masks.tensor.to(self.device)
                    gt = Instances(tuple(pseudo.shape[-2:]))
                    if same:
                        gt.gt_masks = part.bool()
                    else:
                        part_padded = torch.zeros((part.shape[0], h_pad, w_pad), dtype=part.dtype, device=self.device)
                        part_padded[:, :part.shape[1], :part.shape[2]] = part
                        gt.gt_masks = PF.resize_bool_masks(part_padded, image_size, out_size)       # (:156)
                    gt.pred_masks = gt.gt_masks
                    res["gt_masks"] = gt
                out.append((res, mask_resized))
            self.num_test_iterations += 1
            return out


@META_ARCH_REGISTRY.register()
class ProposalGenerationModel(PixelGroupingModel):
    """Pseudo-label generation for proposal learning (reference: part_distillation/proposal_generation_model.py:28-239):
    the same pixel grouping on the object mask of every image, written to disk as COCO-RLE part masks
    (``save_predictions``, :185-199) instead of being evaluated.  Images whose object covers no more backbone pixels than
    there are clusters yield nothing (``_get_superpixels`` returns None, :202-211).  Like the reference, ``forward``
    returns None; ``generate`` returns the per-image results (None for skipped images) for callers that want them."""

    @configurable
    def __init__(self, backbone, size_divisibility: int, dataset_name: str, pixel_mean: Tuple[float],
                 pixel_std: Tuple[float], distance_metric: str = "l2", backbone_feature_key_list: List[str] = ("res4",),
                 num_superpixel_clusters: int = 4, feature_normalize: bool = False, wandb_vis_period: int = 100,
                 debug: bool = False, root_save_path: str = None):
        super().__init__(backbone=backbone, size_divisibility=size_divisibility, pixel_mean=pixel_mean, pixel_std=pixel_std,
                         distance_metric=distance_metric, backbone_feature_key_list=backbone_feature_key_list,
                         num_superpixel_clusters=num_superpixel_clusters, feature_normalize=feature_normalize, debug=debug,
                         wandb_vis_period=wandb_vis_period)
        self.dataset_name = dataset_name
        self.root_save_path = root_save_path

    @classmethod
    def from_config(cls, cfg):
        g = cfg.PROPOSAL_GENERATION
        save_path = None
        try:        # the reference reads MetadataCatalog.get(dataset_name).save_path (:64); needs a real detectron2
            from detectron2.data import MetadataCatalog
            save_path = getattr(MetadataCatalog.get(g.DATASET_NAME), "save_path", None)
        except Exception:
            pass
        return dict(backbone=build_backbone(cfg), size_divisibility=cfg.MODEL.MASK_FORMER.SIZE_DIVISIBILITY,
                    dataset_name=g.DATASET_NAME, pixel_mean=cfg.MODEL.PIXEL_MEAN, pixel_std=cfg.MODEL.PIXEL_STD,
                    distance_metric=g.DISTANCE_METRIC, backbone_feature_key_list=g.BACKBONE_FEATURE_KEY_LIST,
                    num_superpixel_clusters=g.NUM_SUPERPIXEL_CLUSTERS, feature_normalize=g.FEATURE_NORMALIZE,
                    wandb_vis_period=cfg.WANDB.VIS_PERIOD_TEST, debug=g.DEBUG, root_save_path=save_path)

    def get_pixel_grouping(self, feature_per_image, pred_mask):
        """_get_superpixels (:202-211): None when the object has too few backbone pixels to cluster."""
        if int(pred_mask.sum()) <= self.num_superpixel_clusters:
            return None
        return super().get_pixel_grouping(feature_per_image, pred_mask)

    def generate_part_segments(self, feature_per_image, object_mask, object_mask_resized, centroids=None, geometry=None):
        """generate_pseudo_labels (:221-237)."""
        if centroids is None:
            centroids = self.get_pixel_grouping(feature_per_image, object_mask)
            if centroids is None:
                return None
        return super().generate_part_segments(feature_per_image, object_mask, object_mask_resized, centroids, geometry)

    def save_predictions(self, input_per_image, pseudo_label, object_mask):
        """The reference's on-disk record (:185-199); the RLE strings come from pycocotools, as in the reference
        (utils/utils.py:15-33)."""
        import os
        import numpy as np
        if self.root_save_path is None:
            raise RuntimeError("ProposalGenerationModel.save_predictions: no root_save_path (the reference takes it from "
                               "MetadataCatalog.get(dataset_name).save_path)")
        from pycocotools import mask as mask_util
        H, W = object_mask.shape[-2:]
        rles = [mask_util.encode(np.asfortranarray(m.numpy()[:, :, None].astype(np.uint8)))[0] for m in pseudo_label.cpu()]
        for rle in rles:
            rle["counts"] = rle["counts"].decode("utf-8")
        res = {"file_name": input_per_image["file_name"], "file_path": input_per_image["file_path"],
               "class_code": input_per_image["class_code"], "class_name": input_per_image["class_name"],
               "part_mask": [{"segmentation": rle} for rle in rles],
               "object_ratio": int(object_mask.sum()) / (H * W), "height": H, "width": W,
               "class_index": input_per_image["gt_object_class"]}
        folder = os.path.join(self.root_save_path, input_per_image["class_code"])
        os.makedirs(folder, exist_ok=True)
        torch.save(res, os.path.join(folder, input_per_image["file_name"]))

    def generate(self, batched_inputs, save=True):
        results = []
        for x, (r, object_mask) in zip(batched_inputs, self._group(batched_inputs)):
            if r is not None and save:
                self.save_predictions(x, r["proposals"].pred_masks, object_mask)
            results.append(r)
        return results

    def forward(self, batched_inputs):
        assert not self.training, "proposal generation is eval-only."
        self.generate(batched_inputs, save=True)
