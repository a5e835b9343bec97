"""``ProposalModel`` meta-architecture — training branch (reference: part_distillation/proposal_model.py:30-204,313-338)
and eval branch (:205-302,341-432; ``postprocess.ProposalInferenceMixin``).

Registered under the reference's name so ``cfg.MODEL.META_ARCHITECTURE = "ProposalModel"`` resolves here;
same ``from_config`` keys and ``forward(batched_inputs) -> dict of weighted losses``."""
from typing import Tuple

from torch import nn

from .compat import META_ARCH_REGISTRY, build_backbone, build_sem_seg_head, configurable
from .meta_base import Mask2FormerTrainingArch, build_criterion
from .postprocess import ProposalInferenceMixin


@META_ARCH_REGISTRY.register()
class ProposalModel(ProposalInferenceMixin, Mask2FormerTrainingArch):
    @configurable
    def __init__(self, *, backbone, sem_seg_head: nn.Module, criterion: nn.Module, num_queries: int, num_classes: int,
                 size_divisibility: int, pixel_mean: Tuple[float], pixel_std: Tuple[float], test_topk_per_image: int,
                 dataset_name: str = "", use_wandb: bool = True, wandb_vis_period_train: int = 200,
                 wandb_vis_period_test: int = 20, wandb_vis_topk: int = 200, use_unique_per_pixel_label: bool = False,
                 minimum_pseudo_mask_score: float = 0.0, minimum_pseudo_mask_ratio: float = 0.0,
                 apply_masking_with_object_mask: bool = True):
        super().__init__()
        self._init_common(backbone, sem_seg_head, criterion, num_queries, num_classes, size_divisibility, pixel_mean,
                          pixel_std, test_topk_per_image, use_wandb)
        self.dataset_name = dataset_name
        self.wandb_vis_period_train = wandb_vis_period_train
        self.wandb_vis_period_test = wandb_vis_period_test
        self.wandb_vis_topk = wandb_vis_topk
        self.num_train_iterations = 0
        self.num_test_iterations = 0
        self.use_unique_per_pixel_label = use_unique_per_pixel_label
        self.minimum_pseudo_mask_score = minimum_pseudo_mask_score
        self.minimum_pseudo_mask_ratio = minimum_pseudo_mask_ratio
        self.apply_masking_with_object_mask = apply_masking_with_object_mask

    def set_postprocess_type(self, postprocess_type):
        if postprocess_type == "semseg":
            self.use_unique_per_pixel_label = True
        elif postprocess_type in ("prop", "prop-filtered"):
            self.use_unique_per_pixel_label = False
            if postprocess_type == "prop-filtered":
                self.minimum_pseudo_mask_score = 0.3

    def reset_postprocess_type(self, flag, score_thres):
        self.use_unique_per_pixel_label = flag
        self.minimum_pseudo_mask_score = score_thres

    @classmethod
    def from_config(cls, cfg):
        backbone = build_backbone(cfg)
        sem_seg_head = build_sem_seg_head(cfg, backbone.output_shape())
        m = cfg.MODEL.MASK_FORMER
        criterion = build_criterion(cfg, sem_seg_head.num_classes, m.TRAIN_NUM_POINTS, m.TRAIN_NUM_POINTS)
        p = cfg.PROPOSAL_LEARNING
        return dict(backbone=backbone, sem_seg_head=sem_seg_head, criterion=criterion,
                    num_queries=m.NUM_OBJECT_QUERIES, size_divisibility=m.SIZE_DIVISIBILITY,
                    pixel_mean=cfg.MODEL.PIXEL_MEAN, pixel_std=cfg.MODEL.PIXEL_STD,
                    num_classes=cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES,
                    wandb_vis_period_train=cfg.WANDB.VIS_PERIOD_TRAIN, wandb_vis_period_test=cfg.WANDB.VIS_PERIOD_TEST,
                    wandb_vis_topk=cfg.WANDB.VIS_TOPK, use_wandb=not cfg.WANDB.DISABLE_WANDB,
                    dataset_name=cfg.DATASETS.TRAIN[0], test_topk_per_image=cfg.TEST.DETECTIONS_PER_IMAGE,
                    use_unique_per_pixel_label=p.USE_PER_PIXEL_LABEL,
                    apply_masking_with_object_mask=p.APPLY_MASKING_WITH_OBJECT_MASK,
                    minimum_pseudo_mask_ratio=p.MIN_AREA_RATIO, minimum_pseudo_mask_score=p.MIN_SCORE)

    def forward(self, batched_inputs):
        if self.training:
            losses = super().forward(batched_inputs)
            self.num_train_iterations += 1
            return losses
        images = self.preprocess_images(batched_inputs)
        features = self.backbone(images.tensor)
        targets = self.prepare_targets(batched_inputs, images)
        outputs = self.run_head(features, targets)
        processed_results = self.inference(batched_inputs, targets, images, outputs, vis=False)
        self.num_test_iterations += 1
        return processed_results
