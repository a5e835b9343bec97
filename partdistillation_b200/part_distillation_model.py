"""``PartDistillationModel`` meta-architecture — training branch
(reference: part_distillation/part_distillation_model.py:32-226,405-428) and eval branch (:227-283,319-394,431-501;
``postprocess.PartDistillationInferenceMixin``).

Differences from ProposalModel, as in the reference: the head receives the targets (``mask=targets``,
:205) because the float64 classifier needs each image's object class; the matcher and the loss may
use different point counts (TRAIN_NUM_POINTS_MATCH / _LOSS, :130,151); classes are the part labels."""
from typing import Tuple

from torch import nn

from .compat import META_ARCH_REGISTRY, build_backbone, build_sem_seg_head, configurable
from .meta_base import Mask2FormerTrainingArch, build_criterion
from .postprocess import PartDistillationInferenceMixin


@META_ARCH_REGISTRY.register()
class PartDistillationModel(PartDistillationInferenceMixin, Mask2FormerTrainingArch):
    part_distillation = True

    @configurable
    def __init__(self, *, backbone, sem_seg_head: nn.Module, criterion: nn.Module, num_queries: int, num_classes: int,
                 size_divisibility: int, pixel_mean: Tuple[float], pixel_std: Tuple[float], test_topk_per_image: int,
                 train_dataset_name: str = "", use_wandb: bool = True, wandb_vis_period_train: int = 200,
                 wandb_vis_period_test: int = 20, wandb_vis_topk: int = 200, use_unique_per_pixel_label: bool = False,
                 min_pseudo_mask_ratio: float = 0.0, min_pseudo_mask_score: float = 0.0,
                 use_oracle_classifier: bool = False, apply_masking_with_object_mask: bool = True,
                 fg_score_threshold: float = 0.1):
        super().__init__()
        self._init_common(backbone, sem_seg_head, criterion, num_queries, num_classes, size_divisibility, pixel_mean,
                          pixel_std, test_topk_per_image, use_wandb)
        self.train_dataset_name = train_dataset_name
        self.wandb_vis_period_train = wandb_vis_period_train
        self.wandb_vis_period_test = wandb_vis_period_test
        self.wandb_vis_topk = wandb_vis_topk
        self.current_train_iteration = 0
        self.current_test_iteration = 0
        self.fg_score_threshold = fg_score_threshold
        self.use_unique_per_pixel_label = use_unique_per_pixel_label
        self.min_pseudo_mask_ratio = min_pseudo_mask_ratio
        self.min_pseudo_mask_score = min_pseudo_mask_score
        self.use_oracle_classifier = use_oracle_classifier
        self.apply_masking_with_object_mask = apply_masking_with_object_mask
        self.majority_vote_mapping = {}
        self.mode = "train"
        # part_distillation_model.py:91-92
        self.root_save_path = "pseudo_labels/part_labels/part_distillation_predictions/{}/{}_{}/".format(
            train_dataset_name, min_pseudo_mask_score, min_pseudo_mask_ratio)

    def update_majority_vote_mapping(self, mapping_dict):
        for cid, mapping in mapping_dict.items():
            self.majority_vote_mapping[cid] = mapping.to(self.device)

    @classmethod
    def from_config(cls, cfg):
        backbone = build_backbone(cfg)
        sem_seg_head = build_sem_seg_head(cfg, backbone.output_shape())
        m = cfg.MODEL.MASK_FORMER
        pd = cfg.PART_DISTILLATION
        num_classes = pd.NUM_PART_CLASSES
        criterion = build_criterion(cfg, num_classes, m.TRAIN_NUM_POINTS_MATCH, m.TRAIN_NUM_POINTS_LOSS)
        return dict(backbone=backbone, sem_seg_head=sem_seg_head, criterion=criterion,
                    num_queries=m.NUM_OBJECT_QUERIES, size_divisibility=m.SIZE_DIVISIBILITY,
                    pixel_mean=cfg.MODEL.PIXEL_MEAN, pixel_std=cfg.MODEL.PIXEL_STD,
                    wandb_vis_period_train=cfg.WANDB.VIS_PERIOD_TRAIN, wandb_vis_period_test=cfg.WANDB.VIS_PERIOD_TEST,
                    wandb_vis_topk=cfg.WANDB.VIS_TOPK, use_wandb=not cfg.WANDB.DISABLE_WANDB,
                    test_topk_per_image=cfg.TEST.DETECTIONS_PER_IMAGE,
                    use_unique_per_pixel_label=pd.USE_PER_PIXEL_LABEL, train_dataset_name=cfg.DATASETS.TRAIN[0],
                    num_classes=num_classes, min_pseudo_mask_ratio=pd.MIN_AREA_RATIO,
                    min_pseudo_mask_score=pd.MIN_SCORE, use_oracle_classifier=pd.USE_ORACLE_CLASSIFIER,
                    apply_masking_with_object_mask=pd.APPLY_MASKING_WITH_OBJECT_MASK)

    def forward(self, batched_inputs):
        if self.training:
            losses = super().forward(batched_inputs)
            self.current_train_iteration += 1
            return losses
        images = self.preprocess_images(batched_inputs)
        features = self.backbone(images.tensor)
        targets = self.prepare_targets(batched_inputs, images)
        outputs = self.run_head(features, targets)
        processed_results = self.inference(batched_inputs, targets, images, outputs, vis=False)
        self.current_test_iteration += 1
        return processed_results
