"""Config presets: detectron2 defaults + this package's ``add_*_config`` + the values the reference's yaml
chain sets (configs/mask2former/coco/instance-segmentation/{Base-COCO-InstanceSegmentation,
maskformer2_R50_bs16_50ep,swin/maskformer2_swin_base_IN21k_384_bs16_50ep}.yaml and
configs/{proposal_learning,part_distillation}/swinL_IN21K_384_mask2former.yaml)."""
from .compat import get_cfg
from .config import (add_maskformer2_config, add_part_distillation_config, add_proposal_learning_config,
                     add_wandb_config)

SWIN = {
    "swin_t": dict(EMBED_DIM=96, DEPTHS=[2, 2, 6, 2], NUM_HEADS=[3, 6, 12, 24], WINDOW_SIZE=7, PRETRAIN_IMG_SIZE=224),
    "swin_b": dict(EMBED_DIM=128, DEPTHS=[2, 2, 18, 2], NUM_HEADS=[4, 8, 16, 32], WINDOW_SIZE=12, PRETRAIN_IMG_SIZE=384),
    "swin_l": dict(EMBED_DIM=192, DEPTHS=[2, 2, 18, 2], NUM_HEADS=[6, 12, 24, 48], WINDOW_SIZE=12, PRETRAIN_IMG_SIZE=384),
    # test-only trunk: Swin-B's channel plan / 4
    "swin_micro": dict(EMBED_DIM=32, DEPTHS=[1, 1, 2, 1], NUM_HEADS=[1, 2, 4, 8], WINDOW_SIZE=4, PRETRAIN_IMG_SIZE=64),
}


def make_cfg(meta_arch="ProposalModel", backbone="swin_b", num_queries=100, dec_layers=10, num_points=12544,
             importance_sample_ratio=0.75, num_object_classes=1000, num_part_classes=8, num_classes=1,
             device="cuda"):
    cfg = get_cfg()
    add_maskformer2_config(cfg)
    add_wandb_config(cfg)
    add_proposal_learning_config(cfg)
    add_part_distillation_config(cfg)
    cfg.WANDB.DISABLE_WANDB = True
    cfg.MODEL.DEVICE = device
    cfg.MODEL.META_ARCHITECTURE = meta_arch
    s = cfg.MODEL.SEM_SEG_HEAD
    s.NAME = "MaskFormerHead"; s.IGNORE_VALUE = 255; s.NUM_CLASSES = num_classes; s.LOSS_WEIGHT = 1.0
    s.CONVS_DIM = 256; s.MASK_DIM = 256; s.NORM = "GN"
    s.PIXEL_DECODER_NAME = "MSDeformAttnPixelDecoder"
    s.IN_FEATURES = ["res2", "res3", "res4", "res5"]
    s.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES = ["res3", "res4", "res5"]
    s.COMMON_STRIDE = 4; s.TRANSFORMER_ENC_LAYERS = 6
    m = cfg.MODEL.MASK_FORMER
    m.TRANSFORMER_DECODER_NAME = ("PartDistillationTransformerDecoder" if meta_arch == "PartDistillationModel"
                                  else "MultiScaleMaskedTransformerDecoder")
    m.TRANSFORMER_IN_FEATURE = "multi_scale_pixel_decoder"
    m.DEEP_SUPERVISION = True; m.NO_OBJECT_WEIGHT = 0.1
    m.CLASS_WEIGHT = 2.0; m.MASK_WEIGHT = 5.0; m.DICE_WEIGHT = 5.0
    m.HIDDEN_DIM = 256; m.NUM_OBJECT_QUERIES = num_queries; m.NHEADS = 8; m.DROPOUT = 0.0
    m.DIM_FEEDFORWARD = 2048; m.ENC_LAYERS = 0; m.PRE_NORM = False; m.ENFORCE_INPUT_PROJ = False
    m.SIZE_DIVISIBILITY = 32; m.DEC_LAYERS = dec_layers
    m.TRAIN_NUM_POINTS = num_points; m.TRAIN_NUM_POINTS_MATCH = num_points; m.TRAIN_NUM_POINTS_LOSS = num_points
    m.OVERSAMPLE_RATIO = 3.0; m.IMPORTANCE_SAMPLE_RATIO = importance_sample_ratio
    cfg.PART_DISTILLATION.NUM_OBJECT_CLASSES = num_object_classes
    cfg.PART_DISTILLATION.NUM_PART_CLASSES = num_part_classes
    cfg.TEST.DETECTIONS_PER_IMAGE = num_queries
    for k, v in SWIN[backbone].items():
        setattr(cfg.MODEL.SWIN, k, v)
    cfg.MODEL.BACKBONE.NAME = "D2SwinTransformer"
    return cfg
