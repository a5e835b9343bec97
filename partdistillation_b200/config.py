"""Config keys consumed by the hot path, with the reference's defaults (part_distillation/config.py:10-276).

The reference's train scripts call ``add_maskformer2_config(cfg)`` etc. on a detectron2 ``CfgNode``;
these functions install the same keys so the same yaml / command-line overrides apply.  Only the
adders whose keys are read by the training path are provided (mask2former, wandb, proposal learning,
part distillation); the dataset / evaluation adders are outside the hot path.
"""
from .compat import CfgNode as CN

_MASKFORMER2 = {
    "INPUT": dict(DATASET_MAPPER_NAME="mask_former_semantic", COLOR_AUG_SSD=False, SIZE_DIVISIBILITY=-1,
                  IMAGE_SIZE_BASE=640, IMAGE_SIZE=1024, MIN_SCALE=0.1, MAX_SCALE=2.0,
                  CROP=dict(SINGLE_CATEGORY_MAX_AREA=1.0)),
    "SOLVER": dict(WEIGHT_DECAY_EMBED=0.0, OPTIMIZER="ADAMW", BACKBONE_MULTIPLIER=0.1),
    "MODEL": {
        "MASK_FORMER": dict(
            DEEP_SUPERVISION=True, NO_OBJECT_WEIGHT=0.1, CLASS_WEIGHT=1.0, DICE_WEIGHT=1.0, MASK_WEIGHT=20.0,
            NHEADS=8, DROPOUT=0.1, DIM_FEEDFORWARD=2048, ENC_LAYERS=0, DEC_LAYERS=6, PRE_NORM=False,
            HIDDEN_DIM=256, NUM_OBJECT_QUERIES=100, TRANSFORMER_IN_FEATURE="res5", ENFORCE_INPUT_PROJ=False,
            SIZE_DIVISIBILITY=32, TRANSFORMER_DECODER_NAME="MultiScaleMaskedTransformerDecoder",
            TRAIN_NUM_POINTS=112 * 112, TRAIN_NUM_POINTS_MATCH=112 * 112, TRAIN_NUM_POINTS_LOSS=112 * 112,
            OVERSAMPLE_RATIO=3.0, IMPORTANCE_SAMPLE_RATIO=0.75, FREEZE_KEYS=[], QUERY_FEATURE_NORMALIZE=False,
            TEST=dict(SEMANTIC_ON=True, INSTANCE_ON=False, PANOPTIC_ON=False, OBJECT_MASK_THRESHOLD=0.0,
                      OVERLAP_THRESHOLD=0.0, SEM_SEG_POSTPROCESSING_BEFORE_INFERENCE=False)),
        "SEM_SEG_HEAD": dict(
            MASK_DIM=256, TRANSFORMER_ENC_LAYERS=0, PIXEL_DECODER_NAME="BasePixelDecoder",
            DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES=["res3", "res4", "res5"],
            DEFORMABLE_TRANSFORMER_ENCODER_N_POINTS=4, DEFORMABLE_TRANSFORMER_ENCODER_N_HEADS=8),
        "SWIN": dict(
            PRETRAIN_IMG_SIZE=224, PATCH_SIZE=4, EMBED_DIM=96, DEPTHS=[2, 2, 6, 2], NUM_HEADS=[3, 6, 12, 24],
            WINDOW_SIZE=7, MLP_RATIO=4.0, QKV_BIAS=True, QK_SCALE=None, DROP_RATE=0.0, ATTN_DROP_RATE=0.0,
            DROP_PATH_RATE=0.3, APE=False, PATCH_NORM=True, OUT_FEATURES=["res2", "res3", "res4", "res5"],
            USE_CHECKPOINT=False),
    },
}

_WANDB = {
    "WANDB": dict(DISABLE_WANDB=False, GROUP=None, PROJECT="", VIS_PERIOD_TRAIN=200, VIS_PERIOD_TEST=20,
                  RUN_NAME="output", VIS_TOPK=10),
    "DATASETS": dict(DEBUG=False),
    "VIS_OUTPUT_DIR": "",
}

_PROPOSAL_LEARNING = {
    "PROPOSAL_LEARNING": dict(
        MIN_OBJECT_AREA_RATIO=0.001, MIN_AREA_RATIO=0.0, MIN_SCORE=-1.0, DATASET_PATH_LIST=[],
        FILTERED_CODE_PATH_LIST=[], EXCLUDE_CODE_PATH="", PATH_ONLY=False, USE_PER_PIXEL_LABEL=True,
        DATASET_PATH="", LABEL_PERCENTAGE=100, APPLY_MASKING_WITH_OBJECT_MASK=True, POSTPROCESS_TYPES=[],
        DEBUG=False),
}

_PART_DISTILLATION = {
    "PART_DISTILLATION": dict(
        DATASET_PATH="", DATASET_PATH_LIST=[], FILTERED_CODE_PATH_LIST=[], EXCLUDE_CODE_PATH="", PATH_ONLY=False,
        USE_PER_PIXEL_LABEL=True, NUM_PART_CLASSES=8, NUM_OBJECT_CLASSES=1000, MIN_OBJECT_AREA_RATIO=0.001,
        MIN_AREA_RATIO=-1.0, MIN_SCORE=-1.0, USE_ORACLE_CLASSIFIER=False, APPLY_MASKING_WITH_OBJECT_MASK=True,
        TOTAL_PARTITIONS=-1, PARTITION_INDEX=-1, SET_IMAGE_SQUARE=False, DEBUG=False),
}


_PIXEL_GROUPING = {
    "PIXEL_GROUPING": dict(NUM_SUPERPIXEL_CLUSTERS=4, DISTANCE_METRIC="l2", BACKBONE_FEATURE_KEY_LIST=["res4"],
                           FEATURE_NORMALIZE=False, DEBUG=False),
}


_PROPOSAL_GENERATION = {      # config.py:188-205 of the reference
    "PROPOSAL_GENERATION": dict(
        DATASET_NAME="imagenet_22k_train", OBJECT_MASK_TYPE="detic",
        OBJECT_MASK_PATH="pseudo_labels/object_labels/imagenet_22k_train/detic_predictions/", NUM_SUPERPIXEL_CLUSTERS=4,
        DISTANCE_METRIC="l2", FEATURE_NORMALIZE=False, BACKBONE_FEATURE_KEY_LIST=["res4"], TOTAL_PARTITIONS=-1,
        PARTITION_INDEX=-1, BATCH_SIZE=4, WITH_GIVEN_MASK=False, USE_PART_IMAGENET_CLASSES=False,
        FILTERED_CODE_PATH_LIST=[], EXCLUDE_CODE_PATH="", SINGLE_CLASS_CODE="", DEBUG=False),
}


def _install(node, table):
    for key, val in table.items():
        if isinstance(val, dict):
            if key not in node:
                setattr(node, key, CN())
            _install(getattr(node, key), val)
        else:
            setattr(node, key, list(val) if isinstance(val, list) else val)


def add_maskformer2_config(cfg):
    _install(cfg, _MASKFORMER2)


def add_wandb_config(cfg):
    _install(cfg, _WANDB)


def add_proposal_learning_config(cfg):
    _install(cfg, _PROPOSAL_LEARNING)


def add_part_distillation_config(cfg):
    _install(cfg, _PART_DISTILLATION)


def add_pixel_grouping_confing(cfg):          # (sic) the reference's spelling, config.py:255
    _install(cfg, _PIXEL_GROUPING)


def add_proposal_generation_config(cfg):
    _install(cfg, _PROPOSAL_GENERATION)
