"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference (/root/reference) read-only.

Used in the build container (which has /root/reference but no GPU) by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/``
and by ``bench.py --impl reference`` / its ``cpu_baseline`` leg, which time the real reference on the host cores
when a reference tree resolves (see ``_resolve_root``).  Nothing in the product package imports this file.

Mechanism: the reference's ``part_distillation/__init__.py``
imports Detic / pycocotools / pydensecrf consumers, so we never execute it; instead we
register empty namespace packages whose ``__path__`` points at the reference directories
and import the hot-path sub-modules one by one.  detectron2 / fvcore / timm are provided
by ``oracle/shims``.
"""
import importlib
import os
import sys
import types

def _resolve_root():
    """$PD_REFERENCE_ROOT, else the reference checkout of the build container, else the copy baseline/install_reference.sh
    made under baseline/_ref (git-ignored; it is what exists on the GPU box)."""
    env = os.environ.get("PD_REFERENCE_ROOT")
    if env:
        return env
    local = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    for root in ("/root/reference", local):
        if os.path.isdir(os.path.join(root, "part_distillation")):
            return root
    return "/root/reference"


REF_ROOT = _resolve_root()
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_NAMESPACES = [
    "part_distillation",
    "part_distillation.modeling",
    "part_distillation.modeling.backbone",
    "part_distillation.modeling.pixel_decoder",
    "part_distillation.modeling.meta_arch",
    "part_distillation.modeling.transformer_decoder",
    "part_distillation.utils",
]

_IMPORT_ORDER = [
    "part_distillation.modeling.backbone.swin",
    "part_distillation.modeling.pixel_decoder.fpn",
    "part_distillation.modeling.pixel_decoder.msdeformattn",
    "part_distillation.modeling.meta_arch.mask_former_head",
    "part_distillation.modeling.transformer_decoder.mask2former_transformer_decoder",
    "part_distillation.modeling.transformer_decoder.part_distillation_transformer_decoder",
    "part_distillation.modeling.criterion",
    "part_distillation.modeling.matcher",
    "part_distillation.config",
    "part_distillation.proposal_model",
    "part_distillation.part_distillation_model",
]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "part_distillation"))


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path modules."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    # utils/utils.py imports cv2-free helpers only through the stubs above; make sure the
    # optional heavy deps resolve to stubs, not to a half-installed package.
    for name in _NAMESPACES:
        if name in sys.modules:
            continue
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
        m.__package__ = name
        sys.modules[name] = m
    mods = {}
    for name in _IMPORT_ORDER:
        mods[name.rsplit(".", 1)[1]] = importlib.import_module(name)
    ns = types.SimpleNamespace(**mods)
    ns.ops_func = importlib.import_module(
        "part_distillation.modeling.pixel_decoder.ops.functions.ms_deform_attn_func")
    ns.ops_mod = importlib.import_module(
        "part_distillation.modeling.pixel_decoder.ops.modules.ms_deform_attn")
    ns.position_encoding = importlib.import_module(
        "part_distillation.modeling.transformer_decoder.position_encoding")
    _loaded = ns
    return ns


def make_cfg(meta_arch="ProposalModel", backbone="swin_t", num_queries=25, dec_layers=10,
             num_points=12544, importance_sample_ratio=0.75, num_object_classes=1000,
             num_part_classes=8, num_classes=1):
    """cfg = detectron2 defaults + the reference's own add_*_config + the yaml values of
    configs/mask2former/.../maskformer2_R50_bs16_50ep.yaml and the swin variants."""
    ref = load()
    from detectron2.config import get_cfg
    cfg = get_cfg()
    ref.config.add_maskformer2_config(cfg)
    ref.config.add_wandb_config(cfg)
    ref.config.add_proposal_learning_config(cfg)
    ref.config.add_part_distillation_config(cfg)
    cfg.WANDB.DISABLE_WANDB = True
    cfg.MODEL.META_ARCHITECTURE = meta_arch
    s = cfg.MODEL.SEM_SEG_HEAD
    s.NAME = "MaskFormerHead"; s.IGNORE_VALUE = 255; s.NUM_CLASSES = num_classes; s.LOSS_WEIGHT = 1.0
    s.CONVS_DIM = 256; s.MASK_DIM = 256; s.NORM = "GN"
    s.PIXEL_DECODER_NAME = "MSDeformAttnPixelDecoder"
    s.IN_FEATURES = ["res2", "res3", "res4", "res5"]
    s.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES = ["res3", "res4", "res5"]
    s.COMMON_STRIDE = 4; s.TRANSFORMER_ENC_LAYERS = 6
    m = cfg.MODEL.MASK_FORMER
    m.TRANSFORMER_DECODER_NAME = ("PartDistillationTransformerDecoder"
                                  if meta_arch == "PartDistillationModel"
                                  else "MultiScaleMaskedTransformerDecoder")
    m.TRANSFORMER_IN_FEATURE = "multi_scale_pixel_decoder"
    m.DEEP_SUPERVISION = True; m.NO_OBJECT_WEIGHT = 0.1
    m.CLASS_WEIGHT = 2.0; m.MASK_WEIGHT = 5.0; m.DICE_WEIGHT = 5.0
    m.HIDDEN_DIM = 256; m.NUM_OBJECT_QUERIES = num_queries; m.NHEADS = 8; m.DROPOUT = 0.0
    m.DIM_FEEDFORWARD = 2048; m.ENC_LAYERS = 0; m.PRE_NORM = False; m.ENFORCE_INPUT_PROJ = False
    m.SIZE_DIVISIBILITY = 32; m.DEC_LAYERS = dec_layers
    m.TRAIN_NUM_POINTS = num_points; m.TRAIN_NUM_POINTS_MATCH = num_points
    m.TRAIN_NUM_POINTS_LOSS = num_points
    m.OVERSAMPLE_RATIO = 3.0; m.IMPORTANCE_SAMPLE_RATIO = importance_sample_ratio
    cfg.PART_DISTILLATION.NUM_OBJECT_CLASSES = num_object_classes
    cfg.PART_DISTILLATION.NUM_PART_CLASSES = num_part_classes
    cfg.TEST.DETECTIONS_PER_IMAGE = num_queries
    sw = cfg.MODEL.SWIN
    if backbone == "swin_t":
        sw.EMBED_DIM = 96; sw.DEPTHS = [2, 2, 6, 2]; sw.NUM_HEADS = [3, 6, 12, 24]
        sw.WINDOW_SIZE = 7; sw.PRETRAIN_IMG_SIZE = 224
    elif backbone == "swin_b":
        sw.EMBED_DIM = 128; sw.DEPTHS = [2, 2, 18, 2]; sw.NUM_HEADS = [4, 8, 16, 32]
        sw.WINDOW_SIZE = 12; sw.PRETRAIN_IMG_SIZE = 384
    elif backbone == "swin_micro":   # test-only: tiny trunk with the Swin-B channel plan / 8
        sw.EMBED_DIM = 32; sw.DEPTHS = [1, 1, 2, 1]; sw.NUM_HEADS = [1, 2, 4, 8]
        sw.WINDOW_SIZE = 4; sw.PRETRAIN_IMG_SIZE = 64
    else:
        raise ValueError(backbone)
    cfg.MODEL.BACKBONE.NAME = "D2SwinTransformer"
    return cfg


def build_model(cfg, workdir=None):
    """Instantiates the reference meta-architecture (PartDistillationModel mkdirs into cwd)."""
    ref = load()
    from detectron2.modeling import build_model as _bm
    cwd = os.getcwd()
    if workdir is not None:
        os.makedirs(workdir, exist_ok=True)
        os.chdir(workdir)
    try:
        model = _bm(cfg)
    finally:
        os.chdir(cwd)
    return model
