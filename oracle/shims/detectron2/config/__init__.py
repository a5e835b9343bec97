# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

import functools, inspect

class CfgNode(dict):
    """Attribute-access dict (yacs-like); just enough for add_*_config + from_config."""
    def __init__(self, init_dict=None):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)
    def __setattr__(self, name, value):
        self[name] = value
    def clone(self):
        import copy
        return copy.deepcopy(self)

def get_cfg():
    """Only the detectron2 defaults that the hot path reads."""
    c = CfgNode()
    c.VERSION = 2
    c.INPUT = CfgNode(); c.INPUT.CROP = CfgNode(); c.INPUT.FORMAT = "RGB"
    c.SOLVER = CfgNode(); c.DATASETS = CfgNode(); c.TEST = CfgNode()
    c.DATASETS.TRAIN = ("synthetic",); c.DATASETS.TEST = ()
    c.TEST.DETECTIONS_PER_IMAGE = 100
    c.MODEL = CfgNode()
    c.MODEL.DEVICE = "cpu"
    c.MODEL.META_ARCHITECTURE = "ProposalModel"
    c.MODEL.PIXEL_MEAN = [123.675, 116.280, 103.530]
    c.MODEL.PIXEL_STD = [58.395, 57.120, 57.375]
    c.MODEL.BACKBONE = CfgNode(); c.MODEL.BACKBONE.NAME = "D2SwinTransformer"; c.MODEL.BACKBONE.FREEZE_AT = 0
    c.MODEL.SEM_SEG_HEAD = CfgNode()
    s = c.MODEL.SEM_SEG_HEAD
    s.NAME = "MaskFormerHead"; s.IN_FEATURES = ["res2", "res3", "res4", "res5"]; s.IGNORE_VALUE = 255
    s.NUM_CLASSES = 1; s.CONVS_DIM = 256; s.COMMON_STRIDE = 4; s.NORM = "GN"; s.LOSS_WEIGHT = 1.0
    return c

def configurable(init_func=None, *, from_config=None):
    assert init_func is not None and init_func.__name__ == "__init__"
    @functools.wraps(init_func)
    def wrapped(self, *args, **kwargs):
        fc = type(self).from_config
        first = args[0] if args else kwargs.get("cfg", None)
        if isinstance(first, CfgNode):
            explicit = fc(*args, **kwargs)
            init_func(self, **explicit)
        else:
            init_func(self, *args, **kwargs)
    return wrapped
