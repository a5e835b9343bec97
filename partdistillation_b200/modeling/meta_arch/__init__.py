from .mask_former_head import MaskFormerHead  # noqa: F401
