from .swin import D2SwinTransformer  # noqa: F401
