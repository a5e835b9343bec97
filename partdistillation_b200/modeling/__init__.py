from .backbone.swin import D2SwinTransformer  # noqa: F401
from .pixel_decoder.msdeformattn import MSDeformAttnPixelDecoder  # noqa: F401
from .meta_arch.mask_former_head import MaskFormerHead  # noqa: F401
from .transformer_decoder import MultiScaleMaskedTransformerDecoder, PartDistillationTransformerDecoder  # noqa: F401
from .criterion import SetCriterion  # noqa: F401
from .matcher import HungarianMatcher  # noqa: F401
