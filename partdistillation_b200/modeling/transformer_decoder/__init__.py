from .mask2former_transformer_decoder import MultiScaleMaskedTransformerDecoder  # noqa: F401
from .part_distillation_transformer_decoder import PartDistillationTransformerDecoder  # noqa: F401
