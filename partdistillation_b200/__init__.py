"""partdistillation_b200 — B200-native (sm_100a) implementation of PartDistillation's Mask2Former
training hot path.  Importing the package registers the drop-in modules under the reference's
registry names (as part_distillation/__init__.py:6-19 does for the reference):

    META_ARCH_REGISTRY           ProposalModel, PartDistillationModel, PixelGroupingModel, ProposalGenerationModel
    SEM_SEG_HEADS_REGISTRY       MaskFormerHead, MSDeformAttnPixelDecoder
    TRANSFORMER_DECODER_REGISTRY MultiScaleMaskedTransformerDecoder, PartDistillationTransformerDecoder
    BACKBONE_REGISTRY            D2SwinTransformer
"""
from . import modeling  # noqa: F401  (registration side effects)
from .config import (add_maskformer2_config, add_part_distillation_config, add_pixel_grouping_confing,  # noqa: F401
                     add_proposal_generation_config, add_proposal_learning_config, add_wandb_config)
from .part_distillation_model import PartDistillationModel  # noqa: F401
from .pixel_grouping_model import PixelGroupingModel, ProposalGenerationModel  # noqa: F401
from .proposal_model import ProposalModel  # noqa: F401

__all__ = ["ProposalModel", "PartDistillationModel", "PixelGroupingModel", "ProposalGenerationModel",
           "add_pixel_grouping_confing", "add_proposal_generation_config", "add_maskformer2_config", "add_wandb_config",
           "add_proposal_learning_config", "add_part_distillation_config"]
