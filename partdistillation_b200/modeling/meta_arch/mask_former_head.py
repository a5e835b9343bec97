"""MaskFormerHead (reference: modeling/meta_arch/mask_former_head.py:22-143): pixel decoder -> predictor."""
from typing import Dict

from torch import nn

from ...compat import (SEM_SEG_HEADS_REGISTRY, ShapeSpec, build_pixel_decoder, build_transformer_decoder,
                       configurable)


@SEM_SEG_HEADS_REGISTRY.register()
class MaskFormerHead(nn.Module):
    _version = 2

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        # version < 2 checkpoints keep pixel-decoder weights directly under the head (reference :27-56)
        version = local_metadata.get("version", None)
        if version is None or version < 2:
            for k in list(state_dict.keys()):
                if k.startswith(prefix) and not k.startswith(prefix + "predictor") \
                        and not k.startswith(prefix + "pixel_decoder."):
                    state_dict[k.replace(prefix, prefix + "pixel_decoder.", 1)] = state_dict.pop(k)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                      error_msgs)

    @configurable
    def __init__(self, input_shape: Dict[str, ShapeSpec], *, num_classes: int, pixel_decoder: nn.Module,
                 loss_weight: float = 1.0, ignore_value: int = -1, transformer_predictor: nn.Module,
                 transformer_in_feature: str):
        super().__init__()
        self.in_features = [k for k, _ in sorted(input_shape.items(), key=lambda kv: kv[1].stride)]
        self.ignore_value = ignore_value
        self.common_stride = 4
        self.loss_weight = loss_weight
        self.pixel_decoder = pixel_decoder
        self.predictor = transformer_predictor
        self.transformer_in_feature = transformer_in_feature
        self.num_classes = num_classes

    @classmethod
    def from_config(cls, cfg, input_shape: Dict[str, ShapeSpec]):
        feat = cfg.MODEL.MASK_FORMER.TRANSFORMER_IN_FEATURE
        head = cfg.MODEL.SEM_SEG_HEAD
        if feat in ("transformer_encoder", "multi_scale_pixel_decoder"):
            in_ch = head.CONVS_DIM
        elif feat == "pixel_embedding":
            in_ch = head.MASK_DIM
        else:
            in_ch = input_shape[feat].channels
        return dict(
            input_shape={k: v for k, v in input_shape.items() if k in head.IN_FEATURES},
            ignore_value=head.IGNORE_VALUE, num_classes=head.NUM_CLASSES,
            pixel_decoder=build_pixel_decoder(cfg, input_shape), loss_weight=head.LOSS_WEIGHT,
            transformer_in_feature=feat,
            transformer_predictor=build_transformer_decoder(cfg, in_ch, mask_classification=True))

    def forward(self, features, mask=None):
        return self.layers(features, mask)

    def layers(self, features, mask=None):
        mask_features, encoder_features, multi_scale = self.pixel_decoder.forward_features(features)
        if self.transformer_in_feature == "multi_scale_pixel_decoder":
            return self.predictor(multi_scale, mask_features, mask)
        if self.transformer_in_feature == "transformer_encoder":
            assert encoder_features is not None, "Please use the TransformerEncoderPixelDecoder."
            return self.predictor(encoder_features, mask_features, mask)
        if self.transformer_in_feature == "pixel_embedding":
            return self.predictor(mask_features, mask_features, mask)
        return self.predictor(features[self.transformer_in_feature], mask_features, mask)
