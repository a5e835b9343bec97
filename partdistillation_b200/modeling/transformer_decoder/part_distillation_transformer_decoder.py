"""PartDistillation's decoder (reference: part_distillation_transformer_decoder.py:21-254): the
Mask2Former decoder with a float64 classifier over num_part_classes * num_object_classes + 1 columns
of which only the image's object class (P columns) and the no-object column are used.

The reference evaluates the full (B*Q, 256) x (256, P*O+1) float64 product and slices it; here the
P+1 weight rows are gathered per image (functional.class_rows) — same float64 arithmetic, dense
zero-filled weight gradients (so the optimizer sees exactly the reference's gradient tensor)."""
import torch
from torch import nn

from ... import functional as PF
from ...compat import TRANSFORMER_DECODER_REGISTRY, configurable
from .mask2former_transformer_decoder import MultiScaleMaskedTransformerDecoder


@TRANSFORMER_DECODER_REGISTRY.register()
class PartDistillationTransformerDecoder(MultiScaleMaskedTransformerDecoder):
    @configurable
    def __init__(self, in_channels, mask_classification=True, *, num_classes: int, hidden_dim: int, num_queries: int,
                 nheads: int, dim_feedforward: int, dec_layers: int, pre_norm: bool, mask_dim: int,
                 enforce_input_project: bool, num_object_classes: int, num_part_classes: int,
                 query_feature_normalize: bool):
        self._pd_classes = (num_part_classes, num_object_classes)
        MultiScaleMaskedTransformerDecoder.__init__.__wrapped__(
            self, in_channels, mask_classification, num_classes=num_classes, hidden_dim=hidden_dim,
            num_queries=num_queries, nheads=nheads, dim_feedforward=dim_feedforward, dec_layers=dec_layers,
            pre_norm=pre_norm, mask_dim=mask_dim, enforce_input_project=enforce_input_project,
            query_feature_normalize=query_feature_normalize)
        self.num_part_classes = num_part_classes

    def _build_class_embed(self, hidden_dim, num_classes):
        parts, objects = self._pd_classes
        self.class_embed = nn.Linear(hidden_dim, parts * objects + 1).double()

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        ret = MultiScaleMaskedTransformerDecoder.from_config.__func__(cls, cfg, in_channels, mask_classification)
        ret["num_object_classes"] = cfg.PART_DISTILLATION.NUM_OBJECT_CLASSES
        ret["num_part_classes"] = cfg.PART_DISTILLATION.NUM_PART_CLASSES
        return ret

    def _classify(self, decoder_output, targets):
        obj = getattr(targets, "object_classes", None)
        if obj is None:
            obj = PF.host_table([int(t["gt_object_class"]) for t in targets], torch.int32, decoder_output.device)
        return PF.class_rows(decoder_output, self.class_embed.weight, self.class_embed.bias, obj,
                             self.num_part_classes)

    def _extra_outputs(self, out, output):
        out["query_feats"] = output
