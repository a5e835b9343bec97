"""Data-parallel training step (one process per GPU).

What the reference does with detectron2's DDP + AMPTrainer + BaseTrainer.build_optimizer
(base_trainer.py:65-147): per-parameter AdamW groups (backbone LR x0.1, no weight decay on norm /
embedding / relative-position parameters, ``FREEZE_KEYS``), full-model gradient-norm clipping
(CLIP_VALUE 0.01), and a bucketed NCCL gradient all-reduce.  Here all trainable fp32 gradients live
in ONE flat buffer (parameters' ``.grad`` are views into it): the step issues a single
``all_reduce`` over that buffer (plus one for the fp64 classifier of PartDistillation when present),
the clip norm is one reduction over the same buffer, and AdamW is PyTorch's fused multi-tensor kernel.
"""
import torch
import torch.distributed as dist

_NORM_TYPES = (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d, torch.nn.SyncBatchNorm,
               torch.nn.GroupNorm, torch.nn.InstanceNorm1d, torch.nn.InstanceNorm2d, torch.nn.InstanceNorm3d,
               torch.nn.LayerNorm, torch.nn.LocalResponseNorm)


def build_param_groups(model, base_lr=1e-4, weight_decay=0.05, weight_decay_norm=0.0, weight_decay_embed=0.0,
                       backbone_multiplier=0.1, freeze_keys=()):
    """Per-parameter groups with the reference's rules; parameters of modules whose name contains a
    freeze key get ``requires_grad = False`` (base_trainer.py:97-100)."""
    groups, seen = [], set()
    for mod_name, module in model.named_modules():
        for p_name, p in module.named_parameters(recurse=False):
            if not p.requires_grad or p in seen:
                continue
            if any(k in mod_name for k in freeze_keys):
                p.requires_grad = False
                continue
            seen.add(p)
            lr, wd = base_lr, weight_decay
            if "backbone" in mod_name:
                lr = lr * backbone_multiplier
            if "relative_position_bias_table" in p_name or "absolute_pos_embed" in p_name:
                wd = 0.0
            if isinstance(module, _NORM_TYPES):
                wd = weight_decay_norm
            if isinstance(module, torch.nn.Embedding):
                wd = weight_decay_embed
            groups.append({"params": [p], "lr": lr, "weight_decay": wd})
    return groups


class DataParallelTrainer:
    """model(batched_inputs) -> loss dict; backward; ONE flat gradient all-reduce; clip; AdamW."""

    def __init__(self, model, base_lr=1e-4, weight_decay=0.05, clip_norm=0.01, freeze_keys=("backbone", "encoder"),
                 backbone_multiplier=0.1):
        self.model = model
        groups = build_param_groups(model, base_lr, weight_decay, freeze_keys=freeze_keys,
                                    backbone_multiplier=backbone_multiplier)
        self.params = [g["params"][0] for g in groups]
        self.clip_norm = clip_norm
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        # flat gradient buffers, one per dtype (fp32; fp64 only for PartDistillation's class_embed)
        self.flat = {}
        for dt in sorted({p.dtype for p in self.params}, key=str):
            ps = [p for p in self.params if p.dtype == dt]
            buf = torch.zeros(sum(p.numel() for p in ps), dtype=dt, device=ps[0].device)
            o = 0
            for p in ps:
                p.grad = buf[o:o + p.numel()].view_as(p)
                o += p.numel()
            self.flat[dt] = buf
        self.optimizer = torch.optim.AdamW(groups, lr=base_lr, fused=True)
        self.grad_bytes = sum(b.numel() * b.element_size() for b in self.flat.values())

    def zero_grad(self):
        for b in self.flat.values():
            b.zero_()

    def reduce_gradients(self):
        if self.world > 1:
            for b in self.flat.values():
                dist.all_reduce(b)
                b.div_(self.world)

    def clip_and_step(self):
        if self.clip_norm and self.clip_norm > 0:
            sq = sum((b.double() if b.dtype != torch.float64 else b).pow(2).sum() for b in self.flat.values())
            coef = torch.clamp(self.clip_norm / (sq.sqrt() + 1e-6), max=1.0)
            for b in self.flat.values():
                b.mul_(coef.to(b.dtype))
        self.optimizer.step()

    def backward_and_step(self, losses):
        total = sum(losses.values())
        total.backward()
        self.reduce_gradients()
        self.clip_and_step()
        return total

    def step(self, batched_inputs):
        self.zero_grad()
        losses = self.model(batched_inputs)
        return self.backward_and_step(losses), losses
