"""Data-parallel training step (one process per GPU).

What the reference does with detectron2's DDP + AMPTrainer + BaseTrainer.build_optimizer
(base_trainer.py:65-147): per-parameter AdamW groups (backbone LR x0.1, no weight decay on norm /
embedding / relative-position parameters, ``FREEZE_KEYS``), full-model gradient-norm clipping
(CLIP_VALUE 0.01), and a bucketed NCCL gradient all-reduce.  Here all trainable fp32 parameters, gradients and
Adam moments live in flat buffers (``p.data`` / ``p.grad`` are views): the step issues a single ``all_reduce``
over the gradient buffer (plus one for the fp64 classifier of PartDistillation when present), the clip norm is
one reduction kernel over it and clip + AdamW one more (csrc/optim.cu).
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


class WarmupMultiStepLR:
    """Learning-rate factor of detectron2's ``WarmupMultiStepLR`` (detectron2==0.6 solver/lr_scheduler.py, un-vendored;
    selected by SOLVER.LR_SCHEDULER_NAME's default and configured by SOLVER.STEPS / GAMMA / WARMUP_* — the reference builds
    it in base_trainer.py:57-62; recipe values: STEPS (40000, 45000) of 50000 iterations, sh_files/proposal_learning/
    train_multi.sh:55-58; WARMUP_FACTOR 1.0, WARMUP_ITERS 10, Base-COCO-InstanceSegmentation.yaml:24-25):
    ``lr(it) = base_lr * warmup(it) * gamma ** #{milestones <= it}``."""

    def __init__(self, milestones, gamma=0.1, warmup_factor=0.001, warmup_iters=1000, warmup_method="linear"):
        if list(milestones) != sorted(milestones):
            raise ValueError(f"Milestones should be a list of increasing integers. Got {milestones}")
        if warmup_method not in ("constant", "linear"):
            raise ValueError(f"Unknown warmup method: {warmup_method}")
        self.milestones, self.gamma = list(milestones), gamma
        self.warmup_factor, self.warmup_iters, self.warmup_method = warmup_factor, warmup_iters, warmup_method

    def warmup(self, it):
        if it >= self.warmup_iters:
            return 1.0
        if self.warmup_method == "constant":
            return self.warmup_factor
        alpha = it / self.warmup_iters
        return self.warmup_factor * (1 - alpha) + alpha

    def factor(self, it):
        import bisect
        return self.warmup(it) * self.gamma ** bisect.bisect_right(self.milestones, it)


class WarmupPolyLR(WarmupMultiStepLR):
    """detectron2.projects.deeplab ``WarmupPolyLR`` (SOLVER.LR_SCHEDULER_NAME = "WarmupPolyLR", the Mask2Former semantic /
    panoptic configs): ``warmup(it) * (1 - it / max_iters) ** power``, held at ``constant_ending`` once below it."""

    def __init__(self, max_iters, power=0.9, constant_ending=0.0, warmup_factor=0.001, warmup_iters=1000,
                 warmup_method="linear"):
        super().__init__([], 1.0, warmup_factor, warmup_iters, warmup_method)
        self.max_iters, self.power, self.constant_ending = max_iters, power, constant_ending

    def factor(self, it):
        poly = (1.0 - it / self.max_iters) ** self.power
        if self.constant_ending > 0 and poly < self.constant_ending:
            return self.constant_ending          # detectron2 drops the warm-up factor here as well
        return self.warmup(it) * poly


def build_lr_schedule(cfg):
    """SOLVER.* -> schedule object (detectron2.projects.deeplab.build_lr_scheduler, used at base_trainer.py:62)."""
    s = cfg.SOLVER
    name = getattr(s, "LR_SCHEDULER_NAME", "WarmupMultiStepLR")
    warm = dict(warmup_factor=s.WARMUP_FACTOR, warmup_iters=s.WARMUP_ITERS, warmup_method=getattr(s, "WARMUP_METHOD", "linear"))
    if name == "WarmupMultiStepLR":
        return WarmupMultiStepLR([x for x in s.STEPS if x <= s.MAX_ITER], getattr(s, "GAMMA", 0.1), **warm)
    if name == "WarmupPolyLR":
        return WarmupPolyLR(s.MAX_ITER, getattr(s, "POLY_LR_POWER", 0.9), getattr(s, "POLY_LR_CONSTANT_ENDING", 0.0), **warm)
    raise ValueError(f"Unknown LR scheduler: {name}")


class DataParallelTrainer:
    """model(batched_inputs) -> loss dict; backward; ONE flat gradient all-reduce; clip; AdamW.

    fp32 CUDA parameters live in flat buffers (parameter, gradient, exp_avg, exp_avg_sq; every parameter a
    4-element-aligned segment, ``p.data`` / ``p.grad`` are views), so the step after backward is: one
    ``all_reduce`` (world > 1), ``pdb_grad_sumsq`` and ``pdb_adamw_flat`` — two kernel launches instead of one
    fused-AdamW launch per parameter group (the reference builds one group per parameter, base_trainer.py:76-116).
    Parameters of other dtypes (PartDistillation's fp64 classifier) and CPU models (the gloo tests of the
    sharding / all-reduce logic) are stepped by ``torch.optim.AdamW`` with the same clip coefficient.

    Differences from torch.optim.AdamW to know about: (1) the flat kernel updates EVERY segment each step, so a trainable
    parameter that receives no gradient in a step still sees weight decay and moment decay (torch skips ``grad is None``);
    every trainable parameter of the two meta-architectures is used in every training forward, so the hot path never hits
    this.  (2) With world > 1 the constructor broadcasts rank 0's parameters and buffers (``broadcast_parameters``)."""

    def __init__(self, model, base_lr=1e-4, weight_decay=0.05, clip_norm=0.01, freeze_keys=("backbone", "encoder"),
                 backbone_multiplier=0.1, betas=(0.9, 0.999), eps=1e-8, cuda_graph=False, target_bucket=0):
        self.model = model
        self.cuda_graph = bool(cuda_graph)
        # target_bucket = m > 0: every image's K pseudo masks are padded (on the device) to the next multiple of m with empty masks
        # marked gt_classes = -1, and PartDistillation's object class travels as a device scalar, so that batches with different
        # target counts / object classes share ONE step signature (hence one captured graph) per bucket.  The padding slots cost
        # the same for every query in the matcher (the real targets' assignment is unchanged), are "no object" in the class loss,
        # carry zero mask loss and are not counted in num_masks (criterion.py, loss.cu matcher_cost_kernel).  The random point
        # coordinates are drawn per matched pair, padding included, so the draws differ from an unpadded run's (same law).
        self.target_bucket = int(target_bucket)
        if self.target_bucket > 0:
            if not hasattr(model, "target_padding"):
                raise ValueError("target_bucket needs a Mask2FormerTrainingArch model (meta_base.py)")
            model.target_padding = True
        self._graphs = {}
        self._side = None
        self._num_masks = None
        self.pdb_launches = 0               # kernels launched through libpdb200.so, graph replays included
        groups = build_param_groups(model, base_lr, weight_decay, freeze_keys=freeze_keys,
                                    backbone_multiplier=backbone_multiplier)
        self.params = [g["params"][0] for g in groups]
        self.clip_norm = float(clip_norm or 0.0)
        self.betas, self.eps = betas, eps
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.step_count = 0
        flat_groups = [g for g in groups if g["params"][0].dtype == torch.float32 and g["params"][0].is_cuda]
        other_groups = [g for g in groups if not (g["params"][0].dtype == torch.float32 and g["params"][0].is_cuda)]
        self.flat = {}                      # name -> flat gradient buffer (what the all-reduce moves)
        self.flat_param = self.flat_m = self.flat_v = None
        if flat_groups:
            dev = flat_groups[0]["params"][0].device
            starts, o = [], 0
            for g in flat_groups:
                starts.append(o)
                o += (g["params"][0].numel() + 3) // 4 * 4
            self.flat_param = torch.zeros(o, dtype=torch.float32, device=dev)
            self.flat["f32"] = torch.zeros(o, dtype=torch.float32, device=dev)
            self.flat_m = torch.zeros(o, dtype=torch.float32, device=dev)
            self.flat_v = torch.zeros(o, dtype=torch.float32, device=dev)
            for g, st in zip(flat_groups, starts):
                p = g["params"][0]
                n = p.numel()
                self.flat_param[st:st + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param[st:st + n].view(p.shape)
                p.grad = self.flat["f32"][st:st + n].view(p.shape)
            self.seg_start = torch.tensor(starts, dtype=torch.int64, device=dev)
            self.seg_lr = torch.tensor([g["lr"] for g in flat_groups], dtype=torch.float32, device=dev)
            self.seg_wd = torch.tensor([g["weight_decay"] for g in flat_groups], dtype=torch.float32, device=dev)
            self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
            self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        # everything else: flat gradient buffer per dtype + torch AdamW
        self.other_params = [g["params"][0] for g in other_groups]
        for dt in sorted({p.dtype for p in self.other_params}, key=str):
            ps = [p for p in self.other_params if p.dtype == dt]
            buf = torch.zeros(sum(p.numel() for p in ps), dtype=dt, device=ps[0].device)
            o = 0
            for p in ps:
                p.grad = buf[o:o + p.numel()].view_as(p)
                o += p.numel()
            self.flat[str(dt)] = buf
        self.optimizer = None
        if other_groups:
            fused = all(g["params"][0].is_cuda for g in other_groups)
            self.optimizer = torch.optim.AdamW(other_groups, lr=base_lr, betas=betas, eps=eps, fused=fused,
                                               capturable=fused and self.cuda_graph)
        self.grad_bytes = sum(b.numel() * b.element_size() for b in self.flat.values())
        self.lr_schedule = None
        self.iteration = 0                  # optimizer steps taken through step() (graph replays included)
        self.max_graphs = 8                 # captured steps kept alive (least recently used signature is dropped)
        if self.world > 1:
            self.broadcast_parameters()

    def broadcast_parameters(self, src=0):
        """What DistributedDataParallel does at construction (detectron2 wraps the reference's model in it,
        train_net.py -> DefaultTrainer): every rank starts from rank `src`'s parameters and buffers, whatever seed it
        initialised its own randomly-initialised decoder / head with.  Frozen parameters are included (a checkpoint loaded
        on rank 0 only must reach every rank)."""
        if self.flat_param is not None:
            dist.broadcast(self.flat_param, src)
        flat_ids = {id(p) for p in self.params if p.dtype == torch.float32 and p.is_cuda} if self.flat_param is not None else set()
        with torch.no_grad():
            for p in self.model.parameters():
                if id(p) not in flat_ids:
                    dist.broadcast(p.data, src)
            for b in self.model.buffers():
                dist.broadcast(b, src)

    # ------------------------------------------------------------------------------------------ checkpointing
    def state_dict(self):
        """Optimizer / schedule state in torch.optim.AdamW's own layout — one param group per parameter, in the order
        `build_param_groups` emits them (the reference's optimizer has the same structure, base_trainer.py:76-116), state[i] =
        {"step", "exp_avg", "exp_avg_sq"} — so a checkpoint written by the reference's DefaultTrainer loads here and vice
        versa, plus ``iteration`` (the LR-schedule position)."""
        state, groups = {}, []
        flat_i = 0
        other = self.optimizer.state_dict() if self.optimizer is not None else None
        other_i = 0
        for i, p in enumerate(self.params):
            if self.flat_param is not None and p.dtype == torch.float32 and p.is_cuda:
                st, n = int(self.seg_start[flat_i]), p.numel()
                base_lr = float(self.seg_lr_base[flat_i]) if hasattr(self, "seg_lr_base") else float(self.seg_lr[flat_i])
                state[i] = {"step": self.step_dev.detach().to(torch.float32).reshape(()).clone(),
                            "exp_avg": self.flat_m[st:st + n].view(p.shape).clone(),
                            "exp_avg_sq": self.flat_v[st:st + n].view(p.shape).clone()}
                groups.append({"lr": base_lr, "weight_decay": float(self.seg_wd[flat_i]), "betas": tuple(self.betas),
                               "eps": self.eps, "params": [i]})
                flat_i += 1
            else:
                g = other["param_groups"][other_i]
                if other_i in other["state"]:
                    state[i] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in other["state"][other_i].items()}
                groups.append({"lr": float(g.get("base_lr", g["lr"])), "weight_decay": g["weight_decay"], "betas": g["betas"],
                               "eps": g["eps"], "params": [i]})
                other_i += 1
        return {"state": state, "param_groups": groups, "iteration": self.iteration}

    def load_state_dict(self, sd):
        """Inverse of ``state_dict`` (also accepts a plain torch.optim.AdamW state dict with one group per parameter)."""
        if len(sd["param_groups"]) != len(self.params):
            raise ValueError(f"optimizer state has {len(sd['param_groups'])} groups, the model {len(self.params)} parameters")
        flat_i = other_i = 0
        other_state = {}
        steps = []
        for i, p in enumerate(self.params):
            st_i = sd["state"].get(i)
            if self.flat_param is not None and p.dtype == torch.float32 and p.is_cuda:
                st, n = int(self.seg_start[flat_i]), p.numel()
                if st_i is not None:
                    self.flat_m[st:st + n].copy_(st_i["exp_avg"].reshape(-1))
                    self.flat_v[st:st + n].copy_(st_i["exp_avg_sq"].reshape(-1))
                    steps.append(int(st_i["step"]))
                flat_i += 1
            else:
                if st_i is not None:           # copies: torch's load_state_dict aliases tensors of matching dtype / device
                    other_state[other_i] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st_i.items()}
                other_i += 1
        if steps:
            if len(set(steps)) != 1:
                raise ValueError("per-parameter step counts differ: the flat AdamW kernel keeps ONE step counter")
            self.step_dev.fill_(steps[0])
            self.step_count = steps[0]
        if self.optimizer is not None and other_state:
            cur = self.optimizer.state_dict()
            cur["state"] = other_state
            self.optimizer.load_state_dict(cur)
        self.iteration = int(sd.get("iteration", steps[0] if steps else 0))
        self._graphs.clear()                # captured steps hold the old python-side state

    def set_lr_schedule(self, schedule):
        """``schedule.factor(iteration)`` scales every group's learning rate before each step.  The rates live in device
        memory (``seg_lr`` for the flat kernel, tensor ``lr`` for the torch groups), so a captured step picks the new
        value up at replay; graphs captured with python-float rates are dropped."""
        self.lr_schedule = schedule
        if self.flat_param is not None and not hasattr(self, "seg_lr_base"):
            self.seg_lr_base = self.seg_lr.clone()
        if self.optimizer is not None:
            for g in self.optimizer.param_groups:
                if "base_lr" not in g:
                    g["base_lr"] = float(g["lr"])
                    if g["params"][0].is_cuda:
                        g["lr"] = torch.tensor(g["base_lr"], dtype=torch.float32, device=g["params"][0].device)
            self._graphs.clear()

    def _apply_lr_factor(self, factor):
        if self.flat_param is not None:
            torch.mul(self.seg_lr_base, float(factor), out=self.seg_lr)
        if self.optimizer is not None:
            for g in self.optimizer.param_groups:
                if torch.is_tensor(g["lr"]):
                    g["lr"].fill_(g["base_lr"] * float(factor))
                else:
                    g["lr"] = g["base_lr"] * float(factor)

    def current_lr_factor(self):
        return 1.0 if self.lr_schedule is None else float(self.lr_schedule.factor(self.iteration))

    def zero_grad(self):
        for b in self.flat.values():
            b.zero_()

    def reduce_gradients(self):
        """The one data-path collective of the step: SUM all-reduce of the flat gradient buffers (the 1 / world
        scale is applied inside the optimizer kernels)."""
        if self.world > 1:
            for b in self.flat.values():
                dist.all_reduce(b)

    def clip_and_step(self):
        self.step_count += 1
        scale = 1.0 / self.world
        others = [b for k, b in self.flat.items() if k != "f32"]
        if "f32" in self.flat:
            from . import _lib
            from .functional import _stream
            lib = _lib.load()
            g = self.flat["f32"]
            self.step_dev.add_(1)
            _lib.check(lib.pdb_grad_sumsq(g.data_ptr(), g.numel(), scale, self.sumsq.data_ptr(), _stream()), "pdb_grad_sumsq")
            total = self.sumsq
            for b in others:
                total = total + (b.double() * scale).pow(2).sum()
            if others:
                self.sumsq.copy_(total)
            _lib.check(lib.pdb_adamw_flat(self.flat_param.data_ptr(), g.data_ptr(), self.flat_m.data_ptr(),
                                          self.flat_v.data_ptr(), g.numel(), self.seg_start.data_ptr(),
                                          self.seg_lr.data_ptr(), self.seg_wd.data_ptr(), self.seg_start.numel(),
                                          self.betas[0], self.betas[1], self.eps, self.step_dev.data_ptr(), scale, self.clip_norm,
                                          self.sumsq.data_ptr() if self.clip_norm > 0 else None, _stream()),
                       "pdb_adamw_flat")
            from . import functional as PF
            PF.weights_epoch += 1          # trainable weights changed in place: their cached pre-split parts are stale
            sq = self.sumsq[0]
        else:
            sq = sum((b.double() * scale).pow(2).sum() for b in others)
        if others:
            coef = scale
            if self.clip_norm > 0:
                coef = torch.clamp(self.clip_norm / (sq.sqrt() + 1e-6), max=1.0) * scale
            for b in others:
                b.mul_(coef.to(b.dtype) if torch.is_tensor(coef) else coef)
            self.optimizer.step()

    def backward_and_step(self, losses):
        total = sum(losses.values())
        total.backward()
        self.reduce_gradients()
        self.clip_and_step()
        return total

    def _eager_step(self, batched_inputs):
        self.zero_grad()
        losses = self.model(batched_inputs)
        return self.backward_and_step(losses), losses

    def _forward_backward(self, batched_inputs):
        self.zero_grad()
        losses = self.model(batched_inputs)
        total = sum(losses.values())
        total.backward()
        return total, losses

    def _padded_counts(self, batched_inputs):
        """Target slots per image: K rounded up to the bucket, never beyond the number of queries (a padding slot must always
        find a free query, or it would compete with the real targets), and 0 stays 0 (an image without targets has no matcher
        problem at all)."""
        m = self.target_bucket
        q = int(getattr(self.model, "num_queries", 0)) or (1 << 30)
        out = []
        for d in batched_inputs:
            k = int(d["instances"].gt_masks.tensor.shape[0])
            out.append(k if k == 0 or k >= q else min(q, (k + m - 1) // m * m))
        return out

    def _global_num_masks(self, batched_inputs):
        """mean over ranks of the per-rank number of target masks, clamped to >= 1 (criterion.py:248-254), all-reduced
        here, ahead of the step, so that the captured forward/backward holds no collective."""
        crit = getattr(self.model, "criterion", None)
        if crit is None or not hasattr(crit, "external_num_masks"):
            return
        total = sum(int(d["instances"].gt_masks.tensor.shape[0]) for d in batched_inputs)
        if self._num_masks is None:
            self._num_masks = torch.zeros(1, dtype=torch.float32, device=next(self.model.parameters()).device)
            self._num_masks_tmp = torch.zeros_like(self._num_masks)
        self._num_masks_tmp.fill_(float(total))
        dist.all_reduce(self._num_masks_tmp)
        torch.clamp(self._num_masks_tmp / self.world, min=1, out=self._num_masks)
        crit.external_num_masks = self._num_masks

    def step(self, batched_inputs):
        """One training step -> (total loss, loss dict).  With ``cuda_graph=True`` the whole step (forward, loss,
        backward, all-reduce, clip, AdamW) is captured once per batch signature (image shapes and per-image mask
        counts) after two eager steps, and replayed afterwards: the inputs are copied into the graph's static
        buffers, the returned tensors are the graph's static outputs."""
        from . import _lib
        if self.lr_schedule is not None:
            self._apply_lr_factor(self.lr_schedule.factor(self.iteration))
        self.iteration += 1
        raw_inputs = batched_inputs                     # the real target counts (num_masks all-reduce)
        dev = next(self.model.parameters()).device
        if not self.cuda_graph:
            if self.target_bucket > 0:
                batched_inputs = _padded_batch(batched_inputs, dev, self._padded_counts(batched_inputs))
            n0 = _lib.launch_count()
            out = self._eager_step(batched_inputs)
            self.pdb_launches += _lib.launch_count() - n0
            return out
        if self.target_bucket > 0:
            counts = self._padded_counts(batched_inputs)
            sig = tuple((tuple(d["image"].shape), (k,) + tuple(d["instances"].gt_masks.tensor.shape[1:]),
                         "gt_object_class" in d) for d, k in zip(batched_inputs, counts))
        else:
            counts = None
            sig = tuple((tuple(d["image"].shape), tuple(d["instances"].gt_masks.tensor.shape),
                         d.get("gt_object_class")) for d in batched_inputs)
        entry = self._graphs.pop(sig, None) or {"warm": 0}
        self._graphs[sig] = entry                       # most recently used last
        while len(self._graphs) > self.max_graphs:      # bounded: variable per-image mask counts re-capture, they must not leak
            self._graphs.pop(next(iter(self._graphs)))
        if "graph" not in entry:
            if self._side is None:
                self._side = torch.cuda.Stream()
            if entry["warm"] < 2:                       # lets shape-keyed caches fill and cuDNN / NCCL initialise
                # warm-up runs on the capture stream: autograd's AccumulateGrad nodes remember the stream they were
                # created on, and a node living on the default stream would invalidate the capture
                entry["warm"] += 1
                if self.world > 1:
                    self._global_num_masks(raw_inputs)
                if counts is not None:
                    batched_inputs = _padded_batch(batched_inputs, dev, counts)
                n0 = _lib.launch_count()
                cur = torch.cuda.current_stream()
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    out = self._eager_step(batched_inputs)
                cur.wait_stream(self._side)
                self.pdb_launches += _lib.launch_count() - n0
                return out
            static = _padded_batch(batched_inputs, dev, counts) if counts is not None else _clone_batch(batched_inputs, dev)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(graph, stream=self._side):
                # world 1: the whole step; world > 1: forward + backward only — the collectives (num_masks before,
                # the flat-gradient all-reduce after) and the two optimizer kernels stay outside the graph
                total, losses = self._eager_step(static) if self.world == 1 else self._forward_backward(static)
            entry.update(graph=graph, static=static, total=total, losses=losses, launches=_lib.launch_count() - n0)
        if counts is not None:
            _padded_batch(batched_inputs, dev, counts, out=entry["static"])
        else:
            _copy_batch(entry["static"], batched_inputs)
        if self.world > 1:
            self._global_num_masks(raw_inputs)
        entry["graph"].replay()
        self.pdb_launches += entry["launches"]
        if self.world == 1:
            from . import functional as PF
            PF.weights_epoch += 1          # the replayed AdamW changed the weights: cached pre-split low parts are stale
        if self.world > 1:
            n0 = _lib.launch_count()
            self.reduce_gradients()
            self.clip_and_step()
            self.pdb_launches += _lib.launch_count() - n0
        return entry["total"], entry["losses"]


def _padded_batch(batch, device, counts, out=None):
    """Device-resident copy of a batch whose i-th image carries counts[i] >= K_i target slots: the K_i real masks / classes, then
    empty masks with class -1 (see DataParallelTrainer's target_bucket).  PartDistillation's object class is added as an int32
    device scalar ("gt_object_class_dev").  ``out``: a previous result of the same shapes to refill in place (the static inputs
    of a captured step); otherwise new tensors are allocated."""
    res = out if out is not None else []
    for n, (d, kp) in enumerate(zip(batch, counts)):
        src = d["instances"]
        k = int(src.gt_masks.tensor.shape[0])
        if out is None:
            e = dict(d)
            e["image"] = d["image"].to(device, copy=True)
            inst = type(src)(src.image_size)
            t = torch.zeros((kp,) + tuple(src.gt_masks.tensor.shape[1:]), dtype=src.gt_masks.tensor.dtype, device=device)
            inst.gt_masks = src.gt_masks.like(t) if hasattr(src.gt_masks, "like") else type(src.gt_masks)(t)
            inst.gt_classes = torch.full((kp,), -1, dtype=src.gt_classes.dtype, device=device)
            e["instances"] = inst
            if "gt_object_class" in d:
                e["gt_object_class_dev"] = torch.zeros((), dtype=torch.int32, device=device)
            res.append(e)
        else:
            e = res[n]
            e["image"].copy_(d["image"], non_blocking=True)
            if kp > k:
                e["instances"].gt_masks.tensor[k:].zero_()
                e["instances"].gt_classes[k:].fill_(-1)
        if k:
            e["instances"].gt_masks.tensor[:k].copy_(src.gt_masks.tensor, non_blocking=True)
            e["instances"].gt_classes[:k].copy_(src.gt_classes, non_blocking=True)
        if "gt_object_class" in d:
            e["gt_object_class"] = d["gt_object_class"]                 # the reference-format dicts keep the host value
            e["gt_object_class_dev"].fill_(int(d["gt_object_class"]))   # a kernel argument, not a host-to-device copy
    return res


def _clone_batch(batch, device):
    """Device-resident copy of a batch (list of dicts with "image" and "instances") used as a graph's static input."""
    out = []
    for d in batch:
        e = dict(d)
        e["image"] = d["image"].to(device, copy=True)
        src = d["instances"]
        inst = type(src)(src.image_size)
        t = src.gt_masks.tensor.to(device, copy=True)
        inst.gt_masks = src.gt_masks.like(t) if hasattr(src.gt_masks, "like") else type(src.gt_masks)(t)
        inst.gt_classes = src.gt_classes.to(device, copy=True)
        e["instances"] = inst
        out.append(e)
    return out


def _copy_batch(static, batch):
    for s, d in zip(static, batch):
        if s["image"] is not d["image"]:
            s["image"].copy_(d["image"], non_blocking=True)
            s["instances"].gt_masks.tensor.copy_(d["instances"].gt_masks.tensor, non_blocking=True)
            s["instances"].gt_classes.copy_(d["instances"].gt_classes, non_blocking=True)
