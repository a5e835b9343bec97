#!/usr/bin/env python
"""bench.py — train images/sec of the Mask2Former hot path (BASELINE.json configs[1]):
ProposalModel, Swin-B, 100 queries, synthetic 1024x1024 images, 2 images per GPU, forward + loss +
backward + gradient all-reduce + clip + AdamW, recipe freeze (backbone + deformable encoder frozen:
sh_files/proposal_learning/train_multi.sh:8,46 of the reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
``--impl reference`` times the reference's own implementation of the same step on the host CPU: the unmodified
reference through oracle/ref_loader.py when a reference tree resolves (/root/reference, or baseline/_ref on the GPU box),
else the oracle port under oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "train images/sec Swin-B 100q 1024^2 (fwd+loss+bwd+allreduce+AdamW)"
H = W = 1024
PER_GPU_BATCH = 2
K_MASKS = 6
QUERIES = 100
POINTS = 12544


# BASELINE configs[1] (default; the driver's line) and configs[2] (--workload c3)
SPECS = {
    "c2": dict(arch="ProposalModel", size=1024, batch=2, kmasks=(K_MASKS, K_MASKS), amp=False, dtype="f32", metric=METRIC,
               workload="proposal_learning Swin-B 100q 1024x1024 bs=2/GPU fwd+bwd (BASELINE configs[1])", cfg={}),
    "c3": dict(arch="PartDistillationModel", size=640, batch=32, kmasks=(2, 8), amp=True, dtype="bf16",
               metric="train images/sec PartDistillation Swin-B 100q 640^2 bf16 autocast (fwd+loss+bwd+allreduce+AdamW)",
               workload="part_distillation_model Swin-B 100q bf16 autocast, 640x640, bs=32/GPU, 22 000 object x 8 part classes "
                        "(BASELINE configs[2])", cfg=dict(num_object_classes=22000, num_part_classes=8)),
}


# ------------------------------------------------------------------------------------------------
# synthetic data: uint8 images, block label maps -> K disjoint bool masks
# ------------------------------------------------------------------------------------------------
def synth_image_and_masks(seed, h=H, w=W, k=K_MASKS, block=16):
    g = torch.Generator().manual_seed(3000 + seed)
    img = torch.randint(0, 256, (3, h, w), generator=g, dtype=torch.uint8)
    lab = torch.randint(0, k, (h // block, w // block), generator=g)
    lab = lab.repeat_interleave(block, 0).repeat_interleave(block, 1)
    m = torch.stack([lab == i for i in range(k)])
    return img, m[m.flatten(1).any(1)]


def make_batch(rank, n, device=None, pin=False, packed=False, spec=None, k_masks=None):
    """``packed``: hand the masks over as PackedBitMasks (1 bit / pixel) instead of BitMasks bools.
    ``spec``: a SPECS entry (default configs[1]); configs[2] draws K ~ U{2..8} parts per image, part labels arange(K) % 8 and
    an object class ~ U[0, 22000)."""
    from partdistillation_b200.compat import BitMasks, Instances, PackedBitMasks
    if spec is not None and spec["arch"] == "PartDistillationModel":
        return _make_pd_batch(rank, n, device, pin, spec)
    out = []
    for i in range(n):
        img, m = synth_image_and_masks(rank * 1000 + i) if k_masks is None else synth_image_and_masks(rank * 1000 + i, k=k_masks)
        if packed:
            m = PackedBitMasks.from_bool(m).tensor
        if pin:
            img, m = img.pin_memory(), m.pin_memory()
        if device is not None:
            img, m = img.to(device), m.to(device)
        inst = Instances((H, W))
        inst.gt_masks = PackedBitMasks(m, W) if packed else BitMasks(m)
        inst.gt_classes = torch.zeros(m.shape[0], dtype=torch.long, device=m.device)
        out.append({"image": img, "instances": inst, "height": H, "width": W})
    return out


def _make_pd_batch(rank, n, device, pin, spec):
    from partdistillation_b200.compat import BitMasks, Instances
    size = spec["size"]
    out = []
    for i in range(n):
        g = torch.Generator().manual_seed(7000 + rank * 1000 + i)
        k = int(torch.randint(spec["kmasks"][0], spec["kmasks"][1] + 1, (1,), generator=g))
        img, m = synth_image_and_masks(rank * 1000 + i, size, size, k)
        if pin:
            img, m = img.pin_memory(), m.pin_memory()
        if device is not None:
            img, m = img.to(device), m.to(device)
        inst = Instances((size, size))
        inst.gt_masks = BitMasks(m)
        inst.gt_classes = (torch.arange(m.shape[0]) % 8).to(m.device)
        out.append({"image": img, "instances": inst, "height": size, "width": size,
                    "gt_object_class": int(torch.randint(0, spec["cfg"]["num_object_classes"], (1,), generator=g))})
    return out


def batch_bytes(batch):
    return sum(d["image"].numel() * d["image"].element_size()
               + d["instances"].gt_masks.tensor.numel() * d["instances"].gt_masks.tensor.element_size() for d in batch)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = sorted(float(r[1]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][2])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for j, n in enumerate(names):
            if any("Active" in r[5 + j] and "Not" not in r[5 + j] for r in rows):
                out["reasons"].append(n)
        out["power_w_max"] = max(float(r[3]) for r in rows)
        out["samples"] = len(rows)
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own implementation of the step on the host cores
# ------------------------------------------------------------------------------------------------
def make_cpu_reference_step(n_images, state_dict=None):
    """-> (step(), kind, cores, what).  One step = forward + loss + backward + full-model clip + AdamW over `n_images`
    synthetic 1024^2 images (configs[1]: Swin-B, 100 queries, 10 decoder layers, 12 544 points, recipe freeze), all host threads.
    kind "reference": the UNMODIFIED reference ProposalModel (proposal_model.py:177-204) built through oracle/ref_loader.py from
    /root/reference or from baseline/_ref (baseline/install_reference.sh; that copy is what exists on the GPU box), detectron2 /
    fvcore / timm provided by oracle/shims, optimizer = torch.optim.AdamW + clip_grad_norm_ as base_trainer.py:118-147.
    kind "port": oracle/m2f_oracle.py (the restatement pinned to the reference by tests/test_oracle_golden.py) when no
    reference tree resolves."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    import ref_loader as rl
    if rl.available():
        import warnings
        warnings.filterwarnings("ignore")
        rl.load()
        from detectron2.structures import BitMasks, Instances
        cfg = rl.make_cfg("ProposalModel", "swin_b", num_queries=QUERIES, dec_layers=10, num_points=POINTS,
                          importance_sample_ratio=0.0)
        torch.manual_seed(0)
        model = rl.build_model(cfg, os.path.join(tempfile.gettempdir(), "pdb_ref_work")).train()
        trainable = []
        for name, p_ in model.named_parameters():
            if "backbone" in name or "encoder" in name:          # train_multi.sh:8,46 (FREEZE_KEYS)
                p_.requires_grad_(False)
            else:
                trainable.append(p_)
        opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=0.05)
        batch = []
        for i in range(n_images):
            img, m = synth_image_and_masks(i)
            inst = Instances((H, W))
            inst.gt_masks = BitMasks(m)
            inst.gt_classes = torch.zeros(m.shape[0], dtype=torch.long)
            batch.append({"image": img, "instances": inst, "height": H, "width": W})

        def step():
            opt.zero_grad(set_to_none=True)
            losses = model(batch)
            sum(losses.values()).backward()
            torch.nn.utils.clip_grad_norm_(trainable, 0.01)
            opt.step()
        return step, "reference", cores, (f"{n_images} x 1024^2 per step, fwd+loss+bwd+clip+AdamW, the unmodified reference "
                                          f"ProposalModel from {os.path.relpath(rl.REF_ROOT, ROOT) if rl.REF_ROOT.startswith(ROOT) else rl.REF_ROOT}")
    import m2f_oracle as O
    if state_dict is None:
        from partdistillation_b200 import compat, presets
        cfg = presets.make_cfg("ProposalModel", "swin_b", QUERIES, 10, POINTS, 0.0, device="cpu")
        torch.manual_seed(0)
        state_dict = compat.META_ARCH_REGISTRY.get("ProposalModel")(cfg).state_dict()
    sd = {k: v.detach().cpu().clone() for k, v in state_dict.items()}
    for k, v in sd.items():
        if k.startswith("sem_seg_head.") and v.is_floating_point() and "encoder" not in k:
            v.requires_grad_(True)
    trainable = [v for v in sd.values() if v.requires_grad]
    opt = torch.optim.AdamW(trainable, lr=1e-4, weight_decay=0.05)
    hp = dict(num_classes=1, dec_layers=10, num_points_match=POINTS, num_points_loss=POINTS, w_class=2.0,
              w_mask=5.0, w_dice=5.0, eos_coef=0.1, oversample_ratio=3.0, importance_ratio=0.0)
    batch = []
    for i in range(n_images):
        img, m = synth_image_and_masks(i)
        batch.append({"image": img.float(), "gt_masks": m})
    mean, std = [123.675, 116.280, 103.530], [58.395, 57.120, 57.375]

    def step():
        opt.zero_grad(set_to_none=True)
        x = O.prepare_images(batch, mean, std, 32)
        with torch.no_grad():
            feats = O.swin_forward(sd, "backbone.", x, 128, [2, 2, 18, 2], [4, 8, 16, 32], 12)
        losses = O.head_and_loss(sd, feats, O.prepare_targets(batch, H, W), hp)
        sum(losses.values()).backward()
        torch.nn.utils.clip_grad_norm_(trainable, 0.01)
        opt.step()
    return step, "port", cores, f"{n_images} x 1024^2 per step, fwd+loss+bwd+clip+AdamW, oracle/m2f_oracle.py (no reference tree on this box)"


def time_cpu_reference(step, steps, warmup, budget_s=420.0):
    """Seconds per step over `steps` timed steps after `warmup` untimed ones; stops early (after >= 1 timed step) only when
    the wall-clock budget would be exceeded.  Returns (seconds per step, timed steps, warm-up steps done)."""
    t_start = time.perf_counter()
    times, wdone = [], 0
    last = 0.0
    for i in range(warmup + steps):
        if times and time.perf_counter() - t_start + last > budget_s:
            break
        if i < warmup and wdone >= 1 and time.perf_counter() - t_start + last * (1 + steps) > budget_s:
            continue                                     # budget: skip the remaining warm-up steps, keep the timed ones
        t0 = time.perf_counter()
        step()
        last = time.perf_counter() - t0
        if i < warmup:
            wdone += 1
        else:
            times.append(last)
    return sum(times) / len(times), len(times), wdone


# ------------------------------------------------------------------------------------------------
# kernel roofline (measured live, CUDA events on the launching stream, L2 flushed)
# ------------------------------------------------------------------------------------------------
def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel from the committed `ncu --set full` summary
    of THIS round (profiles/r02_kernel_traffic.json, written by tools/ncu_traffic.py from the raw export); (None, None) if the
    summary does not hold the kernel.  bench.py cannot read DRAM counters itself: never measured under a profiler here."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")))[key]
        return float(d["dram_bytes"]), d["source"]
    except Exception:
        return None, None


def _time_kernel(calls, flush, iters=10, warmup=3, per_event=8):
    """Average launch duration in seconds.  `per_event` launches are issued back to back between ONE pair of CUDA events on
    the launching (current) stream and the elapsed time is divided by their number, so the ~3-5 us of event / launch
    bracketing is not charged to every launch.  `calls` are the same kernel over different buffer sets whose combined
    footprint exceeds the 126 MB L2 several times, so no launch finds its inputs cached; an L2-flushing memset runs ahead
    of the first launch (and keeps the GPU busy while the host enqueues the launches)."""
    ts = []
    for i in range(warmup + iters):
        flush.zero_()
        flush.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for j in range(per_event):
            calls[j % len(calls)]()
        e.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ts.append(s.elapsed_time(e) * 1e-3 / per_event)
    return sum(ts) / len(ts)


def kernel_rooflines(device):
    """Live rooflines of the hot kernels at the C2 shapes of one step (N = 2 images per GPU):
      * mask einsum forward  B=2, Q=100, C=256, 256x256      186.9 MB algorithmic
      * MSDeformAttn gather  N=2, levels 32^2/64^2/128^2, S=Lq=21504, M=8, D=32, P=4     137.6 MB
      * MSDeformAttn scatter (backward)                                                    231.2 MB
      * encoder FFN linear 43008 x 256 -> 1024 on the tcgen05 3xTF32 GEMM                  22.5 GFLOP (tensor bound)
    """
    from partdistillation_b200 import functional as fn
    g = torch.Generator().manual_seed(0)
    peak, how = measured_peak()
    try:
        tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        tf_peak = 1590.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    out = []

    def hbm_entry(name, alg_bytes, t):
        ach = alg_bytes / t / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "peak_source": how, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": None, "algorithmic_bytes": alg_bytes,
                "avg_launch_us": round(t * 1e6, 2)}

    NSETS = 4          # buffer sets per kernel: 4 x (134 + 52) MB for the einsum, 4 x 138 MB for the gather, 4 x 220 MB for the linear

    # ---- mask einsum (bqc,bchw->bqhw)
    B, Q, C, HH, WW = PER_GPU_BATCH, QUERIES, 256, H // 4, W // 4
    e = torch.randn(B, Q, C, generator=g).to(device)
    fs = [torch.randn(B, C, HH, WW, generator=g).to(device).contiguous(memory_format=torch.channels_last) for _ in range(NSETS)]
    e_lo = fn.split_lo(e)            # the embed is split once per decoder layer by its own tiny launch; the GEMM is what is timed
    with torch.no_grad():
        t = _time_kernel([(lambda f=f: fn.mask_einsum(e, f, embed_lo=e_lo)) for f in fs], flush)
    out.append(hbm_entry("gemm_tf32x3_kernel (mask einsum fwd)", 4 * (B * Q * C + B * C * HH * WW + B * Q * HH * WW), t))
    out[-1]["traffic"], out[-1]["traffic_source"] = ncu_traffic("mask_einsum_fwd")
    del fs

    # ---- MSDeformAttn gather / scatter at the encoder shape
    shapes = [(H // 32, W // 32), (H // 16, W // 16), (H // 8, W // 8)]
    N, M, D, P, L = PER_GPU_BATCH, 8, 32, 4, 3
    S = sum(h * w for h, w in shapes)
    refs = []
    for (hh, ww) in shapes:
        ys, xs = torch.meshgrid((torch.arange(hh) + 0.5) / hh, (torch.arange(ww) + 0.5) / ww, indexing="ij")
        refs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    ref = torch.cat(refs)[None, :, None, None, None, :]
    norm = torch.tensor([[w_, h_] for h_, w_ in shapes], dtype=torch.float32)[None, None, None, :, None, :]
    sets = []
    for _ in range(NSETS):
        value = torch.randn(N, S, M, D, generator=g).to(device).requires_grad_()
        loc = (ref + (torch.rand(N, S, M, L, P, 2, generator=g) * 2 - 1) * 4.0 / norm).contiguous().to(device).requires_grad_()
        attn = torch.softmax(torch.randn(N, S, M, L * P, generator=g), -1).view(N, S, M, L, P).contiguous().to(device).requires_grad_()
        sets.append((value, loc, attn))
    fwd_bytes = 4 * (N * S * M * D + N * S * M * L * P * 3 + N * S * M * D)
    with torch.no_grad():
        t = _time_kernel([(lambda v=v, lo=lo, a=a: fn.ms_deform_attn(v, shapes, None, lo, a)) for v, lo, a in sets], flush)
    out.append(hbm_entry("msda_fwd_tiled (C2: 3 levels, N=2)", fwd_bytes, t))
    out[-1]["traffic"], out[-1]["traffic_source"] = ncu_traffic("msda_fwd_C2")
    outs = [fn.ms_deform_attn(v, shapes, None, lo, a) for v, lo, a in sets]
    go = torch.randn_like(outs[0])
    t = _time_kernel([(lambda o=o, st=st: torch.autograd.grad(o, st, go, retain_graph=True)) for o, st in zip(outs, sets)], flush)
    out.append(hbm_entry("msda_bwd_tiled (+ grad_value zero fill) (C2)", 2 * fwd_bytes - 4 * N * S * M * D, t))
    out[-1]["traffic"], out[-1]["traffic_source"] = ncu_traffic("msda_bwd_C2")
    del sets, outs

    # ---- BASELINE configs[4] = C5(i): 4-scale pyramid 256^2 .. 32^2, N = 1, Lq = S = 87 040 (two buffer sets of 312 MB each)
    shapes5 = [(H // 4, W // 4), (H // 8, W // 8), (H // 16, W // 16), (H // 32, W // 32)]
    S5, L5 = sum(h * w for h, w in shapes5), 4
    refs = []
    for (hh, ww) in shapes5:
        ys, xs = torch.meshgrid((torch.arange(hh) + 0.5) / hh, (torch.arange(ww) + 0.5) / ww, indexing="ij")
        refs.append(torch.stack((xs.reshape(-1), ys.reshape(-1)), -1))
    ref5 = torch.cat(refs)[None, :, None, None, None, :]
    norm5 = torch.tensor([[w_, h_] for h_, w_ in shapes5], dtype=torch.float32)[None, None, None, :, None, :]
    sets = []
    for _ in range(2):
        value = torch.randn(1, S5, M, D, generator=g).to(device).requires_grad_()
        loc = (ref5 + (torch.rand(1, S5, M, L5, P, 2, generator=g) * 2 - 1) * 4.0 / norm5).contiguous().to(device).requires_grad_()
        attn = torch.softmax(torch.randn(1, S5, M, L5 * P, generator=g), -1).view(1, S5, M, L5, P).contiguous().to(device).requires_grad_()
        sets.append((value, loc, attn))
    fwd5 = 4 * (S5 * M * D + S5 * M * L5 * P * 3 + S5 * M * D)
    with torch.no_grad():
        t = _time_kernel([(lambda v=v, lo=lo, a=a: fn.ms_deform_attn(v, shapes5, None, lo, a)) for v, lo, a in sets], flush)
    out.append(hbm_entry("msda_fwd_tma (C5(i): 4 levels 256^2..32^2, N=1, TMA-staged value tiles)", fwd5, t))
    out[-1]["traffic"], out[-1]["traffic_source"] = ncu_traffic("msda_fwd_C5i")
    outs = [fn.ms_deform_attn(v, shapes5, None, lo, a) for v, lo, a in sets]
    go = torch.randn_like(outs[0])
    t = _time_kernel([(lambda o=o, st=st: torch.autograd.grad(o, st, go, retain_graph=True)) for o, st in zip(outs, sets)], flush)
    out.append(hbm_entry("msda_bwd_tiled (+ grad_value zero fill) (C5(i))", 2 * fwd5 - 4 * S5 * M * D, t))
    out[-1]["traffic"], out[-1]["traffic_source"] = ncu_traffic("msda_bwd_C5i")
    del sets, outs

    # ---- BASELINE configs[3] = C4: pixel grouping, res3 + res4 features (768 ch at 64^2) of a 512^2 image, 4 centroids
    GB = 16                                                                               # images per batched call
    feats = [torch.randn(GB, 768, 64, 64, generator=g).to(device) for _ in range(2)]      # 2 x 201 MB > L2
    cent = torch.randn(GB, 4, 768, generator=g).to(device)
    yy, xx = torch.meshgrid(torch.arange(512), torch.arange(512), indexing="ij")
    gmask = (((yy - 256) ** 2 + (xx - 256) ** 2) < 200 ** 2).to(device)[None].repeat(GB, 1, 1)
    t = _time_kernel([(lambda f=f: fn.group_affinity_batched(f, cent, gmask, "dot")) for f in feats], flush, per_event=4) / GB
    out.append(hbm_entry("group_scores_kernel + group_affinity_kernel (C4: 768 ch 64^2 -> 512^2 labels, 4 centroids; per image of a "
                         "16-image batched call)", 4 * 768 * 64 * 64 + 512 * 512 * 5, t))
    out[-1]["images_per_s"] = round(1.0 / t, 1)
    del feats

    # ---- encoder FFN linear on the tensor cores (3 tf32 MMAs per fp32 product: ceiling = tf32 dense / 3 ~ bf16 peak / 6)
    rows = N * S
    xs_ = [torch.randn(rows, 256, generator=g).to(device) for _ in range(NSETS)]
    w = torch.randn(1024, 256, generator=g).to(device)
    bias = torch.randn(1024, generator=g).to(device)
    with torch.no_grad():
        t = _time_kernel([(lambda x=x: fn.linear(x, w, bias, relu=True)) for x in xs_], flush)
    flops = 2.0 * rows * 1024 * 256
    ach = flops / t / 1e12
    out.append({"kernel": "gemm_tf32x3_kernel (encoder FFN linear1 43008x256->1024, fp32-equivalent flops)", "bound": "tensor",
                "achieved": round(ach, 1), "peak": tf_peak, "peak_source": how + " bf16 burst", "unit": "TFLOP/s",
                "frac": round(ach / tf_peak, 4), "frac_of_3xtf32_ceiling": round(ach / (tf_peak / 6), 4), "traffic": None,
                "algorithmic_flops": flops, "avg_launch_us": round(t * 1e6, 2)})
    return out


def c3_rooflines(device, per_gpu):
    """configs[2]'s dominant kernel class: the bf16 tcgen05 GEMM (csrc/gemm_bf16.cu) at the Swin-B stage-3 MLP shape of a
    640^2 batch (tokens = per_gpu * 40 * 40, 512 -> 2048, GELU in the epilogue), tensor-bound; measured live."""
    from partdistillation_b200 import functional as fn
    g = torch.Generator().manual_seed(0)
    peak, how = measured_peak()
    try:
        tf_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        tf_peak = 1590.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    rows = per_gpu * 40 * 40
    xs = [torch.randn(rows, 512, generator=g).to(device).to(torch.bfloat16) for _ in range(4)]
    w = torch.randn(2048, 512, generator=g).to(device).to(torch.bfloat16)
    bias = torch.randn(2048, generator=g).to(device)
    t = _time_kernel([(lambda x=x: fn.gemm_bf16(x, w, bias, 2)) for x in xs], flush)
    flops = 2.0 * rows * 2048 * 512
    ach = flops / t / 1e12
    return [{"kernel": f"gemm_bf16_kernel (Swin-B stage-3 MLP fc1 {rows}x512->2048 + GELU, bf16 in / bf16 out)", "bound": "tensor",
             "achieved": round(ach, 1), "peak": tf_peak, "peak_source": how + " bf16 burst", "unit": "TFLOP/s",
             "frac": round(ach / tf_peak, 4), "traffic": None, "algorithmic_flops": flops, "avg_launch_us": round(t * 1e6, 2)}]


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """The reference arm: rank 0 alone times the reference's CPU implementation of the same step (same batch of
    PER_GPU_BATCH images, same freeze, clip and optimizer) on all host cores; the other ranks exit."""
    if rank != 0:
        return
    step, kind, cores, what = make_cpu_reference_step(PER_GPU_BATCH)
    t, done, wdone = time_cpu_reference(step, args.steps, args.warmup)
    ips = PER_GPU_BATCH / t
    line = {"impl": "reference", "metric": METRIC, "value": round(ips, 5), "unit": "images/s", "n_gpus": args.gpus,
            "steps": done, "steps_requested": args.steps, "warmup": wdone,
            "ms_per_step": round(t * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "proposal_learning Swin-B 100q 1024x1024 bs=2 fwd+bwd (BASELINE configs[1]) on the host CPU",
                       "global_batch": PER_GPU_BATCH, "queries": QUERIES, "dec_layers": 10, "train_num_points": POINTS,
                       "importance_sample_ratio": 0.0, "freeze_keys": ["backbone", "encoder"],
                       "optimizer": "AdamW + full-model clip 0.01"},
            "cpu_baseline": {"value": round(ips, 5), "unit": "images/s", "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": round(ips, 5), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_c4(args, rank, local_rank, world, device):
    """BASELINE configs[3]: PixelGroupingModel's dot-affinity grouping (pixel_grouping_model.py:129-218) on synthetic Swin-B
    res3 + res4 features (768 channels at 64 x 64) of 512 x 512 images, 4 centroids; images are independent, so ranks are
    replicas over disjoint images (no data-path collective).  Step = one batch of 16 images per GPU through
    pdb_group_affinity_batched (two launches: score maps at feature resolution, then interpolation + argmax).  value: features, centroids and masks resident; e2e: the batch's object masks copied in from pinned
    host memory and the label maps read back every step."""
    from partdistillation_b200 import functional as fn
    IMGS = 16
    g = torch.Generator().manual_seed(100 + rank)
    feats = [torch.randn(IMGS, 768, 64, 64, generator=g).to(device) for _ in range(3)]   # 3 batches x 201 MB: never L2-resident
    cents = torch.randn(IMGS, 4, 768, generator=g).to(device)
    yy, xx = torch.meshgrid(torch.arange(512), torch.arange(512), indexing="ij")
    host_masks = torch.stack([((yy - 256) ** 2 + (xx - 200 - 7 * i) ** 2) < (150 + 5 * i) ** 2 for i in range(IMGS)]).pin_memory()
    dev_masks = host_masks.to(device)
    host_labels = torch.empty((IMGS, 512, 512), dtype=torch.int32).pin_memory()
    state = {"i": 0}

    def step(masks, read_back):
        labs = fn.group_affinity_batched(feats[state["i"] % len(feats)], cents, masks, "dot")
        state["i"] += 1
        if read_back:
            host_labels.copy_(labs, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(e2e):
        for _ in range(max(args.warmup, 3)):
            step(host_masks.to(device, non_blocking=True) if e2e else dev_masks, e2e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter()
        s_.record()
        for _ in range(args.steps):
            step(host_masks.to(device, non_blocking=True) if e2e else dev_masks, e2e)
        e_.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([s_.elapsed_time(e_), wall], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, _ = timed(False)
    clocks = sampler.stop() if sampler else None
    ms_e, wall_e = timed(True)
    ms_e = max(ms_e, wall_e)
    if rank == 0:
        images = IMGS * world * args.steps
        peak, how = measured_peak()
        alg = 4 * 768 * 64 * 64 + 512 * 512 * 5
        per_launch = ms * 1e-3 / (args.steps * IMGS)
        line = {"metric": "pixel grouping images/sec (res3+res4 dot affinity, 512^2 images; BASELINE configs[3])",
                "value": round(images / (ms * 1e-3), 1), "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "pixel_grouping_model res3/res4 dot-affinity grouping, 768 ch x 64^2 features -> 512^2 labels, "
                                       "4 centroids, 16 images per GPU per step (BASELINE configs[3])", "parallelism": f"replicas x{world}",
                           "l2": "features rotate over 3 batches of 201 MB per GPU (> 126 MB L2)"},
                "clocks": clocks,
                "e2e": {"value": round(images / (ms_e * 1e-3), 1), "unit": "images/s", "h2d_bytes_per_step": int(host_masks.numel()),
                        "d2h_bytes_per_step": int(host_labels.numel() * 4)},
                "gpu_launches": args.steps * 2,
                "roofline": {"kernel": "group_affinity_kernel", "bound": "hbm", "achieved": round(alg / per_launch / 1e9, 1), "peak": peak,
                             "peak_source": how, "unit": "GB/s", "frac": round(alg / per_launch / 1e9 / peak, 4), "traffic": None,
                             "algorithmic_bytes": alg, "avg_launch_us": round(per_launch * 1e6, 2),
                             "note": "per image: step time / 16 (two batched launches per step, timed inside the step)"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default): the training step of BASELINE configs[1]; c3: PartDistillation under bf16 autocast "
                         "(configs[2]); c4: the pixel-grouping throughput sweep of configs[3]")
    ap.add_argument("--per-gpu-batch", type=int, default=0, help="images per GPU per step (default: the workload's own)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true", help="run the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--target-bucket", type=int, default=0,
                    help="pad every image's targets to a multiple of this (engine.DataParallelTrainer target_bucket): batches with "
                         "different target counts / object classes replay one CUDA graph")
    ap.add_argument("--distinct-batches", type=int, default=1,
                    help="cycle the steps over this many batches with different target counts (default 1: the fixed BASELINE batch)")
    ap.add_argument("--packed-masks", action="store_true",
                    help="e2e leg: feed the target masks bit-packed (PackedBitMasks, 1 bit/pixel over PCIe) instead of bools")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the hot path has no CPU implementation)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    if args.workload == "c4":
        run_c4(args, rank, local_rank, world, device)
        return
    from partdistillation_b200 import _lib, compat, presets
    from partdistillation_b200.engine import DataParallelTrainer
    spec = SPECS[args.workload]
    per_gpu = args.per_gpu_batch or spec["batch"]
    cfg = presets.make_cfg(spec["arch"], "swin_b", QUERIES, 10, POINTS, 0.0, device=str(device), **spec["cfg"])
    torch.manual_seed(0)                                   # identical random-init weights on every rank
    model = compat.build_model(cfg)
    model.train()
    trainer = DataParallelTrainer(model, base_lr=1e-4, weight_decay=0.05, clip_norm=0.01,
                                  freeze_keys=("backbone", "encoder"), cuda_graph=not args.no_cuda_graph,
                                  target_bucket=args.target_bucket)
    cpu_sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()} if rank == 0 and args.workload == "c2" else None

    dev_batch = make_batch(rank, per_gpu, device=device, spec=spec)
    host_batch = make_batch(rank, per_gpu, pin=True, packed=args.packed_masks, spec=spec)
    if args.distinct_batches > 1:
        # the steps cycle over batches with different numbers of pseudo masks per image (and, configs[2], other object classes):
        # what a real loader delivers.  Without --target-bucket every distinct batch is its own step signature / CUDA graph.
        def variant(j, **kw):
            if spec["arch"] == "PartDistillationModel":
                return make_batch(rank + 101 * j, per_gpu, spec=spec, **kw)
            return make_batch(rank, per_gpu, spec=spec, k_masks=max(1, K_MASKS - j), **kw)
        dev_batch = [variant(j, device=device) for j in range(args.distinct_batches)]
        host_batch = [variant(j, pin=True, packed=args.packed_masks) for j in range(args.distinct_batches)]
    import contextlib
    amp = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if spec["amp"] else contextlib.nullcontext

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, warmup, read_loss):
        if not isinstance(batches[0], list):
            batches = [batches]
        for i in range(warmup * len(batches)):
            with amp():
                total, _ = trainer.step(batches[i % len(batches)])
            if read_loss:
                total.item()
        barrier()
        prof = os.environ.get("PDB_PROFILE") == "1" and not read_loss
        if prof:
            torch.cuda.profiler.start()        # ncu --profile-from-start off: only the timed steps
        l0 = trainer.pdb_launches
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter()
        s.record()
        last = None
        for i in range(steps):
            with amp():
                total, _ = trainer.step(batches[i % len(batches)])
            if read_loss:
                last = total.item()
        e.record()
        barrier()
        if prof:
            torch.cuda.profiler.stop()
        wall = time.perf_counter() - t0
        ms = s.elapsed_time(e)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), trainer.pdb_launches - l0, last

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, _, launches, _ = timed(dev_batch, args.steps, args.warmup, read_loss=False)
    clocks = sampler.stop() if sampler else None
    # packed masks have another batch signature than dev_batch: their graph is captured during the (untimed) warm-up
    ms_e2e, wall_e2e, _, last_loss = timed(host_batch, args.steps, 4 if args.packed_masks else 1, read_loss=True)
    e2e_ms = max(ms_e2e, wall_e2e)          # the loss read-back makes wall clock the honest end-to-end time

    images = per_gpu * world * args.steps
    line = None
    if rank == 0:
        roofs = kernel_rooflines(device) if args.workload == "c2" else c3_rooflines(device, per_gpu)
        roof = roofs[0]
        line = {"metric": spec["metric"], "value": round(images / (ms * 1e-3), 3), "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
                "config": {"workload": spec["workload"],
                           "global_batch": per_gpu * world, "queries": QUERIES, "dec_layers": 10,
                           "train_num_points": POINTS, "importance_sample_ratio": 0.0,
                           "freeze_keys": ["backbone", "encoder"], "optimizer": "AdamW + full-model clip 0.01", "cuda_graph": not args.no_cuda_graph,
                           **({"target_bucket": args.target_bucket} if args.target_bucket else {}),
                           **({"distinct_batches": args.distinct_batches, "step_signatures": len(trainer._graphs)}
                              if args.distinct_batches > 1 else {}),
                           "parallelism": f"dp{world}", "grad_allreduce_bytes": trainer.grad_bytes,
                           "l2": "per-step working set (>1 GB of activations) exceeds the 126 MB L2; kernel roofline "
                                 "run: 8 back-to-back launches per CUDA-event pair rotating over 4 buffer sets (inputs larger than L2), "
                                 "L2 flushed ahead of each group"},
                "clocks": clocks,
                "e2e": {"value": round(images / (e2e_ms * 1e-3), 3), "unit": "images/s",
                        "h2d_bytes_per_step": batch_bytes(host_batch[0] if isinstance(host_batch[0], list) else host_batch),
                        "d2h_bytes_per_step": 4,
                        **({"packed_masks": True} if args.packed_masks else {}),
                        "ms_per_step": round(e2e_ms / args.steps, 3), "last_loss": last_loss},
                "gpu_launches": int(launches),
                "roofline": roof, "roofline_kernels": roofs}
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline and args.workload == "c2":
            step, kind, cores, what = make_cpu_reference_step(PER_GPU_BATCH, cpu_sd)
            t, _, _ = time_cpu_reference(step, steps=1, warmup=0)
            line["cpu_baseline"] = {"value": round(PER_GPU_BATCH / t, 5), "unit": "images/s", "cores": cores, "kind": kind,
                                    "sample": f"ONE step ({t:.1f} s): {what}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
