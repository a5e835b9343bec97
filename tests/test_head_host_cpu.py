"""CPU: the whole hot path — pixel decoder, masked-attention decoder, criterion; forward and backward — of the PRODUCT modules
on CPU tensors, against the golden tensors of the unmodified reference: the body of tests/test_head_gpu.py::
test_head_and_loss_vs_reference_golden (mask / class logits, attention-mask bits, Hungarian indices, losses, gradients).

Every SIMT operator runs its own kernel source built for the host (tests/host_kernels.py: MSDeformAttn gather / scatter,
masked cross-attention, attention-mask build, point sampling, matcher cost, LSAP, point loss, classifier rows, Layer / Group
norms); the tcgen05 GEMM interface alone is a restatement (tests/host_gemm_abi.py).  This checks the host orchestration and
the kernels together where there is no GPU; it is not a product path (the product raises without CUDA)."""
import pytest
import torch

import test_head_gpu as head_tests
from host_kernels import full_host_library, patch_functional


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    return full_host_library(tmp_path_factory)


@pytest.fixture
def on_host(monkeypatch, host_lib):
    patch_functional(monkeypatch, host_lib)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))       # the wrappers take their kernel paths
    monkeypatch.setattr(head_tests, "DEV", "cpu")


@pytest.mark.parametrize("name", ["pd_micro"])       # PartDistillationModel: the ProposalModel path + the float64 classifier rows
def test_head_and_loss_vs_reference_golden(on_host, golden_dir, name):
    head_tests.test_head_and_loss_vs_reference_golden(golden_dir, name)


@pytest.mark.parametrize("name", ["proposal_micro", "pd_micro"])
def test_padded_targets_are_loss_neutral(on_host, golden_dir, name):
    head_tests.test_padded_targets_are_loss_neutral(golden_dir, name)


def test_batched_matching_equals_per_output(on_host, golden_dir):
    head_tests.test_batched_matching_equals_per_output(golden_dir, "pd_micro")


def test_gpu_body_packed_bit_masks_ingestion(monkeypatch, host_lib):
    """tests/test_postprocess_gpu.py::test_packed_bit_masks_ingestion on the host builds: packed targets stay packed and the
    sampling / point-loss kernels read the words bit-identically to the byte layout."""
    import test_postprocess_gpu as gpu_pp
    fn = patch_functional(monkeypatch, host_lib)
    monkeypatch.setattr(gpu_pp, "DEV", "cpu")
    gpu_pp.test_packed_bit_masks_ingestion(fn)


def _eval_inputs(H, W, out, seed=0):
    from partdistillation_b200.compat import BitMasks, Instances
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    obj = ((yy - H / 2) ** 2 / (H * 0.42) ** 2 + (xx - W / 2) ** 2 / (W * 0.38) ** 2) < 1.0
    lab = torch.randint(0, 3, (H // 16, W // 16), generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1)
    parts = torch.stack([(lab == k) & obj for k in range(3)])
    parts = parts[parts.flatten(1).any(1)]
    inst = Instances((H, W))
    inst.gt_masks = BitMasks(obj[None])
    inst.gt_classes = torch.tensor([4])
    pinst = Instances((H, W))
    pinst.gt_masks = BitMasks(parts)
    pinst.gt_classes = torch.arange(parts.shape[0])
    item = dict(size=(H, W), out=out, object_mask=obj[None], part_masks=parts, part_classes=pinst.gt_classes)
    return {"image": torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8), "instances": inst,
            "part_instances": pinst, "height": out[0], "width": out[1]}, item


@pytest.mark.parametrize("arch,per_pixel", [("ProposalModel", False), ("PartDistillationModel", True)])
def test_eval_forward_end_to_end(on_host, arch, per_pixel):
    """model.eval()(batched_inputs) of the registered meta-architectures with a micro Swin trunk and the real head: images of
    different sizes (padding), one evaluated at a different output size.  The post-processing of the model's OWN head outputs
    is compared with the oracle's restatement of the reference's inference() on those outputs."""
    import m2f_oracle as O
    from partdistillation_b200 import compat, presets
    from postprocess_cases import NEAR, oracle_resize
    Q = 8
    cfg = presets.make_cfg(arch, "swin_micro", num_queries=Q, dec_layers=3, num_points=64, num_object_classes=10,
                           num_part_classes=4, device="cpu")
    cfg.TEST.DETECTIONS_PER_IMAGE = Q
    torch.manual_seed(3)
    model = compat.build_model(cfg).eval()
    pd = arch == "PartDistillationModel"
    if pd:
        model.use_unique_per_pixel_label = per_pixel
        model.mode = ""
        model.fg_score_threshold = 0.0
    else:
        model.use_unique_per_pixel_label = per_pixel
    (x0, it0), (x1, it1) = _eval_inputs(96, 128, (144, 192), 0), _eval_inputs(128, 96, (128, 96), 1)
    captured = {}
    head_forward = model.sem_seg_head.forward

    def capture(*a, **k):
        captured["out"] = head_forward(*a, **k)
        return captured["out"]
    model.sem_seg_head.forward = capture
    with torch.no_grad():
        res = model([x0, x1])
    out = captured["out"]
    logits, masks = out["pred_logits"].float(), out["pred_masks"].float()
    assert tuple(masks.shape) == (2, Q, 32, 32) and len(res) == 2
    padded = (128, 128)
    if pd:
        ref = O.pd_inference(logits, masks, [it0, it1], padded, [4, 4], 4, Q, mappings=None, per_pixel=per_pixel,
                             min_ratio=model.min_pseudo_mask_ratio, min_score=model.min_pseudo_mask_score, fg_thr=0.0)
        key = "predictions"
    else:
        ref = O.proposal_inference(logits, masks, [it0, it1], padded, Q, per_pixel=per_pixel,
                                   min_ratio=model.minimum_pseudo_mask_ratio, min_score=model.minimum_pseudo_mask_score)
        key = "proposals"
    for b, (r, e, it) in enumerate(zip(res, ref, (it0, it1))):
        got = r[key]
        assert tuple(got.pred_masks.shape) == tuple(e["pred_masks"].shape), (got.pred_masks.shape, e["pred_masks"].shape)
        assert torch.allclose(got.scores, e["scores"], rtol=1e-5, atol=1e-7)
        assert torch.equal(got.pred_classes, e["pred_classes"])
        dense = oracle_resize(masks[b], padded, it["size"], it["out"])
        near = (dense.abs() < NEAR).any(0)
        if per_pixel:
            P = logits.shape[-1] - 1
            sc = logits[b].softmax(-1)[:, :-1]
            sc = sc.flatten() if pd else sc.topk(1, dim=1)[0].flatten()
            idx = torch.arange(sc.numel()) // (P if pd else 1)
            tom = O.sem_seg_postprocess(O.pad_masks(it["object_mask"], padded).float(), it["size"], *it["out"]).bool()
            top2 = (sc[:, None, None] * (dense[idx] * tom.sum(0, keepdim=True).bool()).sigmoid()).topk(2, dim=0)[0]
            near = near | ((top2[0] - top2[1]) < 1e-5)
        assert not ((got.pred_masks != e["pred_masks"]) & ~near[None]).any()


@pytest.mark.parametrize("counts", [(3, 0, 2), (20, 1)])
def test_criterion_ragged_batches_vs_oracle(on_host, monkeypatch, counts):
    """The body of tests/test_zz_ragged_criterion_gpu.py on the kernels' host builds: SetCriterion on batches with images
    without targets and with more targets than queries, against the oracle on the same random draws."""
    import test_zz_ragged_criterion_gpu as ragged
    monkeypatch.setattr(ragged, "DEV", "cpu")
    ragged.test_criterion_ragged_batches_vs_oracle(counts)


def test_training_steps_end_to_end(on_host, monkeypatch):
    """DataParallelTrainer.step on the registered ProposalModel (micro Swin trunk, recipe freeze of backbone + encoder, eager
    step) for two iterations with a warm-up schedule: images -> backbone -> head -> criterion -> backward -> clip -> flat
    AdamW, every SIMT kernel from its own source.  Frozen parameters stay put, every trainable one moves, the loss is finite
    and the scheduled rates reach the kernel."""
    from partdistillation_b200 import compat, presets
    from partdistillation_b200.compat import BitMasks, Instances
    from partdistillation_b200.engine import DataParallelTrainer, WarmupMultiStepLR
    cfg = presets.make_cfg("ProposalModel", "swin_micro", num_queries=6, dec_layers=2, num_points=32, device="cpu")
    torch.manual_seed(5)
    model = compat.build_model(cfg).train()
    tr = DataParallelTrainer(model, base_lr=1e-3, weight_decay=0.05, clip_norm=0.01, freeze_keys=("backbone", "encoder"))
    assert tr.flat_param is not None and tr.optimizer is None
    sched = WarmupMultiStepLR([1], gamma=0.1, warmup_factor=0.5, warmup_iters=1)
    tr.set_lr_schedule(sched)
    before = {k: v.detach().clone() for k, v in model.named_parameters()}
    g = torch.Generator().manual_seed(1)
    batch = []
    for i in range(1):
        lab = torch.randint(0, 3, (4, 4), generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1)
        m = torch.stack([lab == k for k in range(3)])
        inst = Instances((64, 64))
        inst.gt_masks = BitMasks(m[m.flatten(1).any(1)])
        inst.gt_classes = torch.zeros(len(inst.gt_masks), dtype=torch.long)
        batch.append({"image": torch.randint(0, 256, (3, 64, 64), generator=g, dtype=torch.uint8), "instances": inst,
                      "height": 64, "width": 64})
    totals = []
    for step in range(2):
        total, losses = tr.step(batch)
        assert torch.allclose(tr.seg_lr, tr.seg_lr_base * sched.factor(step))
        assert set(losses) == {f"loss_{n}{s}" for n in ("ce", "mask", "dice") for s in ("", "_0")}
        totals.append(float(total.detach()))
    assert all(torch.isfinite(torch.tensor(totals))) and tr.iteration == 2 and tr.step_count == 2
    frozen = moved = 0
    for k, v in model.named_parameters():
        if k.startswith("backbone.") or "encoder" in k:
            assert torch.equal(v, before[k]), k
            frozen += 1
        elif v.requires_grad:
            assert not torch.equal(v, before[k]), k
            moved += 1
    assert frozen > 10 and moved > 50


def test_swin_backbone_frozen_path_vs_golden(on_host, golden_dir, monkeypatch):
    """f2 on the CPU tier: the frozen-backbone configuration (fused window-attention kernels from their own source, linears
    on the restated GEMM interface) reproduces the reference Swin outputs of the golden fixture — the body of the GPU test."""
    import test_ops_gpu as gpu_tests
    from partdistillation_b200 import functional, presets
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    make_cfg = presets.make_cfg
    monkeypatch.setattr(presets, "make_cfg", lambda *a, **k: make_cfg(*a, **{**k, "device": "cpu"}))
    gpu_tests.test_swin_backbone_frozen_path_vs_golden(functional, golden_dir)
