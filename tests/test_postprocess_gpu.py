"""GPU: the ProposalModel eval branch on bit-packed masks — the post-processing kernels against
the CPU oracle (m2f_oracle.proposal_inference, pinned to goldens of the unmodified reference) through the C ABI, and
the host branch end to end on the golden inputs.

Threshold bits are compared exactly except where the interpolated logit is within fp32 interpolation noise of 0
(|v| < NEAR: the oracle's ATen CPU kernel and this kernel round the source coordinates differently; measured max
difference 1.6e-5 on these shapes)."""
import os
import pytest
import torch
import torch.nn.functional as F

import m2f_oracle as O
from postprocess_cases import NEAR, oracle_resize, run_case, run_pd_case

pytestmark = pytest.mark.gpu
DEV = "cuda"        # tests/test_postprocess_host_cpu.py re-runs these test bodies on the kernels' host build with DEV = "cpu"


@pytest.fixture(scope="module")
def fn():
    from partdistillation_b200 import functional
    return functional


def unpack_words(bits, width):
    """int32 words (..., Ww) on any device -> bool (..., width) with torch ops (independent of pdb_unpack_bits)."""
    b = bits.cpu().to(torch.int64) & 0xFFFFFFFF
    out = ((b[..., None] >> torch.arange(32)) & 1).bool().flatten(-2)
    return out[..., :width]


GEOMETRIES = [
    # (h, w), padded, image size, output size
    ((32, 32), (128, 128), (96, 128), (144, 192)),      # padding cropped + second pass up
    ((32, 32), (128, 128), (128, 112), (128, 112)),     # second pass is the identity, ragged last word (112 = 3.5 words)
    ((40, 56), (160, 224), (150, 200), (75, 101)),      # second pass down, odd width
    ((8, 8), (32, 32), (32, 32), (32, 32)),             # one word per row
]


@pytest.mark.parametrize("geom", GEOMETRIES)
@pytest.mark.parametrize("gated", [True, False])
def test_postprocess_masks_bits_and_label(fn, geom, gated):
    (h, w), padded, image_size, out_size = geom
    g = torch.Generator().manual_seed(5)
    Q, K = 9, 6
    logits = torch.randn(Q, h, w, generator=g) * 2.0
    logits[4] = -3.0
    logits[7, : h // 2] = 0.0
    sel = torch.tensor([7, 0, 4, 2, 8, 5])
    scores = torch.rand(K, generator=g)
    yy, xx = torch.meshgrid(torch.arange(out_size[0]), torch.arange(out_size[1]), indexing="ij")
    gate = ((yy - out_size[0] / 2) ** 2 / (out_size[0] * 0.4) ** 2 + (xx - out_size[1] / 2) ** 2 / (out_size[1] * 0.4) ** 2) < 1
    bits, label = fn.postprocess_masks(logits.cuda(), sel.cuda(), padded, image_size, out_size,
                                       gate=gate.cuda() if gated else None, scores=scores.cuda(), want_bits=True,
                                       want_label=True)
    assert tuple(bits.shape) == (K + 1, out_size[0], (out_size[1] + 31) // 32)
    ref = oracle_resize(logits, padded, image_size, out_size)[sel]
    if gated:
        ref = ref * gate
    got = unpack_words(bits, out_size[1])
    exp = ref > 0
    flips = got[:K] != exp
    assert not (flips & (ref.abs() > NEAR)).any()
    assert flips.float().mean() < 1e-3
    # padding bits of a ragged last word are zero
    full = unpack_words(bits, 32 * bits.shape[-1])
    assert not full[..., out_size[1]:].any()
    # object map row = OR of the rows the kernel wrote (exact), = topk(1, dim=0)[0] > 0 of the oracle up to the flips
    assert torch.equal(got[K], got[:K].any(0))
    # label map: argmax of score * sigmoid; mismatches only at numerical near-ties
    sm = scores[:, None, None] * ref.sigmoid()
    exp_label = sm.argmax(0)
    lab = label.cpu().long()
    bad = lab != exp_label
    top2 = sm.topk(2, dim=0)[0]
    assert not (bad & ((top2[0] - top2[1]) > 1e-5)).any()
    # popcounts and IoU on the packed words against dense counting of the SAME bits
    counts = fn.bits_popcount(bits).cpu()
    assert torch.equal(counts, got.flatten(1).sum(1))


@pytest.mark.parametrize("shape", [((5, 64, 96), (3, 64, 96)), ((70, 33, 45), (67, 33, 45)), ((2, 300, 2100), (1, 300, 2100))])
def test_pack_unpack_popcount_iou(fn, shape):
    g = torch.Generator().manual_seed(8)
    a = torch.rand(*shape[0], generator=g) > 0.7
    b = torch.rand(*shape[1], generator=g) > 0.4
    a[1] = False
    b[0] = ~a[0]
    pa, pb = fn.pack_bits(a.cuda()), fn.pack_bits(b.cuda())
    assert torch.equal(unpack_words(pa, a.shape[-1]), a)
    assert torch.equal(fn.unpack_bits(pa, a.shape[-1]).cpu(), a)
    rows = torch.tensor([a.shape[0] - 1, 0])
    assert torch.equal(fn.unpack_bits(pa, a.shape[-1], rows.cuda()).cpu(), a[rows])
    assert torch.equal(fn.bits_popcount(pa).cpu(), a.flatten(1).sum(1))
    iou = fn.bits_iou(pa, pb).cpu()
    assert iou.dtype == torch.float64
    assert torch.equal(iou, O.mask_iou(a, b))


@pytest.mark.parametrize("geom", [((128, 128), (96, 128), (144, 192)), ((128, 128), (128, 112), (128, 112)),
                                  ((160, 224), (150, 200), (75, 101))])
def test_resize_bool_masks(fn, geom):
    padded, image_size, out_size = geom
    g = torch.Generator().manual_seed(2)
    m = torch.zeros(4, *padded, dtype=torch.bool)
    lab = torch.randint(0, 4, (image_size[0] // 8 + 1, image_size[1] // 8 + 1), generator=g)
    lab = lab.repeat_interleave(8, 0).repeat_interleave(8, 1)[:image_size[0], :image_size[1]]
    for k in range(3):
        m[k, :image_size[0], :image_size[1]] = lab == k           # m[3] stays empty
    got = fn.resize_bool_masks(m.cuda(), image_size, out_size).cpu()
    exp = O.sem_seg_postprocess(m.float(), image_size, *out_size).bool()
    assert got.dtype == torch.bool and got.shape == exp.shape
    assert torch.equal(got, exp)
    assert not got[3].any()


@pytest.mark.parametrize("case", ["prop", "prop_filtered", "prop_nomask", "semseg", "semseg_filtered"])
def test_proposal_inference_vs_golden(fn, golden_dir, case):
    """ProposalInferenceMixin.inference on the golden inputs against the outputs of the UNMODIFIED reference
    (tests/golden/proposal_inference.pt, produced by oracle/make_golden.py)."""
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    run_case(g, case, DEV)


@pytest.mark.parametrize("case", ["prop", "prop_filtered", "prop_none_valid", "semseg", "semseg_filtered", "semseg_oracle_cls"])
def test_pd_inference_vs_golden(fn, golden_dir, case):
    """PartDistillationInferenceMixin.inference against the outputs of the UNMODIFIED reference
    (tests/golden/pd_inference.pt)."""
    g = torch.load(os.path.join(golden_dir, "pd_inference.pt"), weights_only=False)
    run_pd_case(g, case, DEV)


def test_score_threshold_bits(fn):
    """score_bits rows = scores[k] * sigmoid(resized * gate) > thr, against the oracle away from the threshold."""
    g = torch.Generator().manual_seed(12)
    Q, K = 7, 10
    logits = torch.randn(Q, 24, 40, generator=g) * 2.0
    sel = torch.randint(0, Q, (K,), generator=g)              # repeated queries, as (query, class) pairs produce
    scores = torch.rand(K, generator=g) * 0.5 + 0.5
    padded, image_size, out_size = (96, 160), (90, 150), (120, 200)
    gate = torch.rand(*out_size, generator=g) > 0.3
    bits, label, sb = fn.postprocess_masks(logits.cuda(), sel.cuda(), padded, image_size, out_size, gate=gate.cuda(),
                                           scores=scores.cuda(), want_bits=True, want_label=False, score_threshold=0.5)
    assert label is None and tuple(sb.shape) == (K, out_size[0], (out_size[1] + 31) // 32)
    ref = oracle_resize(logits, padded, image_size, out_size)[sel] * gate
    pred = scores[:, None, None] * ref.sigmoid()
    got = unpack_words(sb, out_size[1])
    assert not ((got != (pred > 0.5)) & ((pred - 0.5).abs() > 1e-5)).any()
    _, _, sb0 = fn.postprocess_masks(logits.cuda(), sel.cuda(), padded, image_size, out_size, gate=gate.cuda(),
                                     scores=scores.cuda(), want_bits=False, score_threshold=0.0)
    assert unpack_words(sb0, out_size[1]).all()               # score * sigmoid > 0 everywhere for finite logits
    assert not (unpack_words(bits, out_size[1])[:K] != (ref > 0))[(ref.abs() > NEAR)].any()


def test_postprocess_rejects_bad_arguments(fn):
    x = torch.zeros(3, 8, 8, device="cuda")
    sel = torch.arange(3, device="cuda")
    with pytest.raises(RuntimeError):
        fn.postprocess_masks(x, sel, (32, 32), (40, 32), (32, 32))            # image larger than the padded size
    with pytest.raises(RuntimeError):
        fn.postprocess_masks(x, sel, (32, 32), (32, 32), (32, 32), want_label=True)   # label without scores
    with pytest.raises(RuntimeError):
        fn.postprocess_masks(x.cpu(), sel.cpu(), (32, 32), (32, 32), (32, 32))        # CUDA only


# ------------------------------------------------------------------ pixel grouping at a resized evaluation size
@pytest.mark.parametrize("metric", ["dot", "l2"])
def test_group_affinity_resized_vs_reference_golden(fn, golden_dir, metric):
    """pdb_group_affinity_resized against the segments of the UNMODIFIED reference at padded (96, 128) / image (90, 120) /
    output (135, 180) (tests/golden/pixel_grouping_resized.pt); label flips only at numerical near-ties."""
    g = torch.load(os.path.join(golden_dir, "pixel_grouping_resized.pt"), weights_only=False)
    c = g[metric]
    labels = fn.group_affinity(c["feature"].cuda(), c["centroids"].cuda(), c["mask_resized"].cuda(), metric,
                               geometry=g["geometry"]).cpu().long()
    exp_labels, exp_seg = O.pixel_grouping_segments(c["feature"], c["centroids"], c["mask_resized"], metric, geometry=g["geometry"])
    assert torch.equal(exp_seg, c["binary_mask"])
    scores = O.pixel_grouping_scores(c["feature"], c["centroids"], tuple(c["mask_resized"].shape), metric, g["geometry"])
    top2 = scores.topk(2, dim=0)[0]
    bad = labels != exp_labels
    assert not (bad & ((top2[0] - top2[1]) > 1e-3 * scores.abs().max())).any()
    assert bad.float().mean() < 1e-3
    assert (labels[~c["mask_resized"]] == 0).all()


def test_group_affinity_resized_equals_one_pass_kernel(fn, golden_dir):
    """With padded == image == output size the general-geometry kernel and pdb_group_affinity label the same pixels
    (they differ only in FMA contraction of the interpolation)."""
    from partdistillation_b200 import _lib
    g = torch.load(os.path.join(golden_dir, "pixel_grouping.pt"), weights_only=False)
    c = g["dot"]
    H, W = g["mask_resized"].shape
    feat, cent = c["feature"].cuda().contiguous(), c["centroids"].cuda().contiguous()
    mask = g["mask_resized"].cuda()
    old = fn.group_affinity(feat, cent, mask, "dot")
    new = torch.empty_like(old)
    rc = _lib.load().pdb_group_affinity_resized(feat.data_ptr(), cent.data_ptr(), mask.view(torch.uint8).data_ptr(),
                                                new.data_ptr(), feat.shape[0], cent.shape[0], feat.shape[1], feat.shape[2],
                                                H, W, H, W, H, W, 0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert (old != new).float().mean() < 1e-3


def test_pixel_grouping_model_resized_forward(fn):
    """The registered PixelGroupingModel on a micro Swin trunk with an image that needs padding (120 x 104 -> 128 x 128) and
    an evaluation size of 1.5x: the proposals partition the resized object mask exactly."""
    from partdistillation_b200 import compat, presets
    from partdistillation_b200.config import add_pixel_grouping_confing
    cfg = presets.make_cfg("PixelGroupingModel", "swin_micro", device="cuda")
    add_pixel_grouping_confing(cfg)
    cfg.PIXEL_GROUPING.DISTANCE_METRIC = "l2"
    cfg.PIXEL_GROUPING.BACKBONE_FEATURE_KEY_LIST = ["res3", "res4"]
    torch.manual_seed(0)
    model = compat.build_model(cfg).eval()
    H, W, Ho, Wo = 120, 104, 180, 156
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    obj = (((yy - 60) ** 2 + (xx - 52) ** 2) < 36 ** 2)[None]
    inst = compat.Instances((H, W))
    inst.gt_masks = compat.BitMasks(obj)
    inst.gt_classes = torch.zeros(1, dtype=torch.long)
    img = torch.randint(0, 256, (3, H, W), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
    out = model([{"image": img, "instances": inst, "height": Ho, "width": Wo}])
    pm = out[0]["proposals"].pred_masks.cpu()
    padded = torch.zeros(1, 128, 128)
    padded[:, :H, :W] = obj.float()
    exp_obj = O.sem_seg_postprocess(padded, (H, W), Ho, Wo)[0].bool()
    assert pm.dtype == torch.bool and tuple(pm.shape[1:]) == (Ho, Wo) and 1 <= pm.shape[0] <= 4
    assert torch.equal(pm.any(0), exp_obj) and int(pm.sum()) == int(exp_obj.sum())


def test_postprocess_full_size_properties(fn, Q=100, S=1024, G=6):
    """BASELINE geometry (Q = 100, 256^2 logits -> 1024^2): the packed path against the reference expression evaluated by
    PyTorch on the same GPU (F.interpolate -> gate -> > 0; dense areas and IoU), plus size-independent properties."""
    g = torch.Generator().manual_seed(3)
    low = torch.randn(Q, S // 16, S // 16, generator=g) * 3
    logits = (F.interpolate(low[None], size=(S // 4, S // 4), mode="bicubic")[0]
              + 0.3 * torch.randn(Q, S // 4, S // 4, generator=g)).cuda()
    scores = torch.rand(Q, generator=g).cuda()
    sel = torch.randperm(Q, generator=g).cuda()
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    gate = (((yy - S / 2) ** 2 + (xx - S / 2) ** 2) < (0.4 * S) ** 2).cuda()
    lab = torch.randint(0, G, (S // 16, S // 16), generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1)
    gt = torch.stack([lab == k for k in range(G)]).cuda() & gate
    bits, label = fn.postprocess_masks(logits, sel, (S, S), (S, S), (S, S), gate=gate, scores=scores, want_bits=True,
                                       want_label=True)
    dense = F.interpolate(logits[sel][None], size=(S, S), mode="bilinear", align_corners=False)[0] * gate
    masks = fn.unpack_bits(bits, S)
    flips = masks[:Q] != (dense > 0)
    assert not (flips & (dense.abs() > NEAR)).any()
    assert flips.float().mean() < 1e-4
    assert torch.equal(masks[Q], masks[:Q].any(0))                                   # object map = OR of the candidates
    assert not (masks[:Q] & ~gate).any()                                             # nothing outside the gate
    counts = fn.bits_popcount(bits)
    assert torch.equal(counts, masks.flatten(1).sum(1))
    a = masks[:Q].flatten(1).double()
    b = gt.flatten(1).double()
    inter = a @ b.t()
    union = a.sum(1)[:, None] + b.sum(1)[None] - inter
    exp_iou = torch.where(inter > 0, inter / union.clamp(min=1), torch.zeros_like(inter))
    assert torch.equal(fn.bits_iou(bits[:Q], fn.pack_bits(gt)), exp_iou)
    assert torch.equal(fn.unpack_bits(fn.pack_bits(gt), S), gt)                      # pack / unpack round trip
    # every pixel is owned by exactly one candidate; inside the object the owner's own logit is the largest score * sigmoid
    sm = scores[:, None, None] * dense.sigmoid()
    top2 = sm.topk(2, dim=0)[0]
    bad = label.long() != sm.argmax(0)
    assert not (bad & ((top2[0] - top2[1]) > 1e-5)).any()


def test_packed_bit_masks_ingestion(fn):
    """f3: PackedBitMasks targets (1 bit / pixel over PCIe) give the same padded target buffer as BitMasks bools — widths
    that are not multiples of 32, padding, an image without targets (device-side pdb_unpack_bits)."""
    from partdistillation_b200.compat import BitMasks, ImageList, Instances, PackedBitMasks
    from partdistillation_b200.meta_base import Mask2FormerTrainingArch

    class Arch(Mask2FormerTrainingArch):
        pass
    arch = Arch()
    arch._init_common(torch.nn.Identity(), torch.nn.Identity(), torch.nn.Identity(), 4, 1, 32, (0.0, 0.0, 0.0),
                      (1.0, 1.0, 1.0), 4, False)
    arch.to(DEV)
    g = torch.Generator().manual_seed(4)
    sizes, counts = [(70, 45), (64, 96), (50, 33)], [3, 0, 2]
    plain, packed = [], []
    for (H, W), k in zip(sizes, counts):
        m = torch.rand(k, H, W, generator=g) > 0.5
        for store, masks in ((plain, BitMasks(m)), (packed, PackedBitMasks.from_bool(m))):
            inst = Instances((H, W))
            inst.gt_masks = masks
            inst.gt_classes = torch.zeros(k, dtype=torch.long)
            store.append({"image": torch.zeros(3, H, W), "instances": inst})
    images = ImageList(torch.zeros(3, 3, 96, 96, device=DEV), sizes)
    ta, tb = arch._prepare_pseudo_targets(plain, images), arch._prepare_pseudo_targets(packed, images)
    assert ta.offsets == tb.offsets == [0, 3, 3, 5]
    assert str(tb.packed_masks.device).startswith(DEV)
    # f3: the packed targets STAY packed (int32 words, 1/8 of the bytes) ...
    assert tb.packed_masks.dtype == torch.int32 and tuple(tb.packed_masks.shape) == (5, 96, 3)
    assert torch.equal(fn.unpack_bits(tb.packed_masks, 96).to(torch.uint8), ta.packed_masks)
    assert torch.equal(ta.packed_labels, tb.packed_labels)
    # ... and sampling from the words gives the same values as sampling from the bytes, bit for bit
    g2 = torch.Generator().manual_seed(9)
    coords = (torch.rand(5, 700, 2, generator=g2) * 1.1 - 0.05).to(DEV)
    assert torch.equal(fn.point_sample(ta.packed_masks, coords), fn.point_sample(tb.packed_masks, coords))
    pred = torch.randn(7, 24, 24, generator=g2).to(DEV).requires_grad_()
    pi = torch.tensor([0, 3, 6, 2, 5], device=DEV)
    gi = torch.tensor([4, 0, 2, 1, 3], device=DEV)
    la = fn.point_loss(pred, pi, ta.packed_masks, gi, coords)
    lb = fn.point_loss(pred, pi, tb.packed_masks, gi, coords)
    assert torch.equal(la[0], lb[0]) and torch.equal(la[1], lb[1])
    ga = torch.autograd.grad(la[0].sum() + la[1].sum(), pred)[0]
    gb = torch.autograd.grad(lb[0].sum() + lb[1].sum(), pred)[0]
    assert torch.allclose(ga, gb, rtol=1e-6, atol=1e-7)            # same values; atomics order only
    # the expanded layout is still available
    arch.keep_packed_targets = False
    tc = arch._prepare_pseudo_targets(packed, images)
    assert tc.packed_masks.dtype == torch.uint8 and torch.equal(tc.packed_masks, ta.packed_masks)
