"""CPU: host logic of the ProposalModel eval branch (partdistillation_b200/postprocess.py) with the six C-ABI operators
replaced by torch restatements of their contracts (include/pdb200.h, "Inference post-processing").  This checks the
selection / filtering / labelling code around the kernels against the reference's recorded outputs; the kernels
themselves are checked on the GPU (tests/test_postprocess_gpu.py).  The product has no CPU path: without the patch the
same call raises."""
import os

import pytest
import torch

import m2f_oracle as O
from postprocess_cases import oracle_resize, run_case, run_pd_case


def _pack(m):
    R, H, W = m.shape
    Ww = (W + 31) // 32
    p = torch.zeros((R, H, Ww * 32), dtype=torch.int64)
    p[..., :W] = m.to(torch.int64)
    words = (p.view(R, H, Ww, 32) << torch.arange(32)).sum(-1)
    return torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)


def _unpack(bits, width, rows=None):
    b = bits.to(torch.int64) & 0xFFFFFFFF
    out = ((b[..., None] >> torch.arange(32)) & 1).bool().flatten(-2)[..., :width]
    return out if rows is None else out[rows.long()]


def _postprocess_masks(logits, sel, padded, image_size, out_size, gate=None, scores=None, want_bits=True, want_label=False,
                       score_threshold=None):
    v = oracle_resize(logits, padded, image_size, out_size)[sel.long()]
    if gate is not None:
        v = v * gate
    on = v > 0
    bits = _pack(torch.cat([on, on.any(0, keepdim=True)])) if want_bits else None
    label = (scores[:, None, None] * v.sigmoid()).argmax(0).to(torch.int32) if want_label else None
    if score_threshold is not None:
        return bits, label, _pack(scores[:, None, None] * v.sigmoid() > score_threshold)
    return bits, label


def _popcount(bits):
    return _unpack(bits, 32 * bits.shape[-1]).flatten(1).sum(1)


def _iou(a, b):
    return O.mask_iou(_unpack(a, 32 * a.shape[-1]), _unpack(b, 32 * b.shape[-1]))


@pytest.fixture
def torch_ops(monkeypatch):
    from partdistillation_b200 import functional as fn
    monkeypatch.setattr(fn, "postprocess_masks", _postprocess_masks)
    monkeypatch.setattr(fn, "resize_bool_masks", lambda m, i, o: O.sem_seg_postprocess(m.float(), i, *o).bool())
    monkeypatch.setattr(fn, "pack_bits", _pack)
    monkeypatch.setattr(fn, "unpack_bits", _unpack)
    monkeypatch.setattr(fn, "bits_popcount", _popcount)
    monkeypatch.setattr(fn, "bits_iou", _iou)


@pytest.mark.parametrize("case", ["prop", "prop_filtered", "prop_nomask", "semseg", "semseg_filtered"])
def test_eval_branch_host_logic(torch_ops, golden_dir, case):
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    run_case(g, case, "cpu")


PD_CASES = ["prop", "prop_filtered", "prop_none_valid", "semseg", "semseg_filtered", "semseg_oracle_cls"]


@pytest.mark.parametrize("case", PD_CASES)
def test_pd_eval_branch_host_logic(torch_ops, golden_dir, case):
    g = torch.load(os.path.join(golden_dir, "pd_inference.pt"), weights_only=False)
    run_pd_case(g, case, "cpu")


def test_pack_helpers_round_trip():
    m = torch.rand(3, 5, 70, generator=torch.Generator().manual_seed(0)) > 0.5
    assert torch.equal(_unpack(_pack(m), 70), m)
    assert torch.equal(_popcount(_pack(m)), m.flatten(1).sum(1))


def test_eval_branch_has_no_cpu_path(golden_dir):
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    with pytest.raises(RuntimeError, match="CUDA-only"):
        run_case(g, "prop", "cpu")


# ------------------------------------------------------------------ the kernels' per-pixel arithmetic, compiled for the host
GEOMETRIES = [
    # (h, w), padded, image size, output size
    ((32, 32), (128, 128), (96, 128), (144, 192)),
    ((32, 32), (128, 128), (128, 112), (128, 112)),
    ((40, 56), (160, 224), (150, 200), (75, 101)),
    ((40, 56), (160, 224), (150, 200), (333, 517)),
    ((8, 8), (32, 32), (32, 32), (32, 32)),
    ((256, 256), (1024, 1024), (1024, 1024), (1024, 1024)),      # BASELINE configs[1] geometry
]


@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    """tests/native/postprocess_host.cpp + csrc/postprocess_math.cuh built with g++ into a temporary directory."""
    import ctypes
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path_factory.mktemp("pp_host") / "libpp_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++",
                           os.path.join(here, "native", "postprocess_host.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    ints = [ctypes.c_int] * 9
    lib.pp_host_resize.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + ints
    lib.pp_host_resize.restype = None
    lib.pp_host_resize_masks.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + ints[:7]
    lib.pp_host_resize_masks.restype = None
    return lib


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_kernel_pixel_math_matches_oracle(host_math, geom):
    """The composed two-pass bilinear value of postprocess_math.cuh (host build) against F.interpolate o crop o
    F.interpolate: within fp32 interpolation noise everywhere, identical threshold bits outside |v| < 1e-4."""
    (h, w), padded, image_size, out_size = geom
    g = torch.Generator().manual_seed(5)
    K = 3 if h < 100 else 1
    logits = (torch.randn(K, h, w, generator=g) * 2.0).contiguous()
    logits[0, : h // 2] = 0.0
    out = torch.empty(K, *out_size)
    host_math.pp_host_resize(logits.data_ptr(), out.data_ptr(), K, h, w, *padded, *image_size, *out_size)
    ref = oracle_resize(logits, padded, image_size, out_size)
    assert (out - ref).abs().max() < 5e-5
    flips = (out > 0) != (ref > 0)
    assert not (flips & (ref.abs() > 1e-4)).any()
    assert flips.float().mean() < 1e-3
    assert torch.equal(out[0, : out_size[0] // 4] == 0, ref[0, : out_size[0] // 4] == 0)    # exact zeros stay exact


@pytest.mark.parametrize("geom", GEOMETRIES)
def test_kernel_mask_resize_math_matches_oracle(host_math, geom):
    _, padded, image_size, out_size = geom
    g = torch.Generator().manual_seed(2)
    G = 3
    m = torch.zeros(G, *padded, dtype=torch.uint8)
    lab = torch.randint(0, 4, (image_size[0] // 8 + 1, image_size[1] // 8 + 1), generator=g)
    lab = lab.repeat_interleave(8, 0).repeat_interleave(8, 1)[:image_size[0], :image_size[1]]
    for k in range(G - 1):
        m[k, :image_size[0], :image_size[1]] = (lab == k).to(torch.uint8)
    out = torch.empty(G, *out_size, dtype=torch.uint8)
    host_math.pp_host_resize_masks(m.data_ptr(), out.data_ptr(), G, *padded, *image_size, *out_size)
    exp = O.sem_seg_postprocess(m.float(), image_size, *out_size).bool()
    assert torch.equal(out.bool(), exp)


# ------------------------------------------------------------------ the kernels themselves, compiled for the host
@pytest.fixture(scope="module")
def host_kernels(tmp_path_factory):
    """csrc/postprocess_kernels.cuh compiled with g++ through tests/native/cuda_on_cpu.h (one OS thread per CUDA thread,
    std::barrier for the warp / block collectives), exposed with the ctypes signatures of the C ABI minus the stream."""
    import ctypes
    import subprocess
    from partdistillation_b200 import _lib
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path_factory.mktemp("pp_kernels") / "libpp_kernels_host.so")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-pthread", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(here, "native", "postprocess_kernels_host.cpp"), "-o", so])
    cdll = ctypes.CDLL(so)

    class HostLib:
        """Stands in for the loaded libpdb200.so: pdb_<op>(..., stream) -> host_<op>(...)."""
        def pdb_last_error(self):
            return b"host build"

    lib = HostLib()
    for name in ("postprocess_masks", "resize_masks_u8", "pack_bits", "unpack_bits", "bits_popcount", "bits_intersect",
                 "group_affinity_resized", "group_scores"):
        f = getattr(cdll, "host_" + name)
        res, args = _lib.SIGNATURES["pdb_" + name]
        f.restype, f.argtypes = res, args[:-1]
        setattr(lib, "pdb_" + name, (lambda f: lambda *a: f(*a[:-1]))(f))
    return lib


@pytest.fixture
def kernels_on_host(monkeypatch, host_kernels):
    """Runs partdistillation_b200.functional's post-processing wrappers UNMODIFIED on CPU tensors: the library handle
    is the host build of the kernels, the CUDA-only guard and the stream lookup are disabled for the test."""
    from partdistillation_b200 import _lib
    from partdistillation_b200 import functional as fn
    monkeypatch.setattr(_lib, "load", lambda: host_kernels)
    monkeypatch.setattr(fn, "_need_cuda", lambda *a: None)
    monkeypatch.setattr(fn, "_stream", lambda: None)
    return fn


def _words_to_bool(bits, width):
    return _unpack(bits, width)


SMALL_GEOMETRIES = [
    ((8, 8), (32, 32), (32, 32), (32, 32)),             # second pass is the identity, one word per row
    ((8, 10), (32, 40), (28, 36), (42, 54)),            # padding cropped, second pass up, ragged last word
    ((10, 14), (40, 56), (38, 50), (19, 27)),           # second pass down, less than one word per row
]


@pytest.mark.parametrize("geom", SMALL_GEOMETRIES)
@pytest.mark.parametrize("gated", [True, False])
def test_host_built_kernel_postprocess_masks(kernels_on_host, geom, gated):
    fn = kernels_on_host
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
    bits, label, sbits = fn.postprocess_masks(logits, sel, padded, image_size, out_size, gate=gate if gated else None,
                                              scores=scores, want_bits=True, want_label=True, score_threshold=0.5)
    ref = oracle_resize(logits, padded, image_size, out_size)[sel]
    if gated:
        ref = ref * gate
    got = _words_to_bool(bits, out_size[1])
    flips = got[:K] != (ref > 0)
    assert not (flips & (ref.abs() > 1e-4)).any()
    assert not _words_to_bool(bits, 32 * bits.shape[-1])[..., out_size[1]:].any()       # padding bits of ragged words
    assert torch.equal(got[K], got[:K].any(0))
    sm = scores[:, None, None] * ref.sigmoid()
    top2 = sm.topk(2, dim=0)[0]
    assert not ((label.long() != sm.argmax(0)) & ((top2[0] - top2[1]) > 1e-5)).any()
    assert not ((_words_to_bool(sbits, out_size[1]) != (sm > 0.5)) & ((sm - 0.5).abs() > 1e-5)).any()
    assert torch.equal(fn.bits_popcount(bits), got.flatten(1).sum(1))


@pytest.mark.parametrize("shape", [((5, 16, 96), (3, 16, 96)), ((70, 9, 45), (67, 9, 45)), ((2, 36, 2100), (1, 36, 2100))])
def test_host_built_kernels_pack_popcount_iou(kernels_on_host, shape):
    fn = kernels_on_host
    g = torch.Generator().manual_seed(8)
    a = torch.rand(*shape[0], generator=g) > 0.7
    b = torch.rand(*shape[1], generator=g) > 0.4
    a[1] = False
    b[0] = ~a[0]
    pa, pb = fn.pack_bits(a), fn.pack_bits(b)
    assert torch.equal(_unpack(pa, a.shape[-1]), a)
    assert torch.equal(fn.unpack_bits(pa, a.shape[-1]), a)
    rows = torch.tensor([a.shape[0] - 1, 0])
    assert torch.equal(fn.unpack_bits(pa, a.shape[-1], rows), a[rows])
    assert torch.equal(fn.bits_popcount(pa), a.flatten(1).sum(1))
    assert torch.equal(fn.bits_iou(pa, pb), O.mask_iou(a, b))


@pytest.mark.parametrize("case", ["prop_filtered", "semseg_filtered"])
def test_host_built_kernels_proposal_eval_branch(kernels_on_host, golden_dir, case):
    """Host branch + functional wrappers + the kernels' own code (host build) against the reference's recorded outputs."""
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    run_case(g, case, "cpu")


@pytest.mark.parametrize("case", ["prop", "prop_none_valid", "semseg_oracle_cls"])
def test_host_built_kernels_pd_eval_branch(kernels_on_host, golden_dir, case):
    g = torch.load(os.path.join(golden_dir, "pd_inference.pt"), weights_only=False)
    run_pd_case(g, case, "cpu")


# ------------------------------------------------------------------ eval-mode forward() wiring of the registered classes
class _StubBackbone(torch.nn.Module):
    size_divisibility = 32

    def forward(self, x):
        return {"res2": x}


class _StubHead(torch.nn.Module):
    num_classes = 1

    def __init__(self, outputs):
        super().__init__()
        self.outputs = outputs
        self.calls = []

    def forward(self, features, mask=None):
        self.calls.append(mask)
        return self.outputs


def _batched_inputs(inp, object_classes=None):
    from partdistillation_b200.compat import BitMasks, Instances
    bi = []
    for i, it in enumerate(inp["items"]):
        H, W = it["size"]
        inst = Instances((H, W))
        inst.gt_masks = BitMasks(it["object_mask"])
        inst.gt_classes = torch.tensor([object_classes[i]]) if object_classes else torch.zeros(1, dtype=torch.long)
        pinst = Instances((H, W))
        pinst.gt_masks = BitMasks(it["part_masks"])
        pinst.gt_classes = it["part_classes"]
        bi.append({"image": torch.zeros(3, H, W, dtype=torch.uint8), "instances": inst, "part_instances": pinst,
                   "height": it["out"][0], "width": it["out"][1]})
    return bi


def test_proposal_model_eval_forward_wiring(torch_ops, golden_dir):
    """ProposalModel(...).eval()(batched_inputs) -> the reference's list of {"proposals", "gt_masks"} (proposal_model.py:
    205-218), with the backbone / head replaced by stubs that return the golden head outputs."""
    from partdistillation_b200.proposal_model import ProposalModel
    g = torch.load(os.path.join(golden_dir, "proposal_inference.pt"), weights_only=False)
    inp, c = g["inputs"], g["cases"]["prop_filtered"]
    Q = inp["pred_logits"].shape[1]
    head = _StubHead({"pred_logits": inp["pred_logits"], "pred_masks": inp["pred_masks"]})
    model = ProposalModel(backbone=_StubBackbone(), sem_seg_head=head, criterion=torch.nn.Identity(), num_queries=Q,
                          num_classes=1, size_divisibility=32, pixel_mean=(0.0, 0.0, 0.0), pixel_std=(1.0, 1.0, 1.0),
                          test_topk_per_image=Q, use_wandb=False, use_unique_per_pixel_label=False,
                          minimum_pseudo_mask_score=0.3, minimum_pseudo_mask_ratio=0.05)
    model.eval()
    res = model(_batched_inputs(inp))
    assert model.num_test_iterations == 1 and model.num_train_iterations == 0
    assert head.calls == [None]
    for r, ref in zip(res, c["results"]):
        assert set(r) == {"proposals", "gt_masks"}
        assert tuple(r["proposals"].pred_masks.shape) == ref["pred_shape"]
        assert torch.equal(r["proposals"].pred_classes, ref["pred_classes"])
        assert torch.allclose(r["proposals"].scores, ref["scores"])
    model.set_postprocess_type("semseg")
    res = model(_batched_inputs(inp))
    assert [tuple(r["proposals"].pred_masks.shape) for r in res] == [x["pred_shape"] for x in g["cases"]["semseg_filtered"]["results"]]


def test_part_distillation_model_eval_forward_wiring(torch_ops, golden_dir):
    from partdistillation_b200.part_distillation_model import PartDistillationModel
    g = torch.load(os.path.join(golden_dir, "pd_inference.pt"), weights_only=False)
    inp, c = g["inputs"], g["cases"]["semseg_filtered"]
    Q, P = inp["pred_logits"].shape[1], inp["pred_logits"].shape[2] - 1
    head = _StubHead({"pred_logits": inp["pred_logits"], "pred_masks": inp["pred_masks"]})
    model = PartDistillationModel(backbone=_StubBackbone(), sem_seg_head=head, criterion=torch.nn.Identity(), num_queries=Q,
                                  num_classes=P, size_divisibility=32, pixel_mean=(0.0, 0.0, 0.0), pixel_std=(1.0, 1.0, 1.0),
                                  test_topk_per_image=inp["topk"], use_wandb=False, use_unique_per_pixel_label=True,
                                  min_pseudo_mask_score=0.2, min_pseudo_mask_ratio=0.05)
    model.eval()
    model.mode = "eval"
    model.update_majority_vote_mapping(inp["majority_vote_mapping"])
    res = model(_batched_inputs(inp, inp["object_classes"]))
    assert model.current_test_iteration == 1
    assert len(head.calls) == 1 and [int(t["gt_object_class"]) for t in head.calls[0]] == inp["object_classes"]
    for r, ref in zip(res, c["results"]):
        assert set(r) == {"predictions", "gt_instances", "gt_object_label"}
        assert tuple(r["predictions"].pred_masks.shape) == ref["pred_shape"]
        assert torch.equal(r["predictions"].pred_classes, ref["pred_classes"])
        assert torch.allclose(r["predictions"].scores, ref["scores"])


def test_part_distillation_model_save_mode(torch_ops, golden_dir, tmp_path, monkeypatch):
    """mode == "save": pseudo labels as targets (object mask = union of the parts), one record per image on disk in the
    reference's format (part_distillation_model.py:285-306), through a stand-in pycocotools.encode."""
    import sys
    import types
    from partdistillation_b200.part_distillation_model import PartDistillationModel
    g = torch.load(os.path.join(golden_dir, "pd_inference.pt"), weights_only=False)
    inp = g["inputs"]
    Q, P = inp["pred_logits"].shape[1], inp["pred_logits"].shape[2] - 1
    head = _StubHead({"pred_logits": inp["pred_logits"], "pred_masks": inp["pred_masks"]})
    model = PartDistillationModel(backbone=_StubBackbone(), sem_seg_head=head, criterion=torch.nn.Identity(), num_queries=Q,
                                  num_classes=P, size_divisibility=32, pixel_mean=(0.0, 0.0, 0.0), pixel_std=(1.0, 1.0, 1.0),
                                  test_topk_per_image=inp["topk"], use_wandb=False, use_unique_per_pixel_label=True,
                                  train_dataset_name="synthetic", min_pseudo_mask_score=0.0, min_pseudo_mask_ratio=0.0)
    assert model.root_save_path == "pseudo_labels/part_labels/part_distillation_predictions/synthetic/0.0_0.0/"
    model.root_save_path = str(tmp_path)
    model.eval()
    model.mode = "save"
    fake = types.ModuleType("pycocotools.mask")
    fake.encode = lambda m: [{"size": list(m.shape[:2]), "counts": b"rle%d" % int(m.sum())}]
    parent = types.ModuleType("pycocotools")
    parent.mask = fake
    monkeypatch.setitem(sys.modules, "pycocotools", parent)
    monkeypatch.setitem(sys.modules, "pycocotools.mask", fake)
    bi = []
    for i, (x, it) in enumerate(zip(_batched_inputs(inp, inp["object_classes"]), inp["items"])):
        x["instances"] = x.pop("part_instances")                # training-set format: the pseudo part masks are the instances
        x.update(gt_object_class=inp["object_classes"][i], file_name=f"f{i}", image_id=f"id{i}", class_code="n01")
        bi.append(x)
    res = model(bi)
    assert [int(t["gt_object_class"]) for t in head.calls[0]] == inp["object_classes"]
    for i, r in enumerate(res):
        rec = torch.load(os.path.join(str(tmp_path), "n01", f"id{i}"), weights_only=False)
        pred = r["predictions"]
        assert (rec["height"], rec["width"]) == tuple(pred.pred_masks.shape[1:])
        assert len(rec["part_masks"]) == pred.pred_masks.shape[0] and torch.equal(rec["part_labels"], pred.pred_classes)
        assert [p["segmentation"]["counts"] for p in rec["part_masks"]] == ["rle%d" % int(m.sum()) for m in pred.pred_masks]
        assert rec["object_ratio"] == int(pred.pred_masks.sum()) / pred.pred_masks[0].numel()
        assert torch.allclose(rec["part_area_ratios"].sum(), torch.tensor(1.0))


# ------------------------------------------------------------------ pixel grouping at a resized evaluation size
@pytest.mark.parametrize("metric", ["dot", "l2"])
def test_host_built_kernel_group_affinity_resized(kernels_on_host, golden_dir, metric):
    """pdb_group_affinity_resized's kernel (host build) through functional.group_affinity against the segments the
    UNMODIFIED reference produced (tests/golden/pixel_grouping_resized.pt); label flips only at near-ties."""
    fn = kernels_on_host
    g = torch.load(os.path.join(golden_dir, "pixel_grouping_resized.pt"), weights_only=False)
    c = g[metric]
    labels = fn.group_affinity(c["feature"], c["centroids"], c["mask_resized"], metric, geometry=g["geometry"]).long()
    exp_labels, exp_seg = O.pixel_grouping_segments(c["feature"], c["centroids"], c["mask_resized"], metric, geometry=g["geometry"])
    assert torch.equal(exp_seg, c["binary_mask"])
    scores = O.pixel_grouping_scores(c["feature"], c["centroids"], tuple(c["mask_resized"].shape), metric, g["geometry"])
    top2 = scores.topk(2, dim=0)[0]
    bad = labels != exp_labels
    assert not (bad & ((top2[0] - top2[1]) > 1e-3 * scores.abs().max())).any()
    assert bad.float().mean() < 1e-3
    assert (labels[~c["mask_resized"]] == 0).all()


def test_host_built_kernel_group_affinity_same_size_matches_oracle(kernels_on_host, golden_dir):
    """geometry with padded == image == output size takes the older kernel in the product; the resized kernel's
    one-pass branch must agree with the oracle on the same golden (tests/golden/pixel_grouping.pt)."""
    from partdistillation_b200 import _lib
    g = torch.load(os.path.join(golden_dir, "pixel_grouping.pt"), weights_only=False)
    c = g["dot"]
    H, W = g["mask_resized"].shape
    feat, cent = c["feature"].contiguous(), c["centroids"].contiguous()
    mask = g["mask_resized"].to(torch.uint8).contiguous()
    labels = torch.empty((H, W), dtype=torch.int32)
    rc = _lib.load().pdb_group_affinity_resized(feat.data_ptr(), cent.data_ptr(), mask.data_ptr(), labels.data_ptr(),
                                                feat.shape[0], cent.shape[0], feat.shape[1], feat.shape[2], H, W, H, W, H, W, 0, None)
    assert rc == 0
    exp_labels, _ = O.pixel_grouping_segments(c["feature"], c["centroids"], g["mask_resized"], "dot")
    assert (labels.long() != exp_labels).float().mean() < 1e-3


class _FeatureBackbone(torch.nn.Module):
    size_divisibility = 32

    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def forward(self, x):
        return self.feats


def _grouping_inputs(g, small_object=False):
    from partdistillation_b200.compat import BitMasks, Instances
    padded, image_size, out_size = g["geometry"]
    obj = g["object_mask"].clone()
    if small_object:
        obj[:] = False
        obj[40:44, 50:54] = True          # covers < 1 backbone pixel after the nearest down-sampling
    inst = Instances(image_size)
    inst.gt_masks = BitMasks(obj[None])
    pinst = Instances(image_size)
    pinst.gt_masks = BitMasks(torch.stack([obj & (torch.arange(image_size[1])[None] < 60), obj & (torch.arange(image_size[1])[None] >= 60)]))
    return [{"image": torch.zeros(3, *image_size, dtype=torch.uint8), "instances": inst, "part_instances": pinst,
             "height": out_size[0], "width": out_size[1], "file_name": "img0", "file_path": "/x/img0", "class_code": "n01",
             "class_name": "thing", "gt_object_class": 3}]


def test_pixel_grouping_model_resized_forward(kernels_on_host, golden_dir):
    """PixelGroupingModel.forward at an evaluation size != padded size (the case that used to raise): the grouping kernel
    (host build) against the oracle with the model's own k-means centroids."""
    from partdistillation_b200.pixel_grouping_model import PixelGroupingModel
    g = torch.load(os.path.join(golden_dir, "pixel_grouping_resized.pt"), weights_only=False)
    padded, image_size, out_size = g["geometry"]
    model = PixelGroupingModel(backbone=_FeatureBackbone(g["feats"]), size_divisibility=32, pixel_mean=(0.0, 0.0, 0.0),
                               pixel_std=(1.0, 1.0, 1.0), distance_metric="dot", backbone_feature_key_list=["res3", "res4"],
                               num_superpixel_clusters=4)
    model.eval()
    res = model(_grouping_inputs(g))
    assert len(res) == 1 and set(res[0]) == {"proposals", "gt_masks"}
    pm = res[0]["proposals"].pred_masks
    c = g["dot"]
    assert tuple(pm.shape[-2:]) == out_size
    assert torch.equal(pm.any(0), c["mask_resized"]) and int(pm.sum()) == int(c["mask_resized"].sum())    # a partition of the object
    centroids = model.get_pixel_grouping(c["feature"], c["mask_feat"])
    _, seg = O.pixel_grouping_segments(c["feature"], centroids, c["mask_resized"], "dot", geometry=g["geometry"])
    assert pm.shape == seg.shape and (pm != seg).float().mean() < 1e-3
    gt = res[0]["gt_masks"].gt_masks
    assert tuple(gt.shape) == (2, *out_size) and torch.equal(gt.any(0), c["mask_resized"])


def test_proposal_generation_model(kernels_on_host, golden_dir, tmp_path, monkeypatch):
    """ProposalGenerationModel: registered name, None for objects too small to cluster, the reference's on-disk record
    (proposal_generation_model.py:185-199) through a stand-in pycocotools.encode."""
    import sys
    import types
    import partdistillation_b200 as pkg
    from partdistillation_b200.compat import META_ARCH_REGISTRY
    assert META_ARCH_REGISTRY.get("ProposalGenerationModel") is pkg.ProposalGenerationModel
    g = torch.load(os.path.join(golden_dir, "pixel_grouping_resized.pt"), weights_only=False)
    make = lambda path: pkg.ProposalGenerationModel(
        backbone=_FeatureBackbone(g["feats"]), size_divisibility=32, dataset_name="synthetic", pixel_mean=(0.0, 0.0, 0.0),
        pixel_std=(1.0, 1.0, 1.0), distance_metric="l2", backbone_feature_key_list=["res3", "res4"],
        num_superpixel_clusters=4, root_save_path=path).eval()
    model = make(None)
    assert model.generate(_grouping_inputs(g, small_object=True), save=False) == [None]
    res = model.generate(_grouping_inputs(g), save=False)
    assert res[0]["proposals"].pred_masks.any(0).equal(g["l2"]["mask_resized"])
    with pytest.raises(RuntimeError, match="root_save_path"):
        model(_grouping_inputs(g))
    fake = types.ModuleType("pycocotools.mask")
    fake.encode = lambda m: [{"size": list(m.shape[:2]), "counts": b"rle%d" % int(m.sum())}]
    parent = types.ModuleType("pycocotools")
    parent.mask = fake
    monkeypatch.setitem(sys.modules, "pycocotools", parent)
    monkeypatch.setitem(sys.modules, "pycocotools.mask", fake)
    model = make(str(tmp_path))
    assert model(_grouping_inputs(g)) is None
    rec = torch.load(os.path.join(str(tmp_path), "n01", "img0"), weights_only=False)
    assert rec["height"] == g["geometry"][2][0] and rec["width"] == g["geometry"][2][1] and rec["class_index"] == 3
    assert len(rec["part_mask"]) == res[0]["proposals"].pred_masks.shape[0]
    assert all(isinstance(p["segmentation"]["counts"], str) for p in rec["part_mask"])
    assert rec["object_ratio"] == int(g["l2"]["mask_resized"].sum()) / g["l2"]["mask_resized"].numel()


# ------------------------------------------------------------------ the GPU tests' own bodies on the kernels' host build
@pytest.fixture
def gpu_bodies(monkeypatch, kernels_on_host):
    """tests/test_postprocess_gpu.py with Tensor.cuda() = identity and DEV = "cpu": the very assertions (and tolerances)
    the B200 run will evaluate, checked here first against the kernels' host build."""
    import test_postprocess_gpu as gpu_pp
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(gpu_pp, "DEV", "cpu")
    return gpu_pp, kernels_on_host


@pytest.mark.parametrize("geom", [1, 3])
@pytest.mark.parametrize("gated", [True, False])
def test_gpu_body_postprocess_masks_bits_and_label(gpu_bodies, geom, gated):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_postprocess_masks_bits_and_label(fn, gpu_pp.GEOMETRIES[geom], gated)


def test_gpu_body_pack_unpack_popcount_iou(gpu_bodies):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_pack_unpack_popcount_iou(fn, ((5, 64, 96), (3, 64, 96)))


def test_gpu_body_resize_bool_masks(gpu_bodies):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_resize_bool_masks(fn, ((128, 128), (96, 128), (144, 192)))


def test_gpu_body_score_threshold_bits(gpu_bodies):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_score_threshold_bits(fn)


def test_gpu_body_inference_vs_golden(gpu_bodies, golden_dir):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_proposal_inference_vs_golden(fn, golden_dir, "prop")
    gpu_pp.test_pd_inference_vs_golden(fn, golden_dir, "semseg")


@pytest.mark.parametrize("metric", ["dot", "l2"])
def test_gpu_body_group_affinity_resized(gpu_bodies, golden_dir, metric):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_group_affinity_resized_vs_reference_golden(fn, golden_dir, metric)


def test_gpu_body_full_size_properties_at_reduced_size(gpu_bodies):
    gpu_pp, fn = gpu_bodies
    gpu_pp.test_postprocess_full_size_properties(fn, Q=5, S=64, G=3)


# ------------------------------------------------------------------ f3: bit-packed target ingestion
def test_packed_bit_masks_ingestion(kernels_on_host):
    """PackedBitMasks (1 bit / pixel from the mapper) -> the same padded uint8 target buffer, labels and offsets as BitMasks,
    through the device-side unpack kernel (host build); widths that are not multiples of 32, padding, an image without
    targets; and the static-batch clone / copy of the CUDA-graph path keeps the width."""
    from partdistillation_b200.compat import BitMasks, ImageList, Instances, PackedBitMasks
    from partdistillation_b200.engine import _clone_batch, _copy_batch
    from partdistillation_b200.meta_base import Mask2FormerTrainingArch

    class Arch(Mask2FormerTrainingArch):
        pass
    arch = Arch()
    arch._init_common(torch.nn.Identity(), torch.nn.Identity(), torch.nn.Identity(), 4, 1, 32, (0.0, 0.0, 0.0),
                      (1.0, 1.0, 1.0), 4, False)
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
        assert packed[-1]["instances"].gt_masks.image_size == (H, W)
    images = ImageList(torch.zeros(3, 3, 96, 96), sizes)
    ta, tb = arch._prepare_pseudo_targets(plain, images), arch._prepare_pseudo_targets(packed, images)
    assert ta.offsets == tb.offsets == [0, 3, 3, 5]
    # f3: packed targets stay packed (int32 words); expanded they are the BitMasks buffer
    fn = kernels_on_host
    assert tb.packed_masks.dtype == torch.int32 and tuple(tb.packed_masks.shape) == (5, 96, 3)
    assert torch.equal(fn.unpack_bits(tb.packed_masks, 96).to(torch.uint8), ta.packed_masks) and ta.packed_masks.shape == (5, 96, 96)
    assert torch.equal(ta.packed_labels, tb.packed_labels)
    for a, b in zip(ta, tb):
        assert torch.equal(a["masks"], fn.unpack_bits(b["masks"].tensor, 96).bool())
    arch.keep_packed_targets = False
    static = _clone_batch(packed, torch.device("cpu"))
    assert all(s["instances"].gt_masks.width == p["instances"].gt_masks.width for s, p in zip(static, packed))
    assert all(s["instances"].gt_masks.tensor.data_ptr() != p["instances"].gt_masks.tensor.data_ptr() or len(p["instances"].gt_masks) == 0
               for s, p in zip(static, packed))
    for s in static:
        s["instances"].gt_masks.tensor.zero_()
    _copy_batch(static, packed)
    tc = arch._prepare_pseudo_targets(static, images)
    assert torch.equal(tc.packed_masks, ta.packed_masks)
    static_plain = _clone_batch(plain, torch.device("cpu"))                 # BitMasks path of the clone is unchanged
    assert all(type(s["instances"].gt_masks) is BitMasks for s in static_plain)
    with pytest.raises(ValueError):
        PackedBitMasks(torch.zeros(1, 4, 2, dtype=torch.int32), width=20)


# test_packed_bit_masks_ingestion's GPU body also samples the packed words (point_sample / point_loss): its twin lives in
# tests/test_head_host_cpu.py, which builds the loss kernels too
