"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by executing the UNMODIFIED reference
(/root/reference, imported read-only through oracle/ref_loader.py) on seeded synthetic inputs.

Run in the build container (CPU):  python oracle/make_golden.py
The fixtures travel to the GPU box, where /root/reference does not exist.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader as rl  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
warnings.filterwarnings("ignore")


def golden_msda(ref):
    """ops/test.py protocol (shapes :27-31, seed :34, inputs :39-42) on the reference's production
    arithmetic (ms_deform_attn_core_pytorch), plus a D=32/M=8/L=3 case with autograd gradients."""
    core = ref.ops_func.ms_deform_attn_core_pytorch
    g = {}
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = [(6, 4), (3, 2)]
    S = sum(h * w for h, w in shapes)
    torch.manual_seed(3)
    for tag in ("double", "float"):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        aw = torch.rand(N, Lq, M, L, P) + 1e-5
        aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
        if tag == "double":
            out = core(value.double(), torch.as_tensor(shapes), loc.double(), aw.double())
        else:
            out = core(value, torch.as_tensor(shapes), loc, aw)
        g[f"kat_{tag}"] = dict(shapes=shapes, value=value, loc=loc, attn=aw, out=out)
    # gradient cases: channels per head as in ops/test.py:91 (small ones) through autograd
    for D in (30, 32, 64, 71):
        value = (torch.rand(N, S, M, D) * 0.01).double().requires_grad_()
        loc = torch.rand(N, Lq, M, L, P, 2).double().requires_grad_()
        aw = torch.rand(N, Lq, M, L, P) + 1e-5
        aw = (aw / aw.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().requires_grad_()
        out = core(value, torch.as_tensor(shapes), loc, aw)
        go = torch.rand(out.shape, dtype=torch.float64)
        gv, gl, ga = torch.autograd.grad(out, (value, loc, aw), go)
        g[f"grad_D{D}"] = dict(shapes=shapes, value=value.detach(), loc=loc.detach(), attn=aw.detach(),
                               grad_out=go, out=out.detach(), grad_value=gv, grad_loc=gl, grad_attn=ga)
    # config-like: M=8, D=32, P=4, 3 levels, encoder-style (Lq == S), offsets a few px around centres
    # plus out-of-range locations to exercise the zero padding.
    shapes = [(4, 4), (8, 8), (16, 16)]
    S = sum(h * w for h, w in shapes)
    N, M, D, L, P = 2, 8, 32, 3, 4
    gen = torch.Generator().manual_seed(11)
    value = torch.randn(N, S, M, D, generator=gen).requires_grad_()
    loc = (torch.rand(N, S, M, L, P, 2, generator=gen) * 1.3 - 0.15).requires_grad_()
    aw = torch.softmax(torch.randn(N, S, M, L * P, generator=gen), -1).view(N, S, M, L, P).requires_grad_()
    out = core(value, torch.as_tensor(shapes), loc, aw)
    go = torch.randn(out.shape, generator=gen)
    gv, gl, ga = torch.autograd.grad(out, (value, loc, aw), go)
    g["cfg_like"] = dict(shapes=shapes, value=value.detach(), loc=loc.detach(), attn=aw.detach(), grad_out=go,
                         out=out.detach(), grad_value=gv, grad_loc=gl, grad_attn=ga)
    torch.save(g, os.path.join(OUT, "msda.pt"))
    print("msda.pt", {k: tuple(v["out"].shape) for k, v in g.items()})


HEAD_CASES = {
    # name: (meta_arch, Q, dec_layers, points, importance_ratio, K, B, H, W)
    "proposal_micro": ("ProposalModel", 10, 4, 256, 0.75, 3, 2, 128, 128),
    "proposal_micro_uniform": ("ProposalModel", 12, 3, 192, 0.0, 4, 2, 128, 160),
    "pd_micro": ("PartDistillationModel", 10, 4, 256, 0.75, 3, 2, 128, 128),
}
MICRO_CHANNELS = [32, 64, 128, 256]


def golden_head(ref, name):
    arch, Q, DL, PTS, ratio, K, B, H, W = HEAD_CASES[name]
    pd = arch == "PartDistillationModel"
    cfg = rl.make_cfg(arch, "swin_micro", num_queries=Q, dec_layers=DL, num_points=PTS,
                      importance_sample_ratio=ratio, num_object_classes=50, num_part_classes=8)
    model = rl.build_model(cfg, "/tmp/pd_oracle_work")
    head_sd = {k: v for k, v in model.state_dict().items() if not k.startswith("backbone.")}
    table = synth.table_of(head_sd)
    new_sd = synth.synth_state_dict(table, seed=1)
    missing = model.load_state_dict(new_sd, strict=False)
    assert all(k.startswith("backbone.") or "empty_weight" in k for k in missing.missing_keys), missing
    model.train()
    feats = synth.synth_features(B, H, W, MICRO_CHANNELS, seed=5)
    batch = synth.synth_batch(B, H, W, K, pd, 50, seed=9)
    from detectron2.structures import ImageList, Instances, BitMasks
    bi = []
    for d in batch:
        inst = Instances((H, W))
        inst.gt_masks = BitMasks(d["gt_masks"])
        inst.gt_classes = d["gt_classes"]
        e = {"image": d["image"], "instances": inst, "height": H, "width": W}
        if pd:
            e["gt_object_class"] = d["gt_object_class"]
        bi.append(e)
    il = ImageList(torch.zeros(B, 3, H, W), [(H, W)] * B)
    targets = model.prepare_targets(bi, il)

    attn_masks = []
    pred = model.sem_seg_head.predictor
    orig_fph = pred.forward_prediction_heads

    def fph(*a, **k):
        r = orig_fph(*a, **k)
        attn_masks.append(r[2].clone())
        return r
    pred.forward_prediction_heads = fph
    lsap_calls = []
    orig_lsap = ref.matcher.linear_sum_assignment

    def lsap(C):
        r = orig_lsap(C)
        lsap_calls.append((C.clone(), np.asarray(r[0]).copy(), np.asarray(r[1]).copy()))
        return r
    ref.matcher.linear_sum_assignment = lsap
    matcher_out = []
    orig_mf = model.criterion.matcher.forward

    try:
        with synth.RecordRand() as rr:
            mf, _, ms = model.sem_seg_head.pixel_decoder.forward_features(feats)
            outputs = model.sem_seg_head(feats, mask=targets) if pd else model.sem_seg_head(feats)
            n_head_draws = len(rr.draws)
            indices_all = []
            m = model.criterion.matcher

            class _M(torch.nn.Module):
                def forward(self, o, t):
                    r = m(o, t)
                    indices_all.append([(a.clone(), b.clone()) for a, b in r])
                    return r
            model.criterion.matcher = _M()
            losses = model.criterion(outputs, targets)
            model.criterion.matcher = m
            draws = rr.draws[n_head_draws:]
    finally:
        ref.matcher.linear_sum_assignment = orig_lsap
        pred.forward_prediction_heads = orig_fph
    losses = {k: v * model.criterion.weight_dict[k] for k, v in losses.items()}
    total = sum(losses.values())
    total.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None and not n.startswith("backbone.")}
    keep = ["sem_seg_head.predictor.query_feat.weight", "sem_seg_head.predictor.mask_embed.layers.2.bias",
            "sem_seg_head.predictor.decoder_norm.weight", "sem_seg_head.pixel_decoder.transformer.level_embed",
            "sem_seg_head.pixel_decoder.input_proj.0.0.bias", "sem_seg_head.pixel_decoder.mask_features.bias",
            "sem_seg_head.pixel_decoder.transformer.encoder.layers.0.self_attn.sampling_offsets.bias",
            "sem_seg_head.pixel_decoder.transformer.encoder.layers.5.self_attn.attention_weights.bias",
            "sem_seg_head.predictor.transformer_cross_attention_layers.0.multihead_attn.in_proj_bias",
            "sem_seg_head.predictor.class_embed.bias"]
    g = dict(
        case=dict(arch=arch, Q=Q, dec_layers=DL, points=PTS, importance_ratio=ratio, K=K, B=B, H=H, W=W,
                  channels=MICRO_CHANNELS, weight_seed=1, feature_seed=5, batch_seed=9,
                  num_object_classes=50, num_part_classes=8),
        table=table,
        rand_draws=draws,
        mask_features=mf.detach()[:, ::16, ::2, ::2].clone(),
        mask_features_absmean=mf.detach().abs().mean(),
        multi_scale=[x.detach()[:, ::32].clone() for x in ms],
        pred_logits=[o["pred_logits"].detach() for o in outputs["aux_outputs"]] + [outputs["pred_logits"].detach()],
        pred_masks=[o["pred_masks"].detach() for o in outputs["aux_outputs"]] + [outputs["pred_masks"].detach()],
        decoder_output=outputs["decoder_output"].detach(),
        attn_mask_bits=[np.packbits(a[::8].numpy(), axis=-1) for a in attn_masks],   # heads are replicas
        attn_mask_shapes=[tuple(a.shape) for a in attn_masks],
        lsap=[(c, r, cc) for c, r, cc in lsap_calls],
        indices=indices_all,
        losses={k: v.detach() for k, v in losses.items()},
        grads={k: grads[k] for k in keep if k in grads},
        grad_norms={k: v.double().norm().item() for k, v in grads.items()},
    )
    torch.save(g, os.path.join(OUT, f"head_{name}.pt"))
    sz = os.path.getsize(os.path.join(OUT, f"head_{name}.pt")) / 1e6
    print(f"head_{name}.pt {sz:.2f} MB  total loss {float(total):.6f}  #lsap {len(lsap_calls)}")


def golden_swin(ref):
    cfg = rl.make_cfg("ProposalModel", "swin_micro", num_queries=10, dec_layers=4, num_points=256)
    model = rl.build_model(cfg, "/tmp/pd_oracle_work")
    bb = model.backbone
    table = synth.table_of(bb.state_dict())
    bb.load_state_dict(synth.synth_state_dict(table, seed=2), strict=False)
    bb.eval()
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 96, 128, generator=gen)       # 24x32 tokens: exercises window padding (ws=4 ok) + odd merges
    with torch.no_grad():
        out = bb(x)
    sw = cfg.MODEL.SWIN
    g = dict(table=table, weight_seed=2, x=x, out={k: v for k, v in out.items()},
             cfg=dict(embed_dim=sw.EMBED_DIM, depths=list(sw.DEPTHS), num_heads=list(sw.NUM_HEADS),
                      window_size=sw.WINDOW_SIZE))
    torch.save(g, os.path.join(OUT, "swin_micro.pt"))
    print("swin_micro.pt", {k: tuple(v.shape) for k, v in out.items()})


def golden_pixel_grouping(ref):
    """PixelGroupingModel._prepare_features + generate_part_segments of the UNMODIFIED reference
    (pixel_grouping_model.py:114-125,205-218), with the k-means centroids fixed (get_pixel_grouping stubbed to return
    them) so the fixture pins the up-sample + measure_distance + topk + segment arithmetic, for both metrics."""
    import importlib
    import types
    pg = importlib.import_module("part_distillation.pixel_grouping_model").PixelGroupingModel
    g = torch.Generator().manual_seed(11)
    C3, C4, h, w, H, W, Kc = 16, 24, 12, 10, 48, 40, 4
    feats = {"res3": torch.randn(1, C3, h, w, generator=g), "res4": torch.randn(1, C4, h // 2, w // 2, generator=g)}
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    mask_resized = ((yy - H / 2) ** 2 / (H * 0.4) ** 2 + (xx - W / 2) ** 2 / (W * 0.35) ** 2) < 1.0
    out = {}
    for metric in ("dot", "l2"):
        self = types.SimpleNamespace(backbone_feature_key_list=["res3", "res4"], feature_normalize=False,
                                     distance_metric=metric, num_superpixel_clusters=Kc)
        feature = pg._prepare_features(self, feats)[0]                       # (C3 + C4, h, w)
        centroids = torch.randn(Kc, feature.shape[0], generator=g)
        self.get_pixel_grouping = lambda f, m: centroids
        self.measure_distance = types.MethodType(pg.measure_distance, self)
        resized = torch.nn.functional.interpolate(feature[None], size=(H, W), mode="bilinear", align_corners=False)[0]
        mask_feat = torch.nn.functional.interpolate(mask_resized[None, None].float(), size=(h, w), mode="nearest")[0, 0].bool()
        binary = pg.generate_part_segments(self, {}, feature, resized, mask_feat, mask_resized)
        out[metric] = dict(feature=feature, centroids=centroids, binary_mask=binary)
    out.update(feats=feats, mask_resized=mask_resized)
    torch.save(out, os.path.join(OUT, "pixel_grouping.pt"))
    print("pixel_grouping.pt", {m: tuple(out[m]["binary_mask"].shape) for m in ("dot", "l2")})


INFER_CASES = {
    # name: attribute overrides on the reference ProposalModel (set_postprocess_type / from_config keys,
    # proposal_model.py:55-103,139-168)
    "prop": dict(use_unique_per_pixel_label=False, minimum_pseudo_mask_score=0.0, minimum_pseudo_mask_ratio=0.0,
                 apply_masking_with_object_mask=True),
    "prop_filtered": dict(use_unique_per_pixel_label=False, minimum_pseudo_mask_score=0.3,
                          minimum_pseudo_mask_ratio=0.05, apply_masking_with_object_mask=True),
    "prop_nomask": dict(use_unique_per_pixel_label=False, minimum_pseudo_mask_score=0.0,
                        minimum_pseudo_mask_ratio=0.0, apply_masking_with_object_mask=False),
    "semseg": dict(use_unique_per_pixel_label=True, minimum_pseudo_mask_score=0.0, minimum_pseudo_mask_ratio=0.0,
                   apply_masking_with_object_mask=True),
    "semseg_filtered": dict(use_unique_per_pixel_label=True, minimum_pseudo_mask_score=0.3,
                            minimum_pseudo_mask_ratio=0.05, apply_masking_with_object_mask=True),
}


def synth_inference_inputs(seed=21, Q=12):
    """Two images of different sizes (padding to a multiple of 32), one of them evaluated at a different output
    size (second bilinear pass of sem_seg_postprocess), ground-truth parts inside an elliptic object mask."""
    g = torch.Generator().manual_seed(seed)
    sizes = [(96, 128), (128, 112)]
    outs = [(144, 192), (128, 112)]
    Hp, Wp = 128, 128
    items = []
    for (H, W), (Ho, Wo) in zip(sizes, outs):
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        obj = ((yy - H / 2) ** 2 / (H * 0.42) ** 2 + (xx - W / 2) ** 2 / (W * 0.38) ** 2) < 1.0
        lab = torch.randint(0, 4, (H // 16, W // 16), generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1)
        parts = torch.stack([(lab == k) & obj for k in range(4)])
        parts = parts[parts.flatten(1).any(1)]
        items.append(dict(size=(H, W), out=(Ho, Wo), object_mask=obj[None], part_masks=parts,
                          part_classes=torch.arange(parts.shape[0]) % 3))
    low = torch.randn(2, Q, Hp // 16, Wp // 16, generator=g) * 3.0
    pred_masks = torch.nn.functional.interpolate(low, size=(Hp // 4, Wp // 4), mode="bicubic", align_corners=False)
    pred_masks = pred_masks + 0.3 * torch.randn(2, Q, Hp // 4, Wp // 4, generator=g)
    pred_masks[0, 3] = -4.0                     # a query that claims nothing
    pred_masks[1, 5, :8] = 0.0                  # exact zeros: not > 0
    pred_logits = torch.randn(2, Q, 2, generator=g) * 1.5
    return dict(items=items, padded=(Hp, Wp), pred_masks=pred_masks, pred_logits=pred_logits)


def golden_proposal_inference(ref):
    """ProposalModel eval branch of the UNMODIFIED reference (proposal_model.py:220-302,341-420: inference,
    _prepare_gt_targets, instance_inference, _unique_assignment, match_gt_labels) on synthetic head outputs.
    pycocotools is absent: get_iou_all_cocoapi runs on oracle/shims/pycocotools (restated rleIou semantics)."""
    from detectron2.structures import ImageList, Instances, BitMasks
    Q = 12
    cfg = rl.make_cfg("ProposalModel", "swin_micro", num_queries=Q, dec_layers=3, num_points=64)
    model = rl.build_model(cfg, "/tmp/pd_oracle_work")
    model.eval()
    inp = synth_inference_inputs(Q=Q)
    Hp, Wp = inp["padded"]
    bi = []
    for it in inp["items"]:
        H, W = it["size"]
        inst = Instances((H, W)); inst.gt_masks = BitMasks(it["object_mask"]); inst.gt_classes = torch.zeros(1, dtype=torch.long)
        pinst = Instances((H, W)); pinst.gt_masks = BitMasks(it["part_masks"]); pinst.gt_classes = it["part_classes"]
        bi.append({"image": torch.zeros(3, H, W, dtype=torch.uint8), "instances": inst, "part_instances": pinst,
                   "height": it["out"][0], "width": it["out"][1]})
    il = ImageList(torch.zeros(2, 3, Hp, Wp), [it["size"] for it in inp["items"]])
    outputs = {"pred_logits": inp["pred_logits"], "pred_masks": inp["pred_masks"]}
    g = dict(inputs=inp, cases={})
    for name, over in INFER_CASES.items():
        for k, v in over.items():
            assert hasattr(model, k), k
            setattr(model, k, v)
        with torch.no_grad():
            targets = model.prepare_targets(bi, il)
            res = model.inference(bi, targets, il, outputs, vis=False)
        rec = []
        for r in res:
            p, t = r["proposals"], r["gt_masks"]
            rec.append(dict(pred_masks=np.packbits(p.pred_masks.numpy(), axis=-1), pred_shape=tuple(p.pred_masks.shape),
                            scores=p.scores.clone(), pred_classes=p.pred_classes.clone(),
                            image_size=tuple(p.image_size),
                            gt_masks=np.packbits(t.gt_masks.numpy(), axis=-1), gt_shape=tuple(t.gt_masks.shape),
                            gt_classes=t.gt_classes.clone()))
        g["cases"][name] = dict(overrides=over, results=rec)
    torch.save(g, os.path.join(OUT, "proposal_inference.pt"))
    print("proposal_inference.pt", {n: [r["pred_shape"] for r in c["results"]] for n, c in g["cases"].items()})


PD_INFER_CASES = {
    # name: (mode, attribute overrides on the reference PartDistillationModel; part_distillation_model.py:36-97)
    "prop": ("eval", dict(use_unique_per_pixel_label=False, min_pseudo_mask_score=0.0, min_pseudo_mask_ratio=0.0,
                          apply_masking_with_object_mask=True, use_oracle_classifier=False)),
    "prop_filtered": ("eval", dict(use_unique_per_pixel_label=False, min_pseudo_mask_score=0.2, min_pseudo_mask_ratio=0.9,
                                   apply_masking_with_object_mask=True, use_oracle_classifier=False)),
    "prop_none_valid": ("eval", dict(use_unique_per_pixel_label=False, min_pseudo_mask_score=0.0, min_pseudo_mask_ratio=50.0,
                                     apply_masking_with_object_mask=True, use_oracle_classifier=False)),
    "semseg": ("eval", dict(use_unique_per_pixel_label=True, min_pseudo_mask_score=0.0, min_pseudo_mask_ratio=0.0,
                            apply_masking_with_object_mask=True, use_oracle_classifier=False)),
    "semseg_filtered": ("eval", dict(use_unique_per_pixel_label=True, min_pseudo_mask_score=0.2, min_pseudo_mask_ratio=0.05,
                                     apply_masking_with_object_mask=True, use_oracle_classifier=False)),
    "semseg_oracle_cls": ("", dict(use_unique_per_pixel_label=True, min_pseudo_mask_score=0.0, min_pseudo_mask_ratio=0.0,
                                   apply_masking_with_object_mask=False, use_oracle_classifier=True)),
}


def golden_pd_inference(ref):
    """PartDistillationModel eval branch of the UNMODIFIED reference (part_distillation_model.py:239-283,329-394,
    431-501: inference, _prepare_gt_targets, instance_inference_with_classification,
    _unique_assignment_with_classes, match_gt_labels) on synthetic head outputs."""
    from detectron2.structures import ImageList, Instances, BitMasks
    Q, P, TOPK = 12, 8, 20
    cfg = rl.make_cfg("PartDistillationModel", "swin_micro", num_queries=Q, dec_layers=3, num_points=64,
                      num_object_classes=50, num_part_classes=P)
    cfg.TEST.DETECTIONS_PER_IMAGE = TOPK
    model = rl.build_model(cfg, "/tmp/pd_oracle_work")
    model.eval()
    inp = synth_inference_inputs(seed=23, Q=Q)
    g = torch.Generator().manual_seed(77)
    inp["pred_logits"] = torch.randn(2, Q, P + 1, generator=g) * 2.0
    inp["object_classes"] = [7, 31]
    inp["majority_vote_mapping"] = {7: torch.randint(0, 5, (P,), generator=g), 31: torch.randint(0, 5, (P,), generator=g)}
    inp["topk"] = TOPK
    inp["fg_score_threshold"] = float(model.fg_score_threshold)
    Hp, Wp = inp["padded"]
    bi = []
    for it, oc in zip(inp["items"], inp["object_classes"]):
        H, W = it["size"]
        inst = Instances((H, W)); inst.gt_masks = BitMasks(it["object_mask"]); inst.gt_classes = torch.tensor([oc])
        pinst = Instances((H, W)); pinst.gt_masks = BitMasks(it["part_masks"]); pinst.gt_classes = it["part_classes"]
        bi.append({"image": torch.zeros(3, H, W, dtype=torch.uint8), "instances": inst, "part_instances": pinst,
                   "height": it["out"][0], "width": it["out"][1]})
    il = ImageList(torch.zeros(2, 3, Hp, Wp), [it["size"] for it in inp["items"]])
    outputs = {"pred_logits": inp["pred_logits"], "pred_masks": inp["pred_masks"]}
    import logging
    model.logger = logging.getLogger("oracle")      # the reference never sets it (AttributeError at :192 otherwise)
    model.update_majority_vote_mapping(inp["majority_vote_mapping"])
    out = dict(inputs=inp, cases={})
    for name, (mode, over) in PD_INFER_CASES.items():
        model.mode = mode
        for k, v in over.items():
            assert hasattr(model, k), k
            setattr(model, k, v)
        with torch.no_grad():
            targets = model.prepare_targets(bi, il)
            res = model.inference(bi, targets, il, outputs, vis=False)
        rec = []
        for r in res:
            p, t = r["predictions"], r["gt_instances"]
            rec.append(dict(pred_masks=np.packbits(p.pred_masks.numpy(), axis=-1), pred_shape=tuple(p.pred_masks.shape),
                            scores=p.scores.clone(), pred_classes=p.pred_classes.clone(), image_size=tuple(p.image_size),
                            gt_masks=np.packbits(t.gt_masks.numpy(), axis=-1), gt_shape=tuple(t.gt_masks.shape),
                            gt_classes=t.gt_classes.clone(), gt_object_label=torch.as_tensor(r["gt_object_label"]).clone()))
        out["cases"][name] = dict(mode=mode, overrides=over, results=rec)
    torch.save(out, os.path.join(OUT, "pd_inference.pt"))
    print("pd_inference.pt", {n: [(r["pred_shape"], r["pred_classes"].tolist()) for r in c["results"]] for n, c in out["cases"].items()})


def golden_pixel_grouping_resized(ref):
    """As golden_pixel_grouping, for an evaluation size that differs from the padded size: the features go through the
    reference forward's F.interpolate(-> padded) + sem_seg_postprocess(crop, -> (height, width)) (pixel_grouping_model.py:
    139-160) before generate_part_segments; the object mask through sem_seg_postprocess(...).bool() and the nearest
    down-sampling to the feature size (:158-160)."""
    import importlib
    import types
    from detectron2.modeling.postprocessing import sem_seg_postprocess
    pg = importlib.import_module("part_distillation.pixel_grouping_model").PixelGroupingModel
    g = torch.Generator().manual_seed(13)
    C3, C4, h, w, Kc = 16, 24, 12, 16, 4
    padded, image_size, out_size = (96, 128), (90, 120), (135, 180)
    feats = {"res3": torch.randn(1, C3, h, w, generator=g), "res4": torch.randn(1, C4, h // 2, w // 2, generator=g)}
    yy, xx = torch.meshgrid(torch.arange(image_size[0]), torch.arange(image_size[1]), indexing="ij")
    obj = ((yy - image_size[0] / 2) ** 2 / (image_size[0] * 0.4) ** 2 + (xx - image_size[1] / 2) ** 2 / (image_size[1] * 0.35) ** 2) < 1.0
    masks = torch.zeros(1, *padded)
    masks[0, :image_size[0], :image_size[1]] = obj.float()
    out = {}
    for metric in ("dot", "l2"):
        self = types.SimpleNamespace(backbone_feature_key_list=["res3", "res4"], feature_normalize=False,
                                     distance_metric=metric, num_superpixel_clusters=Kc)
        features = pg._prepare_features(self, feats)                                       # (1, C, h, w)
        features_resized = torch.nn.functional.interpolate(features, size=padded, mode="bilinear", align_corners=False)
        resized = sem_seg_postprocess(features_resized[0], image_size, *out_size)
        mask_resized = sem_seg_postprocess(masks, image_size, *out_size)[0].bool()
        mask_feat = torch.nn.functional.interpolate(masks[None], size=features.shape[-2:], mode="nearest")[0, 0].bool()
        centroids = torch.randn(Kc, features.shape[1], generator=g)
        self.get_pixel_grouping = lambda f, m: centroids
        self.measure_distance = types.MethodType(pg.measure_distance, self)
        binary = pg.generate_part_segments(self, {}, features[0], resized, mask_feat, mask_resized)
        out[metric] = dict(feature=features[0], centroids=centroids, binary_mask=binary, mask_resized=mask_resized,
                           mask_feat=mask_feat)
    out.update(feats=feats, object_mask=obj, geometry=(padded, image_size, out_size))
    torch.save(out, os.path.join(OUT, "pixel_grouping_resized.pt"))
    print("pixel_grouping_resized.pt", {m: tuple(out[m]["binary_mask"].shape) for m in ("dot", "l2")})


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = rl.load()
    if "--pixel-grouping-only" in sys.argv:
        golden_pixel_grouping(ref)
        golden_pixel_grouping_resized(ref)
        return
    if "--pixel-grouping-resized-only" in sys.argv:
        golden_pixel_grouping_resized(ref)
        return
    if "--inference-only" in sys.argv:
        golden_proposal_inference(ref)
        golden_pd_inference(ref)
        return
    golden_proposal_inference(ref)
    golden_pd_inference(ref)
    golden_pixel_grouping(ref)
    golden_pixel_grouping_resized(ref)
    golden_msda(ref)
    golden_swin(ref)
    for name in HEAD_CASES:
        torch.manual_seed(1234)
        golden_head(ref, name)


if __name__ == "__main__":
    main()
