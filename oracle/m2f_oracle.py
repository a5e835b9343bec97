"""TEST INFRASTRUCTURE ONLY — CPU oracle for the Mask2Former training hot path.

A functional restatement (plain torch CPU ops over a ``state_dict``) of the reference
algorithm, written from the reference's behaviour, each function citing the file:line it
follows under /root/reference.  It is pinned against the *unmodified* reference executed
through ``oracle/ref_loader.py`` (see ``oracle/make_golden.py`` and
``tests/test_oracle_golden.py``): golden tensors produced by the reference itself are
committed under ``tests/golden/`` and this file must reproduce them.

Third-party arithmetic not vendored by the reference: detectron2==0.6
``point_sample`` / ``get_uncertain_point_coords_with_randomness`` and scipy
``linear_sum_assignment`` — restated here from their published algorithms
(``point_sample``/``lsap_jv``); no reference test pins them ("parity unpinned" at those two
boundaries; they are cross-checked against the container's torch ``grid_sample`` and SciPy).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` leg may import this module; the product path never does.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------


def linear(sd, prefix, x):
    b = sd.get(prefix + ".bias")
    w = sd[prefix + ".weight"]
    return F.linear(x, w.to(x.dtype), None if b is None else b.to(x.dtype))


def layer_norm(sd, prefix, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"].to(x.dtype), sd[prefix + ".bias"].to(x.dtype), eps)


def group_norm(sd, prefix, x, groups=32, eps=1e-5):
    return F.group_norm(x, groups, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


# --------------------------------------------------------------------------------------
# a11  PositionEmbeddingSine  (transformer_decoder/position_encoding.py:33-56)
# --------------------------------------------------------------------------------------


def position_embedding_sine(B, H, W, num_pos_feats=128, temperature=10000.0, dtype=torch.float32):
    """mask=None, normalize=True, scale=2*pi.  Output (B, 2*num_pos_feats, H, W):
    first half = y embedding, second half = x embedding; even channels sin, odd cos."""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, H + 1, dtype=torch.float32).view(H, 1).expand(H, W)   # cumsum of ones (:38)
    x = torch.arange(1, W + 1, dtype=torch.float32).view(1, W).expand(H, W)   # (:39)
    y = y / (y[-1:, :] + eps) * scale                                        # (:42)
    x = x / (x[:, -1:] + eps) * scale                                        # (:43)
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)  # (:45-46)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)  # (:50-52)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    pos = torch.cat((py, px), dim=2).permute(2, 0, 1)                                  # (:56)
    return pos.unsqueeze(0).expand(B, -1, -1, -1).to(dtype)


# --------------------------------------------------------------------------------------
# bilinear sampling with zero padding, align_corners=False
# (the arithmetic of F.grid_sample as called by ms_deform_attn_func.py:69-70 and by
#  detectron2 point_sample, criterion.py:178-196 / matcher.py:130-140)
# --------------------------------------------------------------------------------------


def _bilinear_zero_pad(img, gx, gy):
    """img (R, C, H, W); gx, gy (R, P) grid coords in [-1, 1] units.  Returns (R, C, P).

    ATen unnormalises with ((g + 1) * size - 1) / 2 and takes the 4 integer neighbours,
    dropping the ones outside the image."""
    R, C, H, W = img.shape
    x = ((gx + 1) * W - 1) / 2
    y = ((gy + 1) * H - 1) / 2
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    wx1 = x - x0
    wy1 = y - y0
    wx0 = 1 - wx1
    wy0 = 1 - wy1
    flat = img.reshape(R, C, H * W)
    out = 0
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = (x0 + dx).long()
            yi = (y0 + dy).long()
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1))
            v = torch.gather(flat, 2, idx[:, None, :].expand(R, C, -1))
            out = out + v * (wy * wx * ok.to(img.dtype))[:, None, :]
    return out


def point_sample(inp, coords):
    """detectron2 point_rend.point_sample(input, coords, align_corners=False):
    inp (R, C, H, W), coords (R, P, 2) as (x, y) in [0, 1] -> (R, C, P)."""
    g = 2.0 * coords - 1.0
    return _bilinear_zero_pad(inp, g[..., 0], g[..., 1])


# --------------------------------------------------------------------------------------
# a5  MSDeformAttn  (ops/functions/ms_deform_attn_func.py:55-75, ops/modules/ms_deform_attn.py:86-131)
# --------------------------------------------------------------------------------------


def ms_deform_attn_core(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N, S, M, D); spatial_shapes [(H, W)]*L; sampling_locations (N, Lq, M, L, P, 2) in
    [0, 1]; attention_weights (N, Lq, M, L, P).  Returns (N, Lq, M*D)."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    grids = 2 * sampling_locations - 1                                   # (:61)
    out = value.new_zeros(N, M, D, Lq)
    start = 0
    for lvl, (H, W) in enumerate(shapes):
        v = value[:, start:start + H * W]                                # (N, HW, M, D)
        start += H * W
        img = v.permute(0, 2, 3, 1).reshape(N * M, D, H, W)              # (:65)
        g = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(N * M, Lq * P, 2)   # (:67)
        s = _bilinear_zero_pad(img, g[..., 0], g[..., 1])                # (N*M, D, Lq*P)  (:69)
        a = attention_weights[:, :, :, lvl].permute(0, 2, 1, 3).reshape(N * M, 1, Lq, P)
        out = out + (s.view(N * M, D, Lq, P) * a).sum(-1).view(N, M, D, Lq)   # (:73-74)
    return out.reshape(N, M * D, Lq).transpose(1, 2).contiguous()


def ms_deform_attn_module(sd, prefix, query, reference_points, input_flatten, spatial_shapes,
                          n_heads=8, n_points=4, core=ms_deform_attn_core):
    """ops/modules/ms_deform_attn.py:98-131 (padding mask None on this path)."""
    N, Lq, C = query.shape
    S = input_flatten.shape[1]
    L = len(spatial_shapes)
    value = linear(sd, prefix + ".value_proj", input_flatten).view(N, S, n_heads, C // n_heads)
    off = linear(sd, prefix + ".sampling_offsets", query).view(N, Lq, n_heads, L, n_points, 2)
    aw = linear(sd, prefix + ".attention_weights", query).view(N, Lq, n_heads, L * n_points)
    aw = F.softmax(aw, -1).view(N, Lq, n_heads, L, n_points)
    normalizer = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=query.dtype)   # (:110)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = core(value, spatial_shapes, loc, aw)
    return linear(sd, prefix + ".output_proj", out)


def encoder_reference_points(spatial_shapes, B, dtype=torch.float32):
    """msdeformattn.py:144-157 with all-valid masks (valid_ratios == 1): pixel centres, the
    same for every level -> (B, S, L, 2) as (x, y)."""
    refs = []
    for (H, W) in spatial_shapes:
        ry = torch.linspace(0.5, H - 0.5, H, dtype=torch.float32)
        rx = torch.linspace(0.5, W - 0.5, W, dtype=torch.float32)
        ry, rx = torch.meshgrid(ry, rx, indexing="ij")
        one = torch.ones(1, dtype=torch.float32)
        refs.append(torch.stack((rx.reshape(-1) / (one * W), ry.reshape(-1) / (one * H)), -1))
    ref = torch.cat(refs, 0)[None].expand(B, -1, -1)
    L = len(spatial_shapes)
    return (ref[:, :, None, :] * torch.ones(1, 1, L, 2)).to(dtype)


def encoder_layer(sd, prefix, src, pos, ref, spatial_shapes, core=ms_deform_attn_core):
    """msdeformattn.py:126-135 (dropout 0): MSDeformAttn -> +res -> LN -> FFN -> +res -> LN."""
    src2 = ms_deform_attn_module(sd, prefix + ".self_attn", src + pos, ref, src, spatial_shapes, core=core)
    src = layer_norm(sd, prefix + ".norm1", src + src2)
    ff = linear(sd, prefix + ".linear2", F.relu(linear(sd, prefix + ".linear1", src)))
    return layer_norm(sd, prefix + ".norm2", src + ff)


# --------------------------------------------------------------------------------------
# a4  MSDeformAttnPixelDecoder.forward_features  (msdeformattn.py:318-362, :65-93)
# --------------------------------------------------------------------------------------


def pixel_decoder_forward_features(sd, prefix, features, enc_layers=6, core=ms_deform_attn_core,
                                   intermediates=None):
    """features: {"res2","res3","res4","res5"} NCHW fp32.  Returns
    (mask_features, out[0], [3 multi-scale features low->high res])."""
    names = ["res5", "res4", "res3"]                       # low -> high resolution (:322)
    srcs, poss, shapes = [], [], []
    B = features["res5"].shape[0]
    for i, n in enumerate(names):
        x = features[n].float()
        y = F.conv2d(x, sd[f"{prefix}.input_proj.{i}.0.weight"], sd[f"{prefix}.input_proj.{i}.0.bias"])
        y = group_norm(sd, f"{prefix}.input_proj.{i}.1", y)
        H, W = y.shape[-2:]
        p = position_embedding_sine(B, H, W)
        shapes.append((H, W))
        srcs.append(y.flatten(2).transpose(1, 2))
        poss.append(p.flatten(2).transpose(1, 2) + sd[f"{prefix}.transformer.level_embed"][i].view(1, 1, -1))
    src = torch.cat(srcs, 1)
    pos = torch.cat(poss, 1)
    ref = encoder_reference_points(shapes, B)
    for l in range(enc_layers):
        src = encoder_layer(sd, f"{prefix}.transformer.encoder.layers.{l}", src, pos, ref, shapes, core=core)
        if intermediates is not None:
            intermediates.append(src)
    out, start = [], 0
    for (H, W) in shapes:                                   # (:331-343)
        out.append(src[:, start:start + H * W].transpose(1, 2).reshape(B, -1, H, W))
        start += H * W
    # one extra FPN level on res2 (:347-355)
    x = features["res2"].float()
    lat = group_norm(sd, f"{prefix}.adapter_1.norm", F.conv2d(x, sd[f"{prefix}.adapter_1.weight"]))
    y = lat + F.interpolate(out[-1], size=lat.shape[-2:], mode="bilinear", align_corners=False)
    y = F.relu(group_norm(sd, f"{prefix}.layer_1.norm", F.conv2d(y, sd[f"{prefix}.layer_1.weight"], padding=1)))
    out.append(y)
    mask_features = F.conv2d(out[-1], sd[f"{prefix}.mask_features.weight"], sd[f"{prefix}.mask_features.bias"])
    return mask_features, out[0], out[:3]


# --------------------------------------------------------------------------------------
# a6/a7  MultiScaleMaskedTransformerDecoder  (mask2former_transformer_decoder.py:370-459)
#        PartDistillationTransformerDecoder  (part_distillation_transformer_decoder.py:141-254)
# --------------------------------------------------------------------------------------


def multihead_attention(sd, prefix, query, key, value, attn_mask=None, nheads=8):
    """nn.MultiheadAttention forward (seq-first): query (Lq, B, C), key/value (Lk, B, C),
    attn_mask bool (B*nheads, Lq, Lk) with True = not allowed."""
    Lq, B, C = query.shape
    Lk = key.shape[0]
    d = C // nheads
    w = sd[prefix + ".in_proj_weight"].to(query.dtype)
    b = sd[prefix + ".in_proj_bias"].to(query.dtype)
    q = F.linear(query, w[:C], b[:C])
    k = F.linear(key, w[C:2 * C], b[C:2 * C])
    v = F.linear(value, w[2 * C:], b[2 * C:])
    q = q.reshape(Lq, B * nheads, d).transpose(0, 1) * (1.0 / math.sqrt(d))
    k = k.reshape(Lk, B * nheads, d).transpose(0, 1)
    v = v.reshape(Lk, B * nheads, d).transpose(0, 1)
    scores = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:
        scores = scores.masked_fill(attn_mask, float("-inf"))
    p = F.softmax(scores, dim=-1)
    o = torch.bmm(p, v).transpose(0, 1).reshape(Lq, B, C)
    return linear(sd, prefix + ".out_proj", o)


def prediction_heads(sd, prefix, output, mask_features, target_size, nheads=8,
                     gt_object_class=None, num_part_classes=None):
    """forward_prediction_heads (:441-459; PD variant part_distillation_transformer_decoder.py:233-254).
    Returns (outputs_class, outputs_mask, attn_mask bool (B*nheads, Q, hw), decoder_output)."""
    dec = layer_norm(sd, prefix + ".decoder_norm", output).transpose(0, 1)            # (B, Q, C)
    if gt_object_class is None:
        cls = linear(sd, prefix + ".class_embed", dec)
    else:
        # fp64 classifier, then keep the object's P columns + the no-object column (:215-230)
        full = linear(sd, prefix + ".class_embed", dec.double())
        P = num_part_classes
        keep = torch.stack([full[i][:, o * P:(o + 1) * P] for i, o in enumerate(gt_object_class)], 0)
        cls = torch.cat([keep, full[:, :, -1:]], -1) + full.sum() * 0
    e = dec
    for i in range(3):
        e = linear(sd, f"{prefix}.mask_embed.layers.{i}", e)
        if i < 2:
            e = F.relu(e)
    B, Q, C = e.shape
    masks = torch.bmm(e, mask_features.flatten(2)).view(B, Q, *mask_features.shape[-2:])   # einsum (:449)
    am = F.interpolate(masks, size=target_size, mode="bilinear", align_corners=False)
    am = (am.sigmoid().flatten(2).unsqueeze(1).repeat(1, nheads, 1, 1).flatten(0, 1) < 0.5).detach()
    return cls, masks, am, dec


def transformer_decoder_forward(sd, prefix, feats, mask_features, num_layers=9, nheads=8,
                                gt_object_class=None, num_part_classes=None, record=None):
    """feats: 3 multi-scale maps (B, C, H_l, W_l) low->high res.  Returns the predictions dict."""
    B = feats[0].shape[0]
    src, pos, sizes = [], [], []
    for i, x in enumerate(feats):
        H, W = x.shape[-2:]
        sizes.append((H, W))
        pos.append(position_embedding_sine(B, H, W).flatten(2).permute(2, 0, 1))
        src.append((x.flatten(2) + sd[prefix + ".level_embed.weight"][i][None, :, None]).permute(2, 0, 1))
    qe = sd[prefix + ".query_embed.weight"].unsqueeze(1).repeat(1, B, 1)
    out = sd[prefix + ".query_feat.weight"].unsqueeze(1).repeat(1, B, 1)
    kw = dict(nheads=nheads, gt_object_class=gt_object_class, num_part_classes=num_part_classes)
    classes, masks = [], []
    c, m, am, dec = prediction_heads(sd, prefix, out, mask_features, sizes[0], **kw)
    classes.append(c); masks.append(m)
    for i in range(num_layers):
        lvl = i % 3
        if record is not None:
            record.setdefault("attn_mask", []).append(am.clone())
        am = am.clone()
        am[am.all(-1)] = False                                                       # (:405)
        p = f"{prefix}.transformer_cross_attention_layers.{i}"
        t2 = multihead_attention(sd, p + ".multihead_attn", out + qe, src[lvl] + pos[lvl], src[lvl], am, nheads)
        out = layer_norm(sd, p + ".norm", out + t2)
        p = f"{prefix}.transformer_self_attention_layers.{i}"
        qk = out + qe
        t2 = multihead_attention(sd, p + ".self_attn", qk, qk, out, None, nheads)
        out = layer_norm(sd, p + ".norm", out + t2)
        p = f"{prefix}.transformer_ffn_layers.{i}"
        t2 = linear(sd, p + ".linear2", F.relu(linear(sd, p + ".linear1", out)))
        out = layer_norm(sd, p + ".norm", out + t2)
        c, m, am, dec = prediction_heads(sd, prefix, out, mask_features, sizes[(i + 1) % 3], **kw)
        classes.append(c); masks.append(m)
    res = {"pred_logits": classes[-1], "pred_masks": masks[-1], "decoder_output": dec,
           "aux_outputs": [{"pred_logits": a, "pred_masks": b} for a, b in zip(classes[:-1], masks[:-1])]}
    if gt_object_class is not None:
        res["query_feats"] = out.permute(1, 0, 2)
    return res


# --------------------------------------------------------------------------------------
# a9  HungarianMatcher  (matcher.py:100-168)
# --------------------------------------------------------------------------------------


def lsap_jv(cost):
    """Rectangular linear sum assignment (shortest augmenting path, Crouse 2016 — the algorithm
    behind scipy.optimize.linear_sum_assignment).  cost (nr, nc) array-like; float64 inside.
    Returns (row_ind, col_ind) sorted by row like SciPy."""
    C = np.asarray(cost, dtype=np.float64)
    transposed = C.shape[0] > C.shape[1]
    if transposed:
        C = C.T
    nr, nc = C.shape
    u = np.zeros(nr); v = np.zeros(nc)
    col4row = -np.ones(nr, dtype=np.int64); row4col = -np.ones(nc, dtype=np.int64)
    for cur in range(nr):
        shortest = np.full(nc, np.inf)
        path = -np.ones(nc, dtype=np.int64)
        SR = np.zeros(nr, dtype=bool); SC = np.zeros(nc, dtype=bool)
        remaining = list(range(nc))[::-1]
        min_val, i, sink = 0.0, cur, -1
        while sink == -1:
            SR[i] = True
            lowest, idx = np.inf, -1
            for it, j in enumerate(remaining):
                r = min_val + C[i, j] - u[i] - v[j]
                if r < shortest[j]:
                    path[j] = i
                    shortest[j] = r
                if shortest[j] < lowest or (shortest[j] == lowest and row4col[j] == -1):
                    lowest = shortest[j]
                    idx = it
            min_val = lowest
            if not np.isfinite(min_val):
                raise ValueError("cost matrix is infeasible")
            j = remaining[idx]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            remaining[idx] = remaining[-1]
            remaining.pop()
        u[cur] += min_val
        for r in range(nr):
            if SR[r] and r != cur:
                u[r] += min_val - shortest[col4row[r]]
        for j in range(nc):
            if SC[j]:
                v[j] -= min_val - shortest[j]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transposed:
        order = np.argsort(col4row, kind="stable")
        return col4row[order], order
    return np.arange(nr), col4row


def matcher_costs(pred_logits_b, pred_masks_b, tgt_labels, tgt_masks, coords, w_class, w_mask, w_dice):
    """Cost matrix for one image (matcher.py:108-158).  coords (1, P, 2)."""
    if pred_logits_b.shape[-1] == 1:
        prob = pred_logits_b.sigmoid()
    else:
        prob = pred_logits_b.softmax(-1)
    cost_class = -prob[:, tgt_labels]
    Q = pred_masks_b.shape[0]
    K = tgt_masks.shape[0]
    t = point_sample(tgt_masks[:, None].to(pred_masks_b.dtype), coords.repeat(K, 1, 1)).squeeze(1).float()
    o = point_sample(pred_masks_b[:, None], coords.repeat(Q, 1, 1)).squeeze(1).float()
    P = o.shape[1]
    pos = F.softplus(-o)            # BCE-with-logits against ones (matcher.py:55-57)
    neg = F.softplus(o)             # against zeros (:58-60)
    cost_mask = (pos @ t.T + neg @ (1 - t).T) / P
    s = o.sigmoid()
    cost_dice = 1 - (2 * (s @ t.T) + 1) / (s.sum(-1)[:, None] + t.sum(-1)[None, :] + 1)
    return w_mask * cost_mask + w_class * cost_class + w_dice * cost_dice


def hungarian_matcher(outputs, targets, num_points, w_class, w_mask, w_dice, rand=torch.rand,
                      record=None, lsap=None):
    """Returns [(idx_pred int64, idx_tgt int64)] per image, ordered by ascending matched cost
    (matcher.py:161-163 — PartDistillation-specific re-ordering)."""
    lsap = lsap or lsap_jv
    out = []
    B = outputs["pred_logits"].shape[0]
    with torch.no_grad():
        for b in range(B):
            coords = rand(1, num_points, 2)
            C = matcher_costs(outputs["pred_logits"][b], outputs["pred_masks"][b],
                              targets[b]["labels"], targets[b]["masks"], coords, w_class, w_mask, w_dice)
            C = C.reshape(C.shape[0], -1).cpu()
            row, col = lsap(C.numpy())
            row = torch.as_tensor(np.asarray(row), dtype=torch.int64, device=C.device)     # CPU indices, as in the reference
            col = torch.as_tensor(np.asarray(col), dtype=torch.int64, device=C.device)
            order = C[row, col].topk(len(row), largest=False)[1]
            if record is not None:
                record.setdefault("cost", []).append(C.clone())
                record.setdefault("coords", []).append(coords.clone())
            out.append((row[order], col[order]))
    return out


# --------------------------------------------------------------------------------------
# a8  SetCriterion  (criterion.py:126-270)
# --------------------------------------------------------------------------------------


def uncertain_point_coords(logits, num_points, oversample_ratio, importance_ratio, rand=torch.rand):
    """detectron2 get_uncertain_point_coords_with_randomness with uncertainty = -|logit|
    (criterion.py:77-91,178-186).  logits (R, 1, H, W).  Two RNG draws, in this order."""
    R = logits.shape[0]
    n = int(num_points * oversample_ratio)
    pc = rand(R, n, 2)
    unc = -point_sample(logits, pc).abs()
    nu = int(importance_ratio * num_points)
    idx = torch.topk(unc[:, 0, :], k=nu, dim=1)[1]
    pc = torch.gather(pc, 1, idx[:, :, None].expand(-1, -1, 2))
    if num_points - nu > 0:
        pc = torch.cat([pc, rand(R, num_points - nu, 2)], 1)
    return pc


def loss_labels(pred_logits, targets, indices, num_classes, eos_coef):
    """criterion.py:126-145."""
    logits = pred_logits.float()
    B, Q, _ = logits.shape
    tc = torch.full((B, Q), num_classes, dtype=torch.int64)
    for b, (i, j) in enumerate(indices):
        tc[b, i] = targets[b]["labels"][j]
    w = torch.ones(num_classes + 1); w[-1] = eos_coef
    return F.cross_entropy(logits.transpose(1, 2), tc, w)


def loss_masks(pred_masks, targets, indices, num_masks, num_points, oversample_ratio,
               importance_ratio, rand=torch.rand, record=None):
    """criterion.py:147-207: point-sampled BCE (mean over points, summed over masks) + dice."""
    src = torch.cat([pred_masks[b, i] for b, (i, _) in enumerate(indices)], 0)[:, None]
    tgt = torch.cat([targets[b]["masks"][j] for b, (_, j) in enumerate(indices)], 0)[:, None].to(src.dtype)
    with torch.no_grad():
        pc = uncertain_point_coords(src, num_points, oversample_ratio, importance_ratio, rand)
        labels = point_sample(tgt, pc).squeeze(1)
    logits = point_sample(src, pc).squeeze(1)
    if record is not None:
        record.setdefault("loss_coords", []).append(pc.clone())
    bce = F.binary_cross_entropy_with_logits(logits, labels, reduction="none").mean(1).sum() / num_masks
    s = logits.sigmoid()
    dice = (1 - (2 * (s * labels).sum(-1) + 1) / (s.sum(-1) + labels.sum(-1) + 1)).sum() / num_masks
    return bce, dice


def set_criterion(outputs, targets, num_classes, num_points_match, num_points_loss,
                  w_class=2.0, w_mask=5.0, w_dice=5.0, eos_coef=0.1, oversample_ratio=3.0,
                  importance_ratio=0.75, world_size=1, rand=torch.rand, record=None):
    """SetCriterion.forward (criterion.py:235-270), unweighted losses dict (30 keys at 10 layers)."""
    num_masks = max(float(sum(len(t["labels"]) for t in targets)) / world_size, 1.0)   # (:248-254)
    losses = {}

    def one(out, suffix):
        idx = hungarian_matcher(out, targets, num_points_match, w_class, w_mask, w_dice, rand, record)
        if record is not None:
            record.setdefault("indices", []).append(idx)
        losses["loss_ce" + suffix] = loss_labels(out["pred_logits"], targets, idx, num_classes, eos_coef)
        bce, dice = loss_masks(out["pred_masks"], targets, idx, num_masks, num_points_loss,
                               oversample_ratio, importance_ratio, rand, record)
        losses["loss_mask" + suffix] = bce
        losses["loss_dice" + suffix] = dice

    one({k: v for k, v in outputs.items() if k != "aux_outputs"}, "")
    for i, aux in enumerate(outputs.get("aux_outputs", [])):
        one(aux, f"_{i}")
    return losses


# --------------------------------------------------------------------------------------
# f2  Swin backbone  (modeling/backbone/swin.py) — eval-mode restatement (DropPath inactive)
# --------------------------------------------------------------------------------------


def _window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _window_reverse(w, ws, H, W):
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def _rel_pos_index(ws):
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0) + (ws - 1)
    return rel[:, :, 0] * (2 * ws - 1) + rel[:, :, 1]


def _shift_mask(Hp, Wp, ws, shift):
    """swin.py:422-447: region ids -> (nW, ws*ws, ws*ws) additive mask of 0 / -100."""
    img = torch.zeros(1, Hp, Wp, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = _window_partition(img, ws).squeeze(-1)
    diff = mw[:, None, :] - mw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def swin_forward(sd, prefix, x, embed_dim, depths, num_heads, window_size, patch_size=4):
    """SwinTransformer.forward (swin.py:655-682) in eval mode: returns {"res2".."res5"} NCHW."""
    ws = window_size
    _, _, H0, W0 = x.shape
    if W0 % patch_size:
        x = F.pad(x, (0, patch_size - W0 % patch_size))
    if H0 % patch_size:
        x = F.pad(x, (0, 0, 0, patch_size - H0 % patch_size))
    x = F.conv2d(x, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], stride=patch_size)
    B, C, H, W = x.shape
    x = layer_norm(sd, prefix + "patch_embed.norm", x.flatten(2).transpose(1, 2))
    rpi = _rel_pos_index(ws).view(-1)
    outs = {}
    for s, depth in enumerate(depths):
        dim = embed_dim * 2 ** s
        nh = num_heads[s]
        Hp = int(math.ceil(H / ws)) * ws
        Wp = int(math.ceil(W / ws)) * ws
        mask = _shift_mask(Hp, Wp, ws, ws // 2)
        for d in range(depth):
            p = f"{prefix}layers.{s}.blocks.{d}"
            shift = 0 if d % 2 == 0 else ws // 2
            h = layer_norm(sd, p + ".norm1", x).view(B, H, W, dim)
            h = F.pad(h, (0, 0, 0, Wp - W, 0, Hp - H))
            if shift:
                h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
            win = _window_partition(h, ws)
            qkv = linear(sd, p + ".attn.qkv", win).reshape(win.shape[0], ws * ws, 3, nh, dim // nh).permute(2, 0, 3, 1, 4)
            q, k, v = qkv[0] * (dim // nh) ** -0.5, qkv[1], qkv[2]
            attn = q @ k.transpose(-2, -1)
            bias = sd[p + ".attn.relative_position_bias_table"][rpi].view(ws * ws, ws * ws, nh).permute(2, 0, 1)
            attn = attn + bias[None]
            if shift:
                nW = mask.shape[0]
                attn = (attn.view(-1, nW, nh, ws * ws, ws * ws) + mask[None, :, None]).view(-1, nh, ws * ws, ws * ws)
            attn = attn.softmax(-1)
            o = (attn @ v).transpose(1, 2).reshape(win.shape[0], ws * ws, dim)
            o = linear(sd, p + ".attn.proj", o)
            h = _window_reverse(o, ws, Hp, Wp)
            if shift:
                h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
            h = h[:, :H, :W].reshape(B, H * W, dim)
            x = x + h
            m = linear(sd, p + ".mlp.fc2", F.gelu(linear(sd, p + ".mlp.fc1", layer_norm(sd, p + ".norm2", x))))
            x = x + m
        o = layer_norm(sd, f"{prefix}norm{s}", x)
        outs[f"res{s + 2}"] = o.view(B, H, W, dim).permute(0, 3, 1, 2).contiguous()
        if s < len(depths) - 1:                                  # PatchMerging (swin.py:316-343)
            p = f"{prefix}layers.{s}.downsample"
            h = x.view(B, H, W, dim)
            if H % 2 or W % 2:
                h = F.pad(h, (0, 0, 0, W % 2, 0, H % 2))
            h = torch.cat([h[:, 0::2, 0::2], h[:, 1::2, 0::2], h[:, 0::2, 1::2], h[:, 1::2, 1::2]], -1)
            H, W = (H + 1) // 2, (W + 1) // 2
            h = layer_norm(sd, p + ".norm", h.view(B, H * W, 4 * dim))
            x = linear(sd, p + ".reduction", h)
    return outs


# --------------------------------------------------------------------------------------
# a1/a2  meta-architecture train branch  (proposal_model.py:177-196, part_distillation_model.py:197-216)
# --------------------------------------------------------------------------------------


def prepare_images(batched_inputs, pixel_mean, pixel_std, size_divisibility=32):
    """(x - mean) / std then zero-pad bottom/right to the batch max rounded up (ImageList.from_tensors)."""
    mean = torch.tensor(pixel_mean).view(-1, 1, 1)
    std = torch.tensor(pixel_std).view(-1, 1, 1)
    imgs = [(x["image"] - mean) / std for x in batched_inputs]
    d = size_divisibility
    H = max(i.shape[-2] for i in imgs); W = max(i.shape[-1] for i in imgs)
    if d > 1:
        H = (H + d - 1) // d * d; W = (W + d - 1) // d * d
    out = imgs[0].new_zeros(len(imgs), 3, H, W)
    for i, im in enumerate(imgs):
        out[i, :, :im.shape[-2], :im.shape[-1]] = im
    return out


def prepare_targets(batched_inputs, H, W, part_distillation=False):
    """_prepare_pseudo_targets (proposal_model.py:313-338; part_distillation_model.py:405-428).
    Each input dict carries "gt_masks" (K, h, w) bool and, for PartDistillation, "gt_classes" and
    "gt_object_class"."""
    tg = []
    for x in batched_inputs:
        m = x["gt_masks"]
        pm = torch.zeros(m.shape[0], H, W, dtype=m.dtype)
        pm[:, :m.shape[1], :m.shape[2]] = m
        if part_distillation:
            tg.append({"labels": x["gt_classes"], "masks": pm, "gt_object_class": int(x["gt_object_class"])})
        else:
            tg.append({"labels": torch.zeros(m.shape[0], dtype=torch.long), "masks": pm})
    return tg


def head_and_loss(sd, features, targets, hp, rand=torch.rand, record=None, core=ms_deform_attn_core):
    """sem_seg_head + criterion + weight_dict scaling from the backbone features onward.
    hp: dict(num_classes, dec_layers, nheads, num_points_match, num_points_loss, w_class, w_mask,
    w_dice, eos_coef, oversample_ratio, importance_ratio, part_distillation, num_part_classes)."""
    mf, _, ms = pixel_decoder_forward_features(sd, "sem_seg_head.pixel_decoder", features, core=core)
    pd = hp.get("part_distillation", False)
    outputs = transformer_decoder_forward(
        sd, "sem_seg_head.predictor", ms, mf, num_layers=hp["dec_layers"] - 1, nheads=hp.get("nheads", 8),
        gt_object_class=[t["gt_object_class"] for t in targets] if pd else None,
        num_part_classes=hp.get("num_part_classes"), record=record)
    if record is not None:
        record["mask_features"] = mf
        record["multi_scale"] = ms
        record["outputs"] = outputs
    losses = set_criterion(outputs, targets, hp["num_classes"], hp["num_points_match"], hp["num_points_loss"],
                           hp["w_class"], hp["w_mask"], hp["w_dice"], hp["eos_coef"], hp["oversample_ratio"],
                           hp["importance_ratio"], 1, rand, record)
    wd = {"loss_ce": hp["w_class"], "loss_mask": hp["w_mask"], "loss_dice": hp["w_dice"]}
    return {k: v * wd[k.split("_")[0] + "_" + k.split("_")[1]] for k, v in losses.items()}


# --------------------------------------------------------------------------------------
# a12  pixel grouping  (pixel_grouping_model.py:139-144 up-sampling, :197-201 measure_distance, :205-218 segments)
# --------------------------------------------------------------------------------------


def pixel_grouping_scores(feature, centroids, out_size, metric="dot", geometry=None):
    """(C, h, w) features -> (Kc, H, W) affinity of every image-resolution pixel to every centroid.  ``geometry`` =
    (padded size, image size, output size): up-sample to the padded size, crop, resize (sem_seg_postprocess, :146-158)."""
    if geometry is None:
        up = F.interpolate(feature[None], size=out_size, mode="bilinear", align_corners=False)[0]     # (:139-144)
    else:
        padded, image_size, out_size = geometry
        up = F.interpolate(feature[None], size=padded, mode="bilinear", align_corners=False)[0]
        up = up[:, :image_size[0], :image_size[1]][None]
        up = F.interpolate(up, size=out_size, mode="bilinear", align_corners=False)[0]
    A = up.flatten(1).t()
    B = centroids
    if metric == "dot":
        d = A @ B.t()                                                                                # (:198-199)
    else:
        d = 2 * A @ B.t() - (A * A).sum(dim=1)[:, None] - (B * B).sum(1, keepdim=True).t()          # (:200-201)
    return d.t().reshape(centroids.shape[0], *out_size)


def pixel_grouping_segments(feature, centroids, mask_resized, metric="dot", geometry=None):
    """generate_part_segments with given centroids (:205-218): -> (label map (H, W) int64, bool (P, H, W))."""
    scores = pixel_grouping_scores(feature, centroids, tuple(mask_resized.shape), metric, geometry)
    labels = torch.zeros(mask_resized.shape, dtype=torch.long)
    labels[mask_resized] = scores[:, mask_resized].argmax(0) + 1
    present = labels[mask_resized].unique()
    return labels, torch.stack([labels == p for p in present]) if len(present) else labels.new_zeros((0, *labels.shape)).bool()


# --------------------------------------------------------------------------------------
# f4  ProposalModel eval branch  (proposal_model.py:220-302 inference / _unique_assignment,
#     :341-366 _prepare_gt_targets, :381-432 instance_inference / match_gt_labels;
#     detectron2 sem_seg_postprocess; pycocotools rleIou via utils/utils.py:35-42)
# --------------------------------------------------------------------------------------


def sem_seg_postprocess(result, img_size, out_h, out_w):
    """detectron2.modeling.postprocessing.sem_seg_postprocess: crop the padding, bilinear resize."""
    result = result[:, :img_size[0], :img_size[1]][None]
    return F.interpolate(result, size=(out_h, out_w), mode="bilinear", align_corners=False)[0]


def pad_masks(masks, padded):
    """_prepare_gt_targets' zero padding to the padded batch size (:349-352,355-358)."""
    out = torch.zeros((masks.shape[0], *padded), dtype=masks.dtype)
    out[:, :masks.shape[1], :masks.shape[2]] = masks
    return out


def mask_iou(pr, gt):
    """get_iou_all_cocoapi (utils/utils.py:35-42): float64 |a & b| / |a | b|, exactly 0 when disjoint (rleIou)."""
    a = pr.flatten(1).to(torch.float64)
    b = gt.flatten(1).to(torch.float64)
    inter = a @ b.t()
    union = a.sum(1)[:, None] + b.sum(1)[None] - inter
    return torch.where(inter > 0, inter / union.clamp(min=1), torch.zeros_like(inter))


def proposal_unique_assignment(masks, scores, per_pixel, min_ratio, min_score):
    """_unique_assignment (:258-302)."""
    obj = masks.topk(1, dim=0)[0] > 0.0
    if per_pixel:
        pred = scores[:, None, None] * masks.sigmoid()
        score_map = pred.topk(1, dim=0)[1]
        ids = score_map.unique()
        new = torch.stack([((score_map == c) & obj)[0] for c in ids]).to(masks.dtype)
        scores = scores[ids]
        valid = new.flatten(1).sum(1) / obj.flatten(1).sum(1) > min_ratio
        if valid.any():
            new, scores = new[valid], scores[valid]
        valid = scores > min_score
        if valid.any():
            new, scores = new[valid], scores[valid]
        return new.bool(), scores
    valid = (masks > 0).flatten(1).sum(1) / obj.flatten(1).sum(1) > min_ratio
    if valid.any():
        masks, scores = masks[valid], scores[valid]
    valid = scores > min_score
    if valid.any():
        masks, scores = masks[valid], scores[valid]
    return masks > 0, scores


def proposal_instance_inference(mask_cls, mask_pred, target_masks, target_object_masks, target_labels, topk,
                                per_pixel=False, min_ratio=0.0, min_score=0.0, gate=True, iou_thr=0.001):
    """instance_inference + match_gt_labels (:381-432) -> (bool masks, scores, gt part labels)."""
    scores = mask_cls.softmax(-1)[:, :-1].topk(1, dim=1)[0].flatten()
    scores, idx = scores.topk(topk, sorted=False)
    mask_pred = mask_pred[idx]
    if gate:
        mask_pred = mask_pred * target_object_masks.sum(dim=0, keepdim=True).bool()         # (:369-376)
    masks, scores = proposal_unique_assignment(mask_pred, scores, per_pixel, min_ratio, min_score)
    ious = mask_iou(masks, target_masks)
    top1, top1_idx = ious.topk(1, dim=1)
    fg = (top1 > iou_thr).flatten()
    labels = target_labels[top1_idx.flatten()[fg]]
    masks, scores = masks[fg], scores[fg]
    if masks.shape[0] == 0:                                                                  # (:402-406)
        masks = torch.zeros((1, *mask_pred.shape[1:]), dtype=torch.bool)
        scores = scores.new_zeros(1)
        labels = labels.new_zeros(1)
    return masks, scores, labels


def proposal_inference(pred_logits, pred_masks, items, padded, topk, **flags):
    """inference (:220-254).  ``items``: per image dict(size, out, object_mask (1,H,W) bool, part_masks (G,H,W) bool,
    part_classes (G,)).  Returns per image dict(pred_masks, scores, pred_classes, gt_masks, gt_classes)."""
    up = F.interpolate(pred_masks, size=padded, mode="bilinear", align_corners=False)
    out = []
    for b, it in enumerate(items):
        size, (oh, ow) = it["size"], it["out"]
        mp = sem_seg_postprocess(up[b], size, oh, ow)
        tm = sem_seg_postprocess(pad_masks(it["part_masks"], padded).float(), size, oh, ow).bool()
        tom = sem_seg_postprocess(pad_masks(it["object_mask"], padded).float(), size, oh, ow).bool()
        masks, scores, labels = proposal_instance_inference(pred_logits[b].to(mp), mp, tm, tom, it["part_classes"],
                                                            topk, **flags)
        out.append(dict(pred_masks=masks, scores=scores, pred_classes=labels, gt_masks=tm,
                        gt_classes=it["part_classes"]))
    return out


# --------------------------------------------------------------------------------------
# f4  PartDistillationModel eval branch  (part_distillation_model.py:239-283 inference,
#     :329-394 match_gt_labels / _unique_assignment_with_classes, :431-501 _prepare_gt_targets /
#     instance_inference_with_classification)
# --------------------------------------------------------------------------------------


def pd_unique_assignment_with_classes(masks, scores, class_labels, per_pixel, min_ratio, min_score):
    """_unique_assignment_with_classes (:346-394), including its quirk in the proposal setting: once the area
    filter keeps anything, the returned masks are ``score * sigmoid(logit) > 0`` instead of ``logit > 0`` (:385-394)."""
    obj = masks.topk(1, dim=0)[0] > 0.0
    pred = scores[:, None, None] * masks.sigmoid()
    if per_pixel:
        score_map = pred.topk(1, dim=0)[1]
        ids = score_map.unique()
        seg = torch.stack([((score_map == c) & obj)[0] for c in ids]).to(masks.dtype)
        scores, class_labels = scores[ids], class_labels[ids]
        new_labels = class_labels.unique()
        new = torch.stack([seg[class_labels == c].sum(dim=0).bool() for c in new_labels]).to(masks.dtype)
        new_scores = torch.stack([scores[class_labels == c].topk(1, dim=0)[0].flatten()[0] for c in new_labels])
        valid = new.flatten(1).sum(1) / obj.flatten(1).sum(1) > min_ratio
        if valid.any():
            new, new_scores, new_labels = new[valid], new_scores[valid], new_labels[valid]
        valid = new_scores > min_score
        if valid.any():
            new, new_scores, new_labels = new[valid], new_scores[valid], new_labels[valid]
        return new.bool(), new_scores, new_labels
    valid = (pred > 0.5).flatten(1).sum(1) / obj.flatten(1).sum(1) > min_ratio
    if valid.any():
        masks, scores, class_labels = pred[valid], scores[valid], class_labels[valid]
    valid = scores > min_score
    if valid.any():
        masks, scores, class_labels = masks[valid], scores[valid], class_labels[valid]
    return masks > 0, scores, class_labels


def pd_instance_inference(mask_cls, mask_pred, target_mask, target_object_mask, target_labels, num_classes, topk,
                          mapping=None, per_pixel=False, min_ratio=0.0, min_score=0.0, gate=True, fg_thr=0.1,
                          oracle_classifier=False):
    """instance_inference_with_classification (:456-501) -> (bool masks, scores, pred classes)."""
    Q = mask_cls.shape[0]
    scores = mask_cls.softmax(-1)[:, :-1]
    labels = torch.arange(num_classes).unsqueeze(0).repeat(Q, 1).flatten(0, 1)
    scores, idx = scores.flatten(0, 1).topk(topk, sorted=False)
    labels = labels[idx]
    if mapping is not None:                                                                   # mode == "eval" (:467-469)
        labels = mapping[labels]
    idx = torch.div(idx, num_classes, rounding_mode="floor")
    mask_pred = mask_pred[idx]
    if gate:
        mask_pred = mask_pred * target_object_mask.sum(dim=0, keepdim=True).bool()           # (:319-326)
    masks, scores, labels = pd_unique_assignment_with_classes(mask_pred, scores, labels, per_pixel, min_ratio, min_score)
    ious = mask_iou(masks, target_mask)                                                       # (:329-343)
    top1, top1_idx = ious.topk(1, dim=1)
    fg = (top1 > fg_thr).flatten()
    gt_labels = target_labels[top1_idx.flatten()[fg]]
    masks, scores, labels = masks[fg], scores[fg], labels[fg]
    if masks.shape[0] == 0:                                                                   # (:481-486)
        masks = torch.zeros((1, *mask_pred.shape[1:]), dtype=torch.bool)
        scores = scores.new_zeros(1)
        labels = scores.new_ones(1).long() * num_classes
        gt_labels = scores.new_ones(1).long() * num_classes
    return masks, scores, (gt_labels if oracle_classifier else labels)


def pd_inference(pred_logits, pred_masks, items, padded, object_classes, num_classes, topk, mappings=None, **flags):
    """inference (:239-283).  ``mappings``: {object class: (num_classes,) long} when mode == "eval", else None."""
    up = F.interpolate(pred_masks, size=padded, mode="bilinear", align_corners=False)
    out = []
    for b, it in enumerate(items):
        size, (oh, ow) = it["size"], it["out"]
        mp = sem_seg_postprocess(up[b], size, oh, ow)
        tm = sem_seg_postprocess(pad_masks(it["part_masks"], padded).float(), size, oh, ow).bool()
        tom = sem_seg_postprocess(pad_masks(it["object_mask"], padded).float(), size, oh, ow).bool()
        mapping = mappings[int(object_classes[b])] if mappings is not None else None
        masks, scores, labels = pd_instance_inference(pred_logits[b].to(mp), mp, tm, tom, it["part_classes"], num_classes,
                                                      topk, mapping=mapping, **flags)
        out.append(dict(pred_masks=masks, scores=scores, pred_classes=labels, gt_masks=tm, gt_classes=it["part_classes"]))
    return out
