"""Hungarian matcher (reference: modeling/matcher.py:74-201).

Same constructor and ``forward(outputs, targets) -> [(index_i, index_j)]`` (int64, ordered by ascending
matched cost — the PartDistillation-specific re-ordering of matcher.py:162-163).  Internally the
B per-image Python iterations, host syncs and SciPy calls of the reference collapse into four
launches per decoder layer: point-sample predictions, point-sample targets, cost matrices, batched LSAP.
The random point sets are still drawn with one ``torch.rand(1, P, 2)`` per image, in the reference's
order, so the CUDA RNG stream is the same as the reference PyTorch path's.
"""
import torch
from torch import nn

from .. import functional as PF
from .targets import pack_targets


class HungarianMatcher(nn.Module):
    def __init__(self, cost_class: float = 1, cost_mask: float = 1, cost_dice: float = 1, num_points: int = 0):
        super().__init__()
        self.cost_class = cost_class
        self.cost_mask = cost_mask
        self.cost_dice = cost_dice
        assert cost_class != 0 or cost_mask != 0 or cost_dice != 0, "all costs cant be 0"
        self.num_points = num_points
        self.rand = torch.rand           # injectable point-coordinate provider (tests replay recorded draws)

    @torch.no_grad()
    def match_packed(self, outputs, targets):
        """-> (pred_idx, tgt_idx) int64 (Ktot,): image b's matches at targets.offsets[b]:..., in ascending
        cost order; pred_idx is the query index inside the image, tgt_idx the target index inside the image."""
        targets = pack_targets(targets)
        logits = outputs["pred_logits"]
        masks = outputs["pred_masks"]
        B, Q = logits.shape[:2]
        dev = masks.device
        if targets.total == 0:
            e = torch.zeros((0,), dtype=torch.int64, device=dev)
            return e, e
        coords = self.draw_coords(B, dev)
        cache = _index_cache(targets, B, Q, dev)
        prob = logits.float().sigmoid() if logits.shape[-1] == 1 else logits.float().softmax(-1)
        pred_pts = PF.point_sample(masks.detach().float().flatten(0, 1), coords, None, cache["img_of_query"])
        tgt_pts = PF.point_sample(targets.packed_masks, coords, None, cache["img_of_target"])
        cost = PF.matcher_cost(pred_pts, tgt_pts, prob.reshape(B * Q, -1), targets.packed_labels, targets.offsets, Q,
                               self.cost_class, self.cost_mask, self.cost_dice)
        return PF.lsap_batched(cost, targets.offsets, Q)

    @torch.no_grad()
    def draw_coords(self, B, dev):
        """The point coordinates of one matcher call: one (1, P, 2) draw per image, in image order (matcher.py:130-131)."""
        return torch.cat([self.rand(1, self.num_points, 2, device=dev) for _ in range(B)], 0)

    @torch.no_grad()
    def match_all(self, outs, targets, coords_list=None):
        """match_packed for several decoder outputs (the main one and the deep-supervision ones) at once: the cost matrices of
        len(outs) * B (output, image) pairs in ONE cost launch and ONE LSAP launch — at B = 2 the per-output launches put 200
        CTAs / 2 warps on 148 SMs, ten times in a row.  ``coords_list``: the matcher coordinates of every output, already drawn
        (the criterion draws all random numbers of the step in the reference's order); drawn here otherwise.  Returns the list of
        match_packed results.  Same arithmetic per (output, image, query, target) as match_packed: identical assignments."""
        targets = pack_targets(targets)
        n = len(outs)
        B, Q = outs[0]["pred_logits"].shape[:2]
        dev = outs[0]["pred_masks"].device
        if coords_list is None:
            coords_list = [self.draw_coords(B, dev) for _ in outs]
        if targets.total == 0:
            e = torch.zeros((0,), dtype=torch.int64, device=dev)
            return [(e, e) for _ in outs]
        cache = _index_cache(targets, B, Q, dev)
        Ktot = targets.offsets[-1]
        res = []
        step = max(1, 255 // B)                     # the kernels take at most 255 cost matrices per launch
        for s in range(0, n, step):
            chunk, coords = outs[s:s + step], coords_list[s:s + step]
            m = len(chunk)
            pred = torch.cat([PF.point_sample(o["pred_masks"].detach().float().flatten(0, 1), c, None, cache["img_of_query"])
                              for o, c in zip(chunk, coords)])
            tgt = torch.cat([PF.point_sample(targets.packed_masks, c, None, cache["img_of_target"]) for c in coords])
            prob = torch.cat([(o["pred_logits"].float().sigmoid() if o["pred_logits"].shape[-1] == 1
                               else o["pred_logits"].float().softmax(-1)).reshape(B * Q, -1) for o in chunk])
            key = ("all", m)
            tiled = cache.get(key)
            if tiled is None:
                offs = [i * Ktot + o for i in range(m) for o in targets.offsets[:-1]] + [m * Ktot]
                tiled = cache[key] = (offs, targets.packed_labels.repeat(m))
            offs, labels = tiled
            cost = PF.matcher_cost(pred, tgt, prob, labels, offs, Q, self.cost_class, self.cost_mask, self.cost_dice)
            pi, ti = PF.lsap_batched(cost, offs, Q)
            res += [(pi[i * Ktot:(i + 1) * Ktot], ti[i * Ktot:(i + 1) * Ktot]) for i in range(m)]
        return res

    @torch.no_grad()
    def forward(self, outputs, targets):
        targets = pack_targets(targets)
        pi, ti = self.match_packed(outputs, targets)
        Q = outputs["pred_logits"].shape[1]
        out = []
        for b in range(len(targets)):
            n = min(Q, targets.offsets[b + 1] - targets.offsets[b])
            s = targets.offsets[b]
            out.append((pi[s:s + n], ti[s:s + n]))
        return out

    def __repr__(self, _repr_indent=4):
        body = [f"cost_class: {self.cost_class}", f"cost_mask: {self.cost_mask}", f"cost_dice: {self.cost_dice}"]
        return "\n".join(["Matcher " + self.__class__.__name__] + [" " * _repr_indent + line for line in body])


def _index_cache(targets, B, Q, dev):
    """int32 helper tables that depend only on the target counts (built once per step)."""
    c = getattr(targets, "_cache", None)
    if c is None or c["Q"] != Q:
        counts = [targets.offsets[b + 1] - targets.offsets[b] for b in range(B)]
        c = dict(Q=Q,
                 img_of_query=PF.host_table([b for b in range(B) for _ in range(Q)], torch.int32, dev),
                 img_of_target=PF.host_table([b for b in range(B) for _ in range(counts[b])], torch.int32, dev))
        targets._cache = c
    return c
