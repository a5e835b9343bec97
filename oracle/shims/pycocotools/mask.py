# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.
#
# pycocotools (2.0.x, un-vendored; `environment.yml` of the reference) is absent from this image.
# `encode` / `iou` / `area` restate the published semantics of its maskApi.c for the one call site on
# the eval path (part_distillation/utils/utils.py:35-42):
#   * encode(fortran uint8/bool (H, W)) -> an opaque per-mask object (here: the mask itself, not RLE);
#   * iou(dt, gt, iscrowd) -> float64 ndarray (len(dt), len(gt)); per pair  i / u  with
#     i = |dt & gt|, u = |dt | gt|  (iscrowd = 0), and `if (i == 0) u = 1` so disjoint or empty
#     masks give exactly 0.0 (rleIou); an empty dt or gt list gives [] (the Python wrapper's early return).
import numpy as np


def _na(*a, **k):
    raise NotImplementedError("pycocotools is an import-time stub in the oracle")


decode = toBbox = frPyObjects = merge = _na


def encode(m):
    m = np.asarray(m)
    if m.ndim != 2:
        raise NotImplementedError("oracle shim: encode() of one (H, W) mask only")
    return {"size": list(m.shape), "_mask": m.astype(bool)}


def area(rles):
    if isinstance(rles, dict):
        return np.uint32(rles["_mask"].sum())
    return np.array([r["_mask"].sum() for r in rles], dtype=np.uint32)


def iou(dt, gt, iscrowd):
    if len(dt) == 0 or len(gt) == 0:
        return []
    if any(iscrowd):
        raise NotImplementedError("oracle shim: iscrowd = 0 only")
    out = np.zeros((len(dt), len(gt)), dtype=np.float64)
    for a, d in enumerate(dt):
        for b, g in enumerate(gt):
            i = int(np.logical_and(d["_mask"], g["_mask"]).sum())
            u = int(np.logical_or(d["_mask"], g["_mask"]).sum())
            if i == 0:
                u = 1
            out[a, b] = float(i) / float(u)
    return out
