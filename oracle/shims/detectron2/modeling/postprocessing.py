# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

from torch.nn import functional as F
def sem_seg_postprocess(result, img_size, output_height, output_width):
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]
