# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

class ColorMode:
    IMAGE = 0
    SEGMENTATION = 1
    IMAGE_BW = 2
class Visualizer:
    def __init__(self, *a, **k):
        raise NotImplementedError("visualisation is out of scope for the oracle")
class GenericMask:
    pass
def _create_text_labels(*a, **k):
    return []
