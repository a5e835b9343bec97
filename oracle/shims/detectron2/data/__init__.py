# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

import types
class _Catalog(dict):
    def get(self, name):
        if name not in self:
            self[name] = types.SimpleNamespace(name=name, class_codes=[])
        return self[name]
MetadataCatalog = _Catalog()
DatasetCatalog = _Catalog()
