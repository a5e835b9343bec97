# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

class Registry:
    """String-keyed class registry (detectron2/fvcore semantics: keyed by __name__)."""
    def __init__(self, name):
        self._name = name
        self._obj_map = {}
    def _do_register(self, name, obj):
        self._obj_map[name] = obj   # oracle reloads modules; last writer wins
    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]
    def __contains__(self, name):
        return name in self._obj_map
