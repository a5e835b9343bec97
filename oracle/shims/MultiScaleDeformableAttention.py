# TEST INFRASTRUCTURE ONLY (oracle shim). Minimal stand-in for the third-party
# symbols the reference hot path touches; it exists so
# /root/reference can be imported read-only in the build container to generate
# golden vectors.  Never imported by the product package.

# Import-time stub: ops/functions/ms_deform_attn_func.py:24-32 requires the module to exist, but the
# reference never calls it (ops/modules/ms_deform_attn.py:28 has the Function import commented out).
def ms_deform_attn_forward(*a, **k):
    raise NotImplementedError
def ms_deform_attn_backward(*a, **k):
    raise NotImplementedError
