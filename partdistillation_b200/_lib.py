"""ctypes binding of csrc/libpdb200.so (the C ABI declared in include/pdb200.h).

There is no fallback: if the library is missing or a call is rejected this raises.  The library is
built in-tree by ``python -m partdistillation_b200.build`` (``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpdb200.so")

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_hp32 = C.POINTER(C.c_int32)    # host int32 array
_hp64 = C.POINTER(C.c_int64)    # host int64 array

# name -> (restype, argtypes); mirrors include/pdb200.h one to one
SIGNATURES = {
    "pdb_abi_version": (_i, []),
    "pdb_last_error": (C.c_char_p, []),
    "pdb_launch_count": (_l, []),
    "pdb_msda_forward": (_i, [_p, _hp64, _hp64, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_msda_backward": (_i, [_p, _hp64, _hp64, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_msda_pack_value_h": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "pdb_msda_forward_h": (_i, [_p, _hp64, _hp64, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_debug_set_msda_path": (_i, [_i]),
    "pdb_mask_einsum_forward": (_i, [_p, _p, _p, _p, _i, _i, _i, _l, _p]),
    "pdb_mask_einsum_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _l, _p]),
    "pdb_gemm_tf32x3": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _l, _l, _l, _l, _l, _l, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_gemm_bf16": (_i, [_p, _p, _p, _p, _i, _i, _i, _l, _l, _l, _i, _i, _i, _i, _i, _p]),
    "pdb_gemm_tf32x3_gated": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _l, _l, _l, _l, _l, _l, _i, _i, _p]),
    "pdb_col_sum": (_i, [_p, _p, _i, _i, _i, _p]),
    "pdb_col_sum_bf16": (_i, [_p, _p, _i, _i, _i, _p]),
    "pdb_set_xattn_passes": (_i, [_i]),
    "pdb_upsample_add_forward": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _l, _p]),
    "pdb_upsample_backward": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_pad_nhwc": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_gemm_small_tf32x3": (_i, [_p, _p, _p, _p, _i, _i, _i, _l, _l, _l, _i, _i, _i, _p]),
    "pdb_gemm_taps_tf32x3": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _l, _l, _l, _l, _i, _hp32, _i, _p]),
    "pdb_gemm_taps_cropped_tf32x3": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _l, _l, _l, _l, _i, _hp32, _i, _i, _i, _p]),
    "pdb_split_lo": (_i, [_p, _p, _l, _p]),
    "pdb_attn_mask_build": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_attn_mask_reset_rows": (_i, [_p, _p, _i, _l, _p]),
    "pdb_masked_xattn_workspace_bytes": (_l, [_i, _i, _i, _i, _i]),
    "pdb_masked_xattn_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pdb_masked_xattn_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pdb_point_sample_forward": (_i, [_p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "pdb_point_sample_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "pdb_matcher_cost": (_i, [_p, _p, _p, _p, _hp32, _p, _i, _i, _i, _i, _f, _f, _f, _p]),
    "pdb_lsap_batched": (_i, [_p, _hp32, _p, _p, _i, _i, _p]),
    "pdb_point_loss_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_point_loss_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_class_rows_forward": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _l, _p]),
    "pdb_window_attention_forward": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p]),
    "pdb_swin_window_attention_forward": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _p]),
    "pdb_swin_window_attention_forward_tc": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _p]),
    "pdb_layer_norm_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _l, _i, _f, _p]),
    "pdb_layer_norm_forward_scaled": (_i, [_p, _p, _p, _l, _p, _p, _p, _p, _p, _p, _l, _i, _f, _i, _p]),
    "pdb_group_norm_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _l, _i, _i, _f, _i, _p]),
    "pdb_group_norm_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _l, _i, _i, _i, _p]),
    "pdb_group_affinity": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_group_affinity_batched": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_group_scores": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "pdb_group_affinity_resized": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_grad_sumsq": (_i, [_p, _l, _f, _p, _p]),
    "pdb_adamw_flat": (_i, [_p, _p, _p, _p, _l, _p, _p, _p, _i, _f, _f, _f, _p, _f, _f, _p, _p]),
    "pdb_postprocess_masks": (_i, [_p, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_resize_masks_u8": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "pdb_pack_bits": (_i, [_p, _p, _i, _i, _i, _p]),
    "pdb_unpack_bits": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "pdb_bits_popcount": (_i, [_p, _p, _i, _l, _p]),
    "pdb_bits_intersect": (_i, [_p, _p, _p, _i, _i, _l, _p]),
    "pdb_class_rows_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _l, _p]),
}

_lib = None


def load():
    """Returns the loaded library (loading it on first use).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m partdistillation_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        if lib.pdb_abi_version() != 1:
            raise RuntimeError(f"libpdb200.so ABI {lib.pdb_abi_version()} != 1; rebuild")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().pdb_last_error().decode()
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def launch_count():
    return int(load().pdb_launch_count())


def host_i64(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])


def host_i32(values):
    return (C.c_int32 * len(values))(*[int(v) for v in values])
