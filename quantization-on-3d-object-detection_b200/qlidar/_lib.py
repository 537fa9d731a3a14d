"""ctypes binding of the C ABI declared in include/qlidar.h.

There is no fallback: if libqlidar_b200.so is missing or a symbol cannot be resolved, importing the ops raises.
Build it with `python quantization-on-3d-object-detection_b200/build.py` (or `__graft_entry__.build()`)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QLIDAR_LIB: test-time override (the ablation build of tools/conv_sweep.py); the product always loads the library next to this file
LIB_PATH = os.environ.get("QLIDAR_LIB") or os.path.join(_HERE, "libqlidar_b200.so")

QL_OK = 0
QL_F16, QL_F32, QL_S8, QL_S32 = 0, 1, 2, 3
QL_Q_CODES_PER_TENSOR, QL_Q_FAKE_PER_CHANNEL, QL_Q_FAKE_PER_TENSOR, QL_Q_FAKE_PER_ROW = 0, 1, 2, 3
TILE_M = 128

_p, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t

# symbol -> (restype, argtypes); must list every function of include/qlidar.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    "ql_abi_version": (C.c_int, []),
    "ql_error_string": (C.c_char_p, [C.c_int]),
    "ql_last_cuda_error": (C.c_char_p, []),
    "ql_num_sms_on_device": (C.c_int, []),
    "ql_hash_capacity": (_i64, [_i64]),
    "ql_hash_build": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _i64, _p]),
    "ql_voxelize_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "ql_voxelize_mean": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _i64, _i64, _p, _i32, _p, _p, _p, _p, _i64, _p, _sz, _p]),
    "ql_voxelize_coords": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _i64, _i64, _p, _i32, _p, _p, _p, _p, _i64, _p, _sz, _p]),
    "ql_voxelize_features": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _i64, _i64, _p, _i32, _p, _p, _p, _p, _i64, _p, _sz, _p]),
    "ql_mean_vfe": (C.c_int, [_p, _p, _i32, _i64, _i32, _i32, _p, _p]),
    "ql_rulebook_num_tiles": (_i64, [_i64]),
    "ql_rulebook_mask_words": (_i32, [_i32]),
    "ql_rulebook_subm": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _i64, _p, _p, _p]),
    "ql_rulebook_strided_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32, _p, _p, _p]),
    "ql_rulebook_strided": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _sz, _p]),
    "ql_rulebook_strided_index": (C.c_int, [_i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "ql_rulebook_strided_ranked": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "ql_renumber_by_key": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p, _sz, _p]),
    "ql_rulebook_subm_ranked": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "ql_rulebook_group_workspace_bytes": (_sz, [_i64]),
    "ql_rulebook_subm_ranked_grouped": (C.c_int, [_p, _i64, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ql_packed_weight_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ql_pack_weights_host": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p]),
    "ql_spconv_mma": (C.c_int, [_p, _i32, _p, _p, _i64, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p, _i32, _p, _p, _p, _p]),
    "ql_spconv_mma_rows": (C.c_int, [_p, _i32, _p, _p, _p, _i64, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _i32, _p, _i32, _p, _p, _p, _p]),
    "ql_spconv_weights_streamed": (_i32, [_i32, _i32, _i32, _i32]),
    "ql_permute_rows": (C.c_int, [_p, _p, _i32, _p, _i64, _p, _p]),
    "ql_stem_conv": (C.c_int, [_p, _i32, _i32, _p, _p, _i64, _p, _i32, _i32, _p, _p, _p, _i32, _p, _i32, _p, _p]),
    "ql_absmax_cols": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p]),
    "ql_quantize_rows": (C.c_int, [_p, _i32, _i64, _p, _i32, _p, _p, _i32, _i32, _p, _p, _p]),
    "ql_sq_prepare_weights": (C.c_int, [_p, _p, _p, C.c_float, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "ql_unfold_absmax": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "ql_unfold_quantize": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _i32, _i32, _p, _p, _p]),
    "ql_bev_densify_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ql_bev_densify": (C.c_int, [_p, _i32, _i32, _p, _i64, _i32, _i32, _i32, _i32, _p, _i32, _p, _sz, _p]),
    "ql_bev_merge2d_workspace_bytes": (_sz, [_i32, _i32, _i32, _i64, _i32, _i32]),
    "ql_bev_merge2d": (C.c_int, [_p, _i32, _i32, _p, _i64, _p, _i32, _i32, _i32, _p, _i32, _p, _i64, _p, _p, _sz, _p]),
    "ql_bev_merge2d_multi": (C.c_int, [_i32, _p, _i32, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _p, _i32, _p, _i32, _i64, _p, _p, _sz, _p]),
    "ql_bev_densify_ranked": (C.c_int, [_p, _i32, _i32, _p, _p, _i64, _p, _i32, _i32, _i32, _i32, _p, _i32, _p, _sz, _p]),
    "ql_voxelize_sorted_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "ql_voxelize_sorted_coords": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _i64, _p, _p, _p, _p, _sz, _p, _sz, _p]),
    "ql_voxelize_sorted_features": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p, _i32, _i32, _i64, _p, _p, _i32, _p, _p, _sz, _p]),
    "ql_centerhead_decode_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ql_centerhead_decode": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, C.c_float, _p, _p, _p, C.c_float, _p,
                                       _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ql_voxelhead_decode_workspace_bytes": (_sz, [_i32, _i32, _i64]),
    "ql_voxelhead_decode": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _i32, _i32, _i32, C.c_float, _p, _p, _p, C.c_float, _p,
                                      _p, _p, _p, _p, _p, _p, _sz, _p]),
    "ql_voxelhead_class_split": (C.c_int, [_p, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "ql_nms_rotated_workspace_bytes": (_sz, [_i32, _i32]),
    "ql_nms_rotated": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _i32, _i32, C.c_float, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
}

_lib = None


class QlidarError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QlidarError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
                "Run `python quantization-on-3d-object-detection_b200/build.py`.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int, what: str) -> None:
    if code != QL_OK:
        l = lib()
        msg = l.ql_error_string(code).decode()
        if code == -2:
            msg += ": " + l.ql_last_cuda_error().decode()
        raise QlidarError(f"{what} failed ({code}): {msg}")
