"""ctypes binding of libmultipoint_b200.so (the C ABI declared in include/multipoint_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing the
product path raises.  Nothing here imports the CPU oracle.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmultipoint_b200.so")

MP_OK, MP_ERR_INVALID, MP_ERR_CUDA, MP_ERR_WORKSPACE, MP_ERR_UNSUPPORTED = 0, -1, -2, -3, -4

_c = ctypes
_vp, _i, _d, _sz, _i64 = _c.c_void_p, _c.c_int, _c.c_double, _c.c_size_t, _c.c_int64

# name -> (restype, argtypes); mirrors include/multipoint_b200.h one to one
SIGNATURES = {
    "mp_version": (_i, []),
    "mp_last_error_string": (_c.c_char_p, []),
    "mp_launch_count": (_c.c_ulonglong, []),
    "mp_profile_begin": (_i, []),
    "mp_profile_end": (_sz, [_c.c_char_p, _sz]),
    "mp_detector_head_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "mp_heatmap_magicleap_f32": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "mp_depth_to_space_f32": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mp_normalize_descriptors_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "mp_transpose_descriptors_f32": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "mp_box_nms_workspace_bytes": (_sz, [_i, _i, _i]),
    "mp_box_nms_f32": (_i, [_vp, _i, _i, _i, _d, _d, _d, _i, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "mp_extract_keypoints_workspace_bytes": (_sz, [_i, _i, _i]),
    "mp_extract_keypoints_f32": (_i, [_vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "mp_sample_descriptors_f32": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "mp_sample_descriptors_split_f32": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "mp_match_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mp_nearest_f32": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mp_match_f32": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _d, _d, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mp_match_split_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _d, _d,
                                _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mp_match_threshold_f32": (_i, [_vp, _i, _vp, _i, _i, _d, _vp, _vp, _vp, _i64, _c.POINTER(_i64), _vp, _sz, _vp]),
    "mp_warp_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "mp_warp_groups_f32": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "mp_valid_mask_u8": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mp_warp_keypoints_i64": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "mp_points_min_dist2_i64": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mp_points_correct_f32": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _c.c_float, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "mp_relu_bn_pad_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "mp_conv1_relu_bn_pad_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "mp_ha_aggregate_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "multipoint_b200: %s is missing.  Build it with `python -m multipoint_b200.build` "
            "(needs nvcc; cross-compiles for sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().mp_last_error_string().decode("utf-8", "replace")


def check(status, what):
    """Map a C status to the exception type the reference would raise for the same misuse."""
    if status == MP_OK:
        return
    msg = "%s failed (%d): %s" % (what, status, last_error())
    if status == MP_ERR_INVALID:
        raise ValueError(msg)
    if status == MP_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def launch_count():
    return int(load().mp_launch_count())


def profile_begin():
    load().mp_profile_begin()


def profile_end():
    """-> {kernel_name: {"launches": n, "total_ms": t}} measured with CUDA events on the launching stream."""
    import json
    lib = load()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.mp_profile_end(buf, len(buf))
    return json.loads(buf.value.decode() or "{}")
