"""ctypes binding of libtvm_b200.so (the C ABI declared in include/tvm_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, an exception is
raised.  The product path never runs on the CPU and never touches `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TVM_B200_LIB lets the tuning scripts load a differently-compiled build of the SAME sources
LIB_PATH = os.environ.get("TVM_B200_LIB") or os.path.join(_HERE, "libtvm_b200.so")

ABI_VERSION = 11

F_EARLY_TERM = 1 << 0
F_MLP_BF16 = 1 << 1
F_NO_SHADE = 1 << 2
F_POINT_SAMPLES = 1 << 3
F_MLP_TC3 = 1 << 4
F_SPLIT_APP = 1 << 5
F_ZERO_UNLIT = 1 << 6
F_MASK_ANYWHERE = 1 << 7
F_GATHER_ONLY = 1 << 8
F_COUNT_FETCH = 1 << 9
F_BWD_RUNS = 1 << 10

_f3 = C.c_float * 3
_f6 = C.c_float * 6
_i3 = C.c_int32 * 3
_l3 = C.c_int64 * 3


class FieldDesc(C.Structure):
    """Mirror of `tvm_field_desc` (include/tvm_b200.h) — field order and types must match exactly."""
    _fields_ = [
        ("aabb", _f6), ("inv_aabb", _f3), ("grid", _i3),
        ("step_size", C.c_float), ("near_t", C.c_float), ("far_t", C.c_float),
        ("density_shift", C.c_float), ("distance_scale", C.c_float), ("weight_thres", C.c_float),
        ("early_term_eps", C.c_float), ("act", C.c_int32),
        ("n_sigma", _i3), ("n_app", _i3), ("app_dim", C.c_int32),
        ("fea_pe", C.c_int32), ("view_pe", C.c_int32), ("feature_c", C.c_int32),
        ("dplane_off", _l3), ("dline_off", _l3), ("aplane_off", _l3), ("aline_off", _l3),
        ("n_factor_floats", C.c_int64),
        ("occ_cells", C.c_void_p), ("occ_dims", _i3), ("occ_lo", _f3), ("occ_inv", _f3),
        ("occ_coarse", C.c_void_p), ("occ_cdims", _i3),
        ("factors", C.c_void_p), ("basis", C.c_void_p), ("mlp", C.c_void_p), ("mlp_tc", C.c_void_p), ("mlp_tc3", C.c_void_p),
    ]


class RefHead(C.Structure):
    """Mirror of `tvm_ref_head` (include/tvm_b200.h)."""
    _fields_ = [
        ("params", C.c_void_p), ("in_c", C.c_int32), ("feature_c", C.c_int32), ("n_pairs", C.c_int32),
        ("l_max", C.c_int32), ("m", C.c_int32 * 32), ("l", C.c_int32 * 32),
        ("rgb_premultiplier", C.c_float), ("rgb_bias", C.c_float), ("rgb_padding", C.c_float),
        ("diffuse_shift", C.c_float), ("rough_shift", C.c_float),
    ]


MAX_PEERS = 8


class ScatterOut(C.Structure):
    """Mirror of `tvm_scatter_out` (include/tvm_b200.h)."""
    _fields_ = [("dst_index", C.c_void_p), ("n_dst", C.c_int32), ("rgb", C.c_void_p * MAX_PEERS),
                ("depth", C.c_void_p * MAX_PEERS)]


class TvmError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_SIGNATURES = {
    "tvm_abi_version": (C.c_int, []),
    "tvm_launch_count": (C.c_ulonglong, []),
    "tvm_error_string": (C.c_char_p, [C.c_int]),
    "tvm_pack_factors": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(_P), C.POINTER(_P), _P, _P]),
    "tvm_unpack_factor_grads": (C.c_int, [C.POINTER(FieldDesc), _P, C.POINTER(_P), C.POINTER(_P), C.c_int, _P]),
    "tvm_unpack_factor_grads_scaled": (C.c_int, [C.POINTER(FieldDesc), _P, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_float,
                                                 C.c_int, _P]),
    "tvm_allreduce_sum_peer": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int64, _P, _P]),
    "tvm_occupancy_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "tvm_occupancy_coarse_offset": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "tvm_pack_occupancy": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "tvm_mlp_pack_floats": (C.c_size_t, [C.POINTER(FieldDesc)]),
    "tvm_pack_mlp": (C.c_int, [C.POINTER(FieldDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "tvm_mlp_tc_pack_bytes": (C.c_size_t, [C.POINTER(FieldDesc)]),
    "tvm_pack_mlp_tc": (C.c_int, [C.POINTER(FieldDesc), _P, _P, _P, _P, _P, _P]),
    "tvm_mlp_tc3_supported": (C.c_int, [C.POINTER(FieldDesc)]),
    "tvm_mlp_tc3_pack_bytes": (C.c_size_t, [C.POINTER(FieldDesc)]),
    "tvm_pack_mlp_tc3": (C.c_int, [C.POINTER(FieldDesc), _P, _P, _P, _P, _P, _P]),
    "tvm_sample_mask": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, C.c_int, _P, C.c_uint32, _P, _P, _P]),
    "tvm_workspace_bytes": (C.c_int, [C.POINTER(FieldDesc), C.c_int64, C.c_uint32, C.POINTER(C.c_size_t)]),
    "tvm_render_fwd": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, C.c_int, _P, _P, C.c_uint32,
                                 _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "tvm_march_bwd": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, C.c_int, _P, C.c_uint32, _P, _P, _P, _P,
                                _P, _P, C.c_size_t, _P]),
    "tvm_shade_fwd": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, _P, C.c_uint32, _P, _P, _P, _P,
                                C.c_size_t, _P]),
    "tvm_shade_fwd_scatter": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, _P, C.c_uint32, C.POINTER(ScatterOut),
                                        _P, _P, C.c_size_t, _P]),
    "tvm_shade_ref_fwd_scatter": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(RefHead), _P, C.c_int64, C.c_int, _P,
                                            C.POINTER(ScatterOut), _P, _P, C.c_size_t, _P]),
    "tvm_shade_bwd": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                C.c_size_t, _P]),
    "tvm_mlp_grad_floats": (C.c_size_t, [C.POINTER(FieldDesc)]),
    "tvm_unpack_mlp_grads": (C.c_int, [C.POINTER(FieldDesc), _P, _P, _P, _P, _P, _P, _P, C.c_int, _P]),
    "tvm_point_density": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, C.c_float, _P, _P]),
    "tvm_ref_head_floats": (C.c_size_t, [C.POINTER(RefHead)]),
    "tvm_ref_head_layout": (C.c_int, [C.POINTER(RefHead), C.POINTER(C.c_int32)]),
    "tvm_shade_ref_fwd": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(RefHead), _P, C.c_int64, C.c_int, _P, _P, _P, _P, _P,
                                    C.c_size_t, _P]),
    "tvm_shade_ref_bwd": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(RefHead), _P, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P, _P]),
    "tvm_point_appfeature": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, _P, _P]),
    "tvm_pixel_rays_fwd": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P, _P, C.c_int, C.c_int64, C.c_int, _P, _P]),
    "tvm_pixel_rays_bwd": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P, _P, C.c_int, C.c_int64, C.c_int, _P, C.c_int,
                                     _P, _P]),
    "tvm_dense_alpha_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "tvm_dense_alpha_mask": (C.c_int, [C.POINTER(FieldDesc), _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                       _P, _P, _P, C.c_size_t, _P]),
    "tvm_resize_factor": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "tvm_rays_hit_box": (C.c_int, [C.POINTER(FieldDesc), _P, C.c_int64, C.c_int, _P, _P]),
    "tvm_workspace_layout": (C.c_int, [C.POINTER(FieldDesc), C.c_int64] + [C.POINTER(C.c_size_t)] * 6),
}


def exported_symbols():
    """Names every C-ABI entry point include/tvm_b200.h declares (used by the symbol test)."""
    return sorted(_SIGNATURES)


def load():
    """Loads libtvm_b200.so once; raises TvmError if it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TvmError(
            f"{LIB_PATH} not found: the CUDA library is the only implementation of this path "
            "(no CPU fallback). Build it with `python -m iffnerf_b200.build`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing -> loud
        fn.restype = res
        fn.argtypes = args
    v = lib.tvm_abi_version()
    if v != ABI_VERSION:
        raise TvmError(f"libtvm_b200.so ABI {v} != binding ABI {ABI_VERSION}; rebuild the library")
    _lib = lib
    return lib


_bench = None
BENCH_LIB_PATH = os.path.join(_HERE, "libtvm_bench.so")


def load_bench():
    """Measurement aids (include/tvm_bench.h) — a separate library, never loaded by the render path."""
    global _bench
    if _bench is None:
        if not os.path.exists(BENCH_LIB_PATH):
            raise TvmError(f"{BENCH_LIB_PATH} not found; build it with `python -m iffnerf_b200.build`")
        lib = C.CDLL(BENCH_LIB_PATH)
        lib.tvm_gather_microbench.restype = C.c_int
        lib.tvm_gather_microbench.argtypes = [_P, C.c_size_t, C.c_int, C.c_int, _P, C.POINTER(C.c_ulonglong), _P]
        _bench = lib
    return _bench


def check(code: int, what: str):
    if code != 0:
        msg = load().tvm_error_string(code)
        raise TvmError(f"{what} failed: [{code}] {msg.decode() if msg else '?'}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array_int(addresses):
    arr = (_P * len(addresses))()
    for i, a in enumerate(addresses):
        arr[i] = int(a)
    return arr


def ptr_array(tensors):
    arr = (_P * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr
