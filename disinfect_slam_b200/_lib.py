"""Loader for libtsdf_b200.so (the C ABI in include/tsdf_b200.h).

There is no CPU fallback anywhere in this package: if the CUDA library is missing this raises,
and every entry point needs a B200 (the library refuses other architectures at tsdf_create).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSDF_B200_LIB") or os.path.join(_HERE, "libtsdf_b200.so")  # the override is for kernel experiments (tools/kbench.py)
CSRC = os.path.join(_HERE, "csrc")
_LIB = None


class TsdfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tsdf_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("pool_blocks", C.c_int32),
                ("table_slots", C.c_int32), ("max_image_pixels", C.c_int32), ("shard_rank", C.c_int32),
                ("shard_count", C.c_int32), ("flags", C.c_int32)]


class HostFrame(C.Structure):
    """tsdf_host_frame (include/tsdf_b200.h)."""
    _fields_ = [("rgb", C.c_void_p), ("depth", C.c_void_p), ("ht", C.c_void_p), ("lt", C.c_void_p),
                ("q", C.c_float * 4), ("t", C.c_float * 3), ("format", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_active_pre", "n_new", "n_visible", "n_updated", "n_carved",
                                         "n_active_post", "n_candidates", "reserved")]


# every symbol include/tsdf_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
_FRAME = [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp]
SYMBOLS = {
    "tsdf_last_error": (C.c_char_p, []),
    "tsdf_abi_version": (_i32, []),
    "tsdf_default_config": (_i32, [C.POINTER(Config)]),
    "tsdf_create": (_i32, [_f32, _f32, C.POINTER(Config), C.POINTER(_vp)]),
    "tsdf_destroy": (_i32, [_vp]),
    "tsdf_integrate": (_i32, _FRAME),
    "tsdf_integrate_async": (_i32, _FRAME),
    "tsdf_integrate_device": (_i32, _FRAME + [_vp]),
    "tsdf_integrate_enqueue": (_i32, _FRAME),
    "tsdf_streams_run": (_i32, [_i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _i32, _vp, _vp, _vp]),
    "tsdf_integrate_u16": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _i32]),
    "tsdf_raycast": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsdf_raycast_async": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsdf_raycast_wait": (_i32, [_vp]),
    "tsdf_raycast_device": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tsdf_raycast_resident": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "tsdf_ipc_export": (_i32, [_vp, _vp]),
    "tsdf_ipc_attach": (_i32, [_vp, _i32, _vp]),
    "tsdf_peer_attach_local": (_i32, [_vp, _i32, _vp]),
    "tsdf_raycast_shared": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "tsdf_mirror_attach": (_i32, [_vp, _i32, _vp, _i32]),
    "tsdf_raycast_shared_scatter": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "tsdf_shared_cache_attach": (_i32, [_vp, _i32, _i32]),
    "tsdf_shared_cache_stats": (_i32, [_vp, C.POINTER(_i64)]),
    "tsdf_alloc_exchange_bytes": (C.c_size_t, [_i32, _i32]),
    "tsdf_alloc_exchange_attach": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "tsdf_gather_valid": (_i32, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "tsdf_gather_in_bound": (_i32, [_vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "tsdf_gather_fetch": (_i32, [_vp, _vp, _i64]),
    "tsdf_gather_device_result": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "tsdf_extract_mesh": (_i32, [_vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "tsdf_mesh_fetch": (_i32, [_vp, _vp, _i64]),
    "tsdf_mesh_device_result": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "tsdf_num_active_blocks": (_i32, [_vp, C.POINTER(_i32)]),
    "tsdf_get_counters": (_i32, [_vp, C.POINTER(Counters)]),
    "tsdf_get_skip_map_stats": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "tsdf_synchronize": (_i32, [_vp]),
    "tsdf_stream": (_vp, [_vp]),
    "tsdf_block_owner": (_i32, [C.c_int16, C.c_int16, C.c_int16, _i32, _i32]),
    "tsdf_hash": (C.c_uint32, [C.c_int16, C.c_int16, C.c_int16]),
    "tsdf_allocate_blocks": (_i32, [_vp, _vp, _i32]),
    "tsdf_delete_blocks": (_i32, [_vp, _vp, _i32]),
    "tsdf_retrieve_voxels": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "tsdf_assign_voxels": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "tsdf_export_blocks": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, C.POINTER(_i32)]),
    "tsdf_host_alloc": (_i32, [C.POINTER(_vp), C.c_size_t]),
    "tsdf_host_free": (_i32, [_vp]),
    "tsdf_set_profiling": (_i32, [_vp, _i32]),
    "tsdf_get_phase_ms": (_i32, [_vp, _vp, _vp]),
    "tsdf_get_totals": (_i32, [_vp, C.POINTER(Counters), C.POINTER(_i64)]),
}


IPC_BLOB_BYTES = 320


def build(force=False):
    """Compile libtsdf_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-s", "-j4"] + (["-B"] if force else [])
    subprocess.check_call(args)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("build did not produce " + LIB_PATH)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for the TSDF path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.tsdf_abi_version() != 1:
            raise RuntimeError("libtsdf_b200.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise TsdfError(rc, lib().tsdf_last_error().decode("utf-8", "replace"))
