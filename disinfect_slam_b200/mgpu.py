"""ctypes binding of libtsdf_b200_mgpu.so (include/tsdf_b200_mgpu.h): the C++ / NCCL multi-GPU data plane of the
engine -- one volume sharded over the GPUs of a node by block ownership.  One `ShardedVolume` per rank (process or
thread); every rank makes the same calls with the same camera arguments.  The binding only forwards pointers: the
frame broadcast, the barrier, the peer-memory ray march and the image all-gather all run inside the library.

When torch is also used in the process, import torch BEFORE this module, so that the NCCL torch bundles is the one
the dynamic linker binds (both libraries carry the soname libnccl.so.2).
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Config, Counters, TsdfError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtsdf_b200_mgpu.so")
ID_BYTES = 128
_LIB = None


class Frame(C.Structure):
    _fields_ = [("rgb", C.c_void_p), ("depth", C.c_void_p), ("ht", C.c_void_p), ("lt", C.c_void_p),
                ("q", C.c_float * 4), ("t", C.c_float * 3), ("reserved", C.c_float)]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
SYMBOLS = {
    "tsdf_mgpu_last_error": (C.c_char_p, []),
    "tsdf_mgpu_unique_id": (_i32, [_vp]),
    "tsdf_mgpu_create": (_i32, [_f32, _f32, C.POINTER(Config), _i32, _i32, _vp, C.POINTER(_vp)]),
    "tsdf_mgpu_destroy": (_i32, [_vp]),
    "tsdf_mgpu_engine": (_vp, [_vp]),
    "tsdf_mgpu_integrate": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp]),
    "tsdf_mgpu_raycast": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "tsdf_mgpu_raycast_composite": (_i32, [_vp, _f32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_vp)]),
    "tsdf_mgpu_fetch_images": (_i32, [_vp, _vp, _vp, _vp]),
    "tsdf_mgpu_gather": (_i32, [_vp, _i32, _vp, _vp, _i64, C.POINTER(_i64)]),
    "tsdf_mgpu_counters": (_i32, [_vp, C.POINTER(Counters), C.POINTER(Counters), C.POINTER(_i64)]),
    "tsdf_mgpu_run_sequence": (_i32, [_vp, _i32, _i32, C.POINTER(Frame), _i32, _i32, _i32, _i32, _i32, _f32, _vp, _i32]),
    "tsdf_mgpu_synchronize": (_i32, [_vp]),
    "tsdf_mgpu_set_profiling": (_i32, [_vp, _i32]),
    "tsdf_mgpu_get_comm_ms": (_i32, [_vp, _vp, _vp]),
}
COMM_PHASES = ("broadcast", "barrier", "allgather", "composite_allreduce", "raycast_shared", "gather_sendrecv", "exchange_barrier")


def lib():
    global _LIB
    if _LIB is None:
        _lib.lib()  # the engine library first (the data plane links against it); raises if it is missing
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise TsdfError(rc, lib().tsdf_mgpu_last_error().decode("utf-8", "replace"))


def unique_id():
    """ncclGetUniqueId as 128 bytes: made on one rank, handed to every rank's ShardedVolume."""
    buf = np.zeros(ID_BYTES, np.uint8)
    check(lib().tsdf_mgpu_unique_id(buf.ctypes.data_as(_vp)))
    return buf.tobytes()


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _f32a(a, n):
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    assert a.size == n
    return a


def _ptr(x):
    """numpy array -> its address; int (device pointer) -> itself; None -> None."""
    if x is None:
        return None
    return x.ctypes.data if isinstance(x, np.ndarray) else int(x)


class ShardedVolume:
    def __init__(self, voxel_size, truncation, rank, world, nccl_id, device=None, pool_blocks=None, table_slots=None,
                 max_image_pixels=None, shard_shift=2):
        self.L = lib()
        self.rank, self.world = int(rank), int(world)
        cfg = Config()
        _lib.check(_lib.lib().tsdf_default_config(C.byref(cfg)))
        if pool_blocks is not None:
            cfg.pool_blocks = int(pool_blocks)
        if table_slots is not None:
            cfg.table_slots = int(table_slots)
        elif pool_blocks is not None:
            cfg.table_slots = max(1 << 16, 1 << int(np.ceil(np.log2(8 * cfg.pool_blocks))))
        if max_image_pixels is not None:
            cfg.max_image_pixels = int(max_image_pixels)
        cfg.device = self.rank if device is None else int(device)
        cfg.flags = int(shard_shift) & 0xF
        self.h = _vp()
        idb = np.frombuffer(nccl_id, np.uint8).copy()
        assert idb.size == ID_BYTES
        check(self.L.tsdf_mgpu_create(voxel_size, truncation, C.byref(cfg), self.rank, self.world, _p(idb), C.byref(self.h)))
        self.engine = self.L.tsdf_mgpu_engine(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.tsdf_mgpu_destroy(self.h)
            self.h = None

    __del__ = close

    def Integrate(self, root, rgb, depth, ht, lt, width, height, max_depth, K, cam_T_world, on_device=False):
        """Planes: numpy arrays (host) or ints (device pointers) on the root rank, None elsewhere."""
        q, t = cam_T_world
        K, q, t = _f32a(K, 4), _f32a(q, 4), _f32a(t, 3)
        check(self.L.tsdf_mgpu_integrate(self.h, root, int(on_device), _ptr(rgb), _ptr(depth), _ptr(ht), _ptr(lt), width, height,
                                         max_depth, _p(K), _p(q), _p(t)))

    def RayCast(self, max_depth, width, height, K, cam_T_world, to_host=True):
        q, t = cam_T_world
        K, q, t = _f32a(K, 4), _f32a(q, 4), _f32a(t, 3)
        a, b, c = _vp(), _vp(), _vp()
        check(self.L.tsdf_mgpu_raycast(self.h, max_depth, width, height, _p(K), _p(q), _p(t), C.byref(a), C.byref(b), C.byref(c)))
        if not to_host:
            return a.value, b.value, c.value
        rgba, normal = np.empty((height, width, 4), np.uint8), np.empty((height, width, 4), np.uint8)
        depth = np.empty((height, width), np.float32)
        check(self.L.tsdf_mgpu_fetch_images(self.h, _p(rgba), _p(normal), _p(depth)))
        return rgba, normal, depth

    def RayCastComposite(self, max_depth, width, height, K, cam_T_world):
        q, t = cam_T_world
        K, q, t = _f32a(K, 4), _f32a(q, 4), _f32a(t, 3)
        k = _vp()
        check(self.L.tsdf_mgpu_raycast_composite(self.h, max_depth, width, height, _p(K), _p(q), _p(t), C.byref(k)))
        return k.value

    def Gather(self, root, bbox=None):
        """GatherValid (bbox None) / GatherVoxels over all shards; records on `root`, empty elsewhere."""
        bb = None if bbox is None else _f32a(tuple(bbox), 6)
        n = _i64(0)
        check(self.L.tsdf_mgpu_gather(self.h, root, _p(bb), None, 0, C.byref(n)))
        out = np.empty((n.value if self.rank == root else 0, 4), np.float32)
        check(self.L.tsdf_mgpu_gather(self.h, root, _p(bb), _p(out) if len(out) else None, len(out), C.byref(n)))
        return out

    def counters(self):
        """(last frame, totals, active blocks) summed over all shards."""
        a, b, n = Counters(), Counters(), _i64(0)
        check(self.L.tsdf_mgpu_counters(self.h, C.byref(a), C.byref(b), C.byref(n)))
        d = lambda c: {k: int(getattr(c, k)) for k, _ in Counters._fields_ if k != "reserved"}  # noqa: E731
        return d(a), d(b), int(n.value)

    def run_sequence(self, root, frames, first, count, width, height, max_depth, K, raycast_mode=1, on_device=True):
        """frames: ctypes array of Frame (see make_frames)."""
        K = _f32a(K, 4)
        check(self.L.tsdf_mgpu_run_sequence(self.h, root, int(on_device), frames, len(frames), first, count, width, height,
                                            max_depth, _p(K), raycast_mode))

    def synchronize(self):
        check(self.L.tsdf_mgpu_synchronize(self.h))

    def set_profiling(self, on=True):
        check(self.L.tsdf_mgpu_set_profiling(self.h, int(on)))

    def comm_ms(self):
        ms, cnt = np.zeros(8, np.float32), np.zeros(8, np.int64)
        check(self.L.tsdf_mgpu_get_comm_ms(self.h, _p(ms), _p(cnt)))
        return {n: float(ms[i]) for i, n in enumerate(COMM_PHASES)}, {n: int(cnt[i]) for i, n in enumerate(COMM_PHASES)}


def make_frames(cams, planes=None):
    """ctypes Frame array from cameras [(q, t)] and, on the root, per-frame (rgb, depth, ht, lt) pointers / arrays."""
    arr = (Frame * len(cams))()
    for i, (q, t) in enumerate(cams):
        arr[i].q[:] = [float(v) for v in q]
        arr[i].t[:] = [float(v) for v in t]
        if planes is not None:
            rgb, depth, ht, lt = planes[i]
            arr[i].rgb, arr[i].depth, arr[i].ht, arr[i].lt = _ptr(rgb), _ptr(depth), _ptr(ht), _ptr(lt)
    return arr
