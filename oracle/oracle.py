"""ctypes wrapper around the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (disinfect_slam_b200) never imports
this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("tsdf_oracle.c", "ref_hash_model.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [f32, f32]
        L.oracle_destroy.argtypes = [vp]
        L.oracle_num_blocks.argtypes = [vp]
        L.oracle_hash.restype = C.c_uint32
        L.oracle_hash.argtypes = [C.c_int16] * 3
        L.oracle_integrate.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp, vp, i32]
        L.oracle_export.argtypes = [vp, vp, vp, vp, vp, i32]
        L.oracle_gather.restype = i64
        L.oracle_gather.argtypes = [vp, vp, vp, i64]
        L.oracle_raycast.argtypes = [vp, f32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
        L.oracle_get_voxel.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        L.oracle_allocate_block.argtypes = [vp, i32, i32, i32]
        L.oracle_delete_block.argtypes = [vp, i32, i32, i32]
        L.oracle_set_voxel.argtypes = [vp, i32, i32, i32, vp, vp, vp]
        L.refhash_create.restype = vp
        L.refhash_create.argtypes = [i32, i32]
        L.refhash_destroy.argtypes = [vp]
        L.refhash_reset_locks.argtypes = [vp]
        L.refhash_num_active.argtypes = [vp]
        L.refhash_hash.restype = C.c_uint32
        L.refhash_hash.argtypes = [vp, i32, i32, i32]
        for f in (L.refhash_allocate, L.refhash_delete, L.refhash_find):
            f.argtypes = [vp, i32, i32, i32]
        L.refhash_pool_acquire.argtypes = [vp]
        L.refhash_pool_release.argtypes = [vp, i32]
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert n is None or a.size == n
    return a


COUNTER_NAMES = ("n_active_pre", "n_new", "n_vis", "n_upd", "n_carved", "n_active_post", "n_requests", "_")


class Oracle:
    """Ideal-set-semantics CPU restatement of TSDFGrid (utils/tsdf/voxel_tsdf.cuh:32-88)."""

    def __init__(self, voxel_size, truncation):
        self.L = lib()
        self.voxel_size, self.truncation = float(voxel_size), float(truncation)
        self.h = self.L.oracle_create(voxel_size, truncation)

    def close(self):
        if self.h:
            self.L.oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def num_blocks(self):
        return self.L.oracle_num_blocks(self.h)

    def integrate(self, rgb, depth, ht, lt, max_depth, K, q, t, want_new_keys=False):
        h, w = depth.shape
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        depth, ht, lt = _f32(depth), _f32(ht), _f32(lt)
        assert rgb.shape == (h, w, 3) and ht.shape == (h, w) and lt.shape == (h, w)
        K, q, t = _f32(K, 4), _f32(q, 4), _f32(t, 3)
        cnt = np.zeros(8, np.int64)
        cap = 1 << 20 if want_new_keys else 0
        nk = np.zeros((cap, 3), np.int16) if want_new_keys else None
        self.L.oracle_integrate(self.h, _p(rgb), _p(depth), _p(ht), _p(lt), w, h, max_depth, _p(K), _p(q), _p(t),
                                _p(cnt), _p(nk), cap)
        out = dict(zip(COUNTER_NAMES, (int(c) for c in cnt)))
        out.pop("_")
        if want_new_keys:
            out["new_keys"] = nk[: out["n_new"]].copy()
        return out

    def export(self, voxels=True):
        n = self.num_blocks()
        keys = np.zeros((n, 3), np.int16)
        if voxels:
            tsdf = np.zeros((n, 512), np.float32)
            rgbw = np.zeros((n, 512, 4), np.uint8)
            prob = np.zeros((n, 512), np.float32)
        else:
            tsdf = rgbw = prob = None
        self.L.oracle_export(self.h, _p(keys), _p(tsdf), _p(rgbw), _p(prob), n)
        return keys, tsdf, rgbw, prob

    def gather(self, bbox=None):
        bb = None if bbox is None else _f32(bbox, 6)
        n = self.L.oracle_gather(self.h, _p(bb), None, 0)
        out = np.zeros((n, 4), np.float32)
        self.L.oracle_gather(self.h, _p(bb), _p(out), n)
        return out

    def raycast(self, max_depth, w, h, K, q, t):
        K, q, t = _f32(K, 4), _f32(q, 4), _f32(t, 3)
        rgba = np.zeros((h, w, 4), np.uint8)
        normal = np.zeros((h, w, 4), np.uint8)
        depth = np.zeros((h, w), np.float32)
        cnt = np.zeros(4, np.int64)
        self.L.oracle_raycast(self.h, max_depth, w, h, _p(K), _p(q), _p(t), _p(rgba), _p(normal), _p(depth), _p(cnt))
        return rgba, normal, depth, {"samples": int(cnt[0]), "block_switches": int(cnt[1]), "hits": int(cnt[2])}

    def get_voxel(self, x, y, z):
        ts, pr = C.c_float(), C.c_float()
        rgbw = np.zeros(4, np.uint8)
        found = self.L.oracle_get_voxel(self.h, x, y, z, C.byref(ts), _p(rgbw), C.byref(pr))
        return bool(found), ts.value, rgbw, pr.value

    def allocate_block(self, bx, by, bz):
        return self.L.oracle_allocate_block(self.h, bx, by, bz)

    def set_voxels(self, points, tsdf=None, rgbw=None, prob=None):
        pts = np.asarray(points, np.int32).reshape(-1, 3)
        for i, (x, y, z) in enumerate(pts.tolist()):
            t = None if tsdf is None else C.byref(C.c_float(float(tsdf[i])))
            p = None if prob is None else C.byref(C.c_float(float(prob[i])))
            c = None if rgbw is None else _p(np.ascontiguousarray(rgbw[i], np.uint8))
            assert self.L.oracle_set_voxel(self.h, x, y, z, t, c, p) == 1

    def delete_block(self, bx, by, bz):
        return self.L.oracle_delete_block(self.h, bx, by, bz)

    def prune_to(self, keys):
        """Delete every block whose coordinate is not in `keys` (n x 3); returns how many were removed."""
        keep = set(map(tuple, np.asarray(keys).tolist()))
        mine = self.export(voxels=False)[0]
        n = 0
        for k in mine.tolist():
            if tuple(k) not in keep:
                n += self.delete_block(*k)
        return n


def hash_block(x, y, z):
    return int(lib().oracle_hash(x, y, z))


class RefHashModel:
    """Sequential model of the reference's VoxelHashTable + VoxelMemPool (see ref_hash_model.c)."""

    def __init__(self, bucket_bits=21, num_block=1 << 18):
        self.L = lib()
        self.h = self.L.refhash_create(bucket_bits, num_block)

    def close(self):
        if self.h:
            self.L.refhash_destroy(self.h)
            self.h = None

    __del__ = close

    def hash(self, x, y, z):
        return int(self.L.refhash_hash(self.h, x, y, z))

    def allocate(self, x, y, z):
        return self.L.refhash_allocate(self.h, x, y, z)

    def delete(self, x, y, z):
        return self.L.refhash_delete(self.h, x, y, z)

    def find(self, x, y, z):
        return self.L.refhash_find(self.h, x, y, z)

    def reset_locks(self):
        self.L.refhash_reset_locks(self.h)

    def num_active(self):
        return self.L.refhash_num_active(self.h)

    def pool_acquire(self):
        return self.L.refhash_pool_acquire(self.h)

    def pool_release(self, b):
        self.L.refhash_pool_release(self.h, b)
