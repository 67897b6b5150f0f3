"""ctypes wrapper around the REFERENCE's own CUDA TSDFGrid rebuilt for sm_100a (oracle/_ref/, see
oracle/build_ref.sh and oracle/ref_harness.cu).

TEST INFRASTRUCTURE ONLY: tests/, tests/golden/make_ref_golden.py and bench.py's `--impl reference`
arm.  The libraries are built in the development container (where /root/reference exists) and
travel to the GPU box prebuilt; nothing here reads /root/reference at run time.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def lib_path(parity=True):
    return os.path.join(_HERE, "_ref", "libref_tsdf_parity.so" if parity else "libref_tsdf.so")


def available(parity=True):
    return os.path.exists(lib_path(parity))


def lib(parity=True):
    if parity not in _LIBS:
        L = C.CDLL(lib_path(parity))
        vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float
        L.ref_create.restype = vp
        L.ref_create.argtypes = [f32, f32]
        L.ref_destroy.argtypes = [vp]
        L.ref_integrate.argtypes = [vp, vp, vp, vp, vp, i32, i32, f32, vp, vp, vp]
        L.ref_raycast.argtypes = [vp, f32, i32, i32, vp, vp, vp, vp, vp]
        L.ref_gather_valid.restype = i64
        L.ref_gather_valid.argtypes = [vp, vp, i64]
        L.ref_gather_voxels.restype = i64
        L.ref_gather_voxels.argtypes = [vp, vp, vp, i64]
        L.ref_num_active.argtypes = [vp]
        L.ref_export.argtypes = [vp, vp, vp, vp, vp, i32, C.POINTER(i32)]
        _LIBS[parity] = L
    return _LIBS[parity]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    assert a.size == n
    return a


class RefTSDFGrid:
    """The reference's TSDFGrid (utils/tsdf/voxel_tsdf.cuh:32-88), fixed 2^18-block pool / 2^22 entries."""

    def __init__(self, voxel_size, truncation, parity=True):
        self.L = lib(parity)
        self.h = self.L.ref_create(voxel_size, truncation)

    def close(self):
        if getattr(self, "h", None):
            self.L.ref_destroy(self.h)
            self.h = None

    __del__ = close

    def integrate(self, rgb, depth, ht, lt, max_depth, K, q, t):
        h, w = depth.shape
        K, q, t = _f32(K, 4), _f32(q, 4), _f32(t, 3)
        rc = self.L.ref_integrate(self.h, _p(rgb), _p(depth), _p(ht), _p(lt), w, h, max_depth, _p(K), _p(q), _p(t))
        assert rc == 0, f"CUDA error {rc} in the reference Integrate"

    def raycast(self, max_depth, w, h, K, q, t, download=True):
        K, q, t = _f32(K, 4), _f32(q, 4), _f32(t, 3)
        rgba = np.zeros((h, w, 4), np.uint8) if download else None
        normal = np.zeros((h, w, 4), np.uint8) if download else None
        rc = self.L.ref_raycast(self.h, max_depth, w, h, _p(K), _p(q), _p(t), _p(rgba), _p(normal))
        assert rc == 0, f"CUDA error {rc} in the reference RayCast"
        return rgba, normal

    def gather(self, bbox=None):
        if bbox is None:
            n = self.L.ref_gather_valid(self.h, None, 0)
            out = np.zeros((n, 4), np.float32)
            self.L.ref_gather_valid(self.h, _p(out), n)
        else:
            bb = _f32(bbox, 6)
            n = self.L.ref_gather_voxels(self.h, _p(bb), None, 0)
            out = np.zeros((n, 4), np.float32)
            self.L.ref_gather_voxels(self.h, _p(bb), _p(out), n)
        return out

    def num_active(self):
        return self.L.ref_num_active(self.h)

    def export(self, voxels=True):
        n = C.c_int(0)
        self.L.ref_export(self.h, None, None, None, None, 0, C.byref(n))
        nb = n.value
        keys = np.zeros((nb, 3), np.int16)
        tsdf = np.zeros((nb, 512), np.float32) if voxels else None
        rgbw = np.zeros((nb, 512, 4), np.uint8) if voxels else None
        prob = np.zeros((nb, 512), np.float32) if voxels else None
        if nb:
            rc = self.L.ref_export(self.h, _p(keys), _p(tsdf), _p(rgbw), _p(prob), nb, C.byref(n))
            assert rc == 0
        return keys, tsdf, rgbw, prob
