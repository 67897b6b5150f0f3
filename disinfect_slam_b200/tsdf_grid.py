"""Python host mirror of the reference's `TSDFGrid` (utils/tsdf/voxel_tsdf.cuh:32-88) on top of the
C ABI of libtsdf_b200.so.  Same method names and argument meaning as the reference class:

    TSDFGrid(voxel_size, truncation)
    Integrate(img_rgb, img_depth, img_ht, img_lt, max_depth, intrinsics, cam_T_world)
    RayCast(max_depth, virtual_cam, cam_T_world) -> rgba, normal (+ hit depth)
    GatherValid() / GatherVoxels(volumn) -> array of VoxelSpatialTSDF records (x, y, z, tsdf)

numpy arrays stand in for cv::Mat, (fx, fy, cx, cy) for CameraIntrinsics<float>, and
(q_xyzw, t_xyz) for SE3<float> (utils/cuda/lie_group.cuh:43-44).  The C++ shim with the
reference's exact C++ signatures is include/tsdf_b200/voxel_tsdf.hpp.  No CPU fallback.
"""
import ctypes as C
from collections import namedtuple

import numpy as np

from . import _lib
from ._lib import Config, Counters, TsdfError, check  # noqa: F401

CameraParams = namedtuple("CameraParams", "intrinsics img_h img_w")  # utils/cuda/camera.cuh:54-68
BoundingCube = namedtuple("BoundingCube", "xmin xmax ymin ymax zmin zmax")  # utils/tsdf/voxel_tsdf.cuh:12-19


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    if a.size != n:
        raise ValueError(f"expected {n} floats, got {a.size}")
    return a


def _pose(cam_T_world):
    q, t = cam_T_world
    return _f32(q, 4), _f32(t, 3)


def hash_block(bx, by, bz):
    """Hash() of the reference (utils/tsdf/voxel_hash.cu:31-35), through the C ABI."""
    return int(_lib.lib().tsdf_hash(bx, by, bz))


def block_owner(keys, shard_count, shard_shift=0):
    """Owner rank of each block coordinate (n x 3) under the engine's sharding rule, through the C ABI."""
    L = _lib.lib()
    return np.array([L.tsdf_block_owner(int(k[0]), int(k[1]), int(k[2]), int(shard_count), int(shard_shift))
                     for k in np.asarray(keys).reshape(-1, 3)], np.int32)


class PinnedArray:
    """numpy view of pinned host memory from tsdf_host_alloc (DMA-able Integrate inputs)."""

    def __init__(self, shape, dtype):
        self.L = _lib.lib()
        self.ptr = C.c_void_p()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        check(self.L.tsdf_host_alloc(C.byref(self.ptr), n))
        buf = (C.c_char * n).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.L.tsdf_host_free(self.ptr)
            self.ptr = None

    __del__ = free


class TSDFGrid:
    def __init__(self, voxel_size, truncation, pool_blocks=None, table_slots=None, max_image_pixels=None, device=None,
                 shard_rank=0, shard_count=1, shard_shift=0, blocking_sync=False):
        self.L = _lib.lib()
        self.voxel_size, self.truncation = float(voxel_size), float(truncation)
        cfg = Config()
        check(self.L.tsdf_default_config(C.byref(cfg)))
        if pool_blocks is not None:
            cfg.pool_blocks = int(pool_blocks)
        if table_slots is not None:
            cfg.table_slots = int(table_slots)
        elif pool_blocks is not None:
            cfg.table_slots = max(1 << 16, 1 << int(np.ceil(np.log2(8 * cfg.pool_blocks))))
        if max_image_pixels is not None:
            cfg.max_image_pixels = int(max_image_pixels)
        if device is not None:
            cfg.device = int(device)
        cfg.shard_rank, cfg.shard_count = int(shard_rank), int(shard_count)
        cfg.flags = (int(shard_shift) & 0xF) | (0x10 if blocking_sync else 0)
        self.cfg = cfg
        self.h = C.c_void_p()
        check(self.L.tsdf_create(voxel_size, truncation, C.byref(cfg), C.byref(self.h)))

    def close(self):
        if getattr(self, "_pin", None) is not None:
            self._pin.free()
            self._pin = None
        if getattr(self, "h", None):
            self.L.tsdf_destroy(self.h)
            self.h = None

    __del__ = close

    # ---- TSDFGrid::Integrate (utils/tsdf/voxel_tsdf.cu:347-375) ------------------------------------
    def _frame_args(self, img_rgb, img_depth, img_ht, img_lt, intrinsics, cam_T_world):
        if img_rgb.dtype != np.uint8 or img_rgb.ndim != 3 or img_rgb.shape[2] != 3:
            raise ValueError("img_rgb must be uint8 HxWx3 (CV_8UC3)")  # assert voxel_tsdf.cu:350
        if img_depth.dtype != np.float32 or img_depth.ndim != 2:
            raise ValueError("img_depth must be float32 HxW (CV_32FC1)")  # assert voxel_tsdf.cu:351
        h, w = img_depth.shape
        if img_rgb.shape[:2] != (h, w):
            raise ValueError("rgb / depth size mismatch")  # asserts voxel_tsdf.cu:352-353
        for a in (img_ht, img_lt):
            if a.dtype != np.float32 or a.shape != (h, w):
                raise ValueError("ht / lt must be float32 HxW")
        for a in (img_rgb, img_depth, img_ht, img_lt):
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError("images must be continuous")
        q, t = _pose(cam_T_world)
        return w, h, _f32(intrinsics, 4), q, t

    def Integrate(self, img_rgb, img_depth, img_ht, img_lt, max_depth, intrinsics, cam_T_world, asynchronous=False):
        w, h, K, q, t = self._frame_args(img_rgb, img_depth, img_ht, img_lt, intrinsics, cam_T_world)
        fn = self.L.tsdf_integrate_async if asynchronous else self.L.tsdf_integrate
        check(fn(self.h, _p(img_rgb), _p(img_depth), _p(img_ht), _p(img_lt), w, h, max_depth, _p(K), _p(q), _p(t)))

    def IntegrateU16(self, img_rgb, depth_u16, ht_u16, lt_u16, depthmap_factor, max_depth, intrinsics, cam_T_world, flags=0):
        """The sensor's own formats (tsdf_integrate_u16): uint16 depth (metres = value / depthmap_factor) and uint16
        probabilities (value / 65535; both None = planes of ones), converted on the GPU with cv::Mat::convertTo's
        arithmetic.  flags: 0 synchronous, 1 = TSDF_FRAME_ASYNC, 2 = TSDF_FRAME_NOWAIT."""
        h, w = depth_u16.shape
        if img_rgb.dtype != np.uint8 or img_rgb.shape != (h, w, 3) or depth_u16.dtype != np.uint16:
            raise ValueError("rgb must be uint8 HxWx3, depth uint16 HxW")
        for a in (ht_u16, lt_u16):
            if a is not None and (a.dtype != np.uint16 or a.shape != (h, w)):
                raise ValueError("ht / lt must be uint16 HxW or None")
        q, t = _pose(cam_T_world)
        K = _f32(intrinsics, 4)
        check(self.L.tsdf_integrate_u16(self.h, _p(img_rgb), _p(depth_u16), _p(ht_u16), _p(lt_u16), w, h, depthmap_factor, max_depth,
                                        _p(K), _p(q), _p(t), int(flags)))

    def IntegrateDevice(self, d_rgb, d_depth, d_ht, d_lt, width, height, max_depth, intrinsics, cam_T_world,
                        after_event=None):
        """Planes already in device memory (raw device pointers as ints)."""
        q, t = _pose(cam_T_world)
        K = _f32(intrinsics, 4)
        check(self.L.tsdf_integrate_device(self.h, d_rgb, d_depth, d_ht, d_lt, width, height, max_depth, _p(K), _p(q),
                                           _p(t), after_event))

    # ---- TSDFGrid::RayCast (utils/tsdf/voxel_tsdf.cu:490-506) --------------------------------------
    def RayCast(self, max_depth, virtual_cam, cam_T_world, want_depth=True, out=None):
        """`out` = optional (rgba, normal, depth) host arrays to fill (e.g. PinnedArray views)."""
        K = _f32(virtual_cam.intrinsics, 4)
        h, w = int(virtual_cam.img_h), int(virtual_cam.img_w)
        q, t = _pose(cam_T_world)
        if out is not None:
            rgba, normal, depth = out
            for a, shp, dt in ((rgba, (h, w, 4), np.uint8), (normal, (h, w, 4), np.uint8), (depth, (h, w), np.float32)):
                if a is not None and (a.shape != shp or a.dtype != dt or not a.flags["C_CONTIGUOUS"]):
                    raise ValueError("RayCast out buffers must be contiguous rgba/normal uint8 HxWx4, depth float32 HxW")
        else:
            rgba = np.empty((h, w, 4), np.uint8)
            normal = np.empty((h, w, 4), np.uint8)
            depth = np.empty((h, w), np.float32) if want_depth else None
        check(self.L.tsdf_raycast(self.h, max_depth, w, h, _p(K), _p(q), _p(t), _p(rgba), _p(normal), _p(depth)))
        return rgba, normal, depth

    def RayCastAsync(self, max_depth, virtual_cam, cam_T_world, out):
        """Pipelined RayCast (tsdf_raycast_async): `out` = (rgba, normal, depth) pinned host arrays (PinnedArray views), valid
        after RayCastWait() / synchronize(); at most two views in flight."""
        K = _f32(virtual_cam.intrinsics, 4)
        h, w = int(virtual_cam.img_h), int(virtual_cam.img_w)
        q, t = _pose(cam_T_world)
        rgba, normal, depth = out
        for a, shp, dt in ((rgba, (h, w, 4), np.uint8), (normal, (h, w, 4), np.uint8), (depth, (h, w), np.float32)):
            if a is not None and (a.shape != shp or a.dtype != dt or not a.flags["C_CONTIGUOUS"]):
                raise ValueError("RayCastAsync out buffers must be contiguous rgba/normal uint8 HxWx4, depth float32 HxW")
        check(self.L.tsdf_raycast_async(self.h, max_depth, w, h, _p(K), _p(q), _p(t), _p(rgba), _p(normal), _p(depth)))

    def RayCastWait(self):
        """Blocks until the images of the oldest outstanding RayCastAsync are in host memory."""
        check(self.L.tsdf_raycast_wait(self.h))

    def RayCastDevice(self, max_depth, virtual_cam, cam_T_world, d_rgba=None, d_normal=None, d_depth=None,
                      d_packed=None):
        K = _f32(virtual_cam.intrinsics, 4)
        q, t = _pose(cam_T_world)
        check(self.L.tsdf_raycast_device(self.h, max_depth, int(virtual_cam.img_w), int(virtual_cam.img_h), _p(K),
                                         _p(q), _p(t), d_rgba, d_normal, d_depth, d_packed))

    # ---- shared-volume RayCast (volume sharded over several engines) --------------------------------------
    def ipc_export(self):
        blob = np.zeros(_lib.IPC_BLOB_BYTES, np.uint8)
        check(self.L.tsdf_ipc_export(self.h, _p(blob)))
        return blob

    def ipc_attach(self, blobs):
        blobs = np.ascontiguousarray(blobs, np.uint8).reshape(-1, _lib.IPC_BLOB_BYTES)
        check(self.L.tsdf_ipc_attach(self.h, len(blobs), _p(blobs)))

    def peer_attach_local(self, grids):
        arr = (C.c_void_p * len(grids))(*[g.h for g in grids])
        check(self.L.tsdf_peer_attach_local(self.h, len(grids), arr))

    def shared_cache_attach(self, stride_blocks, pad_voxels=3):
        """Pulled TSDF cache for the shared-volume RayCast (tsdf_shared_cache_attach); 0 blocks detaches."""
        check(self.L.tsdf_shared_cache_attach(self.h, int(stride_blocks), int(pad_voxels)))

    def shared_cache_fetched(self):
        """Foreign blocks the most recent shared-volume view fetched into the cache."""
        n = C.c_int64(0)
        check(self.L.tsdf_shared_cache_stats(self.h, C.byref(n)))
        return int(n.value)

    def RayCastSharedScatter(self, max_depth, virtual_cam, cam_T_world, tile_first, tile_stride, tile_count, peers_unchanged, dests):
        """tsdf_raycast_shared_scatter: `dests` = [(d_rgba, d_normal, d_depth)] device pointers of every destination."""
        K = _f32(virtual_cam.intrinsics, 4)
        q, t = _pose(cam_T_world)
        n = len(dests)
        arrs = [(C.c_void_p * n)(*[d[i] for d in dests]) for i in range(3)]
        check(self.L.tsdf_raycast_shared_scatter(self.h, max_depth, int(virtual_cam.img_w), int(virtual_cam.img_h), _p(K), _p(q), _p(t),
                                                 int(tile_first), int(tile_stride), int(tile_count), int(bool(peers_unchanged)), n, arrs[0], arrs[1], arrs[2]))

    def RayCastShared(self, max_depth, virtual_cam, cam_T_world, row0, rows, d_rgba, d_normal, d_depth):
        K = _f32(virtual_cam.intrinsics, 4)
        q, t = _pose(cam_T_world)
        check(self.L.tsdf_raycast_shared(self.h, max_depth, int(virtual_cam.img_w), int(virtual_cam.img_h), _p(K), _p(q), _p(t),
                                         int(row0), int(rows), d_rgba, d_normal, d_depth))

    # ---- TSDFGrid::GatherValid / GatherVoxels (utils/tsdf/voxel_tsdf.cu:399-454) -------------------
    def _gather(self, bbox, pinned=False, out=None):
        """`out`: optional caller-owned (n, 4) float32 destination (e.g. a PinnedArray view sized for the largest query);
        the records land in out[:n], which is returned."""
        n = C.c_int64(0)
        if bbox is None:
            check(self.L.tsdf_gather_valid(self.h, None, 0, C.byref(n)))
        else:
            bb = _f32(bbox, 6)
            check(self.L.tsdf_gather_in_bound(self.h, _p(bb), None, 0, C.byref(n)))
        if out is not None:
            if out.dtype != np.float32 or out.ndim != 2 or out.shape[1] != 4 or not out.flags["C_CONTIGUOUS"] or out.shape[0] < n.value:
                raise ValueError(f"gather destination must be C-contiguous float32 (>= {n.value}, 4)")
            out = out[:n.value]
        elif pinned:
            # grow-only pinned result buffer owned by this object: the DMA engine writes it directly (PCIe speed); the
            # returned view is valid until the next pinned gather
            if getattr(self, "_pin", None) is None or self._pin.array.shape[0] < n.value:
                old = 0 if getattr(self, "_pin", None) is None else self._pin.array.shape[0]
                if old:
                    self._pin.free()
                self._pin = PinnedArray((max(n.value, int(1.5 * old), 1 << 16), 4), np.float32)  # geometric growth: pinning memory is slow
            out = self._pin.array[:n.value]
        else:
            out = np.empty((n.value, 4), np.float32)  # pageable: the engine pipelines the copy through pinned staging
        if n.value:
            check(self.L.tsdf_gather_fetch(self.h, _p(out), n.value))
        return out

    def ExtractMesh(self, volumn=None, to_host=True):
        """Triangle mesh of the zero level set from the blocks GatherVoxels(volumn) would select (all blocks if None),
        extracted on the GPU (tsdf_extract_mesh; replaces Query + KrisLibrary meshing,
        examples/ros_camera_driver/ros_offline.cc:258-350).  Returns float32 [n, 3, 3] (triangle, vertex, xyz), or
        with to_host=False only the triangle count (result stays on the device: tsdf_mesh_device_result)."""
        n = C.c_int64(0)
        bb = None if volumn is None else np.asarray(tuple(volumn), np.float32)
        check(self.L.tsdf_extract_mesh(self.h, _p(bb), None, 0, C.byref(n)))
        if not to_host:
            return n.value
        out = np.empty((n.value, 3, 3), np.float32)
        if n.value:
            check(self.L.tsdf_mesh_fetch(self.h, _p(out), n.value))
        return out

    def gather_device(self, bbox=None):
        """Run the selection + emit kernels only; the records stay in the engine's device buffer
        (tsdf_gather_device_result).  Returns the number of voxels selected."""
        n = C.c_int64(0)
        if bbox is None:
            check(self.L.tsdf_gather_valid(self.h, None, 0, C.byref(n)))
        else:
            bb = _f32(tuple(bbox), 6)
            check(self.L.tsdf_gather_in_bound(self.h, _p(bb), None, 0, C.byref(n)))
        return int(n.value)

    def GatherValid(self, pinned=False, out=None):
        return self._gather(None, pinned, out)

    def GatherVoxels(self, volumn, pinned=False, out=None):
        return self._gather(tuple(volumn), pinned, out)

    # ---- bookkeeping / parity access ---------------------------------------------------------------
    def NumActiveBlock(self):  # VoxelHashTable::NumActiveBlock, voxel_hash.cu:200
        n = C.c_int(0)
        check(self.L.tsdf_num_active_blocks(self.h, C.byref(n)))
        return n.value

    def counters(self):
        c = Counters()
        check(self.L.tsdf_get_counters(self.h, C.byref(c)))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_ if k != "reserved"}

    def skip_map_stats(self):
        """(build attempts, real rebuilds) of the RayCast skip map."""
        a, r = C.c_int64(0), C.c_int64(0)
        check(self.L.tsdf_get_skip_map_stats(self.h, C.byref(a), C.byref(r)))
        return int(a.value), int(r.value)

    def synchronize(self):
        check(self.L.tsdf_synchronize(self.h))

    def stream(self):
        return self.L.tsdf_stream(self.h)

    def set_profiling(self, on=True):
        check(self.L.tsdf_set_profiling(self.h, int(on)))

    def phase_ms(self):
        """Device ms per phase summed since set_profiling(True), and the number of timed launches."""
        ms = np.zeros(8, np.float32)
        cnt = np.zeros(8, np.int64)
        check(self.L.tsdf_get_phase_ms(self.h, _p(ms), _p(cnt)))
        names = ("upload", "allocate", "select", "integrate", "raycast", "gather")
        return {n: float(ms[i]) for i, n in enumerate(names)}, {n: int(cnt[i]) for i, n in enumerate(names)}

    def totals(self):
        c = Counters()
        n = C.c_int64(0)
        check(self.L.tsdf_get_totals(self.h, C.byref(c), C.byref(n)))
        d = {k: int(getattr(c, k)) for k, _ in Counters._fields_ if k != "reserved"}
        d["frames"] = int(n.value)
        return d

    def allocate_blocks(self, keys):
        keys = np.ascontiguousarray(keys, np.int16).reshape(-1, 3)
        check(self.L.tsdf_allocate_blocks(self.h, _p(keys), len(keys)))

    def delete_blocks(self, keys):
        keys = np.ascontiguousarray(keys, np.int16).reshape(-1, 3)
        check(self.L.tsdf_delete_blocks(self.h, _p(keys), len(keys)))

    def retrieve(self, points):
        pts = np.ascontiguousarray(points, np.int16).reshape(-1, 3)
        n = len(pts)
        tsdf, prob = np.empty(n, np.float32), np.empty(n, np.float32)
        rgbw, found = np.empty((n, 4), np.uint8), np.empty(n, np.int32)
        check(self.L.tsdf_retrieve_voxels(self.h, _p(pts), n, _p(tsdf), _p(rgbw), _p(prob), _p(found)))
        return tsdf, rgbw, prob, found.astype(bool)

    def assign(self, points, tsdf=None, rgbw=None, prob=None):
        pts = np.ascontiguousarray(points, np.int16).reshape(-1, 3)
        n = len(pts)
        tsdf = None if tsdf is None else _f32(tsdf, n)
        prob = None if prob is None else _f32(prob, n)
        rgbw = None if rgbw is None else np.ascontiguousarray(rgbw, np.uint8).reshape(n, 4)
        check(self.L.tsdf_assign_voxels(self.h, _p(pts), n, _p(tsdf), _p(rgbw), _p(prob)))

    def export(self, voxels=True):
        """All active blocks in canonical (z, y, x) order: keys, tsdf, rgbw, prob."""
        n = C.c_int(0)
        check(self.L.tsdf_export_blocks(self.h, None, None, None, None, 0, C.byref(n)))
        nb = n.value
        keys = np.zeros((nb, 3), np.int16)
        tsdf = np.zeros((nb, 512), np.float32) if voxels else None
        rgbw = np.zeros((nb, 512, 4), np.uint8) if voxels else None
        prob = np.zeros((nb, 512), np.float32) if voxels else None
        if nb:
            check(self.L.tsdf_export_blocks(self.h, _p(keys), _p(tsdf), _p(rgbw), _p(prob), nb, C.byref(n)))
        return keys, tsdf, rgbw, prob


def packed_pinned_frame(rgb, depth, ht=None, lt=None):
    """One pinned block holding the planes of a frame back to back -- [rgb | depth | ht | lt] -- and views of the planes in
    it.  The engine recognises this layout and uploads the frame with ONE DMA transfer (large transfers keep the PCIe
    link much busier than four small ones when image downloads run at the same time).  Returns (PinnedArray, dict of
    views); keep the PinnedArray alive as long as the views are used."""
    parts = [np.ascontiguousarray(rgb), np.ascontiguousarray(depth)] + ([np.ascontiguousarray(ht), np.ascontiguousarray(lt)] if ht is not None else [])
    block = PinnedArray((sum(a.nbytes for a in parts),), np.uint8)
    views, off = [], 0
    for a in parts:
        v = block.array[off:off + a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        views.append(v)
        off += a.nbytes
    d = dict(rgb=views[0], depth=views[1], ht=views[2] if ht is not None else None, lt=views[3] if ht is not None else None)
    return block, d


def packed_pinned_images(height, width, n_sets, with_depth=False):
    """n_sets pinned blocks [rgba | normal (| hit depth)] -- adjacent host images are downloaded with one transfer.
    Returns (blocks, rgba views, normal views, depth views or None)."""
    n = height * width
    blocks = [PinnedArray(((12 if with_depth else 8) * n,), np.uint8) for _ in range(n_sets)]
    rgba = [b.array[:4 * n].reshape(height, width, 4) for b in blocks]
    normal = [b.array[4 * n:8 * n].reshape(height, width, 4) for b in blocks]
    depth = [b.array[8 * n:].view(np.float32).reshape(height, width) for b in blocks] if with_depth else None
    return blocks, rgba, normal, depth


def make_host_frames(per_stream):
    """ctypes array of tsdf_host_frame, stream-major, for run_streams.  per_stream: list (streams) of lists (frames) of
    dicts with rgb / depth / ht / lt (pinned numpy views; ht / lt may be None), q, t; uint16 depth selects the 16-bit format."""
    n_frames = len(per_stream[0])
    arr = (_lib.HostFrame * (len(per_stream) * n_frames))()
    for b, frames in enumerate(per_stream):
        assert len(frames) == n_frames
        for i, f in enumerate(frames):
            x = arr[b * n_frames + i]
            x.rgb, x.depth = f["rgb"].ctypes.data, f["depth"].ctypes.data
            x.ht = None if f.get("ht") is None else f["ht"].ctypes.data
            x.lt = None if f.get("lt") is None else f["lt"].ctypes.data
            x.q[:] = [float(v) for v in f["q"]]
            x.t[:] = [float(v) for v in f["t"]]
            x.format = 1 if f["depth"].dtype == np.uint16 else 0
    return arr


def run_streams(grids, frames, first, count, width, height, max_depth, intrinsics, depthmap_factor=1.0, raycast=True, rgba=None,
                normal=None, hit_depth=None):
    """tsdf_streams_run: the calling thread drives all `grids` (one engine per stream).  rgba / normal / hit_depth: lists of
    2 * len(grids) pinned arrays (two image sets per stream) or None for images that are not downloaded."""
    L = _lib.lib()
    n = len(grids)
    eng = (C.c_void_p * n)(*[g.h for g in grids])
    K = _f32(intrinsics, 4)

    def ptrs(lst):
        if lst is None:
            return None
        assert len(lst) == 2 * n
        return (C.c_void_p * (2 * n))(*[a.ctypes.data for a in lst])

    check(L.tsdf_streams_run(n, eng, frames, len(frames) // n, first, count, width, height, depthmap_factor, max_depth, _p(K), int(raycast),
                             ptrs(rgba), ptrs(normal), ptrs(hit_depth)))
