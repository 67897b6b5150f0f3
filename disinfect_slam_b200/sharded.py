"""Multi-GPU TSDFGrid: the block set sharded across the ranks of a torch.distributed group
(the reference, utils/tsdf/voxel_tsdf.cuh:32-88, is single-GPU and has no counterpart).

One process per GPU.  Every rank owns the blocks whose super-block coordinate hashes to it
(`tsdf_block_owner`, the same mix the kernels use), with its own hash table, pool and skip map:

  Integrate     the rank holding the host frame (root) uploads it once; the planes and the camera are
                broadcast over NVLink (NCCL); every rank enumerates the whole frame but inserts and
                integrates only the blocks it owns -- no further communication.
  RayCast       every rank marches all rays against its shard (foreign space reads as unallocated) and
                emits per ray two keys  float_bits(hit_depth) << 32 | colour ; one MIN all-reduce merges
                them by nearest hit (a miss carries +inf).  Cheap, but not exact where a hit straddles shards.
  RayCastExact  every rank renders 1/N of the image rows over the WHOLE volume: the shards' tables and pools are
                mapped into every process (CUDA IPC) and foreign blocks are read over NVLink inside the march
                kernel (tsdf_raycast_shared); bit-identical to a single-GPU render; one all-gather of the tiles.
  Gather*       every rank gathers its own blocks; the records are collected on the root (sizes first).
  NumActive     SUM all-reduce.

torch.distributed is plumbing only; all arithmetic runs in the engine (libtsdf_b200.so).  The same
class runs on the CPU with the `gloo` backend and an injected backend object -- that is how
tests/test_sharded_gloo.py checks the host logic without GPUs.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import tsdf_grid

MISS_KEY = 0x7F800000 << 32  # +inf hit depth, colour 0


def pack_keys(depth, rgba, normal):
    """(H,W) float32 depth (+inf = miss), (H,W,4) uint8 images -> int64[H*W*2] min-composite keys."""
    d = np.maximum(np.asarray(depth, np.float32), np.float32(0)).view(np.uint32).astype(np.int64).reshape(-1) << 32
    out = np.empty((d.size, 2), np.int64)
    out[:, 0] = d | np.ascontiguousarray(rgba).view(np.uint32).astype(np.int64).reshape(-1)
    out[:, 1] = d | np.ascontiguousarray(normal).view(np.uint32).astype(np.int64).reshape(-1)
    return out.reshape(-1)


def unpack_keys(keys, h, w):
    """int64[H*W*2] keys -> rgba, normal (H,W,4) uint8, depth (H,W) float32 (+inf = miss)."""
    k = np.asarray(keys, np.int64).reshape(h * w, 2)
    depth = (k[:, 0] >> 32).astype(np.uint32).view(np.float32).reshape(h, w)
    rgba = (k[:, 0] & 0xFFFFFFFF).astype(np.uint32).view(np.uint8).reshape(h, w, 4)
    normal = (k[:, 1] & 0xFFFFFFFF).astype(np.uint32).view(np.uint8).reshape(h, w, 4)
    return rgba, normal, depth


class EngineBackend:
    """The CUDA engine of this rank (device tensors in, device tensors out)."""

    def __init__(self, voxel_size, truncation, rank, world, device, shard_shift, **kw):
        self.device = torch.device("cuda", device)
        self.grid = tsdf_grid.TSDFGrid(voxel_size, truncation, device=device, shard_rank=rank, shard_count=world,
                                       shard_shift=shard_shift, **kw)
        self._keys = None
        self.ext = torch.cuda.ExternalStream(self.grid.stream(), device=self.device)  # the engine's own stream
        self._img = {}
        self._attached = False

    def attach_peers(self, group, world):
        """Exchange the CUDA IPC handles of every shard's table / pool and map them (once)."""
        if self._attached:
            return
        mine = torch.from_numpy(self.grid.ipc_export()).to(self.device)
        blobs = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(blobs, mine, group=group)
        self.grid.ipc_attach(torch.stack(blobs).cpu().numpy())
        self._attached = True

    def raycast_tile(self, max_depth, w, h, K, q, t, rank, world, group):
        """Rows [rank * rows, (rank + 1) * rows) of the view over the whole sharded volume, gathered on every rank.
        Everything is enqueued on the engine stream: the 1-element all-reduce in front is the device-side barrier
        that orders every shard's Integrate before any peer reads its voxels, the all-gather behind orders every
        peer's reads before the next Integrate."""
        rows = (h + world - 1) // world
        key = (w, h, world)
        tile = rows * w * 4  # bytes of one rank's rows of one 4-byte-per-pixel image
        if key not in self._img:
            # one packed tile per rank: [rgba rows | normal rows | depth rows], so that ONE all-gather moves all three
            self._img[key] = dict(local=torch.zeros(3 * tile, dtype=torch.uint8, device=self.device),
                                  full=torch.zeros(world * 3 * tile, dtype=torch.uint8, device=self.device),
                                  flag=torch.zeros(1, dtype=torch.int32, device=self.device))
        im = self._img[key]
        cam = tsdf_grid.CameraParams(K, h, w)
        base = im["local"].data_ptr() - rank * tile  # the kernel indexes whole-image pixels: row0 lands on the tile start
        with torch.cuda.stream(self.ext):
            dist.all_reduce(im["flag"], group=group)
            self.grid.RayCastShared(max_depth, cam, (q, t), rank * rows, rows, base, base + tile, base + 2 * tile)
            dist.all_gather_into_tensor(im["full"], im["local"], group=group)
            # un-interleave [rank][image][rows] -> three images; these copies must stay on the engine stream too
            g = im["full"].view(world, 3, tile)
            rgba = g[:, 0].reshape(world * rows, w, 4)[:h]
            normal = g[:, 1].reshape(world * rows, w, 4)[:h]
            depth = g[:, 2].contiguous().view(torch.float32).reshape(world * rows, w)[:h]
        return rgba, normal, depth

    def integrate(self, planes, w, h, max_depth, K, q, t):
        ev = torch.cuda.Event()  # the broadcast ran on torch's stream: the engine stream waits for it on the device
        ev.record(torch.cuda.current_stream(self.device))
        self.grid.IntegrateDevice(planes["rgb"].data_ptr(), planes["depth"].data_ptr(), planes["ht"].data_ptr(),
                                  planes["lt"].data_ptr(), w, h, max_depth, K, (q, t), after_event=ev.cuda_event)

    def raycast_keys(self, max_depth, w, h, K, q, t):
        if self._keys is None or self._keys.numel() != 2 * w * h:
            self._keys = torch.empty(2 * w * h, dtype=torch.int64, device=self.device)
        cam = tsdf_grid.CameraParams(K, h, w)
        self.grid.RayCastDevice(max_depth, cam, (q, t), d_packed=self._keys.data_ptr())
        ev = torch.cuda.Event()  # the collective (torch's stream) waits for the engine stream on the device, not on the host
        ev.record(self.ext)
        torch.cuda.current_stream(self.device).wait_event(ev)
        return self._keys

    def gather(self, bbox):
        g = self.grid.GatherValid() if bbox is None else self.grid.GatherVoxels(bbox)
        return torch.from_numpy(g).to(self.device)

    def num_active(self):
        return self.grid.NumActiveBlock()

    def synchronize(self):
        self.grid.synchronize()

    def close(self):
        self.grid.close()


class ShardedTSDFGrid:
    def __init__(self, voxel_size, truncation, group=None, device=None, root=0, shard_shift=2, backend=None, **engine_kw):
        self.group, self.root = group, root
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.shard_shift = shard_shift
        if backend is None:
            device = torch.cuda.current_device() if device is None else device
            backend = EngineBackend(voxel_size, truncation, self.rank, self.world, device, shard_shift, **engine_kw)
        self.backend = backend
        self.comm_device = backend.device

    # ---- TSDFGrid::Integrate ------------------------------------------------------------------------
    def Integrate(self, img_rgb, img_depth, img_ht, img_lt, max_depth, intrinsics, cam_T_world, width=None, height=None):
        """Root passes the host frame (numpy, like TSDFGrid.Integrate); the other ranks pass None for the four
        images and the camera, plus width / height (or call with the same arrays -- they are ignored)."""
        hdr = torch.zeros(16, dtype=torch.float64)
        if self.rank == self.root:
            h, w = img_depth.shape
            q, t = cam_T_world
            hdr[:14] = torch.tensor([w, h, max_depth, *np.asarray(intrinsics, np.float64)[:4], *np.asarray(q, np.float64)[:4],
                                     *np.asarray(t, np.float64)[:3]], dtype=torch.float64)
        hdr = hdr.to(self.comm_device)
        dist.broadcast(hdr, src=self.root, group=self.group)
        hdr = hdr.cpu().numpy()
        w, h, max_depth = int(hdr[0]), int(hdr[1]), float(np.float32(hdr[2]))
        K, q, t = hdr[3:7].astype(np.float32), hdr[7:11].astype(np.float32), hdr[11:14].astype(np.float32)
        n = w * h
        # one packed buffer [depth | ht | lt | rgb]: a single broadcast of 15 bytes per pixel
        buf = torch.empty(15 * n, dtype=torch.uint8, device=self.comm_device)
        if self.rank == self.root:
            host = np.empty(15 * n, np.uint8)
            host[0:4 * n] = np.ascontiguousarray(img_depth, np.float32).view(np.uint8).reshape(-1)
            host[4 * n:8 * n] = np.ascontiguousarray(img_ht, np.float32).view(np.uint8).reshape(-1)
            host[8 * n:12 * n] = np.ascontiguousarray(img_lt, np.float32).view(np.uint8).reshape(-1)
            host[12 * n:15 * n] = np.ascontiguousarray(img_rgb, np.uint8).reshape(-1)
            buf.copy_(torch.from_numpy(host))
        self.IntegrateBroadcast(buf, w, h, max_depth, K, q, t)

    def IntegrateBroadcast(self, buf, w, h, max_depth, K, q, t):
        """`buf` = packed [depth f32 | ht f32 | lt f32 | rgb u8x3] tensor on the communication device, valid on
        root; camera arguments must be identical on every rank (bench path: frames already device-resident)."""
        n = w * h
        dist.broadcast(buf, src=self.root, group=self.group)
        planes = {"depth": buf[0:4 * n].view(torch.float32), "ht": buf[4 * n:8 * n].view(torch.float32),
                  "lt": buf[8 * n:12 * n].view(torch.float32), "rgb": buf[12 * n:15 * n]}
        self.backend.integrate(planes, w, h, max_depth, K, q, t)
        self._bufs = (getattr(self, "_bufs", []) + [buf])[-3:]  # keep the planes alive until the engine has consumed them

    # ---- TSDFGrid::RayCast ---------------------------------------------------------------------------
    def RayCastKeys(self, max_depth, virtual_cam, cam_T_world):
        q, t = cam_T_world
        h, w = int(virtual_cam.img_h), int(virtual_cam.img_w)
        keys = self.backend.raycast_keys(max_depth, w, h, np.asarray(virtual_cam.intrinsics, np.float32), np.asarray(q, np.float32),
                                         np.asarray(t, np.float32))
        if self.world > 1:
            dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=self.group)  # nearest hit wins
        return keys

    def RayCastExact(self, max_depth, virtual_cam, cam_T_world, to_host=True):
        """Same camera on every rank; each rank renders 1/N of the rows over the whole volume (peer memory over
        NVLink) -- bit-identical to a single-GPU RayCast.  Returns (rgba, normal, hit_depth) on every rank."""
        q, t = cam_T_world
        h, w = int(virtual_cam.img_h), int(virtual_cam.img_w)
        self.backend.attach_peers(self.group, self.world)
        rgba, normal, depth = self.backend.raycast_tile(max_depth, w, h, np.asarray(virtual_cam.intrinsics, np.float32),
                                                        np.asarray(q, np.float32), np.asarray(t, np.float32), self.rank, self.world, self.group)
        if not to_host:
            return rgba, normal, depth
        self.backend.synchronize()
        return rgba.cpu().numpy(), normal.cpu().numpy(), depth.cpu().numpy()

    def RayCast(self, max_depth, virtual_cam, cam_T_world):
        """Same camera on every rank; returns (rgba, normal, hit_depth) on every rank."""
        keys = self.RayCastKeys(max_depth, virtual_cam, cam_T_world)
        return unpack_keys(keys.cpu().numpy(), int(virtual_cam.img_h), int(virtual_cam.img_w))

    # ---- TSDFGrid::GatherValid / GatherVoxels -------------------------------------------------------------
    def _gather(self, bbox):
        mine = self.backend.gather(bbox).reshape(-1, 4)
        sizes = torch.zeros(self.world, dtype=torch.int64, device=self.comm_device)
        sizes[self.rank] = mine.shape[0]
        dist.all_reduce(sizes, group=self.group)
        sizes = sizes.cpu().tolist()
        if self.rank == self.root:
            parts = [torch.empty((s, 4), dtype=torch.float32, device=self.comm_device) for s in sizes]
            parts[self.root] = mine
            reqs = [dist.irecv(parts[r], src=r, group=self.group) for r in range(self.world) if r != self.root and sizes[r]]
            for rq in reqs:
                rq.wait()
            return torch.cat(parts).cpu().numpy()
        if sizes[self.rank]:
            dist.send(mine.contiguous(), dst=self.root, group=self.group)
        return np.empty((0, 4), np.float32)

    def GatherValid(self):
        """All voxels of all shards on the root rank (empty array elsewhere)."""
        return self._gather(None)

    def GatherVoxels(self, volumn):
        return self._gather(tuple(volumn))

    def NumActiveBlock(self):
        n = torch.tensor([self.backend.num_active()], dtype=torch.int64, device=self.comm_device)
        dist.all_reduce(n, group=self.group)
        return int(n.item())

    def synchronize(self):
        self.backend.synchronize()

    def close(self):
        self.backend.close()
