"""bench_legs.py -- the legs of bench.py that measure BASELINE.json's configs 1, 3, 4 and 5 (config 2 is the headline
and lives in bench.py).  Every leg returns a dict that goes under "configs" in bench.py's one JSON line.

  config 1  100 frames of 640x480 at 1 cm through Integrate, then GatherValid to the host
            (examples/tsdf/offline.cc:90,169,185)
  config 3  room-scale volume at 2 cm voxels with ~50 M active voxels, ONE stream: on one GPU through the plain engine,
            on N > 1 GPUs sharded by block ownership through the C++ / NCCL data plane (libtsdf_b200_mgpu.so) --
            Integrate + exact RayCast every frame, Integrate only, and the min-composited RayCast variant; the exact
            image is compared byte for byte with a single-GPU render of the same history
  config 4  batches of 1920x1080 virtual views over the config-3 volume, max_depth 4 m (TSDFSystem::Render,
            modules/tsdf_module.cc:45-49) and 10 m (examples/tsdf/offline.cc:195); sharded: rows split across the
            GPUs + all-gather (exact) and whole views min-composited
  config 5  one independent 640x480 stream per GPU, GatherVoxels with the +-8 m query box of the ROS node
            (configs/config.yaml:4, examples/ros_camera_driver/ros_offline.cc:320-350) to pinned host memory every 10 frames

The synthetic room-scale scene of config 3: a floor of `rooms` identical 4 x 3 x 4 m rooms on a square grid of 4.5 m
pitch.  Every room is scanned with the same lap of 1280x720 frames (the config-2 frame set, so no extra frames are
generated); the camera pose of room r is the lap's pose translated by the room's offset, which is exactly what a
camera moved to that room would record.  A tour of all rooms leaves ~50 M active voxels (16 rooms).
"""
import time

import numpy as np

from disinfect_slam_b200 import synth

ROOM_PITCH = 4.5
VOXEL3, TRUNC3, MAX_DEPTH3 = 0.02, 0.12, 4.0


def quat_to_R(q):
    x, y, z, w = (float(v) for v in q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def room_offsets(rooms):
    side = int(np.ceil(np.sqrt(rooms)))
    return [np.array([ROOM_PITCH * (r % side - (side - 1) / 2), 0.0, ROOM_PITCH * (r // side - (side - 1) / 2)]) for r in range(rooms)]


def tour_cameras(st, n_frames, rooms):
    """cam_T_world of tour step s = room (s // n_frames), lap frame (s % n_frames): p_cam = R (p_world - d) + t."""
    cams = []
    for d in room_offsets(rooms):
        for i in range(n_frames):
            q, t = st["q"][i], st["t"][i]
            t2 = (np.asarray(t, np.float64) - quat_to_R(q) @ d).astype(np.float32)
            cams.append((np.asarray(q, np.float32), t2))
    return cams


class DeviceTimer:
    """CUDA events on a foreign stream (the engine's), via torch.cuda.ExternalStream."""

    def __init__(self, torch, stream_ptr, dev):
        self.torch, self.s = torch, torch.cuda.ExternalStream(stream_ptr, device=dev)
        self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def start(self):
        self.a.record(self.s)

    def stop_ms(self):
        self.b.record(self.s)
        self.b.synchronize()
        return self.a.elapsed_time(self.b)


def nvlink_bytes(local_rank):
    """(tx, rx) bytes this GPU has moved over NVLink so far (NVML throughput counters, KiB granularity), or None."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        out = []
        for fid in (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX):
            v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, pynvml.NVML_NVLINK_MAX_LINKS)])[0]  # scope = all links
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)
        return tuple(out)
    except Exception:
        return None


def max_over_ranks(torch, dist, world, dev, v):
    if world == 1:
        return v
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


# --------------------------------------------------------------------------------------------------
# config 3 + config 4
# --------------------------------------------------------------------------------------------------
def config3_and_4(args, cfg2, st0, d0, rank, world, local_rank, dev, n_frames, W, K, dist):
    import torch
    from disinfect_slam_b200 import mgpu, tsdf_grid
    H, Wd = cfg2.height, cfg2.width
    npx = H * Wd
    # as many rooms as it takes for ~50 M active voxels (BASELINE configs[2]): one room is scanned on rank 0 to learn its
    # block count, the answer is shared with the other ranks
    rooms = args.rooms
    if rooms <= 0:
        nb = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            probe = tsdf_grid.TSDFGrid(VOXEL3, TRUNC3, pool_blocks=1 << 16, table_slots=1 << 19, max_image_pixels=npx, device=local_rank)
            c0 = tour_cameras(st0, n_frames, 1)
            for s_ in range(n_frames):
                probe.IntegrateDevice(d0["rgb"][s_].data_ptr(), d0["depth"][s_].data_ptr(), d0["ht"][s_].data_ptr(), d0["lt"][s_].data_ptr(), Wd, H,
                                      MAX_DEPTH3, np.asarray(st0["K"], np.float32), c0[s_])
            nb[0] = probe.NumActiveBlock()
            probe.close()
        if world > 1:
            dist.broadcast(nb, src=0)
        rooms = int(min(64, max(4, np.ceil(50e6 / 512 / max(int(nb.item()), 1)))))
    cams = tour_cameras(st0, n_frames, rooms)
    n_tour = len(cams)
    Kv = np.asarray(st0["K"], np.float32)
    big = max(npx, 1920 * 1080)
    pool_total = 1 << 18  # the reference's NUM_BLOCK

    def planes(s):
        fi = s % n_frames
        return (d0["rgb"][fi].data_ptr(), d0["depth"][fi].data_ptr(), d0["ht"][fi].data_ptr(), d0["lt"][fi].data_ptr())

    res3 = {"workload": f"config3: floor of {rooms} rooms (4 x 3 x 4 m, {ROOM_PITCH} m pitch) scanned room after room with the 1280x720 lap "
                        f"of {n_frames} frames, voxel {VOXEL3} m, truncation {TRUNC3} m, max_depth {MAX_DEPTH3} m; one stream; timed: {K} frames of a "
                        f"second tour over the finished volume", "rooms": rooms, "n_gpus": world}
    res4 = {"workload": f"config4: batches of {args.views} virtual 1920x1080 views (fx = fy = 1400) from a ring of poses in room 0 over the config-3 volume",
            "views": args.views, "rays_per_view": 1920 * 1080, "n_gpus": world}
    v4 = [synth.Scene(cfg2).virtual_view(j, args.views, 1920, 1080, (1400.0, 1400.0, 959.5, 539.5)) for j in range(args.views)]
    off0 = room_offsets(rooms)[0]
    for v in v4:  # the ring lives in room 0
        v["t"] = (np.asarray(v["t"], np.float64) - quat_to_R(v["q"]) @ off0).astype(np.float32)

    # ---------------- one GPU: the plain engine ----------------
    def single_engine_history(n_steps):
        g = tsdf_grid.TSDFGrid(VOXEL3, TRUNC3, pool_blocks=pool_total, table_slots=1 << 21, max_image_pixels=big, device=local_rank)
        for s in range(n_steps):
            g.IntegrateDevice(*planes(s), Wd, H, MAX_DEPTH3, Kv, cams[s % n_tour])
        return g

    if world == 1:
        g = single_engine_history(n_tour)
        g.synchronize()
        out = dict(rgba=torch.empty((1080, 1920, 4), dtype=torch.uint8, device=dev), normal=torch.empty((1080, 1920, 4), dtype=torch.uint8, device=dev),
                   depth=torch.empty((1080, 1920), dtype=torch.float32, device=dev))
        cam = tsdf_grid.CameraParams(Kv, H, Wd)
        tm = DeviceTimer(torch, g.stream(), dev)

        def run(first, count, raycast):
            for s in range(first, first + count):
                g.IntegrateDevice(*planes(s), Wd, H, MAX_DEPTH3, Kv, cams[s % n_tour])
                if raycast:
                    g.RayCastDevice(MAX_DEPTH3, cam, cams[s % n_tour], out["rgba"].data_ptr(), out["normal"].data_ptr(), out["depth"].data_ptr())

        run(n_tour, W, True)
        g.synchronize()
        g.set_profiling(False)
        tm.start()
        run(n_tour + W, K, True)
        ms_rc = tm.stop_ms()
        tot = g.totals()
        g.set_profiling(False)
        tm.start()
        run(n_tour + W + K, K, False)
        ms_int = tm.stop_ms()
        tot_i = g.totals()
        n_act = g.NumActiveBlock()
        res3.update({"active_voxels": 512 * n_act, "active_blocks": n_act,
                     "integrate_raycast": {"frames_per_s": K / (ms_rc * 1e-3), "us_per_frame": 1e3 * ms_rc / K, "voxel_updates_per_s": tot["n_updated"] / (ms_rc * 1e-3),
                                           "raycast_mrays_per_s": K * npx / (ms_rc * 1e-3) / 1e6, "visible_blocks_per_frame": tot["n_visible"] / K},
                     "integrate_only": {"frames_per_s": K / (ms_int * 1e-3), "us_per_frame": 1e3 * ms_int / K, "voxel_updates_per_s": tot_i["n_updated"] / (ms_int * 1e-3)},
                     "path": "plain engine: tsdf_integrate_device + tsdf_raycast_device, frames resident in HBM, CUDA events on the engine stream"})
        # config 4 on the same volume
        cam4 = tsdf_grid.CameraParams(v4[0]["K"], 1080, 1920)
        for md in (4.0, 10.0):
            for v in v4[:2]:
                g.RayCastDevice(md, cam4, (v["q"], v["t"]), out["rgba"].data_ptr(), out["normal"].data_ptr(), out["depth"].data_ptr())
            g.synchronize()
            tm.start()
            for v in v4:
                g.RayCastDevice(md, cam4, (v["q"], v["t"]), out["rgba"].data_ptr(), out["normal"].data_ptr(), out["depth"].data_ptr())
            ms = tm.stop_ms()
            hits = float(torch.isfinite(out["depth"]).float().mean().item())
            res4[f"max_depth_{md:g}m"] = {"mrays_per_s": args.views * 1920 * 1080 / (ms * 1e-3) / 1e6, "us_per_view": 1e3 * ms / args.views,
                                         "last_view_hit_fraction": hits}
        res4["path"] = "plain engine: tsdf_raycast_device, rgba + normal + hit depth written to HBM"
        g.close()
        return res3, res4

    # ---------------- N > 1: the C++ / NCCL data plane ----------------
    def fresh_id():  # torch.distributed is plumbing: it carries the 128-byte NCCL id (one per volume), nothing else
        idt = torch.zeros(mgpu.ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(mgpu.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        return idt.cpu().numpy().tobytes()

    pool_rank = int(pool_total / world * 1.5)
    n_hist = n_tour + W + K
    frames = mgpu.make_frames([cams[s % n_tour] for s in range(n_hist + 3 * K)],
                              [planes(s) for s in range(n_hist + 3 * K)] if rank == 0 else None)
    import os

    def make_volume(exchange="fused", mirror="push", alloc="owner", tiles="interleave", mode="sharded"):
        os.environ["TSDF_MGPU_MODE"] = mode
        os.environ["TSDF_MGPU_EXCHANGE"] = exchange  # read by tsdf_mgpu_create
        os.environ["TSDF_MGPU_MIRROR"] = mirror
        os.environ["TSDF_MGPU_ALLOC"] = alloc
        os.environ["TSDF_MGPU_TILES"] = tiles
        pool = pool_rank if mode == "sharded" else pool_total
        return mgpu.ShardedVolume(VOXEL3, TRUNC3, rank, world, fresh_id(), device=local_rank, pool_blocks=pool,
                                  table_slots=max(1 << 16, 1 << int(np.ceil(np.log2(4 * pool)))), max_image_pixels=big, shard_shift=2)

    def seq(first, count, mode):
        vol.run_sequence(0, frames, first, count, Wd, H, MAX_DEPTH3, Kv, raycast_mode=mode, on_device=True)

    def timed(first, count, mode):
        vol.synchronize()
        barrier(torch, dist, world)
        vol.set_profiling(True)
        tsdf_grid._lib.check(tsdf_grid._lib.lib().tsdf_set_profiling(vol.engine, 0))  # resets the engine's counter totals
        tm.start()
        seq(first, count, mode)
        ms = max_over_ranks(torch, dist, world, dev, tm.stop_ms())
        vol.synchronize()
        cms, cn = vol.comm_ms()
        _, totals, n_act = vol.counters()
        return ms, cms, cn, totals, n_act

    per = lambda m, n, k: (1e3 * m[k] / n[k]) if n[k] else None  # noqa: E731
    # the conventional exchange first (NCCL barrier + local image + all-gather), for comparison only
    vol = make_volume("nccl", mirror="0", alloc="owner")
    tm = DeviceTimer(torch, tsdf_grid._lib.lib().tsdf_stream(vol.engine), dev)
    seq(0, n_tour, 0)
    seq(n_tour, W, 1)
    ms_n, cms_n, cn_n, _, _ = timed(n_tour + W, K, 1)
    vol.close()
    barrier(torch, dist, world)
    nccl_variant = {"frames_per_s": K / (ms_n * 1e-3), "us_per_frame": 1e3 * ms_n / K, "barrier_allreduce_us": per(cms_n, cn_n, "barrier"),
                    "image_allgather_us": per(cms_n, cn_n, "allgather"), "raycast_shared_kernels_us": per(cms_n, cn_n, "raycast_shared"),
                    "exchange": "TSDF_MGPU_EXCHANGE=nccl: 4-byte ncclAllReduce, march into a local image, grouped in-place ncclAllGather of 12 B/px"}
    # alternative form of the fused plane: allocation pass sharded by image tiles (candidate keys mailed to their owners, one more
    # peer barrier per frame), foreign TSDF planes fetched into a local cache before the march instead of pushed mirrors
    vol = make_volume("fused", mirror="pull", alloc="exchange")
    tm = DeviceTimer(torch, tsdf_grid._lib.lib().tsdf_stream(vol.engine), dev)
    seq(0, n_tour, 0)
    seq(n_tour, W, 1)
    ms_m, cms_m, cn_m, _, _ = timed(n_tour + W, K, 1)
    vol.close()
    barrier(torch, dist, world)
    nomirror_variant = {"frames_per_s": K / (ms_m * 1e-3), "us_per_frame": 1e3 * ms_m / K, "peer_barrier_before_march_us": per(cms_m, cn_m, "barrier"),
                        "march_with_fused_scatter_us": per(cms_m, cn_m, "raycast_shared"), "peer_barrier_after_march_us": per(cms_m, cn_m, "allgather"),
                        "peer_barrier_candidate_exchange_us": per(cms_m, cn_m, "exchange_barrier"),
                        "exchange": "TSDF_MGPU_MIRROR=pull TSDF_MGPU_ALLOC=exchange: pixel rays walked on 1/N of the 32x8 tiles per rank, candidate keys mailed to "
                                    "their owners' inboxes, peer barrier, owners insert; before the march the foreign TSDF planes the view can meet are fetched "
                                    "into a local cache (bulk NVLink reads) instead of being pushed by the integrate kernels"}
    # the product path with one contiguous band of rows per rank instead of round-robin tiles (TSDF_BENCH_BANDS=1: an extra
    # volume build; profiles/r2n_bench_*gpu.json hold it for 2 and 4 GPUs)
    band_variant = None
    if os.environ.get("TSDF_BENCH_BANDS") == "1":
        vol = make_volume(mirror="pull", alloc="exchange", tiles="band")
        tm = DeviceTimer(torch, tsdf_grid._lib.lib().tsdf_stream(vol.engine), dev)
        seq(0, n_tour, 0)
        seq(n_tour, W, 1)
        ms_b, cms_b, cn_b, _, _ = timed(n_tour + W, K, 1)
        vol.close()
        barrier(torch, dist, world)
        band_variant = {"frames_per_s": K / (ms_b * 1e-3), "us_per_frame": 1e3 * ms_b / K, "peer_barrier_before_march_us": per(cms_b, cn_b, "barrier"),
                        "march_with_fused_scatter_us": per(cms_b, cn_b, "raycast_shared"), "peer_barrier_after_march_us": per(cms_b, cn_b, "allgather"),
                        "exchange": "TSDF_MGPU_TILES=band: as the product path, but rank r renders rows [r H/N, (r+1) H/N): fewer foreign blocks to fetch, "
                                    "uneven finishing times"}
    # the product path: owner-filtered allocation, TSDF mirrors kept current by the integrate kernels, round-robin tiles marched
    # locally, rows stored into every rank's images, peer barriers instead of collectives
    vol = make_volume()
    tm = DeviceTimer(torch, tsdf_grid._lib.lib().tsdf_stream(vol.engine), dev)
    seq(0, n_tour, 0)          # the tour that builds the volume
    seq(n_tour, W, 1)          # warm-up with views
    nv0 = nvlink_bytes(local_rank)
    ms_rc, cms, cn, tot, n_act = timed(n_tour + W, K, 1)
    nv1 = nvlink_bytes(local_rank)
    # parity, outside every timed region: the last exact view against a single-GPU engine fed the same history
    img = np.empty((H, Wd, 4), np.uint8), np.empty((H, Wd, 4), np.uint8), np.empty((H, Wd), np.float32)
    mgpu.check(vol.L.tsdf_mgpu_fetch_images(vol.h, img[0].ctypes.data, img[1].ctypes.data, img[2].ctypes.data))
    parity = None
    want = None
    if rank == 0:
        ref = single_engine_history(n_hist)
        want = ref.RayCast(MAX_DEPTH3, tsdf_grid.CameraParams(Kv, H, Wd), cams[(n_hist - 1) % n_tour])
        bad = sum(int((a.view(np.uint8) != b.view(np.uint8)).sum()) for a, b in zip(img, want))
        n_ref = ref.NumActiveBlock()
        parity = "bit-exact" if bad == 0 and n_ref == n_act else f"MISMATCH: {bad} differing bytes, blocks {n_act} vs {n_ref}"
        ref.close()
    barrier(torch, dist, world)
    # where a frame's time goes on this rank's engine stream: a second pass over the same frames with the engine's own phase
    # timers on (not the timed run: the extra event records cost a few microseconds per frame)
    L0 = tsdf_grid._lib.lib()
    tsdf_grid._lib.check(L0.tsdf_set_profiling(vol.engine, 1))
    seq(n_tour + W, K, 1)
    vol.synchronize()
    import ctypes as C
    pms, pcnt = (C.c_float * 6)(), (C.c_int64 * 6)()
    tsdf_grid._lib.check(L0.tsdf_get_phase_ms(vol.engine, pms, pcnt))
    fetched = C.c_int64(0)
    tsdf_grid._lib.check(L0.tsdf_shared_cache_stats(vol.engine, C.byref(fetched)))
    tsdf_grid._lib.check(L0.tsdf_set_profiling(vol.engine, 0))
    phases = {n: (1e3 * pms[i] / pcnt[i] if pcnt[i] else None) for i, n in ((1, "stage_and_allocate_us"), (2, "select_us"), (3, "integrate_us"),
                                                                            (4, "map_march_scatter_us"))}
    if fetched.value:
        phases["tsdf_blocks_fetched_last_view"] = int(fetched.value)
    phases["note"] = f"rank {rank}'s engine stream, CUDA events"
    ms_int, cms_i, cn_i, tot_i, _ = timed(n_hist + K, K, 0)
    ms_cmp, cms_c, cn_c, _, _ = timed(n_hist + 2 * K, K, 2)
    coll = {"frame_broadcast_us": per(cms, cn, "broadcast"), "peer_barrier_candidate_exchange_us": per(cms, cn, "exchange_barrier"),
            "peer_barrier_before_march_us": per(cms, cn, "barrier"),
            "march_with_fused_scatter_us": per(cms, cn, "raycast_shared"), "peer_barrier_after_march_us": per(cms, cn, "allgather"),
            "composite_allreduce_us": per(cms_c, cn_c, "composite_allreduce"),
            "bytes": {"frame_broadcast": 15 * npx, "scattered_per_rank": 12 * ((H + world - 1) // world) * Wd * (world - 1), "composite_allreduce": 16 * npx},
            "note": "CUDA events around each step on its stream (a barrier's time includes waiting for the slowest rank); the frame broadcast "
                    "(grouped ncclBroadcast) runs on its own stream and overlaps the previous frame's kernels; on the engine stream there is no NCCL "
                    "kernel: the image rows travel as posted NVLink stores issued by the march kernel itself"}
    if nv0 and nv1:  # what rank 0's GPU really moved over NVLink per frame (NVML link counters around the timed frames)
        coll["nvlink_bytes_per_frame_rank0"] = {"tx": (nv1[0] - nv0[0]) / K, "rx": (nv1[1] - nv0[1]) / K,
                                                "source": "NVML NVLINK_THROUGHPUT_DATA_TX / RX, all links of rank 0's GPU; rank 0 is the broadcast root, so tx "
                                                          "carries the frame planes once per ring hop plus its image rows to 7 peers, rx the peers' rows and the voxels "
                                                          "its march read from other shards"}
    on_path = {"frame_broadcast_us": coll["frame_broadcast_us"] or 0.0,
               "peer_barriers_us": (coll["peer_barrier_before_march_us"] or 0.0) + (coll["peer_barrier_after_march_us"] or 0.0) +
                                   (coll["peer_barrier_candidate_exchange_us"] or 0.0)}
    res3.update({"active_voxels": 512 * n_act, "active_blocks": n_act, "sharded_parity": parity,
                 "integrate_raycast": {"frames_per_s": K / (ms_rc * 1e-3), "us_per_frame": 1e3 * ms_rc / K, "voxel_updates_per_s": tot["n_updated"] / (ms_rc * 1e-3),
                                       "raycast_mrays_per_s": K * npx / (ms_rc * 1e-3) / 1e6, "visible_blocks_per_frame": tot["n_visible"] / K},
                 "integrate_only": {"frames_per_s": K / (ms_int * 1e-3), "us_per_frame": 1e3 * ms_int / K, "voxel_updates_per_s": tot_i["n_updated"] / (ms_int * 1e-3),
                                    "broadcast_us": per(cms_i, cn_i, "broadcast")},
                 "integrate_raycast_min_composite": {"frames_per_s": K / (ms_cmp * 1e-3), "us_per_frame": 1e3 * ms_cmp / K},
                 "collectives": coll, "limiting_collective": max(on_path, key=on_path.get), "engine_phases": phases,
                 "nccl_exchange_variant": nccl_variant, "pull_cache_candidate_exchange_variant": nomirror_variant, "row_bands_variant": band_variant,
                 "path": "libtsdf_b200_mgpu.so: tsdf_mgpu_run_sequence (C++ loop, NCCL linked directly): grouped ncclBroadcast of the planes from rank 0's HBM, "
                         "owner-filtered allocate + integrate (every updated TSDF value also stored into all ranks' TSDF mirrors: posted NVLink stores), "
                         "peer barrier kernel, skip map from all shards' directories, tsdf_raycast_shared_scatter (1/N of the 8-row tiles, dealt round-robin; "
                         "TSDF samples from the local mirror, hit colours from the owner, finished rays stored into every "
                         "rank's images); CUDA events on the engine stream, max over ranks; voxel updates all-reduced"})
    # ---- not sharded: a replica of the whole volume on every GPU, whole views dealt round-robin (TSDF_MGPU_MODE=replicas) ----
    # Every rank integrates every frame (the broadcast is the only thing the ranks share), view k is rendered by rank
    # k % N into rank 0's memory; no barrier inside the stream.  For volumes that fit one GPU: it multiplies view
    # throughput, not capacity.
    vol_s = vol
    vol = make_volume(mode="replicas")
    tm_s, tm = tm, DeviceTimer(torch, tsdf_grid._lib.lib().tsdf_stream(vol.engine), dev)
    seq(0, n_tour, 0)
    seq(n_tour, W, 1)
    ms_r, cms_r, cn_r, tot_r, n_act_r = timed(n_tour + W, K, 1)
    img_r = np.empty((H, Wd, 4), np.uint8), np.empty((H, Wd, 4), np.uint8), np.empty((H, Wd), np.float32)
    parity_r = None
    if rank == 0:
        mgpu.check(vol.L.tsdf_mgpu_fetch_images(vol.h, img_r[0].ctypes.data, img_r[1].ctypes.data, img_r[2].ctypes.data))
        bad = sum(int((a.view(np.uint8) != b.view(np.uint8)).sum()) for a, b in zip(img_r, want))
        parity_r = "bit-exact" if bad == 0 and n_act_r == n_act else f"MISMATCH: {bad} differing bytes, blocks {n_act_r} vs {n_act}"
    vol.close()
    barrier(torch, dist, world)
    res3["replicated_volume_round_robin_views"] = {
        "frames_per_s": K / (ms_r * 1e-3), "us_per_frame": 1e3 * ms_r / K, "parity": parity_r,
        "raycast_mrays_per_s": K * npx / (ms_r * 1e-3) / 1e6, "frame_broadcast_us": per(cms_r, cn_r, "broadcast"),
        "unique_voxel_updates_per_s": tot_r["n_updated"] / (ms_r * 1e-3),
        "path": "TSDF_MGPU_MODE=replicas: ncclBroadcast of the planes, every GPU integrates every frame into its own copy of the whole volume "
                "(the updates are replicated work: only one copy is counted), view k rendered by GPU k mod N straight into rank 0's image memory; "
                "no barrier between the ranks inside the stream.  Needs the volume to fit one GPU; scales views, not capacity"}
    vol, tm = vol_s, tm_s
    # config 4 on the sharded volume: exact (rows split) and min-composited (whole views per rank)
    for md in (4.0, 10.0):
        r = {}
        for mode in ("exact", "composite"):
            fn = (lambda v: vol.RayCast(md, 1920, 1080, v["K"], (v["q"], v["t"]), to_host=False)) if mode == "exact" else \
                 (lambda v: vol.RayCastComposite(md, 1920, 1080, v["K"], (v["q"], v["t"])))
            for v in v4[:2]:
                fn(v)
            vol.synchronize()
            barrier(torch, dist, world)
            tm.start()
            for v in v4:
                fn(v)
            ms = max_over_ranks(torch, dist, world, dev, tm.stop_ms())
            r[mode] = {"mrays_per_s": args.views * 1920 * 1080 / (ms * 1e-3) / 1e6, "us_per_view": 1e3 * ms / args.views}
        res4[f"max_depth_{md:g}m"] = r
    res4["path"] = ("exact: tsdf_mgpu_raycast (peer barrier, 1/N of the rows per rank over peer memory scattered to every rank, peer barrier); composite: tsdf_mgpu_raycast_composite "
                    "(every rank marches every ray over its shard, ncclAllReduce(min) of 16 B/px)")
    vol.close()
    return res3, res4


# --------------------------------------------------------------------------------------------------
# config 1 and config 5 (one 640x480 stream per GPU)
# --------------------------------------------------------------------------------------------------
def config1_and_5(args, stream1, rank, world, local_rank, dev, dist):
    import torch
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config("config1")
    H, Wd = cfg.height, cfg.width
    npx = H * Wd
    n = len(stream1["rgb"])
    d = {k: torch.from_numpy(np.stack(stream1[k])).to(dev) for k in ("rgb", "depth", "ht", "lt")}
    Kv = np.asarray(stream1["K"], np.float32)

    def make():
        return tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 17, table_slots=1 << 19, max_image_pixels=npx, device=local_rank)

    def integrate(g, i):
        g.IntegrateDevice(d["rgb"][i].data_ptr(), d["depth"][i].data_ptr(), d["ht"][i].data_ptr(), d["lt"][i].data_ptr(), Wd, H, cfg.max_depth, Kv,
                          (stream1["q"][i], stream1["t"][i]))

    res1 = None
    if rank == 0:
        g = make()
        for i in range(min(5, n)):
            integrate(g, i)
        g.close()
        g = make()
        tm = DeviceTimer(torch, g.stream(), dev)
        g.set_profiling(False)
        tm.start()
        for i in range(n):
            integrate(g, i)
        ms = tm.stop_ms()
        tot = g.totals()
        g.GatherValid(pinned=True)  # warm-up: sizes the device result buffer and the pinned destination
        t0 = time.perf_counter()
        rec = g.GatherValid(pinned=True)
        t_pin = time.perf_counter() - t0
        g.GatherValid()  # same warm-up for the pageable path (pins its two staging buffers once)
        t0 = time.perf_counter()
        rec2 = g.GatherValid()
        t_page = time.perf_counter() - t0
        from oracle.compare import canonical_gather  # block order of a gather is unspecified: compare in canonical order (a check, outside every timed region)
        same = bool(np.array_equal(canonical_gather(rec).view(np.uint32), canonical_gather(rec2).view(np.uint32)))
        res1 = {"workload": f"config1: {n} frames 640x480, TUM intrinsics, voxel 0.01 m, Integrate every frame (frames resident in HBM), then GatherValid",
                "frames_per_s": n / (ms * 1e-3), "us_per_frame": 1e3 * ms / n, "voxel_updates_per_s": tot["n_updated"] / (ms * 1e-3),
                "active_blocks": g.NumActiveBlock(),
                "gather_valid_to_host": {"voxels": int(len(rec)), "bytes": int(rec.nbytes), "pinned_destination_ms": 1e3 * t_pin,
                                         "pinned_gbs": rec.nbytes / t_pin / 1e9, "pageable_destination_ms": 1e3 * t_page,
                                         "pageable_gbs": rec2.nbytes / t_page / 1e9, "identical": same,
                                         "note": "select + emit kernels + the 16 B/voxel copy; pinned: one DMA into caller memory; pageable (fresh "
                                                 "numpy array, what the reference's std::vector return is): 16 MB chunks through two pinned buffers, 4 host threads"}}
        g.close()
    barrier(torch, dist, world)
    # config 5: every GPU runs its own stream; GatherVoxels(+-8 m) to pinned host memory every 10 frames
    g = make()
    bbox = tsdf_grid.BoundingCube(-8.0, 8.0, -8.0, 8.0, -8.0, 8.0)
    # the query's destination is allocated once, sized for the largest answer the application expects (here 32 k blocks =
    # 256 MB of records): pinning memory per query would cost more than the query
    dest = tsdf_grid.PinnedArray((32768 * 512, 4), np.float32)
    for i in range(min(3, n)):
        integrate(g, i)
    g.GatherVoxels(bbox, out=dest.array)
    g.close()
    g = make()
    barrier(torch, dist, world)
    t0 = time.perf_counter()
    voxels = 0
    t_gather = 0.0
    for i in range(n):
        integrate(g, i)
        if i % 10 == 9:
            t1 = time.perf_counter()
            voxels += len(g.GatherVoxels(bbox, out=dest.array))
            t_gather += time.perf_counter() - t1
    g.synchronize()
    dt = max_over_ranks(torch, dist, world, dev, time.perf_counter() - t0)
    vs = voxels
    if world > 1:
        t = torch.tensor([voxels], device=dev, dtype=torch.int64)
        dist.all_reduce(t)
        vs = int(t.item())
    g.close()
    dest.free()
    res5 = {"workload": f"config5: one independent 640x480 stream per GPU ({n} frames each, own seed), Integrate every frame + GatherVoxels(+-8 m) into pinned "
                        "host memory every 10 frames; host clock around the whole run, max over ranks",
            "n_gpus": world, "frames_per_s": world * n / dt, "gathered_voxels_per_s": vs / dt, "gather_share_of_time_rank0": t_gather / dt,
            "gathers_per_stream": n // 10}
    return res1, res5
