#!/usr/bin/env python
"""bench.py -- throughput of the voxel-hashed semantic TSDF hot path on B200.

One "step" = one pass of the hot path over one batch of synthetic input: for each of the
`--streams` independent RGB-D streams resident on this GPU, one frame through
TSDFGrid::Integrate followed by one TSDFGrid::RayCast from the same camera (BASELINE.json
configs[1]: aligned 1280x720 L515/ZED-style frames, 5 mm voxels, "Integrate + RayCast every
frame").  Several streams are interleaved per GPU so that the per-step working set (voxel blocks
+ frame planes of all streams) is larger than the 126 MB L2: no stream finds its blocks cached
from its own previous frame, which is what the HBM roofline assumes.

Legs (own arm):
  value     frames already resident in HBM, tsdf_integrate_device + tsdf_raycast_device, CUDA
            events bracketing exactly K steps, max over ranks
  roofline  same frames on fresh engines with per-kernel CUDA events on the engine stream
            (kernels serialised), achieved = algorithmic bytes / kernel time
  e2e       host (pinned) buffers through tsdf_integrate + tsdf_raycast: H2D of every frame and
            D2H of both rendered images + hit depth inside the timed region
  cpu       the scalar/OpenMP CPU oracle on a bounded sample of the same workload (rank 0, N=1)

`--impl reference` runs the reference's OWN CUDA TSDFGrid (utils/tsdf/*.cu, unmodified, rebuilt for
sm_100a into oracle/_ref by oracle/build_ref.sh) through its public Integrate / RayCast API on the
same workload -- the reference has no CPU implementation of this path -- and falls back to the CPU
oracle port only if that library is absent.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time
from dataclasses import replace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from disinfect_slam_b200 import synth  # noqa: E402

L2_BYTES = 126e6
METRIC, UNIT = "integrated_frames_per_s", "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--streams", type=int, default=4, help="independent RGB-D streams interleaved per GPU")
    ap.add_argument("--lap", type=int, default=100, help="distinct frames per stream before the trajectory repeats")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the sharded single-stream leg at N > 1")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the BASELINE configs 1, 3, 4, 5 legs")
    ap.add_argument("--no-config15", action="store_true", help="skip the config 1 / config 5 legs (development runs)")
    ap.add_argument("--config1-frames", type=int, default=100)
    ap.add_argument("--rooms", type=int, default=0, help="config 3: rooms of the synthetic floor (each = one lap of the config-2 frames); 0 = as many as ~50 M active voxels take")
    ap.add_argument("--views", type=int, default=16, help="config 4: 1920x1080 virtual views per batch")
    ap.add_argument("--scale", type=float, default=1.0, help="image scale (debug only; 1.0 = the named config)")
    ap.add_argument("--blocking-sync", default="auto", choices=["auto", "on", "off"],
                    help="host waits of the e2e leg yield the CPU (TSDF_FLAG_BLOCKING_SYNC); auto = when the engine threads of all ranks reach half of the cores")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# synthetic frames (generated before CUDA is touched: the pool forks)
# --------------------------------------------------------------------------------------------------
def _gen_one(job):
    cfg, i = job
    f = synth.Scene(cfg).frame(i)
    return f["rgb"], f["depth"], f["ht"], f["lt"], f["q"], f["t"], f["K"]


def stream_cfg(cfg, rank, b):
    return replace(cfg, seed=cfg.seed + 1000 * rank + 17 * b)


def generate_streams(cfg, rank, n_streams, n_frames):
    import multiprocessing as mp
    jobs = [(stream_cfg(cfg, rank, b), i) for b in range(n_streams) for i in range(n_frames)]
    procs = max(1, min(len(jobs), (os.cpu_count() or 2)))
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(_gen_one, jobs, chunksize=1)
    else:
        res = [_gen_one(j) for j in jobs]
    out = []
    for b in range(n_streams):
        fr = res[b * n_frames:(b + 1) * n_frames]
        out.append(dict(rgb=[r[0] for r in fr], depth=[r[1] for r in fr], ht=[r[2] for r in fr], lt=[r[3] for r in fr],
                        q=[r[4] for r in fr], t=[r[5] for r in fr], K=fr[0][6]))
    return out


# --------------------------------------------------------------------------------------------------
# clocks (NVML polled from a thread while the GPU legs run)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples = []  # (t, sm_mhz, reasons_mask, power_w)
        self.ok = False
        self._stop = threading.Event()
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report it, never fake numbers
            self.err = repr(e)
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), int(mhz), int(rs), pw))
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.ok:
            self.th.start()

    def stop(self):
        self._stop.set()
        if self.ok:
            self.th.join(timeout=1.0)

    def summary(self, windows):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap,
                 "hw_power_brake": nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown,
                 "applications_clocks_setting": nv.nvmlClocksThrottleReasonApplicationsClocksSetting}
        sel = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        where = "timed regions"
        if len(sel) < 3:
            sel, where = self.samples, "whole run (timed regions too short for >= 3 samples)"
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        mask = 0
        for s in sel:
            mask |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in sel), "sm_max_mhz": self.max_mhz,
                "reasons": [k for k, v in names.items() if mask & v], "samples": len(sel), "window": where,
                "power_w_max": max(s[3] for s in sel)}


# --------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline and --impl reference)
# --------------------------------------------------------------------------------------------------
def oracle_threads():
    n = os.environ.get("OMP_NUM_THREADS")
    return int(n) if n and n.isdigit() else (os.cpu_count() or 1)


def oracle_step(o, cfg, st, i):
    o.integrate(st["rgb"][i], st["depth"][i], st["ht"][i], st["lt"][i], cfg.max_depth, st["K"], st["q"][i], st["t"][i])
    return o.raycast(cfg.max_depth, cfg.width, cfg.height, st["K"], st["q"][i], st["t"][i])[3]


def cpu_sample(cfg, st, budget_s, max_frames):
    """Oracle (C, -O2, OpenMP over blocks / rays) on the first frames of stream 0 until the budget is spent."""
    from oracle.oracle import Oracle
    o = Oracle(cfg.voxel_size, cfg.truncation)
    n, t0 = 0, time.perf_counter()
    samples = switches = hits = 0
    while n < max_frames and (n < 2 or time.perf_counter() - t0 < budget_s):
        c = oracle_step(o, cfg, st, n)
        samples, switches, hits = samples + c["samples"], switches + c["block_switches"], hits + c["hits"]
        n += 1
    dt = time.perf_counter() - t0
    o.close()
    # B_ray = sum over rays (4 B per sample + 12 B per block switch) + 8 B per ray written (+ 4 B hit depth)
    rays = cfg.width * cfg.height
    ray_stats = {"frames_counted": n, "samples_per_ray": samples / (n * rays), "block_switches_per_ray": switches / (n * rays),
                 "hit_fraction": hits / (n * rays), "algorithmic_bytes_per_view": (4 * samples + 12 * switches) / n + 12 * rays,
                 "model": "the reference's march counted by the oracle (every sample looked up, 12-byte hash entry per block switch): "
                          "4 B x samples + 12 B x block switches + 12 B per ray written (rgba, normal, hit depth)"}
    return n / dt, n, dt, ray_stats


def run_reference(args, cfg, rank, world):
    """--impl reference.  Preferred: the reference's OWN CUDA TSDFGrid (utils/tsdf/*.cu, unmodified, rebuilt for
    sm_100a with its Release flags -> oracle/_ref/libref_tsdf.so) on this box's GPU 0, through its public
    Integrate / RayCast API, same streams-per-GPU workload as the own arm.  Fallback (library absent): the CPU
    oracle port on the host cores."""
    from oracle import ref_cuda
    if ref_cuda.available(parity=False):
        # one replica set per GPU, like the own arm: every rank pins itself to its GPU BEFORE CUDA is initialised (the
        # reference's TSDFGrid knows nothing about devices: it runs on device 0 of the process)
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = vis.split(",") if vis else None
        os.environ["CUDA_VISIBLE_DEVICES"] = ids[local_rank] if ids and local_rank < len(ids) else str(local_rank)
        import torch
        if torch.cuda.is_available():
            return run_reference_cuda(args, cfg, rank, world)
    if rank != 0:
        return
    from oracle.oracle import Oracle
    n_frames = min(args.lap, args.warmup + args.steps)
    st = generate_streams(cfg, 0, 1, n_frames)[0]
    o = Oracle(cfg.voxel_size, cfg.truncation)
    for i in range(args.warmup):
        oracle_step(o, cfg, st, i % n_frames)
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        oracle_step(o, cfg, st, i % n_frames)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, 1, "one stream, one frame per step (bounded sample of the same workload)"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": oracle_threads(), "kind": "port",
                             "sample": f"{args.steps} frames (Integrate + RayCast each) of stream 0 after {args.warmup} warm-up frames; "
                                       "oracle/tsdf_oracle.c -O2, OpenMP over visible blocks and image rows, allocation pass scalar"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "oracle/_ref absent: CPU oracle port timed instead of the reference's CUDA rebuild"}
    emit(line)


def run_reference_cuda(args, cfg, rank=0, world=1):
    from oracle.ref_cuda import RefTSDFGrid
    B, K, W = args.streams, args.steps, args.warmup
    n_frames = min(args.lap, W + K)
    streams = generate_streams(cfg, rank, B, n_frames)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")  # plumbing: one barrier and one MAX of the wall time
    grids = [RefTSDFGrid(cfg.voxel_size, cfg.truncation, parity=False) for _ in range(B)]
    H, Wd = cfg.height, cfg.width
    npx = H * Wd

    def step(b, i):
        fi = i % n_frames
        st, g = streams[b], grids[b]
        g.integrate(st["rgb"][fi], st["depth"][fi], st["ht"][fi], st["lt"][fi], cfg.max_depth, st["K"], st["q"][fi], st["t"][fi])
        g.raycast(cfg.max_depth, Wd, H, st["K"], st["q"][fi], st["t"][fi], download=True)

    for i in range(W):
        for b in range(B):
            step(b, i)
    start = threading.Barrier(B + 1)

    def worker(b):
        start.wait()
        for i in range(W, W + K):
            step(b, i)

    ths = [threading.Thread(target=worker, args=(b,)) for b in range(B)]
    for t in ths:
        t.start()
    if dist is not None:
        dist.barrier()
    start.wait()
    t0 = time.perf_counter()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    act = [g.num_active() for g in grids]
    if dist is not None:
        import torch
        tt = torch.tensor([dt], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        dist.barrier()
    if rank != 0:
        for g in grids:
            g.close()
        if dist is not None:
            dist.destroy_process_group()
        return
    v = world * B * K / dt
    desc = (f"the reference's own CUDA TSDFGrid (utils/tsdf/*.cu unmodified, -O3 -DNDEBUG, sm_100a), one replica set per GPU on {world} GPU(s) of this box: "
            f"{B} streams per GPU, one host thread each, TSDFGrid::Integrate(host cv::Mat planes) + TSDFGrid::RayCast + download of both images per frame; "
            f"{K} timed steps after {W} warm-up, wall time = max over ranks; active blocks at end (rank 0) {act}")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * dt / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, B, f"{B} independent streams interleaved per GPU, one frame of each per step", world),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": B * world, "kind": "reference", "sample": desc},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 15 * npx * B, "d2h_bytes_per_step": 8 * npx * B},
            "gpu_launches": 0,
            "note": "the reference has no CPU TSDF path: this arm runs its CUDA kernels (oracle/_ref) on one B200; pageable host buffers as in its API"}
    emit(line)
    for g in grids:
        g.close()
    if dist is not None:
        dist.destroy_process_group()


def workload_config(cfg, streams, extra, world=1):
    """Identical for the own arm and the reference arm (the driver compares the two)."""
    return {"workload": f"{cfg.name}: Integrate + RayCast every frame; {extra}", "width": cfg.width, "height": cfg.height,
            "voxel_size_m": cfg.voxel_size, "truncation_m": cfg.truncation, "max_depth_m": cfg.max_depth,
            "block": "8^3 voxels", "streams_per_gpu": streams, "pool_blocks": cfg.pool_blocks,
            "frames_per_step": streams * world, "parallelism": f"replicas x{world} (independent streams, no data-path collective)",
            "l2": "inputs larger than L2: the streams of a GPU are interleaved so that the per-step working set (visible voxel blocks + frame "
                  "planes of all streams, ~400 MB with 4 streams) exceeds the 126 MB L2; no explicit flush"}


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (e.g. NCCL's version banner)
    was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = synth.config(args.workload)
    if args.scale != 1.0:
        cfg = cfg.scaled(args.scale)
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    B, K, W = args.streams, args.steps, args.warmup
    n_frames = min(args.lap, W + K)
    streams = generate_streams(cfg, rank, B, n_frames)  # before CUDA init (fork)
    # BASELINE configs[0] / configs[4]: one 640x480 stream per GPU (its own seed per rank)
    cfg1 = synth.config("config1")
    n1 = 0 if args.no_extra_configs else min(cfg1.n_frames, args.config1_frames)
    stream1 = generate_streams(replace(cfg1, seed=cfg1.seed + 31 * rank), 0, 1, n1)[0] if n1 else None

    import torch
    import torch.distributed as dist
    from disinfect_slam_b200 import tsdf_grid
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, Wd = cfg.height, cfg.width
    npx = H * Wd
    cam = tsdf_grid.CameraParams(streams[0]["K"], H, Wd)

    # more synchronously-waiting host threads than cores (e.g. 8 ranks x 4 streams on 16 cores): yield instead of spinning
    # spinning waits are the fastest while cores are plentiful (1 GPU, 4 threads on 16 cores: 3150 vs 2960 frames/s) and the
    # slowest once half of them would spin (4 GPUs: 4430 spinning vs 5850 yielding)
    oversubscribed = world * B >= 0.5 * (os.cpu_count() or 1) if args.blocking_sync == "auto" else args.blocking_sync == "on"

    def make_engines(blocking=False):
        return [tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots,
                                   max_image_pixels=npx, device=local_rank, blocking_sync=blocking) for _ in range(B)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host stacks (e2e inputs, and the source of the device-resident copies)
    pinned = []
    for st in streams:
        p = {"rgb": tsdf_grid.PinnedArray((n_frames, H, Wd, 3), np.uint8), "depth": tsdf_grid.PinnedArray((n_frames, H, Wd), np.float32),
             "ht": tsdf_grid.PinnedArray((n_frames, H, Wd), np.float32), "lt": tsdf_grid.PinnedArray((n_frames, H, Wd), np.float32)}
        for k in p:
            for i in range(n_frames):
                p[k].array[i] = st[k][i]
        pinned.append(p)
    dres = [{k: torch.from_numpy(p[k].array).to(dev) for k in p} for p in pinned]
    d_out = [dict(rgba=torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev), normal=torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev),
                  depth=torch.empty((H, Wd), dtype=torch.float32, device=dev)) for _ in range(B)]
    torch.cuda.synchronize()

    def device_step(engs, i, serialise=False):
        fi = i % n_frames
        for b, g in enumerate(engs):
            st, d = streams[b], dres[b]
            pose = (st["q"][fi], st["t"][fi])
            g.IntegrateDevice(d["rgb"][fi].data_ptr(), d["depth"][fi].data_ptr(), d["ht"][fi].data_ptr(), d["lt"][fi].data_ptr(),
                              Wd, H, cfg.max_depth, st["K"], pose)
            g.RayCastDevice(cfg.max_depth, cam, pose, d_out[b]["rgba"].data_ptr(), d_out[b]["normal"].data_ptr(),
                            d_out[b]["depth"].data_ptr())
            if serialise:
                g.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    windows = []

    # ---------------- leg 1: device-resident throughput (the `value`) ----------------
    engs = make_engines()
    ext = [torch.cuda.ExternalStream(g.stream(), device=dev) for g in engs]
    for i in range(W):
        device_step(engs, i)
    for g in engs:
        g.synchronize()
        g.set_profiling(False)  # resets the counter totals
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    t_a = time.perf_counter()
    ev0.record(cur)
    for s in ext:
        s.wait_event(ev0)
    for i in range(W, W + K):
        device_step(engs, i)
    for s in ext:
        e = torch.cuda.Event()
        e.record(s)
        cur.wait_event(e)
    ev1.record(cur)
    for g in engs:
        g.synchronize()
    torch.cuda.synchronize()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    barrier()
    tot1 = [g.totals() for g in engs]
    upd_local = sum(t["n_updated"] for t in tot1)
    for g in engs:
        g.close()
    del engs, ext

    # ---------------- leg 2: per-kernel times on the launching stream (roofline) ----------------
    engs = make_engines()
    for i in range(W):
        device_step(engs, i, serialise=True)
    for g in engs:
        g.set_profiling(True)
    t_a = time.perf_counter()
    for i in range(W, W + K):
        device_step(engs, i, serialise=True)
    for g in engs:
        g.synchronize()
    t_b = time.perf_counter()
    windows.append((t_a, t_b))
    ph_ms, ph_n = {}, {}
    tot = {}
    for g in engs:
        m, n = g.phase_ms()
        for k in m:
            ph_ms[k] = ph_ms.get(k, 0.0) + m[k]
            ph_n[k] = ph_n.get(k, 0) + n[k]
        for k, v in g.totals().items():
            tot[k] = tot.get(k, 0) + v
    # GatherValid / GatherVoxels on the volumes just built (BASELINE config 5's mesh-export query): selection +
    # emit kernels timed with CUDA events (result stays on the GPU), then the full call with the D2H of every record
    gather = None
    try:
        for g in engs:
            g.gather_device(None)  # warm-up: grows the engine's result buffer once
        for g in engs:
            g.set_profiling(True)
        n_vox = {"valid": 0, "bound": 0}
        bbox = tsdf_grid.BoundingCube(-1.0, 1.0, -1.5, 1.5, -2.0, 2.0)
        for rep in range(3):
            for g in engs:
                n_vox["valid"] = g.gather_device(None)
        m_valid = sum(g.phase_ms()[0]["gather"] for g in engs) / (3 * len(engs))
        for g in engs:
            g.set_profiling(True)
        for rep in range(3):
            for g in engs:
                n_vox["bound"] = g.gather_device(bbox)
        m_bound = sum(g.phase_ms()[0]["gather"] for g in engs) / (3 * len(engs))
        engs[0].GatherValid()  # warm-up of the host path (pins the two staging buffers once)
        t0 = time.perf_counter()
        rec = engs[0].GatherValid()
        t_host = time.perf_counter() - t0
        engs[0].GatherValid(pinned=True)  # sizes the pinned destination
        t0 = time.perf_counter()
        rec_p = engs[0].GatherValid(pinned=True)
        t_pin = time.perf_counter() - t0
        n_act = engs[0].NumActiveBlock()
        # mesh extraction on the GPU: what replaces "download every voxel + mesh on one CPU core"
        for g in engs:
            g.ExtractMesh(None, to_host=False)  # warm-up: sizes the result buffer
            g.set_profiling(True)
        for rep in range(3):
            for g in engs:
                n_tri = g.ExtractMesh(None, to_host=False)
        m_mesh = sum(g.phase_ms()[0]["gather"] for g in engs) / (3 * len(engs))
        t0 = time.perf_counter()
        tris = engs[0].ExtractMesh(None)
        t_mesh_host = time.perf_counter() - t0
        gb = lambda nv, ms: (8 * n_act + 20 * nv) / (ms * 1e-3) / 1e9 if ms > 0 else None  # noqa: E731
        gather = {"gather_valid": {"voxels": n_vox["valid"], "device_ms": m_valid, "hbm_gbs": gb(n_vox["valid"], m_valid)},
                  "gather_in_bound": {"voxels": n_vox["bound"], "device_ms": m_bound, "hbm_gbs": gb(n_vox["bound"], m_bound)},
                  "gather_valid_to_host": {"voxels": int(len(rec)), "ms": 1e3 * t_host, "d2h_bytes": int(rec.nbytes), "gbs": rec.nbytes / t_host / 1e9,
                                           "note": "fresh pageable numpy destination (what the reference's std::vector return is): select + emit kernels + the "
                                                   "16 B/voxel copy, pipelined through two pinned 16 MB buffers with 4 host threads"},
                  "gather_valid_to_pinned_host": {"voxels": int(len(rec_p)), "ms": 1e3 * t_pin, "d2h_bytes": int(rec_p.nbytes), "gbs": rec_p.nbytes / t_pin / 1e9,
                                                  "note": "caller-provided pinned destination: select + emit kernels + one DMA transfer"},
                  "extract_mesh": {"triangles": int(n_tri), "device_ms": m_mesh, "to_host_ms": 1e3 * t_mesh_host, "d2h_bytes": int(tris.nbytes),
                                   "note": "count + emit kernels over the same block selection as gather_valid (incl. two 8-byte read-backs); "
                                           "to_host adds the 36 B/triangle copy into a pageable numpy array"},
                  "bytes_model": "8 B per directory entry + 4 B read + 16 B write per emitted voxel; device_ms includes the selection pass and its 4-byte count read-back"}
    except Exception as e:  # the gather report is supplementary: never lose the main line over it
        gather = {"error": repr(e)}
    for g in engs:
        g.close()
    del engs

    # ---------------- leg 3: end to end through the host-buffer C ABI ----------------
    # Every step uploads each stream's frame from pinned host memory and brings the rendered rgba + normal images -- what
    # TSDFGrid::RayCast delivers -- back to pinned host memory, all inside the timed region.
    #   e2e       tsdf_streams_run: ONE host thread per GPU drives all streams (tsdf_integrate_enqueue +
    #             tsdf_raycast_async + tsdf_raycast_wait one step later); float32 planes, 15 B/px in, 8 B/px out
    #   e2e_u16   the same loop on the sensor's own formats (tsdf_integrate_u16: 16-bit depth + probabilities, 9 B/px in)
    #   e2e_sync  tsdf_integrate + tsdf_raycast (+ hit depth), blocking, one Python thread per stream: the reference's
    #             call pattern
    def run_e2e_native(fmt16):
        engs = make_engines(blocking=False)
        # every frame = one pinned block [rgb | depth | ht | lt] (what a capture thread fills); the 16-bit planes are
        # what a sensor / log delivers (quantised once, outside the timed region)
        keep, per_stream = [], []
        q16 = lambda x, sc: np.clip(np.rint(x.astype(np.float64) * sc), 0, 65535).astype(np.uint16)  # noqa: E731
        for st in streams:
            fl = []
            for i in range(n_frames):
                if fmt16:
                    blk, d = tsdf_grid.packed_pinned_frame(st["rgb"][i], q16(st["depth"][i], cfg.depth_factor), q16(st["ht"][i], 65535), q16(st["lt"][i], 65535))
                else:
                    blk, d = tsdf_grid.packed_pinned_frame(st["rgb"][i], st["depth"][i], st["ht"][i], st["lt"][i])
                keep.append(blk)
                fl.append(dict(d, q=st["q"][i], t=st["t"][i]))
            per_stream.append(fl)
        frames = tsdf_grid.make_host_frames(per_stream)
        oblk, rgba, normal, _ = tsdf_grid.packed_pinned_images(H, Wd, 2 * B)

        def run(first, count):
            tsdf_grid.run_streams(engs, frames, first, count, Wd, H, cfg.max_depth, streams[0]["K"], depthmap_factor=cfg.depth_factor, raycast=True,
                                  rgba=rgba, normal=normal, hit_depth=None)

        run(0, W)
        barrier()
        t_a = time.perf_counter()
        run(W, K)  # returns after every engine has been synchronised: the last images are in host memory
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        e2e_s = t_b - t_a
        if world > 1:
            tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        checksum = float(sum(int(rgba[2 * b + ((W + K - 1) & 1)][::97, ::89].sum()) for b in range(B)))
        bpp_in = 9 if fmt16 else 15
        res = {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": bpp_in * npx * B, "d2h_bytes_per_step": 8 * npx * B + 128 * B,
               "api": ("tsdf_streams_run: two host threads per GPU drive its %d streams -- %s (no host wait), tsdf_raycast_async, tsdf_raycast_wait one "
                       "step later; every frame is one pinned block [rgb | depth | ht | lt] (one DMA transfer), rgba + normal of every frame reach one "
                       "pinned block of host memory inside the timed region"
                       % (B, "tsdf_integrate_u16(TSDF_FRAME_NOWAIT): uint16 depth / ht / lt converted on the GPU" if fmt16 else "tsdf_integrate_enqueue: float32 planes")),
               "ms_per_step": 1e3 * e2e_s / K, "last_frame_rgba_checksum": checksum, "host_cores": os.cpu_count(), "host_threads_all_ranks": 2 * world,
               "h2d_gbs_per_gpu": B * K * bpp_in * npx / e2e_s / 1e9, "d2h_gbs_per_gpu": B * K * 8 * npx / e2e_s / 1e9}
        for g in engs:
            g.close()
        for p_ in oblk + keep:
            p_.free()
        return res

    def run_e2e(pipelined):
        engs = make_engines(blocking=oversubscribed)
        houts = [[(tsdf_grid.PinnedArray((H, Wd, 4), np.uint8), tsdf_grid.PinnedArray((H, Wd, 4), np.uint8),
                   tsdf_grid.PinnedArray((H, Wd), np.float32)) for _ in range(2)] for _ in range(B)]

        def host_step(b, i, first):
            fi = i % n_frames
            g, p, st = engs[b], pinned[b], streams[b]
            pose = (st["q"][fi], st["t"][fi])
            out = tuple(a.array for a in houts[b][i & 1])
            if pipelined:
                g.Integrate(p["rgb"].array[fi], p["depth"].array[fi], p["ht"].array[fi], p["lt"].array[fi], cfg.max_depth, st["K"], pose,
                            asynchronous=True)
                g.RayCastAsync(cfg.max_depth, cam, pose, out)
                if not first:
                    g.RayCastWait()  # the previous frame's images are now in host memory
            else:
                g.Integrate(p["rgb"].array[fi], p["depth"].array[fi], p["ht"].array[fi], p["lt"].array[fi], cfg.max_depth, st["K"], pose)
                g.RayCast(cfg.max_depth, cam, pose, out=out)

        for i in range(W):
            for b in range(B):
                host_step(b, i, i == 0)
        for g in engs:
            g.synchronize()
        barrier()
        start = threading.Barrier(B + 1)

        def worker(b):
            torch.cuda.set_device(local_rank)
            start.wait()
            for i in range(W, W + K):
                host_step(b, i, i == W)
            engs[b].synchronize()  # the last frame's images too

        ths = [threading.Thread(target=worker, args=(b,)) for b in range(B)]
        for t in ths:
            t.start()
        start.wait()
        t_a = time.perf_counter()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        t_b = time.perf_counter()
        windows.append((t_a, t_b))
        e2e_s = t_b - t_a
        if world > 1:
            tt = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        checksum = float(sum(int(houts[b][(W + K - 1) & 1][0].array[::97, ::89].sum()) for b in range(B)))
        res = {"value": world * B * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 15 * npx * B, "d2h_bytes_per_step": 12 * npx * B + 64 * B,
               "api": ("tsdf_integrate_async + tsdf_raycast_async + tsdf_raycast_wait (pipelined, pinned host buffers; every frame's rgba + normal + "
                       "hit depth reach host memory inside the timed region, waited for one frame later)" if pipelined else
                       "tsdf_integrate + tsdf_raycast (synchronous, pinned host buffers; rgba + normal + hit depth)") + ", one Python host thread per stream"
                      + (", TSDF_FLAG_BLOCKING_SYNC (waiting threads yield: they would occupy half of the host cores or more)" if oversubscribed else ""),
               "ms_per_step": 1e3 * e2e_s / K, "last_frame_rgba_checksum": checksum, "host_cores": os.cpu_count(),
               "engine_threads_all_ranks": world * B}
        for g in engs:
            g.close()
        for hb in houts:
            for pair in hb:
                for a in pair:
                    a.free()
        return res

    e2e = e2e_sync = e2e_u16 = None
    if not args.no_e2e:
        e2e_sync = run_e2e(False)
        e2e_u16 = run_e2e_native(True)
        e2e = run_e2e_native(False)
        # what the e2e number runs against: the host <-> device copy bandwidth of this box, pinned memory, both directions
        # busy at once (256 MB each, best of 5) -- the e2e leg moves 15 B/px in and 8 B/px out per frame
        try:
            nb = 256 << 20
            hp_in, hp_out = torch.empty(nb, dtype=torch.uint8, pin_memory=True), torch.empty(nb, dtype=torch.uint8, pin_memory=True)
            d_in, d_out = torch.empty(nb, dtype=torch.uint8, device=dev), torch.empty(nb, dtype=torch.uint8, device=dev)
            s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            best = {"h2d": 0.0, "d2h": 0.0, "both": 0.0}
            for mode in ("h2d", "d2h", "both"):
                for rep in range(5):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    if mode in ("h2d", "both"):
                        with torch.cuda.stream(s1):
                            d_in.copy_(hp_in, non_blocking=True)
                    if mode in ("d2h", "both"):
                        with torch.cuda.stream(s2):
                            hp_out.copy_(d_out, non_blocking=True)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    best[mode] = max(best[mode], nb / dt / 1e9)
            e2e["pcie_measured_gbs"] = {"h2d_alone": best["h2d"], "d2h_alone": best["d2h"], "each_direction_when_both_run": best["both"]}
            e2e["h2d_gbs_in_e2e"] = e2e["value"] / world * 15 * npx / 1e9
            e2e["d2h_gbs_in_e2e"] = e2e["value"] / world * 8 * npx / 1e9
            del hp_in, hp_out, d_in, d_out
        except Exception as ex:  # supplementary only
            e2e["pcie_measured_gbs"] = {"error": repr(ex)}
    # ---------------- BASELINE configs 1, 3, 4, 5 (bench_legs.py) ----------------
    # config 3 at N > 1 is the sharded measurement: ONE stream whose volume is sharded over all ranks through the
    # C++ / NCCL data plane; at N = 1 the same stream runs on the plain engine
    extra = {}
    if not args.no_extra_configs:
        import bench_legs
        try:
            if not (world > 1 and args.no_sharded):
                extra["config3"], extra["config4"] = bench_legs.config3_and_4(args, cfg, streams[0], dres[0], rank, world, local_rank, dev, n_frames, W, K,
                                                                            dist if world > 1 else None)
        except Exception as ex:  # supplementary legs never lose the headline line
            import traceback
            extra["config3"] = {"error": repr(ex), "trace": traceback.format_exc()[-800:]}
        try:
            if stream1 is not None and not args.no_config15:
                extra["config1"], extra["config5"] = bench_legs.config1_and_5(args, stream1, rank, world, local_rank, dev, dist if world > 1 else None)
        except Exception as ex:
            import traceback
            extra["config1"] = {"error": repr(ex), "trace": traceback.format_exc()[-800:]}
    # exact voxel-update count of the headline leg: summed over the ranks, not extrapolated
    upd_all = upd_local
    if world > 1:
        tt = torch.tensor([upd_local], device=dev, dtype=torch.int64)
        dist.all_reduce(tt)
        upd_all = int(tt.item())
    sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- numbers ----------------
    frames = world * B * K
    value = frames / (ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    n_upd, n_vis, n_new, n_act = tot["n_updated"], tot["n_visible"], tot["n_new"], tot["n_active_pre"]
    integ_bytes = 24 * n_upd + 4 * (512 * n_vis - n_upd)  # 24 B per voxel update + 4 B TSDF read of every other visible voxel
    frame_bytes = 15 * npx * tot["frames"] + 8 * n_act + integ_bytes
    integ_ms = ph_ms.get("integrate", 0.0)
    launches = max(ph_n.get("integrate", 0), 1)
    achieved = integ_bytes / (integ_ms * 1e-3) / 1e9 if integ_ms > 0 else 0.0
    kern_total = sum(ph_ms.get(k, 0.0) for k in ("allocate", "select", "integrate", "raycast"))
    kernels = {k: {"ms_per_launch": ph_ms[k] / max(ph_n[k], 1), "share_of_step_kernel_time": ph_ms[k] / kern_total if kern_total else 0.0}
               for k in ("allocate", "select", "integrate", "raycast")}
    ws = (n_vis * 6144 + 31 * npx * tot["frames"]) / max(K, 1)  # voxel blocks + 15 B/px planes + 16 B/px staging
    traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the roofline kernel, from the committed ncu capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = next((v.get("dram_bytes_per_launch") for k, v in tj.items() if "integrate_carve_kernel" in k and isinstance(v, dict)), None)
    except Exception:
        tj = {}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, B, f"{B} independent streams interleaved per GPU, one frame of each per step", world),
        "working_set_mb_per_step": ws / 1e6,
        "voxel_updates_per_s": upd_all / (ms * 1e-3),
        "voxel_updates_per_frame": n_upd / max(tot["frames"], 1),
        "raycast_mrays_per_s": frames * npx / (ms * 1e-3) / 1e6,
        "integrate_frame_hbm_gbs": frame_bytes / ((ph_ms.get("allocate", 0) + ph_ms.get("select", 0) + integ_ms) * 1e-3) / 1e9 if integ_ms else None,
        "roofline": {"kernel": "integrate_carve_kernel", "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": tj.get("source"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": integ_bytes / launches, "us_per_launch": 1e3 * integ_ms / launches,
                     "launches_timed": launches},
        "kernels": kernels,
        "gather": gather,
        "raycast": {"us_per_view": 1e3 * ph_ms.get("raycast", 0.0) / max(ph_n.get("raycast", 0), 1), "rays_per_view": npx,
                    "mrays_per_s_kernel_only": npx * ph_n.get("raycast", 0) / (ph_ms.get("raycast", 1e-9) * 1e-3) / 1e6,
                    "note": "skip-map maintenance (4 launches, skipped when the block set did not change) + march; instruction-issue bound (ncu: 86 % issue-active, DRAM < 6 % of peak), not an HBM-roofline kernel"},
        "e2e": e2e,
        "e2e_u16": e2e_u16,
        "e2e_sync": e2e_sync,
        # per frame: frame_allocate, select_visible, integrate_carve + skip_fill, skip_mark, 3 x skip_pass, raycast
        "gpu_launches": 9 * B * K,
        "clocks": sampler.summary(windows[:1]),
        "counters_per_frame": {k: tot[k] / max(tot["frames"], 1) for k in ("n_new", "n_visible", "n_updated", "n_carved", "n_active_post")},
    }
    line["configs"] = extra
    if not args.no_cpu_baseline and world == 1:
        v, n, dt, ray_stats = cpu_sample(cfg, streams[0], args.cpu_seconds, n_frames)
        rc_us = line["raycast"]["us_per_view"]
        line["raycast"].update(ray_stats)
        line["raycast"]["algorithmic_gbs"] = ray_stats["algorithmic_bytes_per_view"] / (rc_us * 1e-6) / 1e9 if rc_us else None
        line["raycast"]["cache_hit_rates_ncu"] = tj.get("raycast_kernel_hit_rates")
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": oracle_threads(), "kind": "port",
                                "sample": f"first {n} frames of stream 0 (Integrate + RayCast each), {dt:.1f} s; oracle/tsdf_oracle.c -O2, "
                                          "OpenMP over visible blocks and image rows, allocation pass scalar"}
    else:
        line["cpu_baseline"] = None
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
