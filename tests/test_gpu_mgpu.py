"""GPU tests of the C++ / NCCL multi-GPU data plane (libtsdf_b200_mgpu.so, include/tsdf_b200_mgpu.h).

  * world = 1 (runs on any GPU box): the library loads, creates its communicators, and the sharded code path --
    staging + stream-ordered Integrate, shared-volume RayCast, gather, all-reduced counters -- equals the oracle;
  * world = 2, two PROCESSES (needs 2 GPUs): shards mapped through CUDA IPC, frames broadcast by NCCL from rank 0's
    host memory, exact RayCast with peer loads over NVLink, gather-to-root -- against the oracle, bit for bit;
  * world = 2, two THREADS of one pure C++ process (tests/cpp/mgpu_threads_main.cc: no Python, no torch): the same
    through plain peer access, the whole stream driven by one tsdf_mgpu_run_sequence call per rank.
"""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from disinfect_slam_b200 import synth  # noqa: E402
from oracle import compare  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

pytestmark = pytest.mark.gpu
CFG, N_FRAMES, SHIFT = "small", 4, 2


def n_gpus():
    import torch
    return torch.cuda.device_count()


def oracle_run(cfg, n_frames):
    sc = synth.Scene(cfg)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    f, oc = None, None
    for i in range(n_frames):
        f = sc.frame(i)
        oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
    return sc, o, f, oc


def run_rank(rank, world, nccl_id, cfg_name, n_frames):
    """One rank's program; returns what the parent compares."""
    from disinfect_slam_b200 import mgpu, tsdf_grid
    cfg = synth.config(cfg_name)
    sc = synth.Scene(cfg)
    v = mgpu.ShardedVolume(cfg.voxel_size, cfg.truncation, rank, world, nccl_id, device=rank, pool_blocks=cfg.pool_blocks,
                           table_slots=cfg.table_slots, max_image_pixels=cfg.width * cfg.height, shard_shift=SHIFT)
    v.set_profiling(True)
    f = None
    for i in range(n_frames):
        f = sc.frame(i)
        if rank == 0:
            v.Integrate(0, f["rgb"], f["depth"], f["ht"], f["lt"], cfg.width, cfg.height, cfg.max_depth, f["K"], (f["q"], f["t"]))
        else:
            v.Integrate(0, None, None, None, None, cfg.width, cfg.height, cfg.max_depth, f["K"], (f["q"], f["t"]))
        if i == 1:  # a view in the middle of the stream: the barriers must order it against the neighbouring frames
            mid = v.RayCast(cfg.max_depth, cfg.width, cfg.height, f["K"], (f["q"], f["t"]))
    exact = v.RayCast(cfg.max_depth, cfg.width, cfg.height, f["K"], (f["q"], f["t"]))
    exact10 = v.RayCast(10.0, cfg.width, cfg.height, f["K"], (f["q"], f["t"]))
    last, totals, n_active = v.counters()
    gathered = v.Gather(0)
    bbox = (-0.5, 1.2, -1.4, 0.6, -2.2, 0.9)
    bounded = v.Gather(0, bbox)
    ms, cnt = v.comm_ms()
    # the shard itself, through the engine handle of this rank
    g = tsdf_grid.TSDFGrid.__new__(tsdf_grid.TSDFGrid)
    g.L, g.h = tsdf_grid._lib.lib(), v.engine
    keys, tsdf, rgbw, prob = tsdf_grid.TSDFGrid.export(g)
    g.h = None
    out = dict(keys=keys, tsdf=tsdf, rgbw=rgbw, prob=prob, mid=mid, exact=exact, exact10=exact10, last=last, totals=totals,
               n_active=n_active, gathered=gathered, bounded=bounded, comm_calls=cnt)
    v.synchronize()
    v.close()
    return out


def check_replicas_against_oracle(res, world, cfg, n_frames):
    """TSDF_MGPU_MODE=replicas: every rank holds the whole volume; view k is rendered by rank k % world into rank 0's memory."""
    sc = synth.Scene(cfg)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    for i in range(n_frames):
        f = sc.frame(i)
        oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        if i == 1:
            ref_mid = o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3]
    ok, ot, oc_, op = o.export()
    for r in range(world):
        assert np.array_equal(res[r]["keys"], ok), f"replica {r}: block set"
        assert np.array_equal(res[r]["tsdf"].view(np.uint32), ot.view(np.uint32)) and np.array_equal(res[r]["rgbw"], oc_)
        assert np.abs(res[r]["prob"] - op).max() <= compare.PROB_TOL
        last = res[r]["last"]
        assert (last["n_new"], last["n_visible"], last["n_updated"], last["n_carved"]) == (oc["n_new"], oc["n_vis"], oc["n_upd"], oc["n_carved"])
        assert res[r]["n_active"] == len(ok)
    # the images live on rank 0 whoever rendered them: views 0 (rank 0), 1 (rank 1), 2 (rank 0)
    compare.compare_raycast(res[0]["mid"], ref_mid, "mid-stream view (rendered by rank 0)")
    compare.compare_raycast(res[0]["exact"], o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3], "view rendered by rank 1")
    compare.compare_raycast(res[0]["exact10"], o.raycast(10.0, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3], "max_depth 10 view")
    assert compare.compare_gather(res[0]["gathered"], o.gather(), "replica GatherValid")["tsdf_bit_exact"]
    assert [res[r]["comm_calls"]["raycast_shared"] for r in range(world)] == [2, 1]
    assert res[0]["comm_calls"]["broadcast"] == n_frames and res[0]["comm_calls"]["exchange_barrier"] == 0


def check_against_oracle(res, world, cfg, n_frames, exchange_barriers=None):
    from disinfect_slam_b200 import tsdf_grid
    sc = synth.Scene(cfg)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    for i in range(n_frames):
        f = sc.frame(i)
        oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        if i == 1:
            ref_mid = o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3]
    ok, ot, oc_, op = o.export()
    keys = np.concatenate([res[r]["keys"] for r in range(world)])
    order = compare.key_order(keys)
    assert np.array_equal(keys[order], ok), "union of the shards != single-volume block set"
    assert np.array_equal(np.concatenate([res[r]["tsdf"] for r in range(world)])[order].view(np.uint32), ot.view(np.uint32))
    assert np.array_equal(np.concatenate([res[r]["rgbw"] for r in range(world)])[order], oc_)
    assert np.abs(np.concatenate([res[r]["prob"] for r in range(world)])[order] - op).max() <= compare.PROB_TOL
    ref = o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3]
    ref10 = o.raycast(10.0, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3]
    for r in range(world):
        assert (tsdf_grid.block_owner(res[r]["keys"], world, SHIFT) == r).all()
        if world > 1:
            assert len(res[r]["keys"]) > 0.25 * len(ok)
        compare.compare_raycast(res[r]["mid"], ref_mid, f"mid-stream view, rank {r}")
        compare.compare_raycast(res[r]["exact"], ref, f"exact view, rank {r}")
        compare.compare_raycast(res[r]["exact10"], ref10, f"exact view max_depth 10, rank {r}")
        for a, b in zip(res[r]["exact"], res[0]["exact"]):  # every rank holds the same assembled images
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
        # all-reduced counters = the single-volume counters of the oracle
        last = res[r]["last"]
        assert (last["n_new"], last["n_visible"], last["n_updated"], last["n_carved"]) == (oc["n_new"], oc["n_vis"], oc["n_upd"], oc["n_carved"])
        assert res[r]["n_active"] == len(ok) and res[r]["totals"]["n_updated"] >= last["n_updated"]
    assert compare.compare_gather(res[0]["gathered"], o.gather(), "sharded GatherValid")["tsdf_bit_exact"]
    bbox = (-0.5, 1.2, -1.4, 0.6, -2.2, 0.9)
    rep = compare.compare_gather(res[0]["bounded"], o.gather(bbox), "sharded GatherVoxels")
    assert rep["tsdf_bit_exact"] and rep["n_voxels"] > 0
    if world > 1:
        c = res[0]["comm_calls"]
        assert c["broadcast"] == n_frames and c["barrier"] == 3 and c["allgather"] == 3 and c["raycast_shared"] == 3
        if exchange_barriers is not None:  # candidate exchange: one barrier per frame, between allocate and the owners' inserts
            assert c["exchange_barrier"] == exchange_barriers


def test_data_plane_with_one_shard_matches_oracle(tsdf_lib):
    from disinfect_slam_b200 import mgpu
    cfg = synth.config(CFG)
    res = {0: run_rank(0, 1, mgpu.unique_id(), CFG, N_FRAMES)}
    check_against_oracle(res, 1, cfg, N_FRAMES)


VARIANTS = {
    # the default: peer-barrier kernels, owner-filtered allocation on every rank, TSDF mirrors pushed by the integrate
    # kernels, 8-row tiles dealt round-robin, rows stored into every rank's images by the march kernel
    "fused": {},
    # allocation pass sharded by image tiles with the candidate keys mailed to their owners; foreign TSDF planes fetched
    # into a local cache before the march instead of mirrors
    "fused-pull-cache-exchange": {"TSDF_MGPU_MIRROR": "pull", "TSDF_MGPU_ALLOC": "exchange"},
    # the same with one contiguous band of rows per rank
    "fused-pull-cache-bands": {"TSDF_MGPU_MIRROR": "pull", "TSDF_MGPU_ALLOC": "exchange", "TSDF_MGPU_TILES": "band"},
    # no local TSDF copy at all: every foreign sample is a load over NVLink
    "fused-remote-loads": {"TSDF_MGPU_MIRROR": "0"},
    # ncclAllReduce barrier + in-place ncclAllGather (kept for comparison)
    "nccl": {"TSDF_MGPU_EXCHANGE": "nccl"},
    # not sharded at all: every GPU integrates every frame into its own copy of the whole volume, whole views are dealt
    # round-robin (throughput for volumes that fit one GPU)
    "replicas": {"TSDF_MGPU_MODE": "replicas"},
}


def _proc(rank, world, nccl_id, q, variant):
    try:
        os.environ.update(VARIANTS[variant])  # read by tsdf_mgpu_create
        q.put((rank, run_rank(rank, world, nccl_id, CFG, N_FRAMES)))
    except Exception:
        import traceback
        q.put((rank, {"error": traceback.format_exc()}))


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_two_processes_nccl_and_cuda_ipc_match_oracle(tsdf_lib, variant):
    """Every variant of the data plane (see VARIANTS) gives the single-volume result, bit for bit."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2); the bench's sharded leg carries the driver-visible parity check")
    import multiprocessing as mp
    from disinfect_slam_b200 import mgpu
    world = 2
    nccl_id = mgpu.unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_proc, args=(r, world, nccl_id, q, variant)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, d = q.get(timeout=240)
        if "error" in d:
            for p in procs:
                p.kill()
            pytest.fail(f"rank {r} failed:\n{d['error']}")
        res[r] = d
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if variant == "replicas":
        check_replicas_against_oracle(res, world, synth.config(CFG), N_FRAMES)
    else:
        check_against_oracle(res, world, synth.config(CFG), N_FRAMES,
                             exchange_barriers=N_FRAMES if "exchange" in variant or "bands" in variant else 0)


def test_two_threads_of_a_pure_cpp_process_match_oracle(tsdf_lib, tmp_path):
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "mgpu_threads")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} missing: run __graft_entry__.build()")
    world = 2 if n_gpus() >= 2 else 1  # one thread per GPU; with a single GPU the same program runs as one shard
    cfg = synth.config(CFG)
    sc, o, f, oc = oracle_run(cfg, N_FRAMES)
    frames = tmp_path / "frames.bin"
    with open(frames, "wb") as fh:
        fh.write(struct.pack("<4i3f4f6f", N_FRAMES, cfg.width, cfg.height, world, cfg.voxel_size, cfg.truncation, cfg.max_depth,
                             *[float(np.float32(k)) for k in cfg.K], *([0.0] * 6)))
        for i in range(N_FRAMES):
            fr = sc.frame(i)
            fh.write(np.concatenate([fr["q"], fr["t"]]).astype(np.float32).tobytes())
            for k in ("rgb", "depth", "ht", "lt"):
                fh.write(fr[k].tobytes())
    out = tmp_path / "out.bin"
    res = subprocess.run([exe, str(frames), str(out)], capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout + res.stderr
    raw = open(out, "rb").read()
    n = struct.unpack("<q", raw[:8])[0]
    p = 8
    got = np.frombuffer(raw[p:p + 16 * n], np.float32).reshape(n, 4); p += 16 * n
    npx = cfg.width * cfg.height
    rgba = np.frombuffer(raw[p:p + 4 * npx], np.uint8).reshape(cfg.height, cfg.width, 4); p += 4 * npx
    normal = np.frombuffer(raw[p:p + 4 * npx], np.uint8).reshape(cfg.height, cfg.width, 4); p += 4 * npx
    depth = np.frombuffer(raw[p:p + 4 * npx], np.float32).reshape(cfg.height, cfg.width); p += 4 * npx
    last = struct.unpack("<8q", raw[p:p + 64]); p += 64
    n_active = struct.unpack("<q", raw[p:p + 8])[0]
    assert compare.compare_gather(got, o.gather(), "C++ threads GatherValid")["tsdf_bit_exact"]
    compare.compare_raycast((rgba, normal, depth), o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3], "C++ threads view")
    assert (last[1], last[2], last[3], last[4]) == (oc["n_new"], oc["n_vis"], oc["n_upd"], oc["n_carved"]) and n_active == o.num_blocks()
