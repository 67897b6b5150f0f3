"""GPU parity tests (run on a B200 with `pytest -m gpu`): the CUDA engine, called through the C ABI,
against the CPU oracle on the same seeded inputs, plus the reference's own unit tests
(utils/tests/voxel_hash_test.cu, voxel_mem_test.cu) replayed through the ABI's test hooks."""
import json
import os

import numpy as np
import pytest

from disinfect_slam_b200 import synth
from oracle import compare
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hash_kat.json")))


@pytest.fixture(scope="module")
def grid_mod(tsdf_lib):
    from disinfect_slam_b200 import tsdf_grid
    return tsdf_grid


def make_grid(tg, cfg, **kw):
    return tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots, **kw)


# ---- reference unit tests replayed through the C ABI -------------------------------------------------
def test_hash_single(grid_mod):
    # voxel_hash_test.cu:56-92
    g = grid_mod.TSDFGrid(0.01, 0.06, pool_blocks=1 << 10)
    g.allocate_blocks([[1, 1, 1]])
    assert g.NumActiveBlock() == 1
    tsdf, rgbw, prob, found = g.retrieve([[8, 8, 8]])
    assert found[0] and tsdf[0] == -1.0 and rgbw[0, 3] == 0 and abs(prob[0] - 0.5) < 1e-7
    tsdf, rgbw, prob, found = g.retrieve([[0, 0, 0]])  # unallocated -> default voxel
    assert not found[0] and rgbw[0, 3] == 0 and tsdf[0] == 1.0 and prob[0] == 0.0
    g.allocate_blocks([[0, 0, 0]])
    pts = [[0, 0, i] for i in range(8)]
    g.assign(pts, rgbw=[[i, i, i, i] for i in range(8)])
    assert g.NumActiveBlock() == 2
    _, rgbw, _, found = g.retrieve(pts)
    assert found.all() and np.array_equal(rgbw, np.repeat(np.arange(8, dtype=np.uint8)[:, None], 4, 1))
    g.close()


def test_hash_multiple(grid_mod):
    # voxel_hash_test.cu:94-126
    n = GOLD["multiple_blocks"]
    g = grid_mod.TSDFGrid(0.01, 0.06, pool_blocks=1 << 10)
    g.allocate_blocks([[i, i, i] for i in range(n)])
    assert g.NumActiveBlock() == n
    pts = [[i * 8, i * 8, i * 8] for i in range(n)]
    vals = np.repeat(np.arange(n, dtype=np.uint8)[:, None], 4, 1)
    g.assign(pts, rgbw=vals)
    _, rgbw, _, found = g.retrieve(pts)
    assert found.all() and np.array_equal(rgbw, vals)
    keys, _, _, _ = g.export(voxels=False)
    assert np.array_equal(keys, np.repeat(np.arange(n, dtype=np.int16)[:, None], 3, 1))
    g.close()


def test_hash_collision(grid_mod, tsdf_lib):
    # voxel_hash_test.cu:128-180.  The reference inserts one block per bucket per pass (2, 3, 4 active
    # blocks, pinned by the oracle's RefHashModel); the lock-free table allocates every request in the
    # first pass -- the documented ideal-set semantics -- and further passes are no-ops.
    blocks = GOLD["collision_blocks"]
    for b in blocks[:3]:
        assert tsdf_lib.tsdf_hash(*b) == GOLD["num_bucket"] - 1
    g = grid_mod.TSDFGrid(0.01, 0.06, pool_blocks=1 << 10)
    counts = []
    for _ in range(3):
        g.allocate_blocks(blocks)
        counts.append(g.NumActiveBlock())
    assert counts == [4, 4, 4]
    pts = [[b[0] * 8, b[1] * 8, b[2] * 8] for b in blocks]
    vals = np.repeat(np.arange(4, dtype=np.uint8)[:, None], 4, 1)
    g.assign(pts, rgbw=vals)
    _, rgbw, _, found = g.retrieve(pts)
    assert found.all() and np.array_equal(rgbw, vals)
    # delete one of the colliding blocks: the others stay reachable across the tombstone
    g.delete_blocks([blocks[0]])
    assert g.NumActiveBlock() == 3
    _, rgbw, _, found = g.retrieve(pts)
    assert list(found) == [False, True, True, True] and np.array_equal(rgbw[1:], vals[1:])
    g.allocate_blocks([blocks[0]])  # re-insert reuses the tombstone
    assert g.NumActiveBlock() == 4
    g.close()


def test_pool_reacquire_resets(grid_mod):
    # voxel_mem_test.cu:38-90: release does not clobber, re-acquire resets weight
    g = grid_mod.TSDFGrid(0.01, 0.06, pool_blocks=64, table_slots=1 << 10)
    blocks = [[i, 0, -i] for i in range(8)]
    g.allocate_blocks(blocks)
    pts = [[b[0] * 8 + 3, 1, b[2] * 8 + 2] for b in blocks]
    g.assign(pts, rgbw=[[9, 9, 9, i + 1] for i in range(8)], tsdf=[0.25] * 8)
    g.delete_blocks(blocks)
    assert g.NumActiveBlock() == 0
    g.allocate_blocks(blocks)
    tsdf, rgbw, prob, found = g.retrieve(pts)
    assert found.all() and (rgbw[:, 3] == 0).all() and (tsdf == -1.0).all()
    g.close()


def test_pool_exhaustion_is_an_error_not_ub(grid_mod):
    # the reference device-asserts (voxel_mem.cu:39); the ABI returns TSDF_E_POOL_EXHAUSTED
    g = grid_mod.TSDFGrid(0.01, 0.06, pool_blocks=16, table_slots=1 << 8)
    with pytest.raises(grid_mod.TsdfError) as ei:
        g.allocate_blocks([[i, 1, 2] for i in range(40)])
    assert ei.value.code == -3
    g.close()


def test_invalid_images_rejected(grid_mod):
    g = grid_mod.TSDFGrid(0.01, 0.06, pool_blocks=64, table_slots=1 << 10, max_image_pixels=64 * 48)
    rgb = np.zeros((48, 64, 3), np.uint8)
    d = np.zeros((48, 64), np.float32)
    pose = ([0, 0, 0, 1], [0, 0, 0])
    g.Integrate(rgb, d, d, d, 4.0, (50, 50, 32, 24), pose)  # all-invalid depth: nothing allocated
    assert g.NumActiveBlock() == 0
    with pytest.raises(ValueError):
        g.Integrate(rgb[:, :32], d, d, d, 4.0, (50, 50, 32, 24), pose)
    big = np.zeros((96, 128), np.float32)
    with pytest.raises(grid_mod.TsdfError):
        g.Integrate(np.zeros((96, 128, 3), np.uint8), big, big, big, 4.0, (50, 50, 32, 24), pose)
    assert g.GatherValid().shape == (0, 4)  # empty volume
    g.close()


# ---- Integrate / Gather / RayCast parity against the oracle ------------------------------------------
def run_pair(tg, cfg, n_frames, asynchronous=False, check_every=1):
    sc = synth.Scene(cfg)
    g = make_grid(tg, cfg, max_image_pixels=cfg.width * cfg.height)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    reports = []
    last = None
    for i in range(n_frames):
        f = sc.frame(i)
        last = f
        oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]),
                    asynchronous=asynchronous)
        if asynchronous and i + 1 < n_frames:
            continue
        ec = g.counters()
        assert ec["n_new"] == oc["n_new"], (i, ec, oc)
        assert ec["n_visible"] == oc["n_vis"], (i, ec, oc)
        assert ec["n_updated"] == oc["n_upd"], (i, ec, oc)
        assert ec["n_carved"] == oc["n_carved"], (i, ec, oc)
        assert ec["n_active_post"] == oc["n_active_post"] == g.NumActiveBlock(), (i, ec, oc)
        if (i % check_every) == 0 or i + 1 == n_frames:
            reports.append(compare.compare_volumes(g.export(), o.export(), label=f"{cfg.name} frame {i}"))
    return sc, g, o, reports, last


def test_integrate_parity_tiny_every_frame(grid_mod):
    cfg = synth.config("tiny")
    sc, g, o, reports, _ = run_pair(grid_mod, cfg, 8)
    assert all(r["tsdf_bit_exact"] for r in reports), reports
    assert max(r["prob_max_abs"] for r in reports) <= compare.PROB_TOL
    g.close()


def test_integrate_gather_raycast_parity_small(grid_mod):
    cfg = synth.config("small")
    sc, g, o, reports, f = run_pair(grid_mod, cfg, 6, check_every=5)
    assert all(r["tsdf_bit_exact"] for r in reports), reports
    # GatherValid / GatherVoxels (voxel_tsdf.cu:399-454)
    r = compare.compare_gather(g.GatherValid(), o.gather(), "GatherValid")
    assert r["tsdf_bit_exact"] and r["n_voxels"] == o.num_blocks() * 512
    for bbox in [(-0.5, 0.5, -1.0, 0.5, -2.1, 2.1), (-8, 8, -8, 8, -8, 8), (0.003, 1.0, -1.49, 0.3, -0.33, 2.5),
                 (5, 6, 5, 6, 5, 6)]:
        compare.compare_gather(g.GatherVoxels(grid_mod.BoundingCube(*bbox)), o.gather(bbox), f"GatherVoxels{bbox}")
    # RayCast (voxel_tsdf.cu:232-307) from the last camera and from a virtual view, max_depth 4 and 10
    for view, md in [((f["q"], f["t"]), cfg.max_depth), ((f["q"], f["t"]), 10.0)]:
        cam = grid_mod.CameraParams(f["K"], cfg.height, cfg.width)
        er = g.RayCast(md, cam, view)
        orr = o.raycast(md, cfg.width, cfg.height, f["K"], view[0], view[1])
        rep = compare.compare_raycast(er, orr[:3], f"RayCast md={md}")
        assert rep["hits"] > 0.5 * rep["rays"]
    v = sc.virtual_view(1, 4, width=200, height=120, K=(150.0, 150.0, 99.5, 59.5))
    cam = grid_mod.CameraParams(v["K"], v["height"], v["width"])
    compare.compare_raycast(g.RayCast(4.0, cam, (v["q"], v["t"])),
                            o.raycast(4.0, v["width"], v["height"], v["K"], v["q"], v["t"])[:3], "RayCast virtual")
    g.close()


def test_async_pipeline_matches_sync(grid_mod):
    cfg = synth.config("tiny")
    _, g, o, reports, _ = run_pair(grid_mod, cfg, 6, asynchronous=True)
    assert reports and reports[-1]["tsdf_bit_exact"]
    g.close()


def test_device_input_path_and_sharding(grid_mod):
    """tsdf_integrate_device (planes already on the GPU) and block-ownership sharding: the union of the
    shards' block sets equals the single-engine set and every shard's voxels equal the unsharded ones."""
    import torch
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    full = make_grid(grid_mod, cfg)
    shards = [make_grid(grid_mod, cfg, shard_rank=r, shard_count=3) for r in range(3)]
    for i in range(4):
        f = sc.frame(i)
        full.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        d = {k: torch.from_numpy(f[k]).cuda() for k in ("rgb", "depth", "ht", "lt")}
        torch.cuda.synchronize()
        for s in shards:
            s.IntegrateDevice(d["rgb"].data_ptr(), d["depth"].data_ptr(), d["ht"].data_ptr(), d["lt"].data_ptr(),
                              cfg.width, cfg.height, cfg.max_depth, f["K"], (f["q"], f["t"]))
        for s in shards:
            s.synchronize()
    fk, ft, fc, fp = full.export()
    parts = [s.export() for s in shards]
    assert sum(len(p[0]) for p in parts) == len(fk) and all(len(p[0]) > 0 for p in parts)
    keys = np.concatenate([p[0] for p in parts])
    order = compare.key_order(keys)
    assert np.array_equal(keys[order], fk)
    assert np.array_equal(np.concatenate([p[1] for p in parts])[order].view(np.uint32), ft.view(np.uint32))
    assert np.array_equal(np.concatenate([p[2] for p in parts])[order], fc)
    for g in [full] + shards:
        g.close()


def test_pipelined_raycast_matches_the_synchronous_call(tsdf_lib):
    """tsdf_integrate_async + tsdf_raycast_async + tsdf_raycast_wait (two views in flight, copies on their own stream)
    deliver exactly the images of the blocking tsdf_integrate + tsdf_raycast sequence."""
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    cam = tsdf_grid.CameraParams(sc.frame(0)["K"], cfg.height, cfg.width)
    mk = lambda: tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)  # noqa: E731
    a, b = mk(), mk()
    n = 6
    bufs = [tuple(tsdf_grid.PinnedArray(s, d) for s, d in (((cfg.height, cfg.width, 4), np.uint8), ((cfg.height, cfg.width, 4), np.uint8),
                                                            ((cfg.height, cfg.width), np.float32))) for _ in range(2)]
    want, got = [], []
    for i in range(n):
        f = sc.frame(i)
        a.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        want.append([x.copy() for x in a.RayCast(cfg.max_depth, cam, (f["q"], f["t"]))])
    for i in range(n):
        f = sc.frame(i)
        b.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]), asynchronous=True)
        b.RayCastAsync(cfg.max_depth, cam, (f["q"], f["t"]), tuple(x.array for x in bufs[i & 1]))
        if i > 0:
            b.RayCastWait()  # frame i - 1 is complete, frame i may still be in flight
            got.append([x.array.copy() for x in bufs[(i - 1) & 1]])
    b.synchronize()
    got.append([x.array.copy() for x in bufs[(n - 1) & 1]])
    for i in range(n):
        for w, g in zip(want[i], got[i]):
            assert np.array_equal(w.view(np.uint8), g.view(np.uint8)), i
    # a blocking call after pipelined ones sees a quiescent engine
    f = sc.frame(n)
    assert np.array_equal(b.RayCast(cfg.max_depth, cam, (f["q"], f["t"]))[0], a.RayCast(cfg.max_depth, cam, (f["q"], f["t"]))[0])
    a.close(); b.close()
