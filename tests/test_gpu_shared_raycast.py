"""GPU tests of the shared-volume RayCast (tsdf_raycast_shared): a volume sharded over several engines, each
rendering its rows of the view while reading the other shards' tables and pools.  On one GPU the shards live in
one process (tsdf_peer_attach_local); across GPUs the same kernel reads peer memory over NVLink through CUDA IPC
(tests/test_gpu_sharded.py).  Unlike min-compositing, the result must equal the single-volume render exactly."""
import numpy as np
import pytest
import torch

from disinfect_slam_b200 import synth
from oracle import compare
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,shift", [(3, 0), (4, 2), (8, 1)])
def test_shared_raycast_equals_single_volume(tsdf_lib, world, shift):
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config("small")
    sc = synth.Scene(cfg)
    shards = [tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 14, table_slots=1 << 16, shard_rank=r, shard_count=world,
                                 shard_shift=shift, max_image_pixels=cfg.width * cfg.height) for r in range(world)]
    o = Oracle(cfg.voxel_size, cfg.truncation)
    f = None
    for i in range(4):
        f = sc.frame(i)
        o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        for g in shards:
            g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    assert sum(g.NumActiveBlock() for g in shards) == o.num_blocks()
    for g in shards:
        g.peer_attach_local(shards)
    views = [(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])]
    v = sc.virtual_view(1, 5, width=250, height=141, K=(170.0, 170.0, 124.5, 70.0))
    views.append((10.0, v["width"], v["height"], v["K"], v["q"], v["t"]))
    for md, w, h, K, q, t in views:
        rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        normal = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        depth = torch.zeros((h, w), dtype=torch.float32, device="cuda")
        rows = (h + world - 1) // world
        cam = tsdf_grid.CameraParams(K, h, w)
        for r, g in enumerate(shards):  # shard r renders rows [r * rows, (r + 1) * rows) of the whole volume
            g.RayCastShared(md, cam, (q, t), r * rows, rows, rgba.data_ptr(), normal.data_ptr(), depth.data_ptr())
        for g in shards:
            g.synchronize()
        got = (rgba.cpu().numpy(), normal.cpu().numpy(), depth.cpu().numpy())
        rep = compare.compare_raycast(got, o.raycast(md, w, h, K, q, t)[:3], f"shared raycast world={world} shift={shift} {w}x{h}")
        assert rep["hits"] > 0.4 * rep["rays"]
    # a shard's own (local) RayCast still works afterwards: the union map is discarded, the local one rebuilt
    own = shards[0].RayCast(cfg.max_depth, tsdf_grid.CameraParams(f["K"], cfg.height, cfg.width), (f["q"], f["t"]))
    assert np.isfinite(own[2]).sum() > 0
    for g in shards:
        g.close()


@pytest.mark.parametrize("world,shift,pad", [(2, 1, 3), (4, 0, 3), (8, 2, 0), (3, 1, -4)])
def test_pulled_tsdf_cache_and_row_bands_equal_single_volume(tsdf_lib, world, shift, pad):
    """tsdf_shared_cache_attach + tsdf_raycast_shared_scatter with one contiguous band of 8-row tiles per shard: the
    foreign TSDF planes a band can meet are fetched before the march and sampled locally.  Whatever the frustum test
    lists -- pad 3: everything needed; pad -4: every block shrunk to its centre, so the blocks a band merely touches are
    dropped and their samples take the fallback to the owner -- the image equals the single-volume render; entries survive across views (peers_unchanged) and are dropped by a frame."""
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config("small")
    sc = synth.Scene(cfg)
    shards = [tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 14, table_slots=1 << 16, shard_rank=r, shard_count=world,
                                 shard_shift=shift, max_image_pixels=cfg.width * cfg.height) for r in range(world)]
    o = Oracle(cfg.voxel_size, cfg.truncation)

    def integrate(i):
        f = sc.frame(i)
        o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        for g in shards:
            g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        return f

    for i in range(3):
        f = integrate(i)
    for g in shards:
        g.peer_attach_local(shards)
        g.shared_cache_attach(1 << 14, pad)

    def render(md, w, h, K, q, t, unchanged, what):
        rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        normal = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        depth = torch.zeros((h, w), dtype=torch.float32, device="cuda")
        cam = tsdf_grid.CameraParams(K, h, w)
        tiles = (h + 7) // 8
        per = (tiles + world - 1) // world
        for r, g in enumerate(shards):
            g.RayCastSharedScatter(md, cam, (q, t), r * per, 1, per, unchanged, [(rgba.data_ptr(), normal.data_ptr(), depth.data_ptr())])
        for g in shards:
            g.synchronize()
        got = (rgba.cpu().numpy(), normal.cpu().numpy(), depth.cpu().numpy())
        rep = compare.compare_raycast(got, o.raycast(md, w, h, K, q, t)[:3], f"{what} world={world} shift={shift} pad={pad}")
        assert rep["hits"] > 0.4 * rep["rays"]
        return [g.shared_cache_fetched() for g in shards]

    total = o.num_blocks()
    fetched = render(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"], False, "first view")
    if pad >= 0:
        assert 0 < max(fetched) < total  # a band needs a part of the volume, not all of it
        again = render(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"], True, "same view, cache kept")
        assert sum(again) == 0  # nothing changed, everything needed is already there
    v = sc.virtual_view(1, 5, width=250, height=141, K=(170.0, 170.0, 124.5, 70.0))
    render(10.0, v["width"], v["height"], v["K"], v["q"], v["t"], True, "second view, cache kept")
    render(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"], True, "first view again, cache kept")
    f = integrate(3)  # the volume changes: what the caches hold is stale and must not be used
    render(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"], False, "view after another frame")
    for g in shards:
        g.shared_cache_attach(0)
    render(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"], False, "cache detached")
    for g in shards:
        g.close()
