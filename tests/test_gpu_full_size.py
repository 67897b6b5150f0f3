"""GPU parity at BASELINE.json's full sizes (the tiny / small scenes of the other suites keep the CPU oracle to a
few seconds; here a few frames of each named config run through both):
  config 1  640x480 TUM intrinsics, 1 cm voxels: Integrate + GatherValid
  config 2  1280x720 L515/ZED-style, 5 mm voxels: Integrate + RayCast from the same camera
  config 3  room-scale hall, 2 cm voxels, 1280x720: Integrate + bounded GatherVoxels + mesh of the same box
  config 4  1920x1080 virtual views over a pre-built volume
  config 5  (per-stream part) GatherVoxels with the +-8 m query box of the reference's ROS node
plus size-independent properties on a longer run: counters add up, re-integrating is deterministic, a gather of
the whole volume equals the block export."""
import numpy as np
import pytest

from disinfect_slam_b200 import synth
from oracle import compare
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tg(tsdf_lib):
    from disinfect_slam_b200 import tsdf_grid
    return tsdf_grid


def run(tg, cfg, frames, pool=1 << 17, table=1 << 19):
    sc = synth.Scene(cfg)
    g = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=pool, table_slots=table, max_image_pixels=1920 * 1080)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    f = None
    for i in frames:
        f = sc.frame(i)
        oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        ec = g.counters()
        assert (ec["n_new"], ec["n_visible"], ec["n_updated"], ec["n_carved"], ec["n_active_post"]) == \
               (oc["n_new"], oc["n_vis"], oc["n_upd"], oc["n_carved"], oc["n_active_post"]), (i, ec, oc)
    return sc, g, o, f


def test_config1_integrate_and_gather_valid(tg):
    cfg = synth.config("config1")
    sc, g, o, f = run(tg, cfg, range(0, 20, 4))
    rep = compare.compare_volumes(g.export(), o.export(), "config1")
    assert rep["tsdf_bit_exact"] and rep["n_blocks_engine"] > 3000
    assert compare.compare_gather(g.GatherValid(), o.gather(), "config1 GatherValid")["tsdf_bit_exact"]
    bbox = (-8.0, 8.0, -8.0, 8.0, -8.0, 8.0)  # configs/config.yaml:4 of the reference
    assert compare.compare_gather(g.GatherVoxels(tg.BoundingCube(*bbox)), o.gather(bbox), "config1 GatherVoxels")["n_voxels"] > 0
    g.close()


def test_config2_integrate_and_raycast(tg):
    cfg = synth.config("config2")
    sc, g, o, f = run(tg, cfg, (0, 1))
    assert compare.compare_volumes(g.export(), o.export(), "config2")["tsdf_bit_exact"]
    cam = tg.CameraParams(f["K"], cfg.height, cfg.width)
    rep = compare.compare_raycast(g.RayCast(cfg.max_depth, cam, (f["q"], f["t"])),
                                  o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3], "config2 RayCast")
    assert rep["hits"] > 0.9 * rep["rays"]
    g.close()


def test_config3_room_scale_integrate_bounded_gather_and_mesh(tg):
    from oracle import mesh_oracle
    cfg = synth.config("config3")
    sc, g, o, f = run(tg, cfg, (0, 40, 80))  # three views far apart on the 5 m trajectory of the 16 x 6 x 16 m hall
    assert compare.compare_volumes(g.export(), o.export(), "config3")["tsdf_bit_exact"]
    bbox = (-8.0, 0.0, -3.0, 3.0, -8.0, 8.0)  # half of the hall
    rep = compare.compare_gather(g.GatherVoxels(tg.BoundingCube(*bbox)), o.gather(bbox), "config3 GatherVoxels")
    assert rep["tsdf_bit_exact"] and 0 < rep["n_voxels"] < g.NumActiveBlock() * 512
    keys, tsdf, rgbw, _ = o.export()
    got, want = g.ExtractMesh(bbox), mesh_oracle.extract_mesh(keys, tsdf, rgbw, cfg.voxel_size, bbox)
    assert got.shape == want.shape and len(got) > 10000
    assert np.array_equal(mesh_oracle.canonical_triangles(got).view(np.uint32), mesh_oracle.canonical_triangles(want).view(np.uint32))
    g.close()


def test_integrate_at_the_maximum_image_size(tg):
    """1920 x 1080 is the largest frame the reference accepts (MAX_IMG_SIZE, voxel_tsdf.cu:10-12): one such frame through
    Integrate + RayCast equals the oracle; one pixel more is refused with TSDF_E_INVALID instead of overrunning."""
    cfg = synth.config("config2").scaled(1.5, "config2_1920x1080")
    assert (cfg.width, cfg.height) == (1920, 1080)
    sc, g, o, f = run(tg, cfg, (0,))
    assert compare.compare_volumes(g.export(), o.export(), "1080p")["tsdf_bit_exact"]
    cam = tg.CameraParams(f["K"], cfg.height, cfg.width)
    rep = compare.compare_raycast(g.RayCast(cfg.max_depth, cam, (f["q"], f["t"])),
                                  o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3], "1080p RayCast")
    assert rep["rays"] == 1920 * 1080 and rep["hits"] > 0.9 * rep["rays"]
    big = np.zeros((1081, 1920), np.float32)
    with pytest.raises(Exception, match="max_image_pixels"):
        g.Integrate(np.zeros((1081, 1920, 3), np.uint8), big, big, big, cfg.max_depth, f["K"], (f["q"], f["t"]))
    g.close()


def test_config4_full_hd_virtual_views(tg):
    cfg1, cfg4 = synth.config("config1"), synth.config("config4")
    sc, g, o, f = run(tg, cfg1, (0, 5, 10))
    for j, md in ((0, 4.0), (3, 10.0)):
        v = sc.virtual_view(j, 8, width=cfg4.width, height=cfg4.height, K=cfg4.K)
        cam = tg.CameraParams(v["K"], v["height"], v["width"])
        rep = compare.compare_raycast(g.RayCast(md, cam, (v["q"], v["t"])),
                                      o.raycast(md, v["width"], v["height"], v["K"], v["q"], v["t"])[:3], f"config4 view {j}")
        assert rep["rays"] == 1920 * 1080
    g.close()


def test_size_independent_properties_on_a_longer_run(tg):
    cfg = synth.config("config1")
    sc = synth.Scene(cfg)
    grids = [tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 17, table_slots=1 << 19, max_image_pixels=cfg.width * cfg.height)
             for _ in range(2)]
    active = 0
    for i in range(0, 60, 2):
        f = sc.frame(i)
        for g in grids:
            g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]), asynchronous=(g is grids[1]))
        c = grids[0].counters()
        assert c["n_active_pre"] == active and c["n_active_post"] == active + c["n_new"] - c["n_carved"]
        assert c["n_updated"] <= 512 * c["n_visible"] and c["n_visible"] <= c["n_active_pre"] + c["n_new"]
        active = c["n_active_post"]
    a, b = grids[0].export(), grids[1].export()  # synchronous and pipelined runs: bit-identical volumes
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    keys, tsdf, _, _ = a
    assert len(keys) == active == grids[0].NumActiveBlock()
    gathered = compare.canonical_gather(grids[0].GatherValid()).reshape(len(keys), 512, 4)
    assert np.array_equal(gathered[..., 3].view(np.uint32), tsdf.view(np.uint32))  # the gather IS the volume's TSDF plane
    origin = keys.astype(np.float32) * np.float32(8) * np.float32(cfg.voxel_size)
    assert np.array_equal(gathered[:, 0, :3], (keys.astype(np.int32) * 8).astype(np.float32) * np.float32(cfg.voxel_size)) and origin.shape == (len(keys), 3)
    for g in grids:
        g.close()
