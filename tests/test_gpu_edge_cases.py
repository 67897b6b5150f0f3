"""GPU edge cases of the engine against the oracle: hand-built sparse volumes that force a coarse RayCast skip
map and views from outside the block AABB, tombstone garbage collection (rehash) in the middle of a run,
pool exhaustion inside Integrate, ragged image sizes, repeated / empty frames."""
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


def slab_volume(tg, blocks, voxel_size=0.02, truncation=0.12):
    """Blocks filled with a planar SDF ramp along z (surface in the middle of each block), colours and probabilities."""
    g = tg.TSDFGrid(voxel_size, truncation, pool_blocks=1 << 10, table_slots=1 << 12, max_image_pixels=320 * 240)
    o = Oracle(voxel_size, truncation)
    g.allocate_blocks(blocks)
    rng = np.random.RandomState(3)
    for b in blocks:
        o.allocate_block(*b)
        pts = np.array([[b[0] * 8 + x, b[1] * 8 + y, b[2] * 8 + z] for z in range(8) for y in range(8) for x in range(8)], np.int32)
        tsdf = np.clip((3.5 - (pts[:, 2] - b[2] * 8)) / 6.0, -1, 1).astype(np.float32)  # > 0 in front (low z), < 0 behind
        rgbw = np.concatenate([rng.randint(0, 255, (512, 3)), np.full((512, 1), 7)], 1).astype(np.uint8)
        prob = rng.uniform(0.05, 0.95, 512).astype(np.float32)
        g.assign(pts, tsdf=tsdf, rgbw=rgbw, prob=prob)
        o.set_voxels(pts, tsdf=tsdf, rgbw=rgbw, prob=prob)
    return g, o


@pytest.mark.parametrize("blocks", [
    [[0, 0, 10], [1, 0, 10], [0, 1, 10], [1, 1, 10], [-1, -1, 10], [0, 0, 30]],            # compact: skip-map shift 0
    [[0, 0, 10], [1, 0, 10], [0, 1, 10], [2000, 3, 12], [-1800, -900, 10], [0, 0, 11]],    # 3800 x 900 blocks: shift >= 1
    [[4000, 4000, 4000], [-4000, -4000, -4000], [0, 0, 10], [1, 0, 10], [0, 1, 10]],        # the whole short range: large shift
])
def test_raycast_on_hand_built_sparse_volumes(tg, blocks):
    g, o = slab_volume(tg, blocks)
    K = (300.0, 300.0, 159.5, 119.5)
    cam = tg.CameraParams(K, 240, 320)
    ident = np.array([0, 0, 0, 1], np.float32)
    yaw = np.array([0, np.sin(0.2), 0, np.cos(0.2)], np.float32)
    views = [(ident, np.zeros(3, np.float32), 4.0), (ident, np.array([-0.08, -0.05, 0.3], np.float32), 10.0),
             (yaw, np.array([0.3, 0.0, -0.2], np.float32), 4.0), (ident, np.array([0.0, 0.0, 80.0], np.float32), 6.0)]
    hits = 0
    for q, t, md in views:
        er = g.RayCast(md, cam, (q, t))
        orr = o.raycast(md, 320, 240, np.array(K, np.float32), q, t)
        rep = compare.compare_raycast(er, orr[:3], f"blocks {blocks[3]} view {t} md {md}")
        hits += rep["hits"]
    assert hits > 1000
    assert compare.compare_gather(g.GatherValid(), o.gather(), "sparse gather")["tsdf_bit_exact"]
    g.close()


def test_raycast_while_the_map_layout_changes(tg):
    """The skip map is rebuilt without a fill pass: its two byte planes swap roles and the rebuild restores the plane it
    will mark into next.  A volume that GROWS between views changes the layout every time -- from 3.4 M cells at one
    block per cell, to 0.8 M coarser cells (the cell count shrinks), back up to 3.4 M coarser cells, then to the
    largest shift -- and a view in between deletes blocks again.  Every view must equal the oracle."""
    near = [[0, 0, 10], [1, 0, 10], [0, 1, 10], [1, 1, 10], [-1, -1, 10]]
    g, o = slab_volume(tg, near + [[149, 149, 159]])                       # AABB 151 x 151 x 150 cells, shift 0
    K = (300.0, 300.0, 159.5, 119.5)
    cam = tg.CameraParams(K, 240, 320)
    ident = np.array([0, 0, 0, 1], np.float32)
    views = [(ident, np.zeros(3, np.float32), 4.0), (ident, np.array([-0.08, -0.05, 0.3], np.float32), 10.0)]
    rng = np.random.RandomState(11)

    def add(blocks):
        g.allocate_blocks(blocks)
        for b in blocks:
            o.allocate_block(*b)
            pts = np.array([[b[0] * 8 + x, b[1] * 8 + y, b[2] * 8 + z] for z in range(8) for y in range(8) for x in range(8)], np.int32)
            tsdf = np.clip((3.5 - (pts[:, 2] - b[2] * 8)) / 6.0, -1, 1).astype(np.float32)
            rgbw = np.concatenate([rng.randint(0, 255, (512, 3)), np.full((512, 1), 7)], 1).astype(np.uint8)
            g.assign(pts, tsdf=tsdf, rgbw=rgbw)
            o.set_voxels(pts, tsdf=tsdf, rgbw=rgbw)

    def check(label):
        hits = 0
        for q, t, md in views:
            hits += compare.compare_raycast(g.RayCast(md, cam, (q, t)), o.raycast(md, 320, 240, np.array(K, np.float32), q, t)[:3], label)["hits"]
        assert hits > 500, label

    check("shift 0, 3.4 M cells")
    add([[300, 100, 100]])                                                   # 302 x 151 x 150 blocks: shift 1, 0.86 M cells
    check("shift 1, fewer cells than before")
    add([[300, 300, 309], [2, 2, 10]])                                      # 302 x 302 x 300 blocks at shift 1: 3.4 M cells
    check("shift 1, as many cells as the first layout")
    g.delete_blocks([[1, 1, 10]])
    o.delete_block(1, 1, 10)
    check("after deleting a block (same layout)")
    add([[-3000, 5, 7], [0, 1, 11]])                                        # a much larger shift
    check("large shift")
    attempts, rebuilds = g.skip_map_stats()
    assert rebuilds == 5 and attempts == 5  # one rebuild per change, none for the second view of a pair
    g.close()


def test_raycast_at_the_edge_of_the_short_range(tg):
    """A camera 655 m from the origin at 2 cm voxels: sample coordinates pass -32768 and the reference's float -> short
    cast saturates (utils/tsdf/voxel_mem.cuh via .cast<short>()).  The engine picks its clamping ray-cast variant for
    such views (the usual variant has no clamp); the images must still equal the oracle's."""
    blocks = [[0, 0, -4096], [1, 0, -4096], [0, 1, -4096], [-1, 0, -4096], [0, -1, -4096], [0, 0, 10]]
    g, o = slab_volume(tg, blocks)
    K = (300.0, 300.0, 159.5, 119.5)
    cam = tg.CameraParams(K, 240, 320)
    ident = np.array([0, 0, 0, 1], np.float32)
    hits = 0
    for tz in (655.5, 655.9, 657.0):  # cam_T_world translation: the camera centre is at world z = -tz
        t = np.array([-0.05, -0.03, tz], np.float32)
        er = g.RayCast(4.0, cam, (ident, t))
        orr = o.raycast(4.0, 320, 240, np.array(K, np.float32), ident, t)
        hits += compare.compare_raycast(er, orr[:3], f"edge of range tz {tz}")["hits"]
    assert hits > 1000
    g.close()


def test_integrate_with_divisors_outside_the_shared_reciprocal_range(tg):
    """max_depth >= 2^20 (or a truncation / weight outside (2^-20, 2^20)) switches Integrate to its true-division kernel
    variant; a depth equal to max_depth on a fresh voxel gives weight 0 -> 0 / 0 like the reference (cold fallback of
    the usual variant).  Both must equal the oracle bit for bit."""
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    for max_depth, poison in ((2.0e6, False), (cfg.max_depth, True)):
        g = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
        o = Oracle(cfg.voxel_size, cfg.truncation)
        for i in range(3):
            f = sc.frame(i)
            depth = f["depth"].copy()
            if poison:  # some pixels exactly at max_depth: weight_new == 0, combined weight 0 on never-seen voxels
                depth[::7, ::5] = np.where(depth[::7, ::5] > 0, np.float32(max_depth), 0)
            oc = o.integrate(f["rgb"], depth, f["ht"], f["lt"], max_depth, f["K"], f["q"], f["t"])
            g.Integrate(f["rgb"], depth, f["ht"], f["lt"], max_depth, f["K"], (f["q"], f["t"]))
            ec = g.counters()
            assert (ec["n_new"], ec["n_visible"], ec["n_updated"], ec["n_carved"]) == (oc["n_new"], oc["n_vis"], oc["n_upd"], oc["n_carved"]), (max_depth, i)
        ek, et, er, ep = g.export()
        ok, ot, orr, op = o.export()
        assert np.array_equal(ek, ok)
        nan = np.isnan(ot)
        assert nan.any() == poison and np.array_equal(np.isnan(et), nan), max_depth   # 0 / 0 voxels: NaN in both (payloads are platform-specific)
        assert np.array_equal(et[~nan].view(np.uint32), ot[~nan].view(np.uint32)), max_depth
        assert np.array_equal(er, orr) and np.array_equal(np.isnan(ep), np.isnan(op))
        assert np.abs(ep[~np.isnan(op)] - op[~np.isnan(op)]).max() <= 1e-5
        g.close()


def test_rehash_in_the_middle_of_a_run(tg):
    """Fill more than half of a small table with live + tombstoned slots, then integrate: the garbage collection
    (clear + re-insert, launched when a frame retires) must leave every block reachable and the volume exactly
    as the oracle has it."""
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    g = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 12, table_slots=1 << 13, max_image_pixels=cfg.width * cfg.height)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    far = lambda r: [[i % 60 - 30, 2000 + r, i // 60] for i in range(1500)]  # far outside every view
    for r in range(3):  # 4500 distinct keys pass through the 8192-slot table: > 4096 non-empty slots
        g.allocate_blocks(far(r))
        if r < 2:
            g.delete_blocks(far(r))
    for k in far(2):
        o.allocate_block(*k)
    pts = [[k[0] * 8 + 1, k[1] * 8 + 2, k[2] * 8 + 3] for k in far(2)]
    vals = np.linspace(-0.5, 0.5, len(pts)).astype(np.float32)
    g.assign(pts, tsdf=vals)
    o.set_voxels(pts, tsdf=vals)
    for i in range(4):  # frames 0.. retire -> rehash launched -> later frames insert into the rebuilt table
        f = sc.frame(i)
        oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        ec = g.counters()
        assert (ec["n_new"], ec["n_visible"], ec["n_updated"], ec["n_carved"]) == (oc["n_new"], oc["n_vis"], oc["n_upd"], oc["n_carved"]), i
    tsdf, _, _, found = g.retrieve(pts)
    assert found.all() and np.array_equal(tsdf, vals)
    assert not g.retrieve([[k[0] * 8, k[1] * 8, k[2] * 8] for k in far(0)])[3].any()
    assert compare.compare_volumes(g.export(), o.export(), "after rehash")["tsdf_bit_exact"]
    cam = tg.CameraParams(f["K"], cfg.height, cfg.width)
    compare.compare_raycast(g.RayCast(cfg.max_depth, cam, (f["q"], f["t"])), o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])[:3], "after rehash")
    g.close()


def test_pool_exhaustion_inside_integrate_is_reported(tg):
    cfg = synth.config("tiny")
    f = synth.Scene(cfg).frame(0)
    g = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=256, table_slots=1 << 10, max_image_pixels=cfg.width * cfg.height)
    with pytest.raises(tg.TsdfError) as ei:
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    assert ei.value.code == -3 and g.NumActiveBlock() <= 256
    g.close()


def test_ragged_sizes_repeated_and_empty_frames(tg):
    """Image sizes that are not multiples of the 32 x 8 CTA tile, the same frame twice, an all-invalid frame."""
    base = synth.config("tiny")
    for w, h in ((161, 119), (97, 61), (33, 9)):
        cfg = base.scaled(1.0, name="ragged")
        cfg = type(cfg)(**{**cfg.__dict__, "width": w, "height": h, "K": (120.0, 120.0, (w - 1) / 2, (h - 1) / 2)})
        sc = synth.Scene(cfg)
        g = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 13, table_slots=1 << 15, max_image_pixels=w * h)
        o = Oracle(cfg.voxel_size, cfg.truncation)
        frames = [sc.frame(0), sc.frame(0), sc.frame(1)]
        empty = dict(frames[2])
        empty["depth"] = np.zeros_like(empty["depth"])
        frames.append(empty)
        for f in frames:
            o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
            g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        assert compare.compare_volumes(g.export(), o.export(), f"{w}x{h}")["tsdf_bit_exact"]
        cam = tg.CameraParams(f["K"], h, w)
        compare.compare_raycast(g.RayCast(cfg.max_depth, cam, (f["q"], f["t"])), o.raycast(cfg.max_depth, w, h, f["K"], f["q"], f["t"])[:3], f"{w}x{h}")
        g.close()


def test_blocking_sync_flag_gives_identical_results(tg):
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    a = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
    b = tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots, blocking_sync=True)
    for i in range(3):
        f = sc.frame(i)
        for g in (a, b):
            g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    for x, y in zip(a.export(), b.export()):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    cam = tg.CameraParams(f["K"], cfg.height, cfg.width)
    for x, y in zip(a.RayCast(cfg.max_depth, cam, (f["q"], f["t"])), b.RayCast(cfg.max_depth, cam, (f["q"], f["t"]))):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
    assert np.array_equal(a.GatherValid().view(np.uint32), b.GatherValid().view(np.uint32)) or a.GatherValid().shape == b.GatherValid().shape
    a.close()
    b.close()
