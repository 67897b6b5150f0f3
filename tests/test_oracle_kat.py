"""CPU tests: pin the oracle against the reference's own known-answer tests
(utils/tests/voxel_hash_test.cu, utils/tests/voxel_mem_test.cu) and check its invariants."""
import json
import os

import numpy as np
import pytest

from disinfect_slam_b200 import synth
from oracle.oracle import Oracle, RefHashModel, hash_block

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hash_kat.json")))


def test_hash_known_answers():
    # voxel_hash_test.cu:130-135 + formula rows
    for row in GOLD["hash"]:
        assert hash_block(*row["block"]) == row["bucket"], row
    m = RefHashModel()
    for row in GOLD["hash"]:
        assert m.hash(*row["block"]) == row["bucket"]


def test_ref_model_single():
    # voxel_hash_test.cu:56-92
    m = RefHashModel()
    assert m.allocate(1, 1, 1) == 1
    assert m.num_active() == 1
    assert m.find(1, 1, 1) >= 0
    assert m.find(0, 0, 0) == -1  # retrieve of an unallocated block -> default voxel
    m.reset_locks()
    assert m.allocate(0, 0, 0) == 1
    assert m.num_active() == 2


def test_ref_model_multiple():
    # voxel_hash_test.cu:94-126: 128 diagonal blocks in one launch, "assume no collision"
    m = RefHashModel()
    n = GOLD["multiple_blocks"]
    assert all(m.allocate(i, i, i) == 1 for i in range(n))
    m.reset_locks()
    assert m.num_active() == n
    idx = [m.find(i, i, i) for i in range(n)]
    assert len(set(idx)) == n and min(idx) >= 0


def test_ref_model_collision_sequence():
    # voxel_hash_test.cu:128-155: one insertion per bucket per pass -> 2, 3, 4 active blocks
    m = RefHashModel()
    counts = []
    for _ in range(3):
        for b in GOLD["collision_blocks"]:
            m.allocate(*b)
        m.reset_locks()
        counts.append(m.num_active())
    assert counts == GOLD["collision_active_counts_reference"]
    pool_idx = [m.find(*b) for b in GOLD["collision_blocks"]]
    assert min(pool_idx) >= 0 and len(set(pool_idx)) == 4
    # delete the list head / a list node and look the others up again (voxel_hash.cu:122-171)
    assert m.delete(*GOLD["collision_blocks"][1]) == 1
    m.reset_locks()
    assert m.find(*GOLD["collision_blocks"][1]) == -1
    assert m.find(*GOLD["collision_blocks"][0]) >= 0 and m.find(*GOLD["collision_blocks"][2]) >= 0
    assert m.num_active() == 3


def test_ref_model_pool():
    # voxel_mem_test.cu:38-90: distinct blocks, release then re-acquire returns the same indices
    m = RefHashModel(num_block=1 << 10)
    got = [m.pool_acquire() for _ in range(8)]
    assert len(set(got)) == 8
    for b in got:
        m.pool_release(b)
    again = [m.pool_acquire() for _ in range(8)]
    assert sorted(again) == sorted(got)


def test_ideal_oracle_allocates_every_request():
    # the oracle (and the new engine) use ideal set semantics: all 4 colliding blocks exist after one pass
    o = Oracle(0.01, 0.06)
    for b in GOLD["collision_blocks"]:
        o.allocate_block(*b)
    assert o.num_blocks() == 4
    found, tsdf, rgbw, prob = o.get_voxel(33 * 8, 180 * 8, 42 * 8)
    assert found and tsdf == -1.0 and rgbw[3] == 0 and prob == 0.5  # voxel_mem.cu:43-51
    found, tsdf, rgbw, prob = o.get_voxel(5, 5, 5 + 8 * 100)
    assert not found and tsdf == 1.0 and rgbw[3] == 0 and prob == 0.0  # voxel_types.cu:3-11


@pytest.fixture(scope="module")
def tiny_run():
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    counters, frames = [], []
    for i in range(4):
        f = sc.frame(i)
        frames.append(f)
        counters.append(o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"]))
    return cfg, sc, o, counters, frames


def test_oracle_integrate_invariants(tiny_run):
    cfg, sc, o, counters, frames = tiny_run
    for c in counters:
        assert c["n_active_post"] == c["n_active_pre"] + c["n_new"] - c["n_carved"]
        assert c["n_vis"] >= c["n_new"] > 0 and c["n_upd"] > 0
    keys, tsdf, rgbw, prob = o.export()
    assert len(keys) == o.num_blocks() == counters[-1]["n_active_post"]
    assert not np.isnan(tsdf).any() and tsdf.min() >= -1.0 and tsdf.max() <= 1.0
    assert rgbw[..., 3].max() <= 40  # weight clamp, voxel_tsdf.cu:192
    assert prob.min() > 0.0 and prob.max() < 1.0
    # every surviving block that was visible has a voxel with |tsdf| < .9 or was not visible in the last frame;
    # canonical order is strictly increasing
    k = keys.astype(np.int64)
    lin = (k[:, 2] * 65536 + k[:, 1]) * 65536 + k[:, 0]
    assert (np.diff(lin) > 0).all()


def test_oracle_gather_semantics(tiny_run):
    cfg, sc, o, counters, frames = tiny_run
    g = o.gather()
    assert g.shape == (o.num_blocks() * 512, 4)
    keys, tsdf, _, _ = o.export()
    # download_tsdf_kernel: position = grid * voxel_size, x fastest (voxel_tsdf.cu:34-46)
    b0 = g[:512]
    assert np.allclose(b0[0, :3], keys[0].astype(np.float32) * 8 * np.float32(cfg.voxel_size))
    assert np.array_equal(b0[:, 3], tsdf[0])
    assert np.isclose(b0[1, 0] - b0[0, 0], cfg.voxel_size) and b0[8, 1] > b0[0, 1] and b0[64, 2] > b0[0, 2]
    # bounding box: only blocks fully inside, inclusive, after truncation to short (voxel_tsdf.cuh:21-26)
    bbox = (-0.5, 0.5, -2.0, 2.0, -3.0, 3.0)
    gb = o.gather(bbox)
    scale = np.float32(1.0 / cfg.voxel_size)
    lim = [int(np.float32(v) * scale) for v in bbox]
    vox = keys.astype(np.int64) * 8
    inside = ((vox[:, 0] >= lim[0]) & (vox[:, 0] + 7 <= lim[1]) & (vox[:, 1] >= lim[2]) & (vox[:, 1] + 7 <= lim[3]) &
              (vox[:, 2] >= lim[4]) & (vox[:, 2] + 7 <= lim[5]))
    assert 0 < inside.sum() < len(keys)
    assert gb.shape[0] == inside.sum() * 512
    assert o.gather((10, 11, 10, 11, 10, 11)).shape[0] == 0


def test_oracle_raycast_matches_scene_depth(tiny_run):
    cfg, sc, o, counters, frames = tiny_run
    f = frames[-1]
    rgba, normal, depth, cnt = o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])
    hit = np.isfinite(depth)
    assert hit.mean() > 0.5 and cnt["hits"] == hit.sum()
    both = hit & (f["depth"] > 0) & (f["depth"] < cfg.max_depth - 0.2)
    err = np.abs(depth[both] - f["depth"][both])
    # the zero crossing of a TSDF fused from the same view lies within ~a voxel of the input depth
    assert np.median(err) < 1.5 * cfg.voxel_size
    assert (rgba[hit][:, 3] == 255).all() and (rgba[~hit] == 0).all() and (normal[~hit] == 0).all()


def test_generator_is_deterministic():
    cfg = synth.config("tiny")
    a, b = synth.Scene(cfg).frame(3), synth.Scene(cfg).frame(3)
    for k in ("rgb", "depth", "ht", "lt", "q", "t"):
        assert np.array_equal(a[k], b[k])
    assert abs(float(np.linalg.norm(a["q"].astype(np.float64))) - 1.0) < 1e-6
    assert (a["depth"] == 0).mean() > 0.005 and a["ht"].min() >= 0.02 and a["ht"].max() <= 0.98
