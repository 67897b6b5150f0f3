"""Mesh extraction.  PARITY UNPINNED against the reference: its mesher is KrisLibrary
(external, absent) -- see oracle/mesh_oracle.py.  CPU tests hold the oracle to size-independent properties of a correct
iso-surface; GPU tests hold the engine's table-driven kernel to the oracle's direct statement of the rule, bit for bit."""
import numpy as np
import pytest

from oracle import mesh_oracle as M


sphere_blocks = M.sphere_volume


def test_oracle_sphere_is_closed_oriented_and_has_the_right_volume():
    keys, tsdf, rgbw, vs = sphere_blocks()
    tris = M.extract_mesh(keys, tsdf, rgbw, vs)
    pr = M.mesh_properties(tris)
    assert pr["triangles"] > 30000
    assert pr["edges_shared_by_2"] == pr["edges"] and pr["edges_unbalanced"] == 0      # watertight, consistently oriented
    assert abs(pr["volume"] / (4 / 3 * np.pi * 0.9 ** 3) - 1) < 5e-3                     # positive: normals point outwards
    assert abs(pr["area"] / (4 * np.pi * 0.9 ** 2) - 1) < 5e-3
    r = np.linalg.norm(tris.reshape(-1, 3).astype(np.float64) - np.array([0.013, -0.021, 0.007]), axis=1)
    assert np.abs(r - 0.9).max() < 0.05 * 0.2                                            # vertices on the zero crossing
    # triangle order does not matter to the canonical form
    assert np.array_equal(M.canonical_triangles(tris), M.canonical_triangles(tris[::-1].copy()))


def test_oracle_skips_unobserved_and_unallocated_cells():
    keys, tsdf, rgbw, vs = sphere_blocks(observed=lambda c: c[:, 0] < 0.31)   # nothing observed beyond x = 0.31 m
    tris = M.extract_mesh(keys, tsdf, rgbw, vs)
    assert len(tris) > 1000 and tris[:, :, 0].max() < 0.31                      # an open surface, cut at the boundary
    pr = M.mesh_properties(tris)
    assert pr["edges_unbalanced"] > 0 and pr["edges_shared_by_2"] < pr["edges"]   # it has a rim
    # dropping blocks = unallocated neighbours: cells reaching into them vanish, nothing else changes
    keep = keys[:, 2] < 1
    part = M.extract_mesh(keys[keep], tsdf[keep], rgbw[keep], vs)
    assert 0 < len(part) < len(tris)
    full = {t.tobytes() for t in M.canonical_triangles(tris)}
    assert all(t.tobytes() in full for t in M.canonical_triangles(part))


def test_oracle_bbox_selects_like_gather():
    keys, tsdf, rgbw, vs = sphere_blocks()
    bbox = (-0.81, 0.74, -2.0, 2.0, -2.0, 2.0)   # x voxels -16 .. 14: x blocks -2 .. 0 lie fully inside (inclusive voxel bound)
    sel = M.select_blocks(keys, vs, bbox)
    assert sorted(set(keys[sel][:, 0].tolist())) == [-2, -1, 0]
    tris = M.extract_mesh(keys, tsdf, rgbw, vs, bbox)
    # cells are based at voxels of selected blocks: x voxel index <= 7 -> every vertex has x <= (7 + 1 + .5) * vs
    assert len(tris) > 0 and tris[:, :, 0].max() <= 8.5 * vs + 1e-6


def test_every_case_of_the_table_matches_the_rule():
    """The engine's generated table (tools/gen_mesh_table.py -> csrc/mesh_table.inc) against the oracle's direct
    application of the rule, all 256 inside masks."""
    import os
    import re
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "disinfect_slam_b200", "csrc", "mesh_table.inc")
    rows = [list(map(int, m.group(1).split(","))) for m in re.finditer(r"^\{([0-9,]+)\},$", open(inc).read(), re.M)]
    assert len(rows) == 256
    canon = lambda tris: sorted(min(t[i:] + t[:i] for i in range(3)) for t in tris)  # noqa: E731
    for mask, row in enumerate(rows):
        n = row[0]
        tab = [tuple((c >> 3, c & 7) for c in row[1 + 3 * t:4 + 3 * t]) for t in range(n)]
        assert canon(tab) == canon([tuple(t) for t in M.cell_triangles(mask)]), mask


@pytest.mark.gpu
def test_engine_mesh_equals_oracle_on_a_sphere(tsdf_lib):
    from disinfect_slam_b200 import tsdf_grid
    keys, tsdf, rgbw, vs = sphere_blocks(observed=lambda c: c[:, 0] < 0.61)
    g = tsdf_grid.TSDFGrid(vs, 0.3, pool_blocks=1 << 10, table_slots=1 << 12, max_image_pixels=64 * 64)
    g.allocate_blocks(keys)
    k = np.arange(512)
    for i, (bx, by, bz) in enumerate(keys.tolist()):
        pts = np.stack([bx * 8 + (k & 7), by * 8 + ((k >> 3) & 7), bz * 8 + (k >> 6)], -1).astype(np.int32)
        g.assign(pts, tsdf=tsdf[i], rgbw=rgbw[i])
    for bbox in (None, (-0.81, 0.74, -2.0, 2.0, -0.41, 2.0)):
        got = g.ExtractMesh(bbox)
        want = M.extract_mesh(keys, tsdf, rgbw, vs, bbox)
        assert got.shape == want.shape and len(got) > 1000
        assert np.array_equal(M.canonical_triangles(got).view(np.uint32), M.canonical_triangles(want).view(np.uint32))
    assert g.ExtractMesh((5.0, 6.0, 5.0, 6.0, 5.0, 6.0)).shape == (0, 3, 3)
    g.close()


@pytest.mark.gpu
def test_engine_mesh_equals_oracle_after_integration(tsdf_lib):
    """The real thing: integrate synthetic frames, then mesh the volume; engine == oracle on the identical volume, and the
    surface sits on the scene (vertices within a voxel diagonal of a sign change are implied by construction)."""
    from disinfect_slam_b200 import synth, tsdf_grid
    from oracle.oracle import Oracle
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    for i in range(4):
        f = sc.frame(i)
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
    keys, tsdf, rgbw, _ = o.export()
    for bbox in (None, (-1.0, 1.0, -1.4, 1.0, -2.5, 2.5)):
        got = g.ExtractMesh(bbox)
        want = M.extract_mesh(keys, tsdf, rgbw, cfg.voxel_size, bbox)
        assert got.shape == want.shape and len(got) > 5000
        assert np.array_equal(M.canonical_triangles(got).view(np.uint32), M.canonical_triangles(want).view(np.uint32))
    assert g.ExtractMesh(to_host=False) == len(M.extract_mesh(keys, tsdf, rgbw, cfg.voxel_size))
    g.close()
