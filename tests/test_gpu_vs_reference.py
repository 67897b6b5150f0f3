"""GPU parity tests of the engine (through the C ABI) against the REFERENCE's own CUDA TSDFGrid rebuilt
for sm_100a (oracle/_ref/libref_tsdf_parity.so, built by oracle/build_ref.sh in the development
container from /root/reference/utils/tsdf/*.cu, unmodified, -fmad=false; shipped prebuilt).

Same protocol as tests/test_oracle_vs_reference_golden.py: the reference's bucket-lock table delays
the allocation of blocks that lose a lock (utils/tsdf/voxel_hash.cu:83-88); those are "don't care",
every other block must match bit for bit; RayCast / Gather are compared on identical volumes.
"""
import numpy as np
import pytest

from disinfect_slam_b200 import synth
from oracle import compare, ref_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("ref_parity_lib")]  # a missing oracle/_ref FAILS (conftest.py), it does not skip


def keyset(k):
    return set(map(tuple, np.asarray(k).tolist()))


@pytest.mark.parametrize("name,n_frames", [("tiny", 6), ("small", 4)])
def test_integrate_matches_reference_kernels(tsdf_lib, name, n_frames):
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config(name)
    sc = synth.Scene(cfg)
    g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
    r = ref_cuda.RefTSDFGrid(cfg.voxel_size, cfg.truncation, parity=True)
    dont_care = set()
    for i in range(n_frames):
        f = sc.frame(i)
        r.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        ek, et, ec, ep = g.export()
        rk, rt, rc, rp = r.export()
        es, rs = keyset(ek), keyset(rk)
        assert rs <= es, f"frame {i}: reference-only blocks {sorted(rs - es)[:5]}"
        assert r.num_active() == len(rk) and g.NumActiveBlock() == len(ek)
        dont_care |= (es - rs)
        assert len(dont_care) <= 0.08 * len(es)
        ei = {k: j for j, k in enumerate(map(tuple, ek.tolist()))}
        clean = np.array([tuple(k) not in dont_care for k in rk.tolist()])
        sel = np.array([ei[tuple(k)] for k in rk.tolist()])[clean]
        assert clean.mean() >= 0.95
        rt, rc, rp = rt[clean], rc[clean], rp[clean]
        assert np.array_equal(et[sel].view(np.uint32), rt.view(np.uint32)), f"frame {i}: TSDF not bit-identical"
        assert np.array_equal(ec[sel][..., 3], rc[..., 3]), f"frame {i}: weights differ"
        seen = rc[..., 3] > 0  # colour of never-updated voxels is stale pool memory in the reference
        assert np.array_equal(ec[sel][..., :3][seen], rc[..., :3][seen]), f"frame {i}: colours differ"
        assert np.abs(ep[sel].astype(np.float64) - rp).max() <= compare.PROB_TOL
    g.close()
    r.close()


def test_raycast_and_gather_match_reference_kernels(tsdf_lib):
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config("small")
    sc = synth.Scene(cfg)
    g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
    r = ref_cuda.RefTSDFGrid(cfg.voxel_size, cfg.truncation, parity=True)
    f = sc.frame(0)
    r.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
    g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    rk = r.export(voxels=False)[0]
    extra = sorted(keyset(g.export(voxels=False)[0]) - keyset(rk))
    if extra:
        g.delete_blocks(extra)  # make the volumes identical: drop what the reference's lock losers lack
    ek, et, _, _ = g.export()
    assert np.array_equal(ek, rk) and np.array_equal(et.view(np.uint32), r.export()[1].view(np.uint32))

    def check(img, ref, what):
        hit, rhit = img[..., 3] > 0, ref[..., 3] > 0
        assert np.array_equal(hit, rhit), f"{what}: hit masks differ in {(hit != rhit).sum()} rays"
        d = np.abs(img.astype(np.int32) - ref.astype(np.int32))
        assert d.max(initial=0) <= 1 and (d.max(-1) > 0).mean() <= 1e-3, f"{what}: max {d.max()}, frac {(d.max(-1) > 0).mean():.2e}"

    views = [(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"]), (10.0, cfg.width, cfg.height, f["K"], f["q"], f["t"])]
    v = sc.virtual_view(2, 5, width=256, height=144, K=(180.0, 180.0, 127.5, 71.5))
    views.append((4.0, v["width"], v["height"], v["K"], v["q"], v["t"]))
    for md, w, h, K, q, t in views:
        rgba, normal, depth = g.RayCast(md, tsdf_grid.CameraParams(K, h, w), (q, t))
        rr, rn = r.raycast(md, w, h, K, q, t)
        check(rgba, rr, f"rgba md={md} {w}x{h}")
        check(normal, rn, f"normal md={md} {w}x{h}")
        assert np.array_equal(np.isfinite(depth), rr[..., 3] > 0)
    a = compare.canonical_gather(g.GatherValid())
    b = compare.canonical_gather(r.gather())
    assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    bbox = (-0.5, 1.2, -1.4, 0.6, -2.2, 0.9)
    a = compare.canonical_gather(g.GatherVoxels(tsdf_grid.BoundingCube(*bbox)))
    b = compare.canonical_gather(r.gather(bbox))
    assert len(a) > 0 and a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    g.close()
    r.close()


def test_reference_as_shipped_agrees_within_tolerance(tsdf_lib):
    """The bit-exact pin above is the reference compiled WITHOUT FMA contraction.  The reference as its own build
    compiles it (nvcc's default contraction, oracle/_ref/libref_tsdf.so -- the timing baseline of `bench.py --impl
    reference`) rounds a few projections differently, so a voxel now and then takes the neighbouring pixel or a block
    at the edge of the truncation band appears one frame earlier or later.  That cannot be bit-identical; this test
    bounds how far apart the two volumes are after six frames, the reference's lock losers included (blocks
    it allocates a frame late, 2.4 % here, carry one update less): no reference-only block, and on the common blocks
    nearly all voxels equal to 1e-5 of the truncation-normalised TSDF with equal weights.  Measured on B200:
    0.76 % of the observed voxels differ by more than 1e-5, 0.34 % by more than 1e-2, weights differ on 0.74 %."""
    from disinfect_slam_b200 import tsdf_grid
    from oracle import ref_cuda as rc_mod
    if not rc_mod.available(False):
        pytest.fail("oracle/_ref/libref_tsdf.so is missing: run __graft_entry__.build() where /root/reference exists")
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
    r = rc_mod.RefTSDFGrid(cfg.voxel_size, cfg.truncation, parity=False)
    for i in range(6):
        f = sc.frame(i)
        r.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    ek, et, ec, ep = g.export()
    rk, rt, rcol, rp = r.export()
    es, rs = keyset(ek), keyset(rk)
    common = sorted(es & rs)
    ei = {k: j for j, k in enumerate(map(tuple, ek.tolist()))}
    ri = {k: j for j, k in enumerate(map(tuple, rk.tolist()))}
    a = np.array([ei[k] for k in common]); b = np.array([ri[k] for k in common])
    seen = (ec[a][..., 3] > 0) & (rcol[b][..., 3] > 0)
    dt = np.abs(et[a].astype(np.float64) - rt[b])[seen]
    stats = {
        "blocks_engine": len(es), "blocks_reference": len(rs),
        "reference_only_frac": len(rs - es) / max(len(rs), 1), "engine_only_frac": len(es - rs) / max(len(es), 1),
        "tsdf_frac_above_1e-5": float((dt > 1e-5).mean()), "tsdf_frac_above_1e-2": float((dt > 1e-2).mean()), "tsdf_max": float(dt.max(initial=0)),
        "weight_differs_frac": float((ec[a][..., 3] != rcol[b][..., 3]).mean()),
        "colour_differs_by_more_than_1_frac": float((np.abs(ec[a][..., :3].astype(int) - rcol[b][..., :3].astype(int)).max(-1)[seen] > 1).mean()),
        "prob_max": float(np.abs(ep[a].astype(np.float64) - rp[b])[seen].max(initial=0)),
    }
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "reference_default_flags_stats.json"), "w") as fh:
            json.dump(stats, fh)
    g.close()
    r.close()
    assert stats["reference_only_frac"] <= 0.02 and stats["engine_only_frac"] <= 0.10, stats  # (engine-only includes the reference's lock losers)
    assert stats["tsdf_frac_above_1e-5"] <= 0.03 and stats["tsdf_frac_above_1e-2"] <= 0.015, stats
    assert stats["weight_differs_frac"] <= 0.03 and stats["colour_differs_by_more_than_1_frac"] <= 0.03, stats
