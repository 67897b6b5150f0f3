"""Parity comparison helpers (TEST INFRASTRUCTURE: tests/, smoke(), bench checks only).

Tolerances (BASELINE.json north_star / BASELINE.md 4):
  block-coordinate sets ........ bit-exact
  TSDF ......................... |d| <= 1e-5 (the engine reproduces the reference arithmetic, so the
                                 tests additionally assert bit-equality and report it)
  weight, RGB .................. exact
  semantic probability ......... L-inf <= 1e-5 (engine stores logits; libm vs CUDA logf/expf)
  gather lists ................. exact after canonical ordering
  raycast hit mask / depth ..... exact; colours within +-1 where the semantic overlay is active
"""
import numpy as np

TSDF_TOL = 1e-5
PROB_TOL = 1e-5


def key_order(keys):
    """Canonical order: ascending (z, y, x) signed block coordinate."""
    k = np.asarray(keys, np.int64)
    return np.lexsort((k[:, 0], k[:, 1], k[:, 2]))


def compare_volumes(eng_export, ora_export, label=""):
    """Both exports are (keys[n,3], tsdf[n,512], rgbw[n,512,4], prob[n,512]) in canonical order."""
    ek, et, ec, ep = eng_export
    ok, ot, oc, op = ora_export
    res = {"n_blocks_engine": len(ek), "n_blocks_oracle": len(ok)}
    assert len(ek) == len(ok), f"{label}: block count differs: engine {len(ek)} oracle {len(ok)}"
    assert np.array_equal(ek, ok), f"{label}: block coordinate sets differ"
    if et is not None:
        d = np.abs(et.astype(np.float64) - ot.astype(np.float64))
        res["tsdf_max_abs"] = float(d.max()) if d.size else 0.0
        res["tsdf_bit_exact"] = bool(np.array_equal(et.view(np.uint32), ot.view(np.uint32)))
        assert res["tsdf_max_abs"] <= TSDF_TOL, f"{label}: TSDF differs by {res['tsdf_max_abs']}"
        res["weight_exact"] = bool(np.array_equal(ec[..., 3], oc[..., 3]))
        res["rgb_exact"] = bool(np.array_equal(ec[..., :3], oc[..., :3]))
        assert res["weight_exact"], f"{label}: weights differ"
        assert res["rgb_exact"], f"{label}: colours differ"
        dp = np.abs(ep.astype(np.float64) - op.astype(np.float64))
        res["prob_max_abs"] = float(dp.max()) if dp.size else 0.0
        assert res["prob_max_abs"] <= PROB_TOL, f"{label}: probability differs by {res['prob_max_abs']}"
    return res


def canonical_gather(g):
    """Sort a gather result (n*512 records of x,y,z,tsdf) by the block's first voxel (z,y,x)."""
    g = np.asarray(g, np.float32).reshape(-1, 512, 4)
    first = g[:, 0, :3].astype(np.float64)
    order = np.lexsort((first[:, 0], first[:, 1], first[:, 2]))
    return g[order].reshape(-1, 4)


def compare_gather(eng, ora, label=""):
    assert eng.shape == ora.shape, f"{label}: gather size differs: {eng.shape} vs {ora.shape}"
    a, b = canonical_gather(eng), canonical_gather(ora)
    assert np.array_equal(a[:, :3].view(np.uint32), b[:, :3].view(np.uint32)), f"{label}: gather positions differ"
    d = np.abs(a[:, 3].astype(np.float64) - b[:, 3].astype(np.float64))
    m = float(d.max()) if d.size else 0.0
    assert m <= TSDF_TOL, f"{label}: gather TSDF differs by {m}"
    return {"n_voxels": int(len(a)), "tsdf_max_abs": m,
            "tsdf_bit_exact": bool(np.array_equal(a[:, 3].view(np.uint32), b[:, 3].view(np.uint32)))}


def compare_raycast(eng, ora, label=""):
    """eng / ora = (rgba, normal, depth)."""
    er, en, ed = eng
    orr, on, od = ora
    ehit, ohit = np.isfinite(ed), np.isfinite(od)
    assert np.array_equal(ehit, ohit), f"{label}: hit masks differ in {(ehit != ohit).sum()} rays"
    assert np.array_equal(ed[ehit].view(np.uint32), od[ohit].view(np.uint32)), f"{label}: hit depths differ"
    dr = np.abs(er.astype(np.int32) - orr.astype(np.int32))
    dn = np.abs(en.astype(np.int32) - on.astype(np.int32))
    assert dr.max(initial=0) <= 1 and dn.max(initial=0) <= 1, f"{label}: colour differs by more than 1"
    frac = float(((dr.max(-1) > 0) | (dn.max(-1) > 0)).mean()) if dr.size else 0.0
    assert frac <= 1e-3, f"{label}: {frac:.2e} of rays differ in colour"
    return {"rays": int(ed.size), "hits": int(ehit.sum()), "colour_mismatch_frac": frac}
