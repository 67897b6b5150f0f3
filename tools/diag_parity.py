import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from disinfect_slam_b200 import synth, tsdf_grid
from oracle.oracle import Oracle
cfg = synth.config("tiny"); sc = synth.Scene(cfg)
g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
o = Oracle(cfg.voxel_size, cfg.truncation)
for i in range(2):
    f = sc.frame(i)
    g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
    ek, et, ec, ep = g.export(); ok, ot, oc_, op = o.export()
    print("frame", i, "keys equal", np.array_equal(ek, ok), g.counters()["n_updated"], oc["n_upd"])
    if np.array_equal(ek, ok):
        dt = et.view(np.uint32) != ot.view(np.uint32)
        print("  tsdf bit mismatches", dt.sum(), "max abs", np.abs(et - ot).max(), "weight mism", (ec[..., 3] != oc_[..., 3]).sum(), "rgb mism", (ec[..., :3] != oc_[..., :3]).sum(), "prob max", np.abs(ep - op).max())
        if dt.sum():
            b, v = np.nonzero(dt)
            for k in range(min(5, len(b))):
                print("   ", ek[b[k]], v[k], et[b[k], v[k]], ot[b[k], v[k]], ec[b[k], v[k]], oc_[b[k], v[k]])
