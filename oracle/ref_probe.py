import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
from disinfect_slam_b200 import synth, tsdf_grid
from oracle import compare
from oracle.oracle import Oracle
from oracle.ref_cuda import RefTSDFGrid
cfg = synth.config(sys.argv[1] if len(sys.argv) > 1 else "tiny")
sc = synth.Scene(cfg)
r = RefTSDFGrid(cfg.voxel_size, cfg.truncation, parity=True)
o = Oracle(cfg.voxel_size, cfg.truncation)
g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
for i in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    f = sc.frame(i)
    t0=time.time(); r.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"]); t1=time.time()
    oc = o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
    g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
    rk, rt, rc, rp = r.export()
    ok, ot, oc_, op = o.export()
    rs = set(map(tuple, rk.tolist())); os_ = set(map(tuple, ok.tolist()))
    print(f"frame {i}: ref {len(rk)} blocks ({(t1-t0)*1e3:.1f} ms), oracle {len(ok)}; only-ref {len(rs-os_)} only-oracle {len(os_-rs)}", oc)
    common = sorted(rs & os_, key=lambda k: (k[2], k[1], k[0]))
    ri = {tuple(k): j for j, k in enumerate(rk.tolist())}; oi = {tuple(k): j for j, k in enumerate(ok.tolist())}
    a = np.array([ri[k] for k in common]); b = np.array([oi[k] for k in common])
    dt = np.abs(rt[a] - ot[b]); dw = (rc[a][..., 3] != oc_[b][..., 3]); dc = (rc[a][..., :3] != oc_[b][..., :3]) & (oc_[b][..., 3:4] > 0)
    dp = np.abs(rp[a] - op[b])
    bad_blocks = (dt.max(1) > 0) | dw.any(1)
    print(f"   common {len(common)}: tsdf max {dt.max():.3g} (exact blocks {(dt.max(1)==0).sum()}), weight mismatches {dw.sum()}, rgb mismatches {dc.sum()}, prob max {dp.max():.3g}, blocks with any diff {bad_blocks.sum()}")
rr = r.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])
orr = o.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])
print("raycast rgba diff px:", (rr[0] != orr[0]).any(-1).sum(), "normal diff px:", (rr[1] != orr[1]).any(-1).sum(), "of", cfg.width*cfg.height)
gv = r.gather(); og = o.gather()
print("gather", gv.shape, og.shape)
ys, xs = np.nonzero((rr[0] != orr[0]).any(-1))
for k in range(0, len(ys), max(1, len(ys)//8)):
    y, x = ys[k], xs[k]
    print((y, x), "ref rgba", rr[0][y, x], "ora rgba", orr[0][y, x], "ref nrm", rr[1][y, x], "ora nrm", orr[1][y, x])
same = ((rr[0] == orr[0]).all(-1)) & (orr[0][..., 3] > 0)
print("hit pixels identical:", same.sum())
d = np.abs(rr[0].astype(int) - orr[0].astype(int)); print("max abs diff rgba per channel", d.reshape(-1, 4).max(0), "normal", np.abs(rr[1].astype(int) - orr[1].astype(int)).reshape(-1, 4).max(0))
