#!/usr/bin/env python
"""What does the shared-volume ray march cost beyond the local one, with the NVLink taken out?  (development probe)
Config 2 frames into (a) one engine and (b) two shards that live on the SAME GPU (tsdf_peer_attach_local), with and
without a TSDF mirror; the full 1280x720 view is then rendered by tsdf_raycast_device resp. tsdf_raycast_shared and
timed with the engine's phase timers.  Any difference is kernel overhead of raycast_kernel<true, *>, not link latency."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disinfect_slam_b200 import synth  # noqa: E402


def main():
    import torch
    from disinfect_slam_b200 import tsdf_grid
    cfg = synth.config("config2")
    sc = synth.Scene(cfg)
    H, W = cfg.height, cfg.width
    n = 10
    frames = [sc.frame(i) for i in range(n)]
    mk = lambda **kw: tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=1 << 17, table_slots=1 << 19, max_image_pixels=H * W, **kw)  # noqa: E731
    one = mk()
    for mode in ("no mirror", "mirror"):
        shards = [mk(shard_rank=r, shard_count=2, shard_shift=2) for r in range(2)]
        mirrors = [torch.zeros(2 * (1 << 17) * 512, dtype=torch.float32, device="cuda") for _ in range(2)]
        if mode == "mirror":
            for g in shards:
                ptrs = (C.c_void_p * 2)(*[m.data_ptr() for m in mirrors])
                tsdf_grid.check(g.L.tsdf_mirror_attach(g.h, 2, ptrs, 1 << 17))
        for f in frames:
            for g in shards + ([one] if mode == "no mirror" else []):
                g.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], (f["q"], f["t"]))
        for g in shards:
            g.peer_attach_local(shards)
        f = frames[-1]
        cam = tsdf_grid.CameraParams(f["K"], H, W)
        out = [torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda"), torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda"),
               torch.zeros((H, W), dtype=torch.float32, device="cuda")]
        ref = [torch.zeros_like(o) for o in out]
        if mode == "no mirror":
            for rep in range(3):
                one.RayCastDevice(cfg.max_depth, cam, (f["q"], f["t"]), ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr())
            one.synchronize(); one.set_profiling(True)
            for rep in range(10):
                one.RayCastDevice(cfg.max_depth, cam, (f["q"], f["t"]), ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr())
            ms, cnt = one.phase_ms()
            print(f"local volume, tsdf_raycast_device:           {1e3 * ms['raycast'] / cnt['raycast']:7.1f} us per view (map cached)")
            keep_ref = [r.clone() for r in ref]
        g = shards[0]
        for rep in range(3):
            g.RayCastShared(cfg.max_depth, cam, (f["q"], f["t"]), 0, H, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
        g.synchronize(); g.set_profiling(True)
        for rep in range(10):
            g.RayCastShared(cfg.max_depth, cam, (f["q"], f["t"]), 0, H, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
        ms, cnt = g.phase_ms()
        same = all(bool((a == b).all()) for a, b in zip(out, keep_ref))
        print(f"2 shards on one GPU ({mode:9s}), raycast_shared: {1e3 * ms['raycast'] / cnt['raycast']:7.1f} us per view (map cached), identical: {same}")
        for g in shards:
            g.close()


if __name__ == "__main__":
    main()
