#!/usr/bin/env python
"""Where does the end-to-end leg lose time?  Variants of tsdf_streams_run on the config-2 workload (development probe,
not part of bench.py): with / without the image download, with / without RayCast, float32 / 16-bit planes, 4 / 8 streams."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disinfect_slam_b200 import synth  # noqa: E402
import bench  # noqa: E402


def main():
    cfg = synth.config("config2")
    n_frames, W, K = 20, 4, 40
    B = int(os.environ.get("PROBE_STREAMS", "8"))
    streams = bench.generate_streams(cfg, 0, B, n_frames)
    from disinfect_slam_b200 import tsdf_grid
    H, Wd = cfg.height, cfg.width
    npx = H * Wd
    keep = []

    def pin(a):
        p = tsdf_grid.PinnedArray(a.shape, a.dtype)
        p.array[...] = a
        keep.append(p)  # the view does not own the pinned allocation
        return p

    f32, u16, f32p, u16p = [], [], [], []
    q16 = lambda x, s: np.clip(np.rint(x.astype(np.float64) * s), 0, 65535).astype(np.uint16)  # noqa: E731
    for st in streams:
        a, b, ap, bp = [], [], [], []
        for i in range(n_frames):
            rgb = pin(st["rgb"][i])
            a.append(dict(rgb=rgb.array, depth=pin(st["depth"][i]).array, ht=pin(st["ht"][i]).array, lt=pin(st["lt"][i]).array, q=st["q"][i], t=st["t"][i]))
            b.append(dict(rgb=rgb.array, depth=pin(q16(st["depth"][i], cfg.depth_factor)).array, ht=pin(q16(st["ht"][i], 65535)).array,
                          lt=pin(q16(st["lt"][i], 65535)).array, q=st["q"][i], t=st["t"][i]))
            blk, d = tsdf_grid.packed_pinned_frame(st["rgb"][i], st["depth"][i], st["ht"][i], st["lt"][i])
            keep.append(blk)
            ap.append(dict(d, q=st["q"][i], t=st["t"][i]))
            blk, d = tsdf_grid.packed_pinned_frame(st["rgb"][i], q16(st["depth"][i], cfg.depth_factor), q16(st["ht"][i], 65535), q16(st["lt"][i], 65535))
            keep.append(blk)
            bp.append(dict(d, q=st["q"][i], t=st["t"][i]))
        f32.append(a)
        u16.append(b)
        f32p.append(ap)
        u16p.append(bp)
    rgba = [tsdf_grid.PinnedArray((H, Wd, 4), np.uint8) for _ in range(2 * B)]
    normal = [tsdf_grid.PinnedArray((H, Wd, 4), np.uint8) for _ in range(2 * B)]
    oblk, rgba_p, normal_p, _ = tsdf_grid.packed_pinned_images(H, Wd, 2 * B)

    def run(name, per_stream, nb, raycast, download):
        engs = [tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots, max_image_pixels=npx) for _ in range(nb)]
        frames = tsdf_grid.make_host_frames(per_stream[:nb])
        if download == 2:
            kw = dict(depthmap_factor=cfg.depth_factor, raycast=raycast, rgba=rgba_p[:2 * nb], normal=normal_p[:2 * nb])
        else:
            kw = dict(depthmap_factor=cfg.depth_factor, raycast=raycast, rgba=[p.array for p in rgba[:2 * nb]] if download else None,
                      normal=[p.array for p in normal[:2 * nb]] if download else None)
        tsdf_grid.run_streams(engs, frames, 0, W, Wd, H, cfg.max_depth, streams[0]["K"], **kw)
        t0 = time.perf_counter()
        tsdf_grid.run_streams(engs, frames, W, K, Wd, H, cfg.max_depth, streams[0]["K"], **kw)
        dt = time.perf_counter() - t0
        print(f"{name:46s} streams {nb}  {nb * K / dt:8.1f} frames/s  {1e3 * dt / K:7.3f} ms/step", flush=True)
        for g in engs:
            g.close()

    for nb in sorted({4, B}):
        print("TSDF_STREAMS_THREADS", os.environ.get("TSDF_STREAMS_THREADS"))
        run("u16 integrate + raycast + download", u16, nb, True, True)
        run("u16 integrate + raycast, no download", u16, nb, True, False)
        run("u16 integrate only", u16, nb, False, False)
        run("f32 integrate only", f32, nb, False, False)
        run("f32 integrate + raycast + download", f32, nb, True, True)
        run("u16 PACKED integrate + raycast + PACKED download", u16p, nb, True, 2)
        run("f32 PACKED integrate + raycast + PACKED download", f32p, nb, True, 2)
        run("u16 PACKED integrate only", u16p, nb, False, False)
        run("f32 PACKED integrate only", f32p, nb, False, False)


if __name__ == "__main__":
    main()
