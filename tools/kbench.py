#!/usr/bin/env python
"""Kernel micro-benchmark for experiments (development tool, not part of bench.py).

Runs BASELINE config 2 (1280x720, 5 mm) on `--streams` interleaved engines exactly like bench.py's roofline leg --
kernels serialised, CUDA events on the engine stream through the profiling API -- and prints microseconds per launch
of every phase, plus digests of the volumes and of a few rendered views.  Different builds of the library (selected
with TSDF_B200_LIB=<path>, see csrc/Makefile `EXTRA`) must print identical digests: a variant that is faster but not
bit-identical is not a candidate.

    TSDF_B200_LIB=disinfect_slam_b200/libtsdf_x1.so python tools/kbench.py --tag x1
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from disinfect_slam_b200 import synth  # noqa: E402
import bench  # noqa: E402

CACHE = "/tmp/kbench_frames.npz"


def frames(cfg, B, n):
    if os.path.exists(CACHE):
        z = np.load(CACHE, allow_pickle=True)
        if int(z["B"]) == B and int(z["n"]) == n:
            return z["streams"].tolist()
    st = bench.generate_streams(cfg, 0, B, n)
    np.savez(CACHE, B=B, n=n, streams=np.array(st, dtype=object))
    return st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="base")
    ap.add_argument("--streams", type=int, default=4)
    ap.add_argument("--frames", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--workload", default="config2")
    args = ap.parse_args()
    cfg = synth.config(args.workload)
    B, n, W = args.streams, args.frames, args.warmup
    streams = frames(cfg, B, n)
    import torch
    from disinfect_slam_b200 import tsdf_grid
    dev = torch.device("cuda", 0)
    H, Wd = cfg.height, cfg.width
    cam = tsdf_grid.CameraParams(streams[0]["K"], H, Wd)
    dres = [{k: torch.from_numpy(np.stack(st[k])).to(dev) for k in ("rgb", "depth", "ht", "lt")} for st in streams]
    out = [dict(rgba=torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev), normal=torch.empty((H, Wd, 4), dtype=torch.uint8, device=dev),
                depth=torch.empty((H, Wd), dtype=torch.float32, device=dev)) for _ in range(B)]
    engs = [tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots, max_image_pixels=H * Wd)
            for _ in range(B)]
    dig = hashlib.blake2b(digest_size=8)

    def step(i, digest=False):
        for b, g in enumerate(engs):
            st, d = streams[b], dres[b]
            pose = (st["q"][i], st["t"][i])
            g.IntegrateDevice(d["rgb"][i].data_ptr(), d["depth"][i].data_ptr(), d["ht"][i].data_ptr(), d["lt"][i].data_ptr(), Wd, H, cfg.max_depth,
                              st["K"], pose)
            g.RayCastDevice(cfg.max_depth, cam, pose, out[b]["rgba"].data_ptr(), out[b]["normal"].data_ptr(), out[b]["depth"].data_ptr())
            g.synchronize()
            if digest:
                for k in ("rgba", "normal", "depth"):
                    dig.update(out[b][k].cpu().numpy().tobytes())

    for i in range(W):
        step(i)
    for g in engs:
        g.set_profiling(True)
    for i in range(W, n):
        step(i, digest=(i % 7 == 0))
    ms, cnt = {}, {}
    for g in engs:
        m, c = g.phase_ms()
        for k in m:
            ms[k] = ms.get(k, 0.0) + m[k]
            cnt[k] = cnt.get(k, 0) + c[k]
    keys, tsdf, rgbw, prob = engs[0].export()
    vol = hashlib.blake2b(keys.tobytes() + tsdf.tobytes() + rgbw.tobytes(), digest_size=8).hexdigest()
    res = {"tag": args.tag, "lib": os.environ.get("TSDF_B200_LIB", "default"),
           "us": {k: round(1e3 * ms[k] / max(cnt[k], 1), 2) for k in ("allocate", "select", "integrate", "raycast")},
           "views_digest": dig.hexdigest(), "volume_digest": vol, "blocks": int(len(keys))}
    print(json.dumps(res), flush=True)
    for g in engs:
        g.close()


if __name__ == "__main__":
    main()
