#!/usr/bin/env python
"""Generates tests/golden/ref_tiny.npz from the REFERENCE's own CUDA TSDFGrid (oracle/_ref/
libref_tsdf_parity.so = /root/reference/utils/tsdf/*.cu rebuilt for sm_100a with -fmad=false, see
oracle/build_ref.sh).  Needs a GPU:

    gpurun -- 'python tests/golden/make_ref_golden.py gpurun_out/ref_tiny.npz'
    cp gpurun_out/ref_tiny.npz tests/golden/

The fixture pins the CPU oracle (and through it the engine) to outputs of the reference itself:
per frame the reference's block-coordinate set with a 64-bit digest of every block's TSDF and RGBW
planes (bit-exact quantities), a sample of semantic probabilities, and -- after frame 0 -- the
reference's RayCast images and GatherValid / GatherVoxels digests on its own volume.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from disinfect_slam_b200 import synth  # noqa: E402

CONFIG, BBOX = "tiny", (-1.0, 1.5, -1.2, 0.9, -2.5, 0.4)
VIRTUAL = dict(width=96, height=64, K=(80.0, 80.0, 47.5, 31.5))
# Frame sequence: four frames of the moving camera, then the camera dwells on three neighbouring poses (60 frames) long enough for
# the weight clamp `min(roundf(w), 40)` (voxel_tsdf.cu:192) and dozens of colour re-quantisations to act.  Full
# per-block digests are stored at the SNAPSHOT frames only; the block-coordinate set is stored for every frame (as
# a delta), because a block the reference allocated late is "don't care" from that frame on.
SEQUENCE = [0, 1, 2, 3] + [i % 3 for i in range(4, 64)]
SNAPSHOTS = (0, 1, 2, 3, 15, 31, 47, 63)
N_FRAMES = len(SEQUENCE)


def key_rows(keys):
    return set(map(tuple, np.asarray(keys).reshape(-1, 3).tolist()))


def rows_array(rows):
    return np.array(sorted(rows), np.int16).reshape(-1, 3)


def digest_rows(a):
    """uint64 digest per leading row of a C-contiguous array (bytes of the row)."""
    a = np.ascontiguousarray(a)
    return np.array([int.from_bytes(hashlib.blake2b(a[i].tobytes(), digest_size=8).digest(), "little") for i in range(len(a))],
                    dtype=np.uint64)


def gather_digest(g):
    """Order-independent digest of a gather result: blocks sorted canonically, then hashed."""
    from oracle.compare import canonical_gather
    c = canonical_gather(g)
    return np.frombuffer(hashlib.blake2b(c.tobytes(), digest_size=16).digest(), dtype=np.uint64).copy(), len(c)


def main(out):
    from oracle.ref_cuda import RefTSDFGrid
    cfg = synth.config(CONFIG)
    sc = synth.Scene(cfg)
    r = RefTSDFGrid(cfg.voxel_size, cfg.truncation, parity=True)
    d = {"config": np.array(CONFIG), "n_frames": np.array(N_FRAMES), "bbox": np.array(BBOX, np.float32)}
    d["sequence"] = np.array(SEQUENCE, np.int32)
    d["snapshots"] = np.array(SNAPSHOTS, np.int32)
    rng = np.random.RandomState(7)
    prev = set()
    for i in range(N_FRAMES):
        f = sc.frame(SEQUENCE[i])
        r.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
        now = key_rows(r.export(voxels=False)[0])
        d[f"added_{i}"], d[f"removed_{i}"] = rows_array(now - prev), rows_array(prev - now)
        prev = now
        if i not in SNAPSHOTS:
            continue
        keys, tsdf, rgbw, prob = r.export()
        d[f"keys_{i}"] = keys
        d[f"tsdf_digest_{i}"] = digest_rows(tsdf)
        # colour of never-updated voxels (weight 0) is stale pool memory in the reference (voxel_mem.cu:48 resets
        # only the weight): digest it as 0, which is what the oracle and the engine define
        rgbw = rgbw.copy()
        rgbw[rgbw[..., 3] == 0] = 0
        d[f"rgbw_digest_{i}"] = digest_rows(rgbw)
        d[f"weight_sum_{i}"] = rgbw[..., 3].astype(np.int64).sum(1)
        d[f"weight40_frac_{i}"] = np.array((rgbw[..., 3] == 40).sum() / max((rgbw[..., 3] > 0).sum(), 1))
        sb = rng.randint(0, len(keys), 4000)
        sv = rng.randint(0, 512, 4000)
        d[f"prob_sample_idx_{i}"] = np.stack([sb, sv], 1).astype(np.int32)
        d[f"prob_sample_{i}"] = prob[sb, sv]
        d[f"tsdf_sample_{i}"] = tsdf[sb, sv]
        if i == 0:
            rgba, normal = r.raycast(cfg.max_depth, cfg.width, cfg.height, f["K"], f["q"], f["t"])
            d["raycast_rgba_0"], d["raycast_normal_0"] = rgba, normal
            rgba, normal = r.raycast(10.0, cfg.width, cfg.height, f["K"], f["q"], f["t"])
            d["raycast10_rgba_0"], d["raycast10_normal_0"] = rgba, normal
            v = sc.virtual_view(1, 4, **VIRTUAL)
            rgba, normal = r.raycast(cfg.max_depth, v["width"], v["height"], v["K"], v["q"], v["t"])
            d["raycast_virtual_rgba_0"], d["raycast_virtual_normal_0"] = rgba, normal
            d["gather_valid_digest_0"], n = gather_digest(r.gather())
            d["gather_valid_n_0"] = np.array(n)
            d["gather_bound_digest_0"], n = gather_digest(r.gather(BBOX))
            d["gather_bound_n_0"] = np.array(n)
            d["num_active_0"] = np.array(r.num_active())
    np.savez_compressed(out, **d)
    print("wrote", out, os.path.getsize(out), "bytes;", {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items() if k.startswith("keys")},
          {k: float(v) for k, v in d.items() if k.startswith("weight40")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "ref_tiny.npz"))
