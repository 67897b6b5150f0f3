"""CPU check of the geometry behind tsdf_shared_cache_attach (csrc/kernels_raycast.cu: pull_select_kernel).

The kernel lists a foreign block for fetching unless all eight corners of its box, grown by `pad` voxels, lie beyond
one of the six planes of the pyramid that contains the launch's rays.  The march falls back to the owner's memory for
a block that was not fetched, so this test is not about correctness of the image (tests/test_gpu_shared_raycast.py has
that) but about the claim that with pad = 3 nothing ever takes the fallback: here the select rule is restated in
float32 numpy and compared with the blocks that the samples of a band of rows can touch -- every sample position the
reference's march can take (voxel_tsdf.cu:243-262: pos += step, max_depth / step samples), rounded to its nearest voxel,
plus the +-1 voxel neighbours of the gradient (voxel_tsdf.cu:277-291)."""
import numpy as np
import pytest

from disinfect_slam_b200 import synth


def quat_rotate(q, v):  # Eigen's _transformVector, float32 (tsdf_device.cuh: qrot); v: (3,) or (n, 3)
    qv = q[:3].astype(np.float32)
    v = np.asarray(v, np.float32)
    uv = np.cross(qv, v).astype(np.float32)
    uv = (uv + uv).astype(np.float32)
    return (v + q[3] * uv + np.cross(qv, uv)).astype(np.float32)


def pose_inverse(q, t):
    qi = np.array([-q[0], -q[1], -q[2], q[3]], np.float32) / np.float32((q * q).sum())
    return qi, quat_rotate(qi, (-t).astype(np.float32))


def selected(blocks, cfg_voxel, K, q, t, w, row0, rows, max_depth, step, pad):
    """pull_select_kernel's verdict for block coordinates `blocks` (n, 3)."""
    fx, fy, cx, cy = (np.float32(v) for v in K)
    kinv = (np.float32(1) / fx, np.float32(1) / fy, -cx / fx, -cy / fy)
    xa, xb = kinv[0] * np.float32(-1) + kinv[2], kinv[0] * np.float32(w) + kinv[2]
    ya, yb = kinv[1] * np.float32(row0 - 1) + kinv[3], kinv[1] * np.float32(row0 + rows) + kinv[3]
    xlo, xhi, ylo, yhi = min(xa, xb), max(xa, xb), min(ya, yb), max(ya, yb)
    zfar = np.float32(max_depth) + np.float32(2) * np.float32(step)
    beyond = np.full(len(blocks), 0x3F, np.uint32)
    for c in range(8):
        off = np.array([(7 + pad) if c & 1 else -pad, (7 + pad) if c & 2 else -pad, (7 + pad) if c & 4 else -pad])
        wpt = ((blocks * 8 + off).astype(np.float32) * np.float32(cfg_voxel)).astype(np.float32)
        p = (quat_rotate(q, wpt) + t).astype(np.float32)  # cam_T_world
        inside = np.zeros(len(blocks), np.uint32)
        inside |= np.where(~(p[:, 0] - xlo * p[:, 2] < 0), 1, 0).astype(np.uint32)
        inside |= np.where(~(p[:, 0] - xhi * p[:, 2] > 0), 2, 0).astype(np.uint32)
        inside |= np.where(~(p[:, 1] - ylo * p[:, 2] < 0), 4, 0).astype(np.uint32)
        inside |= np.where(~(p[:, 1] - yhi * p[:, 2] > 0), 8, 0).astype(np.uint32)
        inside |= np.where(~(p[:, 2] < 0), 16, 0).astype(np.uint32)
        inside |= np.where(~(p[:, 2] > zfar), 32, 0).astype(np.uint32)
        beyond &= ~inside
    return beyond == 0


def touched_blocks(cfg_voxel, K, q, t, w, row0, rows, max_depth, step):
    """Block coordinates of every voxel a sample of rows [row0, row0 + rows) can read (march + gradient neighbours)."""
    fx, fy, cx, cy = (np.float32(v) for v in K)
    qi, ti = pose_inverse(q, t)  # world_T_cam
    ys, xs = np.mgrid[row0:row0 + rows, 0:w]
    pc = np.stack([(xs.astype(np.float32) - cx) / fx, (ys.astype(np.float32) - cy) / fy, np.ones_like(xs, np.float32)], -1).reshape(-1, 3)
    d = (pc / np.sqrt((pc * pc).sum(-1, keepdims=True))).astype(np.float32)
    dw = quat_rotate(qi, d)
    stepv = (dw * np.float32(step) / np.float32(cfg_voxel)).astype(np.float32)
    pos = np.broadcast_to((ti / np.float32(cfg_voxel)).astype(np.float32), stepv.shape).copy()
    out = []
    for _ in range(int(np.ceil(max_depth / step))):
        v = np.rint(pos).astype(np.int64)  # (ties are irrelevant at this granularity: the neighbours below cover them)
        for dx in (-1, 1):
            for ax in range(3):
                vv = v.copy()
                vv[:, ax] += dx
                out.append(np.unique(vv >> 3, axis=0))
        out.append(np.unique(v >> 3, axis=0))
        pos = (pos + stepv).astype(np.float32)
    return np.unique(np.concatenate(out), axis=0)


@pytest.mark.parametrize("frame,world", [(0, 2), (7, 8)])
def test_padded_pyramid_test_lists_every_block_a_band_can_touch(frame, world):
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    f = sc.frame(frame)
    q, t = np.asarray(f["q"], np.float32), np.asarray(f["t"], np.float32)
    step = cfg.truncation / 2
    tiles = (cfg.height + 7) // 8
    per = (tiles + world - 1) // world
    n_checked = 0
    for r in range(world):
        row0 = r * per * 8
        rows = min(per * 8, cfg.height - row0)
        if rows <= 0:
            continue
        touched = touched_blocks(cfg.voxel_size, f["K"], q, t, cfg.width, row0, rows, cfg.max_depth, step)
        sel = selected(touched, cfg.voxel_size, f["K"], q, t, cfg.width, row0, rows, cfg.max_depth, step, pad=3)
        assert sel.all(), f"band {r}: {int((~sel).sum())} of {len(touched)} touched blocks would not be fetched, e.g. {touched[~sel][:3].tolist()}"
        n_checked += len(touched)
        # and the test is not vacuous: far more blocks of a box around the camera are rejected than listed
        box = np.stack(np.meshgrid(*[np.arange(-40, 40, 3)] * 3, indexing="ij"), -1).reshape(-1, 3) + np.rint(pose_inverse(q, t)[1] / cfg.voxel_size / 8).astype(np.int64)
        assert selected(box, cfg.voxel_size, f["K"], q, t, cfg.width, row0, rows, cfg.max_depth, step, pad=3).mean() < 0.5
    assert n_checked > 100
