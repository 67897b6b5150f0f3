"""On-disk formats of the reference's offline TSDF replay, so recorded sequences can
drive the engine and volumes can be dumped the way the reference does:

  <logdir>/trajectory.txt         one line per frame: `id` + 12 floats = row-major 3x4 posecam_T_world, left-multiplied
                                  by the config's `Extrinsics` to give cam_T_world (examples/tsdf/offline.cc:36-64)
  <logdir>/<id>_rgb.png           8-bit colour, BGR on disk as cv::imwrite leaves it (offline.cc:72,163)
  <logdir>/<id>_depth.png         16-bit, metres = value / depthmap_factor (offline.cc:73,77)
  <logdir>/<id>_ht.png, _no_ht.png  16-bit probabilities, value / 65535; absent -> ht = 0, lt = 1 (offline.cc:74-83)
  /tmp/data.bin-style dump        raw VoxelSpatialTSDF records {float x, y, z, tsdf} (offline.cc:184-190)

Pure host code (numpy + cv2 for PNG I/O).  The pixel conversions mirror cv::Mat::convertTo of a 16-bit image (float32
pixel times the scale factor cast to float32); the rotation-matrix -> quaternion conversion mirrors Eigen 3.3's
QuaternionBase::operator=(Matrix3) in float32, which is what SE3<float>(Matrix<float,3,4>) runs
(utils/cuda/lie_group.cuh:19-20).
"""
import os

import numpy as np


def quat_from_rotation(R):
    """Eigen 3.3 quaternion-from-matrix, float32 arithmetic; returns (x, y, z, w)."""
    m = np.asarray(R, np.float32).reshape(3, 3)
    f = np.float32
    t = f(m[0, 0] + m[1, 1]) + m[2, 2]
    q = np.zeros(4, np.float32)  # x, y, z, w
    if t > f(0):
        t = np.sqrt(f(t + f(1.0)))
        q[3] = f(0.5) * t
        t = f(0.5) / t
        q[0] = f(m[2, 1] - m[1, 2]) * t
        q[1] = f(m[0, 2] - m[2, 0]) * t
        q[2] = f(m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(f(f(f(m[i, i] - m[j, j]) - m[k, k]) + f(1.0)))
        q[i] = f(0.5) * t
        t = f(0.5) / t
        q[3] = f(m[k, j] - m[j, k]) * t
        q[j] = f(m[j, i] + m[i, j]) * t
        q[k] = f(m[k, i] + m[i, k]) * t
    return q


def rotation_from_quat(q):
    x, y, z, w = (float(v) for v in q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def _cross(a, b):
    f = np.float32
    return np.array([f(f(a[1] * b[2]) - f(a[2] * b[1])), f(f(a[2] * b[0]) - f(a[0] * b[2])), f(f(a[0] * b[1]) - f(a[1] * b[0]))],
                    np.float32)


def quat_rotate(q, v):
    """Eigen 3.3 QuaternionBase::_transformVector in float32 (q = x, y, z, w): uv = 2 (q.vec x v); v + w uv + q.vec x uv."""
    q, v = np.asarray(q, np.float32), np.asarray(v, np.float32)
    uv = _cross(q[:3], v)
    uv = (uv + uv).astype(np.float32)
    c = _cross(q[:3], uv)
    return ((v + (q[3] * uv).astype(np.float32)).astype(np.float32) + c).astype(np.float32)


def quat_multiply(a, b):
    """Eigen 3.3 quaternion product a * b, generic (scalar) path, float32, operands as (x, y, z, w)."""
    f = np.float32
    ax, ay, az, aw = (f(v) for v in a)
    bx, by, bz, bw = (f(v) for v in b)
    w = f(f(f(f(aw * bw) - f(ax * bx)) - f(ay * by)) - f(az * bz))
    x = f(f(f(f(aw * bx) + f(ax * bw)) + f(ay * bz)) - f(az * by))
    y = f(f(f(f(aw * by) + f(ay * bw)) + f(az * bx)) - f(ax * bz))
    z = f(f(f(f(aw * bz) + f(az * bw)) + f(ax * by)) - f(ay * bx))
    return np.array([x, y, z, w], np.float32)


def se3_compose(a, b):
    """SE3<float>::operator* (utils/cuda/lie_group.cuh:38-40): (Ra Rb, Ra tb + ta); poses as (q_xyzw, t)."""
    (qa, ta), (qb, tb) = a, b
    return quat_multiply(qa, qb), (quat_rotate(qa, tb) + np.asarray(ta, np.float32)).astype(np.float32)


def se3_from_matrix(m):
    """SE3<float>(Matrix4f / Matrix<float, 3, 4>) (lie_group.cuh:16-20): rotation block -> quaternion, last column."""
    m = np.asarray(m, np.float32).reshape(-1, 4)
    return quat_from_rotation(m[:3, :3]), m[:3, 3].copy()


def read_trajectory(logdir, extrinsics=None):
    """[(id, q_xyzw float32[4], t float32[3])] = cam_T_world of every frame of trajectory.txt.

    The file stores posecam_T_world (the tracking camera); the reference turns it into the depth camera's pose with
    the config's `Extrinsics` (row-major 4x4, depthcam_T_posecam): `extrinsics * SE3<float>(tmp)`
    (examples/tsdf/offline.cc:36-62).  `extrinsics` = that 4x4 (any shape with 16 values) or None for identity, which is
    what get_extrinsics returns when the key is absent.  All four shipped L515 configs carry a non-identity one."""
    ext = None if extrinsics is None else se3_from_matrix(np.asarray(extrinsics, np.float32).reshape(4, 4))
    out = []
    with open(os.path.join(logdir, "trajectory.txt")) as fh:
        for line in fh:
            v = line.split()
            if len(v) < 13:
                continue
            q, t = se3_from_matrix(np.array([float(s) for s in v[1:13]], np.float32).reshape(3, 4))
            if ext is not None:
                q, t = se3_compose(ext, (q, t))
            out.append((int(v[0]), q, t))
    return out


def convert_scale(scale):
    """The factor cv::Mat::convertTo(CV_32FC1, scale) really multiplies by: for a 16-bit source OpenCV's cvtScale works
    in float32 -- the double `scale` is cast to float once (modules/core/src/convert_scale.simd.hpp, wtype = float;
    verified against cv2 4.13 on all 65536 inputs, tests/test_replay_formats.py)."""
    return np.float32(scale)


def _convert(img_u16, scale):
    """cv::Mat::convertTo(CV_32FC1, scale) of a CV_16UC1 image: float32(pixel) * float32(scale), one rounding."""
    return img_u16.astype(np.float32) * convert_scale(scale)


def read_frame(logdir, frame_id, depthmap_factor):
    """One frame as the reference's get_images_by_id + cvtColor(BGR2RGB) produce it."""
    import cv2
    p = lambda suffix: os.path.join(logdir, f"{frame_id}_{suffix}.png")  # noqa: E731
    bgr = cv2.imread(p("rgb"))
    depth_raw = cv2.imread(p("depth"), cv2.IMREAD_UNCHANGED)
    if bgr is None or depth_raw is None:
        raise FileNotFoundError(f"frame {frame_id}: rgb / depth png missing in {logdir}")
    rgb = np.ascontiguousarray(bgr[..., ::-1])
    depth = _convert(depth_raw, 1.0 / depthmap_factor)
    ht_raw = cv2.imread(p("ht"), cv2.IMREAD_UNCHANGED) if os.path.exists(p("ht")) else None
    lt_raw = cv2.imread(p("no_ht"), cv2.IMREAD_UNCHANGED) if os.path.exists(p("no_ht")) else None
    if ht_raw is not None and lt_raw is not None:
        ht, lt = _convert(ht_raw, 1.0 / 65535), _convert(lt_raw, 1.0 / 65535)
    else:
        ht, lt = np.zeros_like(depth), np.ones_like(depth)
    return dict(rgb=rgb, depth=depth, ht=ht, lt=lt)


def read_log(logdir, depthmap_factor, extrinsics=None):
    """Iterate the frames of a reference-format log: dicts with rgb, depth, ht, lt, q, t, id."""
    for frame_id, q, t in read_trajectory(logdir, extrinsics):
        f = read_frame(logdir, frame_id, depthmap_factor)
        f.update(q=q, t=t, id=frame_id)
        yield f


def write_log(logdir, frames, depthmap_factor):
    """Write frames (dicts with rgb, depth, ht, lt, q, t) in the reference's log format.  Depth and probabilities
    are quantised to 16 bits, exactly as a recorded sequence is."""
    import cv2
    os.makedirs(logdir, exist_ok=True)
    with open(os.path.join(logdir, "trajectory.txt"), "w") as fh:
        for i, f in enumerate(frames):
            R = rotation_from_quat(f["q"])
            m = np.concatenate([R, np.asarray(f["t"], np.float64).reshape(3, 1)], 1)
            fh.write(f"{i} " + " ".join(repr(float(v)) for v in m.reshape(-1)) + "\n")
            cv2.imwrite(os.path.join(logdir, f"{i}_rgb.png"), np.ascontiguousarray(f["rgb"][..., ::-1]))
            d16 = np.clip(np.rint(f["depth"].astype(np.float64) * depthmap_factor), 0, 65535).astype(np.uint16)
            cv2.imwrite(os.path.join(logdir, f"{i}_depth.png"), d16)
            for key, suffix in (("ht", "ht"), ("lt", "no_ht")):
                if f.get(key) is not None:
                    p16 = np.clip(np.rint(f[key].astype(np.float64) * 65535), 0, 65535).astype(np.uint16)
                    cv2.imwrite(os.path.join(logdir, f"{i}_{suffix}.png"), p16)


def replay(grid, logdir, depthmap_factor, intrinsics, max_depth=4.0, limit=None, extrinsics=None):
    """Integrate every frame of a log into `grid` (a TSDFGrid); returns the number of frames integrated.
    `extrinsics`: the config's `Extrinsics` 4x4 (see read_trajectory)."""
    n = 0
    for f in read_log(logdir, depthmap_factor, extrinsics):
        grid.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], max_depth, intrinsics, (f["q"], f["t"]))
        n += 1
        if limit is not None and n >= limit:
            break
    return n


def save_tsdf_dump(path, records):
    """The "Save TSDF" dump: raw {x, y, z, tsdf} float32 records (examples/tsdf/offline.cc:184-190)."""
    np.ascontiguousarray(records, np.float32).reshape(-1, 4).tofile(path)


def load_tsdf_dump(path):
    return np.fromfile(path, np.float32).reshape(-1, 4)


# ---- mesh export (the ROS nodes publish shape_msgs/Mesh: vertices + triangle indices,
# examples/ros_camera_driver/ros_offline.cc:296-312) ----------------------------------------------------------
def weld_mesh(tris):
    """Triangle soup float32 [n, 3, 3] (TSDFGrid.ExtractMesh) -> (vertices float32 [m, 3], triangles int32 [n, 3]).
    Vertices shared by neighbouring cells are bit-identical by construction, so welding is an exact de-duplication;
    degenerate triangles (a crossing exactly at a voxel centre) are dropped."""
    t = np.ascontiguousarray(tris, np.float32).reshape(-1, 3, 3)
    verts, inv = np.unique(t.reshape(-1, 3).view(np.uint32), axis=0, return_inverse=True)
    idx = inv.reshape(-1, 3).astype(np.int32)
    keep = (idx[:, 0] != idx[:, 1]) & (idx[:, 1] != idx[:, 2]) & (idx[:, 0] != idx[:, 2])
    return verts.view(np.float32), idx[keep]


def save_ply(path, tris):
    """Binary little-endian PLY of the welded mesh."""
    verts, idx = weld_mesh(tris)
    with open(path, "wb") as f:
        f.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(verts)}\nproperty float x\nproperty float y\n"
                 f"property float z\nelement face {len(idx)}\nproperty list uchar int vertex_indices\nend_header\n").encode())
        f.write(np.ascontiguousarray(verts, "<f4").tobytes())
        rec = np.empty(len(idx), dtype=[("n", "u1"), ("v", "<i4", 3)])
        rec["n"], rec["v"] = 3, idx
        f.write(rec.tobytes())
    return len(verts), len(idx)
