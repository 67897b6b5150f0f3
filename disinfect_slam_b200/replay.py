"""On-disk formats of the reference's offline TSDF replay, so recorded sequences can
drive the engine and volumes can be dumped the way the reference does:

  <logdir>/trajectory.txt         one line per frame: `id` + 12 floats = row-major 3x4 cam_T_world
                                  (examples/tsdf/offline.cc:45-64)
  <logdir>/<id>_rgb.png           8-bit colour, BGR on disk as cv::imwrite leaves it (offline.cc:72,163)
  <logdir>/<id>_depth.png         16-bit, metres = value / depthmap_factor (offline.cc:73,77)
  <logdir>/<id>_ht.png, _no_ht.png  16-bit probabilities, value / 65535; absent -> ht = 0, lt = 1 (offline.cc:74-83)
  /tmp/data.bin-style dump        raw VoxelSpatialTSDF records {float x, y, z, tsdf} (offline.cc:184-190)

Pure host code (numpy + cv2 for PNG I/O).  The pixel conversions mirror cv::Mat::convertTo (double scale factor,
result rounded to float32); the rotation-matrix -> quaternion conversion mirrors Eigen 3.3's
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


def read_trajectory(logdir):
    """[(id, q_xyzw float32[4], t float32[3])] from trajectory.txt."""
    out = []
    with open(os.path.join(logdir, "trajectory.txt")) as fh:
        for line in fh:
            v = line.split()
            if len(v) < 13:
                continue
            m = np.array([float(s) for s in v[1:13]], np.float32).reshape(3, 4)
            out.append((int(v[0]), quat_from_rotation(m[:, :3]), m[:, 3].copy()))
    return out


def _convert(img_u16, scale):
    """cv::Mat::convertTo(CV_32FC1, scale): double multiply, rounded to float32."""
    return (img_u16.astype(np.float64) * scale).astype(np.float32)


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


def read_log(logdir, depthmap_factor):
    """Iterate the frames of a reference-format log: dicts with rgb, depth, ht, lt, q, t, id."""
    for frame_id, q, t in read_trajectory(logdir):
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


def replay(grid, logdir, depthmap_factor, intrinsics, max_depth=4.0, limit=None):
    """Integrate every frame of a log into `grid` (a TSDFGrid); returns the number of frames integrated."""
    n = 0
    for f in read_log(logdir, depthmap_factor):
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
