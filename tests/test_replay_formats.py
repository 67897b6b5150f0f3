"""CPU tests of the reference's on-disk formats (disinfect_slam_b200/replay.py): log round trip, pixel
conversion rules, rotation-matrix -> quaternion conversion, TSDF dump records; plus a GPU test that a log
replayed through the engine equals the oracle fed with the same decoded frames."""
import numpy as np
import pytest

from disinfect_slam_b200 import replay, synth
from oracle.oracle import Oracle


def test_quaternion_from_rotation_all_branches():
    rng = np.random.RandomState(5)
    qs = [np.array([0, 0, 0, 1.0]), np.array([1.0, 0, 0, 0.001]), np.array([0, 1.0, 0, 0.001]), np.array([0, 0.001, 1.0, 0])]
    qs += [rng.randn(4) for _ in range(50)]
    for q in qs:
        q = q / np.linalg.norm(q)
        R = replay.rotation_from_quat(q)
        q2 = replay.quat_from_rotation(R).astype(np.float64)
        assert abs(np.linalg.norm(q2) - 1) < 1e-5
        assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 2e-6, (q, q2)


def test_log_round_trip(tmp_path):
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    frames = [sc.frame(i) for i in range(3)]
    replay.write_log(str(tmp_path), frames, cfg.depth_factor)
    got = list(replay.read_log(str(tmp_path), cfg.depth_factor))
    assert [g["id"] for g in got] == [0, 1, 2]
    for f, g in zip(frames, got):
        assert np.array_equal(g["rgb"], f["rgb"])
        # synthetic depth is already quantised to 1 / depth_factor: the 16-bit round trip is the convertTo rule
        d16 = np.rint(f["depth"].astype(np.float64) * cfg.depth_factor)
        assert np.array_equal(g["depth"], d16.astype(np.float32) * np.float32(1.0 / cfg.depth_factor))
        assert np.abs(g["depth"] - f["depth"]).max() < 1e-6
        assert np.abs(g["ht"] - f["ht"]).max() <= 0.5 / 65535 + 1e-7 and g["ht"].dtype == np.float32
        assert min(np.abs(g["q"] - f["q"]).max(), np.abs(g["q"] + f["q"]).max()) < 1e-6 and np.abs(g["t"] - f["t"]).max() < 1e-6
    # a log without probability images: ht = 0, lt = 1 (examples/tsdf/offline.cc:80-82)
    for f in frames:
        f["ht"] = f["lt"] = None
    replay.write_log(str(tmp_path / "nop"), frames, cfg.depth_factor)
    g = next(replay.read_log(str(tmp_path / "nop"), cfg.depth_factor))
    assert (g["ht"] == 0).all() and (g["lt"] == 1).all()


# `Extrinsics` of configs/zed_native_l515.yaml:19-24 (l515 -> zed, about 4 cm and 2 degrees)
L515_EXTRINSICS = [0.9995, -0.0097, 0.0289, -37.172e-3, 0.009, 0.9997, 0.0247, -22.6717e-3,
                   -0.0291, -0.0245, 0.9993, 2.4699e-3, 0, 0, 0, 1]


def test_trajectory_extrinsics_premultiply(tmp_path):
    """read_trajectory(extrinsics) = extrinsics * SE3(row) (examples/tsdf/offline.cc:36-62), checked against the
    float64 matrix product; identity / None leave the poses unchanged."""
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    frames = [sc.frame(i) for i in range(4)]
    replay.write_log(str(tmp_path), frames, cfg.depth_factor)
    plain = replay.read_trajectory(str(tmp_path))
    ident = replay.read_trajectory(str(tmp_path), np.eye(4))
    ext = replay.read_trajectory(str(tmp_path), L515_EXTRINSICS)
    E = np.array(L515_EXTRINSICS, np.float64).reshape(4, 4)
    # the YAML matrix is only approximately a rotation: SE3(Matrix4f) goes through a quaternion, which normalises nothing
    # but drops the non-rotation part, so compare against the rotation that quaternion stands for
    qe, te = replay.se3_from_matrix(E)
    Re = replay.rotation_from_quat(qe.astype(np.float64) / np.linalg.norm(qe.astype(np.float64)))
    moved = 0.0
    for (i0, q0, t0), (i1, q1, t1), (i2, q2, t2) in zip(plain, ident, ext):
        assert i0 == i1 == i2
        assert min(np.abs(q1 - q0).max(), np.abs(q1 + q0).max()) < 1e-6 and np.abs(t1 - t0).max() < 1e-6
        R0 = replay.rotation_from_quat(q0.astype(np.float64))
        R_want, t_want = Re @ R0, Re @ t0.astype(np.float64) + te.astype(np.float64)
        # Eigen normalises nothing: the product carries the (1 + 1e-5) norm of the YAML quaternion; compare directions
        R_got = replay.rotation_from_quat(q2.astype(np.float64) / np.linalg.norm(q2.astype(np.float64)))
        assert abs(np.linalg.norm(q2.astype(np.float64)) - np.linalg.norm(qe.astype(np.float64))) < 1e-6
        assert np.abs(R_got - R_want).max() < 5e-6, np.abs(R_got - R_want).max()
        assert np.abs(t2 - t_want).max() < 5e-6
        moved = max(moved, np.abs(t2 - t0).max())
    assert moved > 0.01  # the offset of a few centimetres really is applied
    # frames carry the composed pose
    f = next(replay.read_log(str(tmp_path), cfg.depth_factor, L515_EXTRINSICS))
    assert np.array_equal(f["q"], ext[0][1]) and np.array_equal(f["t"], ext[0][2])


def test_convert_rule_equals_opencv_on_every_16_bit_value():
    """cv::Mat::convertTo(CV_32FC1, 1. / scale) for CV_16UC1 (examples/tsdf/offline.cc:77-80) multiplies in float32; cv2
    exposes the same cvtScale kernel through normalize(NORM_MINMAX), whose scale factor is known in closed form."""
    import cv2
    src = np.arange(65536, dtype=np.uint16).reshape(256, 256)
    beta = 65535.0 / 5000.0
    got = cv2.normalize(src, None, alpha=0.0, beta=beta, norm_type=cv2.NORM_MINMAX, dtype=cv2.CV_32F)
    assert np.array_equal(got, replay._convert(src, beta / 65535.0))
    assert not np.array_equal(got, (src.astype(np.float64) * (beta / 65535.0)).astype(np.float32))  # the double rule is NOT it


def test_tsdf_dump_records(tmp_path):
    cfg = synth.config("tiny")
    f = synth.Scene(cfg).frame(0)
    o = Oracle(cfg.voxel_size, cfg.truncation)
    o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, f["K"], f["q"], f["t"])
    rec = o.gather()
    p = tmp_path / "data.bin"
    replay.save_tsdf_dump(str(p), rec)
    assert p.stat().st_size == 16 * len(rec)
    assert np.array_equal(replay.load_tsdf_dump(str(p)).view(np.uint32), rec.view(np.uint32))


@pytest.mark.gpu
def test_replayed_log_matches_oracle(tmp_path, tsdf_lib):
    from disinfect_slam_b200 import tsdf_grid
    from oracle import compare
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    replay.write_log(str(tmp_path), [sc.frame(i) for i in range(4)], cfg.depth_factor)
    g = tsdf_grid.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)
    assert replay.replay(g, str(tmp_path), cfg.depth_factor, cfg.K, cfg.max_depth) == 4
    o = Oracle(cfg.voxel_size, cfg.truncation)
    for f in replay.read_log(str(tmp_path), cfg.depth_factor):
        o.integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, np.array(cfg.K, np.float32), f["q"], f["t"])
    assert compare.compare_volumes(g.export(), o.export(), "replayed log")["tsdf_bit_exact"]
    g.close()


def test_mesh_welding_and_ply(tmp_path):
    """ExtractMesh's triangle soup -> indexed mesh (what shape_msgs/Mesh carries, ros_offline.cc:296-312): exact welding of the
    bit-identical shared vertices; Euler characteristic 2 for the closed sphere of tests/test_mesh.py."""
    from oracle import mesh_oracle
    keys, tsdf, rgbw, vs = mesh_oracle.sphere_volume()
    tris = mesh_oracle.extract_mesh(keys, tsdf, rgbw, vs)
    verts, idx = replay.weld_mesh(tris)
    edges = np.unique(np.sort(np.concatenate([idx[:, [0, 1]], idx[:, [1, 2]], idx[:, [2, 0]]]), 1), axis=0)
    assert len(verts) - len(edges) + len(idx) == 2            # V - E + F of a sphere
    assert len(idx) <= len(tris) and np.isin(verts[idx].reshape(-1, 9).view(np.uint32), tris.reshape(-1, 9).view(np.uint32)).all()
    nv, nf = replay.save_ply(tmp_path / "m.ply", tris)
    raw = open(tmp_path / "m.ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert f"element vertex {nv}".encode() in head and f"element face {nf}".encode() in head
    assert len(body) == nv * 12 + nf * 13
    assert np.array_equal(np.frombuffer(body[:nv * 12], "<f4").reshape(nv, 3), verts)
