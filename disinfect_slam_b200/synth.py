"""Deterministic synthetic RGB-D + semantic-probability sequences.

Analytic closed room (axis-aligned box, camera inside) with spheres and boxes so that every
pixel has a closed-form depth; camera on a smooth closed trajectory; RGB and the ht/lt
probability planes are procedural functions of the world hit point.  All geometry is evaluated
in float64 and only the final planes / poses are cast to the types the TSDFGrid API takes
(rgb uint8 HxWx3, depth/ht/lt float32 HxW in metres / probabilities, pose as unit quaternion
(x,y,z,w) + translation of cam_T_world, utils/cuda/lie_group.cuh:43-44).

Pure host code (numpy); used by tests/, bench.py and the examples so that the oracle, the
reference rebuild and the engine all consume identical bytes.
"""
from dataclasses import dataclass, field, replace

import numpy as np

# configs/TUM_RGBD_rgbd_1.yaml:11-14,44 and configs/zed_native_l515.yaml:27-31 of the reference
TUM_K = (517.306408, 516.469215, 318.643040, 255.313989)
L515_HALF_K = (456.8617, 456.77127, 322.1042, 187.79485)


@dataclass(frozen=True)
class SceneConfig:
    name: str
    width: int
    height: int
    K: tuple  # fx, fy, cx, cy
    voxel_size: float
    truncation: float
    max_depth: float
    n_frames: int
    room_half: tuple = (2.0, 1.5, 2.0)  # half extents in metres (x, y-up, z)
    depth_factor: float = 5000.0
    traj_radius: float = 0.8
    seed: int = 0xD15F
    invalid_frac: float = 0.02
    n_objects: int = 6
    pool_blocks: int = 1 << 18
    table_slots: int = 1 << 21
    extra: dict = field(default_factory=dict)

    def scaled(self, s, name=None):
        """Same scene at image scale s (intrinsics scale with the image)."""
        fx, fy, cx, cy = self.K
        return replace(self, name=name or f"{self.name}_x{s:g}", width=int(round(self.width * s)),
                       height=int(round(self.height * s)), K=(fx * s, fy * s, cx * s, cy * s))


def config(name):
    """The BASELINE.json configs plus small variants for CPU-sized tests."""
    c1 = SceneConfig("config1_tum_640x480_1cm", 640, 480, TUM_K, 0.01, 0.06, 4.0, 100)
    k2 = tuple(2.0 * v for v in L515_HALF_K)
    c2 = SceneConfig("config2_l515_1280x720_5mm", 1280, 720, k2, 0.005, 0.03, 4.0, 100, depth_factor=4000.0,
                     pool_blocks=1 << 18, table_slots=1 << 20)  # the reference's NUM_BLOCK; a lap peaks near 70 k blocks
    c2h = SceneConfig("config2_l515_640x360_5mm", 640, 360, L515_HALF_K, 0.005, 0.03, 4.0, 100, depth_factor=4000.0,
                      pool_blocks=1 << 20, table_slots=1 << 23)
    c3 = SceneConfig("config3_room_1280x720_2cm", 1280, 720, k2, 0.02, 0.12, 4.0, 200, room_half=(8.0, 3.0, 8.0),
                     depth_factor=4000.0, traj_radius=5.0, n_objects=24, extra={"look": "out"})
    c4 = SceneConfig("config4_raycast_1920x1080", 1920, 1080, (1400.0, 1400.0, 959.5, 539.5), 0.01, 0.06, 4.0, 32)
    tiny = SceneConfig("tiny_160x120_2cm", 160, 120, tuple(v * 0.25 for v in TUM_K), 0.02, 0.12, 4.0, 12,
                       pool_blocks=1 << 15, table_slots=1 << 18)
    small = SceneConfig("small_320x240_1cm", 320, 240, tuple(v * 0.5 for v in TUM_K), 0.01, 0.06, 4.0, 10,
                        pool_blocks=1 << 17, table_slots=1 << 20)
    table = {c.name: c for c in (c1, c2, c2h, c3, c4, tiny, small)}
    alias = {"config1": c1, "config2": c2, "config2_half": c2h, "config3": c3, "config4": c4, "tiny": tiny,
             "small": small}
    return table.get(name) or alias[name]


def _hash32(a):
    """Counter-based PRNG (32-bit mix) on uint32 arrays."""
    a = np.asarray(a, dtype=np.uint32).copy()
    a ^= a >> np.uint32(16)
    a *= np.uint32(0x7FEB352D)
    a ^= a >> np.uint32(15)
    a *= np.uint32(0x846CA68B)
    a ^= a >> np.uint32(16)
    return a


class Scene:
    def __init__(self, cfg: SceneConfig):
        self.cfg = cfg
        rng = np.random.RandomState(cfg.seed & 0x7FFFFFFF)
        hx, hy, hz = cfg.room_half
        self.half = np.array([hx, hy, hz], np.float64)
        n_s = cfg.n_objects // 2
        n_b = cfg.n_objects - n_s
        # objects sit on/near the floor, away from the camera ring
        self.sph_c = np.stack([rng.uniform(-0.8 * hx, 0.8 * hx, n_s), rng.uniform(-hy, -0.3 * hy, n_s),
                               rng.uniform(-0.8 * hz, 0.8 * hz, n_s)], 1)
        self.sph_r = rng.uniform(0.12, 0.35, n_s)
        bc = np.stack([rng.uniform(-0.85 * hx, 0.85 * hx, n_b), rng.uniform(-hy, -0.2 * hy, n_b),
                       rng.uniform(-0.85 * hz, 0.85 * hz, n_b)], 1)
        bh = rng.uniform(0.1, 0.3, (n_b, 3))
        self.box_lo, self.box_hi = bc - bh, bc + bh

    # ---- camera ---------------------------------------------------------------------------
    def pose(self, i):
        """cam_T_world for frame i as (q_xyzw float32[4], t float32[3]) plus float64 (R_wc, c)."""
        cfg = self.cfg
        a = 2.0 * np.pi * i / max(cfg.n_frames, 1)
        r = cfg.traj_radius
        c = np.array([r * np.cos(a), 0.15 * np.sin(2 * a), r * np.sin(a)])
        # look across the room centre towards the far wall, with a slow pitch oscillation
        target = np.array([-1.5 * r * np.cos(a + 0.3), -0.3 + 0.25 * np.sin(3 * a), -1.5 * r * np.sin(a + 0.3)])
        if cfg.extra.get("look") == "out":
            # halls much larger than max_depth: face the NEAR wall instead (slightly ahead of the direction of travel),
            # so that a lap sweeps every wall within range
            target = np.array([3.0 * r * np.cos(a + 0.25), -0.2 + 1.2 * np.sin(3 * a), 3.0 * r * np.sin(a + 0.25)])
        f = target - c
        f /= np.linalg.norm(f)
        up = np.array([0.0, 1.0, 0.0])
        right = np.cross(f, up)
        right /= np.linalg.norm(right)
        down = np.cross(f, right)
        R_wc = np.stack([right, down, f], 1)  # columns: camera x (right), y (down), z (forward) in world
        R_cw = R_wc.T
        t = -R_cw @ c
        q = _quat_from_R(R_cw)
        return q.astype(np.float32), t.astype(np.float32), R_wc, c

    # ---- rendering -------------------------------------------------------------------------
    def _cast(self, o, d):
        """First hit distance t (in units of d, d.z_cam == 1 so t is z-depth) for rays o + t d."""
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / d
            # room: we are inside, exit distance
            t1 = (-self.half - o) * inv
            t2 = (self.half - o) * inv
            t_room = np.min(np.maximum(t1, t2), axis=-1)
            best = t_room
            # spheres
            for c, r in zip(self.sph_c, self.sph_r):
                oc = o - c
                A = np.sum(d * d, -1)
                B = np.sum(d * oc, -1)
                Cc = np.sum(oc * oc, -1) - r * r
                disc = B * B - A * Cc
                t = (-B - np.sqrt(np.where(disc > 0, disc, np.nan))) / A
                best = np.where((t > 1e-6) & (t < best), t, best)
            # boxes (slab method)
            for lo, hi in zip(self.box_lo, self.box_hi):
                ta = (lo - o) * inv
                tb = (hi - o) * inv
                tn = np.max(np.minimum(ta, tb), -1)
                tf = np.min(np.maximum(ta, tb), -1)
                ok = (tn <= tf) & (tn > 1e-6)
                best = np.where(ok & (tn < best), tn, best)
        return best

    def frame(self, i):
        """Returns dict(rgb, depth, ht, lt, q, t, K) for frame i."""
        cfg = self.cfg
        W, H = cfg.width, cfg.height
        fx, fy, cx, cy = cfg.K
        q, t, R_wc, c = self.pose(i)
        xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        d_cam = np.stack([(xs - cx) / fx, (ys - cy) / fy, np.ones_like(xs)], -1)
        d_w = d_cam @ R_wc.T
        z = self._cast(c[None, None, :], d_w)
        hit = c + d_w * z[..., None]
        # sensor-like quantisation, then float32 metres
        depth = (np.round(z * cfg.depth_factor) / cfg.depth_factor).astype(np.float32)
        depth[depth == np.float32(cfg.max_depth)] = 0.0  # avoid the reference's w == 0 -> 0/0 edge
        idx = (np.arange(W * H, dtype=np.uint32)).reshape(H, W)
        rnd = _hash32(idx * np.uint32(2654435761) + np.uint32((cfg.seed * 7919 + i * 104729) & 0xFFFFFFFF))
        depth[rnd < np.uint32(cfg.invalid_frac * 4294967295.0)] = 0.0
        # procedural colour + semantic fields of the hit point
        chk = ((np.floor(hit[..., 0] * 4) + np.floor(hit[..., 1] * 4) + np.floor(hit[..., 2] * 4)) % 2) * 0.5 + 0.5
        rgb = np.stack([128 + 100 * np.sin(7.0 * hit[..., 0] + 1.0), 128 + 100 * np.sin(5.0 * hit[..., 1] + 2.0),
                        128 + 100 * np.sin(6.0 * hit[..., 2] + 3.0)], -1) * chk[..., None]
        rgb = np.clip(np.round(rgb), 0, 255).astype(np.uint8)
        s = np.sin(2.1 * hit[..., 0] + 0.3) * np.cos(1.7 * hit[..., 2] - 0.5) * np.cos(1.3 * hit[..., 1])
        ht = np.clip(0.5 + 0.45 * s, 0.02, 0.98).astype(np.float32)
        lt = (np.float32(1.0) - ht).astype(np.float32)
        return dict(rgb=np.ascontiguousarray(rgb), depth=np.ascontiguousarray(depth), ht=ht, lt=lt, q=q, t=t,
                    K=np.array(cfg.K, np.float32))

    def virtual_view(self, j, n_views, width=None, height=None, K=None):
        """Pose + intrinsics of the j-th of n_views RayCast views (a second ring, phase shifted)."""
        q, t, _, _ = self.pose((j + 0.5) * self.cfg.n_frames / max(n_views, 1))
        return dict(q=q, t=t, K=np.array(K if K is not None else self.cfg.K, np.float32),
                    width=width or self.cfg.width, height=height or self.cfg.height)


def _quat_from_R(R):
    """Rotation matrix -> unit quaternion (x, y, z, w), float64, w >= 0."""
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        w = 0.25 * s
        x = (R[2, 1] - R[1, 2]) / s
        y = (R[0, 2] - R[2, 0]) / s
        z = (R[1, 0] - R[0, 1]) / s
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        w = (R[2, 1] - R[1, 2]) / s
        x = 0.25 * s
        y = (R[0, 1] + R[1, 0]) / s
        z = (R[0, 2] + R[2, 0]) / s
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        w = (R[0, 2] - R[2, 0]) / s
        x = (R[0, 1] + R[1, 0]) / s
        y = 0.25 * s
        z = (R[1, 2] + R[2, 1]) / s
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        w = (R[1, 0] - R[0, 1]) / s
        x = (R[0, 2] + R[2, 0]) / s
        y = (R[1, 2] + R[2, 1]) / s
        z = 0.25 * s
    q = np.array([x, y, z, w])
    q /= np.linalg.norm(q)
    if q[3] < 0:
        q = -q
    return q
