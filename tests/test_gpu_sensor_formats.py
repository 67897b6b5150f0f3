"""GPU tests of the host-side entry points added for end-to-end throughput:

  * tsdf_integrate_u16: 16-bit depth / probability planes converted inside the allocation kernel with
    cv::Mat::convertTo's arithmetic (examples/tsdf/offline.cc:72-83) -- the volume must be bit-identical to converting
    on the host (disinfect_slam_b200.replay rule, itself pinned to cv2) and calling Integrate, and equal the oracle;
  * no probability planes (ht = lt = NULL) == planes of ones (modules/tsdf_module.cc:28-33);
  * tsdf_integrate_enqueue / tsdf_streams_run: several engines driven by one host thread, nothing waited for inside the
    loop -- same volumes and images as the blocking calls.
"""
import numpy as np
import pytest

from disinfect_slam_b200 import replay, synth
from oracle import compare
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tg(tsdf_lib):
    from disinfect_slam_b200 import tsdf_grid
    return tsdf_grid


def sensor_frame(f, depth_factor):
    d16 = np.clip(np.rint(f["depth"].astype(np.float64) * depth_factor), 0, 65535).astype(np.uint16)
    h16 = np.clip(np.rint(f["ht"].astype(np.float64) * 65535), 0, 65535).astype(np.uint16)
    l16 = np.clip(np.rint(f["lt"].astype(np.float64) * 65535), 0, 65535).astype(np.uint16)
    return d16, h16, l16


def equal_exports(a, b):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))


def test_u16_planes_equal_host_conversion_and_oracle(tg):
    cfg = synth.config("tiny")
    sc = synth.Scene(cfg)
    mk = lambda: tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots)  # noqa: E731
    g16, g32, gone, gnone = mk(), mk(), mk(), mk()
    o = Oracle(cfg.voxel_size, cfg.truncation)
    for i in range(5):
        f = sc.frame(i)
        d16, h16, l16 = sensor_frame(f, cfg.depth_factor)
        flags = (0, 1, 2)[i % 3]  # synchronous, async, no-wait: same volume
        g16.IntegrateU16(f["rgb"], d16, h16, l16, cfg.depth_factor, cfg.max_depth, f["K"], (f["q"], f["t"]), flags=flags)
        if flags == 2:
            g16.synchronize()  # the (pageable) test buffers go out of scope below
        depth = replay._convert(d16, 1.0 / float(np.float32(cfg.depth_factor)))
        ht, lt = replay._convert(h16, 1.0 / 65535), replay._convert(l16, 1.0 / 65535)
        g32.Integrate(f["rgb"], depth, ht, lt, cfg.max_depth, f["K"], (f["q"], f["t"]))
        o.integrate(f["rgb"], depth, ht, lt, cfg.max_depth, f["K"], f["q"], f["t"])
        # without probability planes: identical to planes of ones
        gnone.IntegrateU16(f["rgb"], d16, None, None, cfg.depth_factor, cfg.max_depth, f["K"], (f["q"], f["t"]))
        ones = np.ones_like(depth)
        gone.Integrate(f["rgb"], depth, ones, ones, cfg.max_depth, f["K"], (f["q"], f["t"]))
    g16.synchronize()
    assert equal_exports(g16.export(), g32.export())
    assert equal_exports(gnone.export(), gone.export())
    rep = compare.compare_volumes(g16.export(), o.export(), "u16 planes")
    assert rep["tsdf_bit_exact"] and rep["n_blocks_engine"] > 500
    assert (gnone.export()[3] == 0.5).all()  # ln 1 - ln 1 = 0 at every update: the probability never leaves 0.5
    for g in (g16, g32, gone, gnone):
        g.close()


def test_one_thread_drives_several_streams(tg):
    """tsdf_streams_run (enqueue without waiting + pipelined views) against blocking Integrate + RayCast per stream, for
    float32 and 16-bit frames."""
    cfg = synth.config("tiny")
    H, W = cfg.height, cfg.width
    n_streams, n_frames, count = 3, 4, 7  # 7 steps over 4 frames: the lap repeats
    cam = tg.CameraParams(np.array(cfg.K, np.float32), H, W)
    from dataclasses import replace as dc_replace
    for fmt16 in (False, True):
        scenes = [synth.Scene(dc_replace(cfg, seed=cfg.seed + 7 * b)) for b in range(n_streams)]
        pins, per_stream, plain = [], [], []
        for sc in scenes:
            fl, pl = [], []
            for i in range(n_frames):
                f = sc.frame(i)
                if fmt16:
                    d16, h16, l16 = sensor_frame(f, cfg.depth_factor)
                    planes = dict(rgb=f["rgb"], depth=d16, ht=h16, lt=l16)
                    pl.append(dict(rgb=f["rgb"], depth=replay._convert(d16, 1.0 / float(np.float32(cfg.depth_factor))),
                                   ht=replay._convert(h16, 1.0 / 65535), lt=replay._convert(l16, 1.0 / 65535), q=f["q"], t=f["t"]))
                else:
                    planes = dict(rgb=f["rgb"], depth=f["depth"], ht=f["ht"], lt=f["lt"])
                    pl.append(dict(planes, q=f["q"], t=f["t"]))
                d = {}
                for k, a in planes.items():
                    p = tg.PinnedArray(a.shape, a.dtype)
                    p.array[...] = a
                    pins.append(p)
                    d[k] = p.array
                d.update(q=f["q"], t=f["t"])
                fl.append(d)
            per_stream.append(fl)
            plain.append(pl)
        frames = tg.make_host_frames(per_stream)
        mk = lambda: tg.TSDFGrid(cfg.voxel_size, cfg.truncation, pool_blocks=cfg.pool_blocks, table_slots=cfg.table_slots, max_image_pixels=H * W)  # noqa: E731
        grids = [mk() for _ in range(n_streams)]
        outs = {k: [tg.PinnedArray((H, W, 4) if k != "depth" else (H, W), np.uint8 if k != "depth" else np.float32) for _ in range(2 * n_streams)]
                for k in ("rgba", "normal", "depth")}
        tg.run_streams(grids, frames, 0, count, W, H, cfg.max_depth, cfg.K, depthmap_factor=cfg.depth_factor, raycast=True,
                       rgba=[p.array for p in outs["rgba"]], normal=[p.array for p in outs["normal"]], hit_depth=[p.array for p in outs["depth"]])
        for b in range(n_streams):
            ref = mk()
            for i in range(count):
                f = plain[b][i % n_frames]
                ref.Integrate(f["rgb"], f["depth"], f["ht"], f["lt"], cfg.max_depth, cfg.K, (f["q"], f["t"]))
                want = ref.RayCast(cfg.max_depth, cam, (f["q"], f["t"]))
            assert equal_exports(grids[b].export(), ref.export()), (fmt16, b)
            o = 2 * b + ((count - 1) & 1)  # image set of the last step
            for k, w_ in zip(("rgba", "normal", "depth"), want):
                assert np.array_equal(outs[k][o].array.view(np.uint8), w_.view(np.uint8)), (fmt16, b, k)
            ref.close()
        for g in grids:
            g.close()
        for p in pins + [q for v in outs.values() for q in v]:
            p.free()
