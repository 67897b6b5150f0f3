#!/bin/bash
# ncu evidence for round 2 (run under gpurun; results under gpurun_out/, summaries are copied to profiles/ by hand).
#   profile_r2.sh one   : launch list + full captures of the top kernels on one GPU
#   profile_r2.sh peer  : full capture of raycast_kernel<true,*> (peer loads + fused scatter over NVLink) on 2 GPUs
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extra-configs --no-e2e --no-cpu-baseline"
if [ "${1:-one}" = "one" ]; then
  $CMD > gpurun_out/r2_prof_bench.json 2> gpurun_out/r2_prof_bench.err || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
  ncu --set full --import-source on --clock-control none -k regex:"integrate_carve_kernel|raycast_kernel|frame_allocate_kernel|skip_pass_kernel|skip_mark_kernel|select_visible" \
      --launch-skip 60 -c 12 -o gpurun_out/prof_r2_full $CMD > gpurun_out/r2_ncu_full.log 2>&1
  ls -la gpurun_out/prof_r2_full.ncu-rep
else
  ncu --query-metrics 2>/dev/null | grep -i -E "nvl|peer" | head -60 > gpurun_out/r2_peer_metric_names.txt
  python - <<'PY'
import os, struct, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from disinfect_slam_b200 import synth
cfg = synth.config("config2"); sc = synth.Scene(cfg); n = 6
with open("/tmp/peer_frames.bin", "wb") as fh:
    fh.write(struct.pack("<4i3f4f6f", n, cfg.width, cfg.height, 2, cfg.voxel_size, cfg.truncation, cfg.max_depth, *[float(np.float32(k)) for k in cfg.K], *([0.0] * 6)))
    for i in range(n):
        f = sc.frame(i)
        fh.write(np.concatenate([f["q"], f["t"]]).astype(np.float32).tobytes())
        for k in ("rgb", "depth", "ht", "lt"):
            fh.write(f[k].tobytes())
PY
  tests/cpp/_build/mgpu_threads /tmp/peer_frames.bin /tmp/peer_out.bin || exit 1
  ncu --set full --import-source on --clock-control none --metrics regex:".*aperture_peer.*",regex:"nvl.*bytes.*" -k regex:raycast_kernel --launch-skip 6 -c 2 \
      -o gpurun_out/prof_r2_peer tests/cpp/_build/mgpu_threads /tmp/peer_frames.bin /tmp/peer_out.bin > gpurun_out/r2_ncu_peer.log 2>&1
  tail -5 gpurun_out/r2_ncu_peer.log
  ls -la gpurun_out/prof_r2_peer.ncu-rep
fi
