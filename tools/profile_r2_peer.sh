#!/bin/bash
# ncu capture of raycast_kernel<true,*> (peer loads + fused scatter over NVLink) on 2 GPUs, through the pure C++
# thread-per-GPU driver.  Several attempts: kernel replay / application replay, with and without NVLink counters.
cd ${GRAFT_REPO_ROOT:-.}
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
BASE="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum"
NVL="nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum"
i=0
for spec in "kernel:$BASE" "application:$BASE" "application:$BASE,$NVL" "range:$BASE,$NVL"; do
  i=$((i+1)); mode=${spec%%:*}; mets=${spec#*:}
  timeout 240 ncu --metrics $mets --replay-mode $mode --clock-control none --cache-control none -k regex:raycast_kernel --launch-skip 6 -c 2 \
      -o gpurun_out/prof_r2_peer_$i -f tests/cpp/_build/mgpu_threads /tmp/peer_frames.bin /tmp/peer_out.bin > gpurun_out/r2_ncu_peer_$i.log 2>&1
  echo "attempt $i ($mode): rc=$?"; tail -2 gpurun_out/r2_ncu_peer_$i.log
done
ls -la gpurun_out/ | grep peer
