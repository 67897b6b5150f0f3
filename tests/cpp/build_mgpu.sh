#!/bin/bash
# Builds tests/cpp/_build/mgpu_threads: the pure C++ (one thread per GPU, NCCL linked through libtsdf_b200_mgpu.so)
# driver of the multi-GPU data plane.  Needs nothing from /root/reference.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
OUT="$HERE/_build"; mkdir -p "$OUT"
g++ -O2 -std=c++17 -Wall -I"$ROOT/include" -o "$OUT/mgpu_threads" "$HERE/mgpu_threads_main.cc" \
    -L"$ROOT/disinfect_slam_b200" -ltsdf_b200_mgpu -ltsdf_b200 -lpthread \
    -Wl,-rpath,'$ORIGIN/../../../disinfect_slam_b200' -Wl,-rpath-link,"$ROOT/disinfect_slam_b200"
echo "built $OUT/mgpu_threads"
