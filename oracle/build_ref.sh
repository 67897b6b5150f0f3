#!/bin/bash
# Rebuilds the reference's own TSDF CUDA path for sm_100a from the sources where they lie under
# /root/reference (never copied into this repo), against the stand-in headers in oracle/ref_shim
# (Eigen 3.3.9, OpenCV, spdlog and GL are not installed in this image).  Outputs go to oracle/_ref/
# only (git-ignored, shipped to the GPU box):
#   libref_tsdf.so          -O3 -DNDEBUG, nvcc's default FMA contraction: the reference as its own
#                           Release build would be on this GPU (the timing baseline)
#   libref_tsdf_parity.so   same + -fmad=false: IEEE float32 without contraction, comparable bit for
#                           bit with the CPU oracle and the new engine
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
[ -d "$REF/utils/tsdf" ] || { echo "build_ref.sh: $REF/utils/tsdf not found (reference absent): keeping prebuilt $OUT" >&2; exit 0; }
mkdir -p "$OUT"
SRCS="$REF/utils/tsdf/voxel_tsdf.cu $REF/utils/tsdf/voxel_hash.cu $REF/utils/tsdf/voxel_mem.cu $REF/utils/tsdf/voxel_types.cu $HERE/ref_harness.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -DNDEBUG -std=c++14 -rdc=true --expt-relaxed-constexpr -lineinfo -w -Xcompiler -fPIC -I$HERE/ref_shim -I$REF -shared"
newer() { [ ! -e "$1" ] && return 0; for f in $SRCS "$HERE"/ref_shim/Eigen/Dense "$HERE"/build_ref.sh; do [ "$f" -nt "$1" ] && return 0; done; return 1; }
if newer "$OUT/libref_tsdf.so"; then $NVCC $FLAGS -o "$OUT/libref_tsdf.so" $SRCS; fi
if newer "$OUT/libref_tsdf_parity.so"; then $NVCC $FLAGS -fmad=false -o "$OUT/libref_tsdf_parity.so" $SRCS; fi
echo "reference rebuild: $(ls "$OUT")"
