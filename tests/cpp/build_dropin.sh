#!/bin/bash
# Compiles the reference's modules/tsdf_module.cc UNMODIFIED (from /root/reference) against the drop-in
# header include/tsdf_b200/compat/utils/tsdf/voxel_tsdf.cuh and links it to libtsdf_b200.so.
# Eigen / OpenCV / spdlog / GL are not installed here: oracle/ref_shim provides the same stand-ins the
# reference rebuild uses.  Output: tests/cpp/_build/dropin_tsdf_module (git-ignored, shipped to the GPU box).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
[ -d "$REF/modules" ] || { echo "build_dropin.sh: $REF/modules not found: keeping prebuilt binary" >&2; exit 0; }
OUT="$HERE/_build"; mkdir -p "$OUT"
INC="-I$ROOT/include/tsdf_b200/compat -I$ROOT/include -I$ROOT/oracle/ref_shim -I$REF -I/usr/local/cuda/include"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++14 -w --expt-relaxed-constexpr $INC -Xcompiler -fPIC \
    -c "$REF/utils/tsdf/voxel_types.cu" -o "$OUT/voxel_types.o"
g++ -O2 -std=c++17 -w $INC -c "$REF/modules/tsdf_module.cc" -o "$OUT/tsdf_module.o"
g++ -O2 -std=c++17 -w $INC -c "$HERE/dropin_main.cc" -o "$OUT/dropin_main.o"
g++ -o "$OUT/dropin_tsdf_module" "$OUT/dropin_main.o" "$OUT/tsdf_module.o" "$OUT/voxel_types.o" \
    -L"$ROOT/disinfect_slam_b200" -ltsdf_b200 -L/usr/local/cuda/lib64 -lcudart -lpthread \
    -Wl,-rpath,'$ORIGIN/../../../disinfect_slam_b200' -Wl,-rpath,/usr/local/cuda/lib64
echo "built $OUT/dropin_tsdf_module"
# the same driver on the B200-native TSDFSystem (include/tsdf_b200/tsdf_system.hpp through the shadowing
# modules/tsdf_module.h): the reference's tsdf_module.cc is NOT part of this binary
g++ -O2 -std=c++17 -w -I"$ROOT/include/tsdf_b200/compat_system" $INC -c "$HERE/dropin_main.cc" -o "$OUT/dropin_native_main.o"
g++ -o "$OUT/dropin_native_system" "$OUT/dropin_native_main.o" "$OUT/voxel_types.o" \
    -L"$ROOT/disinfect_slam_b200" -ltsdf_b200 -L/usr/local/cuda/lib64 -lcudart -lpthread \
    -Wl,-rpath,'$ORIGIN/../../../disinfect_slam_b200' -Wl,-rpath,/usr/local/cuda/lib64
echo "built $OUT/dropin_native_system"
g++ -O2 -std=c++17 -w -I"$ROOT/include/tsdf_b200/compat_system" $INC -c "$HERE/native_errors_main.cc" -o "$OUT/native_errors_main.o"
g++ -o "$OUT/native_system_errors" "$OUT/native_errors_main.o" "$OUT/voxel_types.o" \
    -L"$ROOT/disinfect_slam_b200" -ltsdf_b200 -L/usr/local/cuda/lib64 -lcudart -lpthread \
    -Wl,-rpath,'$ORIGIN/../../../disinfect_slam_b200' -Wl,-rpath,/usr/local/cuda/lib64
echo "built $OUT/native_system_errors"
