"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU and
exports every symbol include/tsdf_b200.h declares (no compute calls here)."""
import ctypes as C
import os
import re

from disinfect_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="tsdf_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsdf_[a-z_0-9]+)\s*\(", src)))


def test_header_and_binding_agree(tsdf_lib):
    names = declared_functions()
    assert len(names) >= 25
    assert sorted(_lib.SYMBOLS) == names
    for n in names:
        assert hasattr(tsdf_lib, n), f"libtsdf_b200.so does not export {n}"


def test_abi_version_and_default_config(tsdf_lib):
    assert tsdf_lib.tsdf_abi_version() == 1
    cfg = _lib.Config()
    assert tsdf_lib.tsdf_default_config(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(_lib.Config) == 32
    assert cfg.pool_blocks == 1 << 18 and cfg.table_slots == 1 << 21 and cfg.max_image_pixels == 1920 * 1080
    assert C.sizeof(_lib.Counters) == 64


def test_hash_through_cabi_matches_reference_formula(tsdf_lib):
    # pure host function of the ABI: utils/tsdf/voxel_hash.cu:31-35, utils/tests/voxel_hash_test.cu:130-135
    assert tsdf_lib.tsdf_hash(33, 180, 42) == tsdf_lib.tsdf_hash(61, 16, 170) == tsdf_lib.tsdf_hash(63, 171, 45) == (1 << 21) - 1
    assert tsdf_lib.tsdf_hash(-1, -1, -1) == 505009


def test_invalid_arguments_are_reported_not_crashed(tsdf_lib):
    h = C.c_void_p()
    assert tsdf_lib.tsdf_create(-1.0, 0.06, None, C.byref(h)) == -1  # TSDF_E_INVALID
    assert b"voxel_size" in tsdf_lib.tsdf_last_error()
    bad = _lib.Config()
    tsdf_lib.tsdf_default_config(C.byref(bad))
    bad.table_slots = 1000  # not a power of two
    assert tsdf_lib.tsdf_create(0.01, 0.06, C.byref(bad), C.byref(h)) == -1
    assert tsdf_lib.tsdf_destroy(None) == 0


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_multi_gpu_data_plane_header_and_binding_agree(tsdf_lib):
    """libtsdf_b200_mgpu.so (NCCL linked directly) loads without a GPU and exports every symbol of
    include/tsdf_b200_mgpu.h; it links NCCL and the engine, not torch."""
    import subprocess
    from disinfect_slam_b200 import mgpu
    L = mgpu.lib()
    names = declared_functions("tsdf_b200_mgpu.h")
    assert sorted(mgpu.SYMBOLS) == names and len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"libtsdf_b200_mgpu.so does not export {n}"
    assert C.sizeof(mgpu.Frame) == 64
    needed = subprocess.run(["readelf", "-d", mgpu.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl.so" in needed and "libtsdf_b200.so" in needed and "torch" not in needed
    assert L.tsdf_mgpu_destroy(None) == 0


def test_cpp_host_headers_compile(tmp_path):
    """The header-only C++17 mirrors (TSDFGrid incl. IntegrateU16, ShardedVolume over the NCCL data plane) compile with
    nothing but the C ABI headers -- no Eigen, OpenCV, CUDA or torch headers."""
    import subprocess
    src = tmp_path / "hdr_check.cc"
    src.write_text('#include "tsdf_b200/sharded_volume.hpp"\n#include "tsdf_b200/voxel_tsdf.hpp"\n#include "tsdf_b200/tsdf_system.hpp"\n'
                   'int main() { return sizeof(tsdf_b200::ShardedVolume) + sizeof(tsdf_b200::TSDFGrid) > 0 ? 0 : 1; }\n')
    res = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "o.o")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
