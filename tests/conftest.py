import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def tsdf_lib():
    """Build (if needed) and load libtsdf_b200.so; fails loudly if the CUDA extension is missing."""
    from disinfect_slam_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def ref_parity_lib():
    """The reference's own CUDA kernels rebuilt for sm_100a (oracle/_ref, built by __graft_entry__.build() in the
    development container and shipped prebuilt).  The strongest parity pin must never turn into a silent skip: a GPU
    run without the artefact FAILS."""
    from oracle import ref_cuda
    if not ref_cuda.available(True):
        pytest.fail("oracle/_ref/libref_tsdf_parity.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "where /root/reference exists (the .so travels to the GPU box with the snapshot)")
    return ref_cuda


@pytest.fixture(scope="session")
def dropin_binaries():
    """tests/cpp/_build (the reference's modules/tsdf_module.cc compiled unmodified on the engine); missing = failure."""
    d = os.path.join(ROOT, "tests", "cpp", "_build")
    need = [os.path.join(d, n) for n in ("dropin_tsdf_module", "dropin_native_system", "native_system_errors")]
    missing = [n for n in need if not os.path.exists(n)]
    if missing:
        pytest.fail(f"{missing} missing: run __graft_entry__.build() where /root/reference exists")
    return d
