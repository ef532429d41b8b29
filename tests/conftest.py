"""pytest configuration: registers the `gpu` marker and puts the product package and the oracle
wrapper on sys.path.  `-m "not gpu"` runs everywhere; `-m gpu` needs a B200 and the built
libsnn_b200.so and must never read /root/reference."""
import os
import sys

import pytest

# tests/test_gpu_strips.py steps several partition handles of ONE process on ONE device; their step kernels wait for each other on
# the device, so they must never share a hardware work queue (a kernel queued behind the kernel that waits for it would never
# start).  32 queues instead of the default 8; must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200) and the built libsnn_b200.so")


def pytest_sessionstart(session):
    """The built libraries are git-ignored: a freshly restored checkout has none.  Build them once (nvcc cross-compiles without
    a GPU, about a minute) instead of failing every test; a box without nvcc keeps the loud OSError of load_library()."""
    need = [os.path.join(ROOT, "spiking-neural-networks_b200", "libsnn_b200.so"), os.path.join(ROOT, "oracle", "libsnn_oracle.so")]
    if all(os.path.exists(p) for p in need):
        return
    try:
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    except Exception as exc:  # noqa: BLE001
        print(f"[conftest] could not build the native libraries: {exc!r}", file=sys.stderr)


def _cuda_device_usable():
    """True when the product library loads and sees a CUDA device (no compute call)."""
    try:
        import ctypes
        from snn_b200 import _capi
        n = ctypes.c_int32()
        return _capi.load_library().snn_device_count(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing 400 of them.  An explicit
    `-m gpu` run is left alone: there a missing device or library must fail loudly (no silent fallback)."""
    if "gpu" in (config.getoption("-m") or "").replace("not gpu", ""):
        return
    if _cuda_device_usable():
        return
    skip = pytest.mark.skip(reason="no CUDA device / libsnn_b200.so not usable")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lattice_factory():
    from oracle_api import OracleBackend
    return lambda model, ntk, rck, rows, cols: OracleBackend(model, ntk, rck, rows=rows, cols=cols)


@pytest.fixture(scope="session")
def oracle_network_factory():
    from oracle_api import OracleBackend
    return lambda model, ntk, rck, tk, rf: OracleBackend(model, ntk, rck, tk, rf)
