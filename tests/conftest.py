"""pytest configuration: registers the `gpu` marker and puts the product package and the oracle
wrapper on sys.path.  `-m "not gpu"` runs everywhere; `-m gpu` needs a B200 and the built
libsnn_b200.so and must never read /root/reference."""
import os
import sys

import pytest

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


@pytest.fixture(scope="session")
def oracle_lattice_factory():
    from oracle_api import OracleBackend
    return lambda model, ntk, rck, rows, cols: OracleBackend(model, ntk, rck, rows=rows, cols=cols)


@pytest.fixture(scope="session")
def oracle_network_factory():
    from oracle_api import OracleBackend
    return lambda model, ntk, rck, tk, rf: OracleBackend(model, ntk, rck, tk, rf)
