"""The C-ABI library loads and exports every symbol include/snn_b200.h declares (no compute calls),
and the product fails loudly — never silently falls back — when no CUDA device is usable."""
import ctypes as C
import os
import re

import pytest

from snn_b200 import _capi as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "snn_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"SNN_API\s+[\w\s\*]+?\b(snn_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    assert len(names) >= 65
    for must in ("snn_lattice_create", "snn_lattice_set_field", "snn_lattice_set_graph_dense", "snn_lattice_set_graph_csr",
                 "snn_lattice_set_graph_grid", "snn_lattice_set_plasticity", "snn_lattice_run", "snn_lattice_get_field",
                 "snn_lattice_get_grid_history", "snn_lattice_destroy", "snn_network_add_lattice",
                 "snn_network_add_spike_train_lattice", "snn_network_connect_dense", "snn_network_run"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = K.load_library()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/snn_b200.h but not exported"
        assert name in K.SIGNATURES, f"{name} has no ctypes prototype"
    assert set(K.SIGNATURES) <= set(declared_symbols())
    assert lib.snn_abi_version() == 1


def test_status_strings_mirror_gpu_error_display():
    """backend/src/error/mod.rs:241-256."""
    lib = K.load_library()
    expect = {1: "Could not compile program", 2: "Could not compile kernel", 3: "Could not create buffer",
              4: "Could not write to buffer", 5: "Could not read buffer", 6: "Could not wait for event",
              7: "Could not get device", 8: "Could not queue", 19: "Dimensions do not match"}
    for code, msg in expect.items():
        assert lib.snn_status_string(code).decode() == msg


def test_partition_begin_is_a_balanced_cover():
    lib = K.load_library()
    for rows in (0, 1, 7, 8, 100, 3163):
        for world in (1, 2, 3, 8):
            b = [lib.snn_partition_begin(rows, world, r) for r in range(world + 1)]
            assert b[0] == 0 and b[-1] == rows
            sizes = [b[i + 1] - b[i] for i in range(world)]
            assert all(s >= 0 for s in sizes) and max(sizes) - min(sizes) <= 1
    assert lib.snn_lattice_ipc_blob_size() >= 128


def _has_cuda():
    lib = K.load_library()
    n = C.c_int32()
    return lib.snn_device_count(C.byref(n)) == 0 and n.value > 0


def test_no_cpu_fallback_without_a_device():
    if _has_cuda():
        pytest.skip("a CUDA device is present")
    lib = K.load_library()
    d = K.LatticeDesc(C.sizeof(K.LatticeDesc), K.MODEL_IZH, 0, 0, 4, 4, -1, 0, 1)
    h = C.c_void_p()
    st = lib.snn_lattice_create(C.byref(d), C.byref(h))
    assert st == K.SNN_GPU_GET_DEVICE_FAILURE and not h
    assert b"no CPU fallback" in lib.snn_lattice_last_error(None)
    nd = K.NetworkDesc(C.sizeof(K.NetworkDesc), K.MODEL_IZH, 0, 0, 0, 0, -1)
    assert lib.snn_network_create(C.byref(nd), C.byref(h)) == K.SNN_GPU_GET_DEVICE_FAILURE
    import snn_b200 as S
    with pytest.raises(S.SnnError) as ei:
        lat = S.Lattice(S.IzhikevichNeuron)
        lat.populate(S.IzhikevichNeuron(), 2, 2)
    assert ei.value.status == K.SNN_GPU_GET_DEVICE_FAILURE


def test_bad_descriptors_are_rejected():
    lib = K.load_library()
    h = C.c_void_p()
    d = K.LatticeDesc(3, K.MODEL_IZH, 0, 0, 4, 4, -1, 0, 1)
    assert lib.snn_lattice_create(C.byref(d), C.byref(h)) == K.SNN_INVALID_ARGUMENT
    d = K.LatticeDesc(C.sizeof(K.LatticeDesc), 99, 0, 0, 4, 4, -1, 0, 1)
    assert lib.snn_lattice_create(C.byref(d), C.byref(h)) == K.SNN_INVALID_ARGUMENT
    assert lib.snn_lattice_destroy(None) == 0
