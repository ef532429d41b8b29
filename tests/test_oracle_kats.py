"""Pins the CPU oracle against the reference's own known-answer vectors and behavioural tests.

Each test cites the reference file:line it re-states (paths relative to /root/reference/backend).
No GPU, no product code: this validates the checker itself.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_api as O
from oracle_api import OracleBackend

f32 = np.float32
IZH, QIF, HH = 4, 1, 7


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_chemical_inputs_kat_reference_vectors():
    """src/neuron/gpu_lattices/mod.rs:3332-3343 inputs, :3406-3407 expected outputs.
    Node 1 is the spike train of the reference test; its flags/t are the spike_train_* vectors."""
    L = O.lib()
    n, T = 2, 3
    connections = np.array([0, 0, 1, 0], np.uint32)
    weights = np.array([0.0, 0.0, 0.7, 0.0], f32)
    flags = np.array([1, 1, 0, 1, 1, 0], np.uint32)            # neuron flags, then spike_train_flags
    t = np.array([0.5, 1.0, 0.0, 0.2, 0.4, 0.0], f32)          # t, then spike_train_t
    counts, res = np.zeros(n * T, f32), np.zeros(n * T, f32)
    L.orc_chemical_inputs_dense(_p(connections), _p(weights), _p(flags), _p(t), n, T, _p(counts), _p(res))
    assert counts.tolist() == [1.0, 1.0, 0.0, 0.0, 0.0, 0.0]
    assert res.tolist() == [float(f32(0.2) * f32(0.7)), float(f32(0.4) * f32(0.7)), 0.0, 0.0, 0.0, 0.0]


def test_chemical_inputs_through_network_path():
    """Same vectors through the network stepping path (neuron/mod.rs:2169-2212): neuron 0 with an AMPA+NMDA
    ApproximateReceptor takes r = averaged t after one chemical step."""
    be = OracleBackend(QIF, 0, 0, 1, 0)  # train_kind 1 = rate spike trains
    be.add_lattice(0, 1, 1)
    be.add_train_lattice(1, 1, 1)
    be.set_field(1, "neurotransmitters$flags", [1, 1, 0])
    be.set_field(1, "neurotransmitters$t", [0.2, 0.4, 0.0])
    be.set_field(0, "receptors$flags", [1, 1, 0])
    be.connect_dense(1, 0, np.array([[1]], np.uint32), np.array([[0.7]], f32))
    be.set_option(0, 0)
    be.set_option(1, 1)
    be.run(1)
    assert be.get_field(0, "receptors$AMPA$r$kinetics$r")[0] == f32(0.2) * f32(0.7)
    assert be.get_field(0, "receptors$NMDA$r$kinetics$r")[0] == f32(0.4) * f32(0.7)
    assert be.get_field(0, "receptors$GABA$r$kinetics$r")[0] == 0.0


@pytest.mark.parametrize("rate", [0, 100, 200, 300, 400, 500])
def test_rate_spike_train_expected_rate(rate):
    """tests/rate_spike_train.rs:27-52: spikes == ITERATIONS / (rate / dt) within 1; rate 0 never fires."""
    iterations = 10_000
    be = OracleBackend(IZH, 0, 0, 1, 0)
    be.add_train_lattice(0, 1, 1)
    be.fill_field(0, "rate", rate)
    be.set_option(4, 1, 0)
    be.run(iterations)
    spikes = int(be.spike_history(0).sum())
    if rate == 0:
        assert spikes == 0
    else:
        assert abs(spikes - iterations / (rate / 0.1)) <= 1.0


def test_rate_spike_train_spacing():
    """tests/rate_spike_train.rs:54-72: rate 100, dt 1 -> spikes exactly when (i+1) % 100 == 0 (i > 0)."""
    be = OracleBackend(IZH, 0, 0, 1, 0)
    be.add_train_lattice(0, 1, 1)
    be.fill_field(0, "rate", 100.0)
    be.fill_field(0, "dt", 1.0)
    be.set_option(4, 1, 0)
    be.run(1001)
    s = be.spike_history(0)[:, 0]
    for i in range(1001):
        assert bool(s[i]) == (i != 0 and (i + 1) % 100 == 0), i
    # a network holding only spike-train lattices still steps them (tests/rate_spike_train_lattices.rs:62-90)
    assert be.get_field(0, "last_firing_time")[0] == 999


def test_adjacency_matrix_doc_test():
    """src/graph/mod.rs:112-137 (AdjacencyMatrix) and :947-972 (AdjacencyList): identical answers."""
    L = O.lib()
    g = L.orc_adjmat_create()
    for x, y in [(0, 0), (0, 1), (1, 2)]:
        L.orc_adjmat_add_node(g, x, y)
    assert L.orc_adjmat_edit_weight(g, 0, 0, 0, 1, 1, 0.5) == 0
    assert L.orc_adjmat_edit_weight(g, 1, 2, 0, 1, 1, 1.0) == 0
    assert L.orc_adjmat_edit_weight(g, 0, 1, 4, 4, 1, 1.0) != 0
    has, w = C.c_int32(), C.c_float()
    assert L.orc_adjmat_lookup_weight(g, 0, 0, 0, 1, C.byref(has), C.byref(w)) == 0 and has.value == 1 and w.value == 0.5
    assert L.orc_adjmat_lookup_weight(g, 0, 1, 0, 0, C.byref(has), C.byref(w)) == 0 and has.value == 0
    assert L.orc_adjmat_lookup_weight(g, 3, 3, 0, 0, C.byref(has), C.byref(w)) != 0
    buf = np.zeros(16, np.uint32)
    k = L.orc_adjmat_incoming(g, 0, 1, _p(buf), 8)
    assert {tuple(buf[2 * i:2 * i + 2]) for i in range(k)} == {(0, 0), (1, 2)}
    k = L.orc_adjmat_outgoing(g, 1, 2, _p(buf), 8)
    assert {tuple(buf[2 * i:2 * i + 2]) for i in range(k)} == {(0, 1)}
    assert L.orc_adjmat_edit_weight(g, 0, 0, 0, 1, 0, 0.0) == 0
    assert L.orc_adjmat_lookup_weight(g, 0, 0, 0, 1, C.byref(has), C.byref(w)) == 0 and has.value == 0
    k = L.orc_adjmat_incoming(g, 0, 1, _p(buf), 8)
    assert {tuple(buf[2 * i:2 * i + 2]) for i in range(k)} == {(1, 2)}
    L.orc_adjmat_destroy(g)


def test_graph_semantics_in_network_storage():
    """Same doc-test semantics on the oracle's stepping graph: Some(0.0) is a connection, None is not
    (neuron/mod.rs:722-727: the averaging denominator counts zero-weight edges)."""
    be = OracleBackend(IZH, rows=1, cols=3)
    conn = np.array([[0, 1, 0], [0, 0, 0], [0, 1, 0]], np.uint32)
    w = np.array([[0, 0.0, 0], [0, 0, 0], [0, 2.0, 0]], f32)
    be.connect_dense(0, 0, conn, w)
    c2, w2 = be.get_connection_dense(0, 0)
    assert (c2 == conn).all() and (w2 == w).all()
    be.set_field(0, "current_voltage", [-60.0, -65.0, -50.0])
    be.fill_field(0, "gap_conductance", 10.0)
    be.fill_field(0, "c_m", 1.0)
    be.run(1)
    # I = (10*(-60+65)*0 + 10*(-50+65)*2) / 2 = 150 (two in-edges, one of weight zero)
    v, wv, I = f32(-65.0), f32(30.0), f32(150.0)
    dv = (((((f32(0.04) * (v * v)) + (f32(5) * v)) + f32(140)) - wv) + I) * (f32(0.1) / f32(1.0))
    assert be.get_field(0, "current_voltage")[1] == v + dv


def test_identical_all_to_all_neurons_follow_isolated_trajectory():
    """tests/gpu_connection_behavior.rs:51-95: 3x3 identical QIF, all-to-all weight 2, gap 10: the gap
    current is exactly zero, so every neuron follows the isolated neuron's trajectory."""
    n = 9
    be = OracleBackend(QIF, rows=3, cols=3)
    be.fill_field(0, "gap_conductance", 10.0)
    conn = (1 - np.eye(n)).astype(np.uint32)
    be.connect_dense(0, 0, conn, (conn * 2.0).astype(f32))
    be.set_option(3, 1, 0)
    iso = OracleBackend(QIF, rows=1, cols=1)
    iso.fill_field(0, "gap_conductance", 10.0)
    iso.set_option(3, 1, 0)
    be.run(1000)
    iso.run(1000)
    h, hi = be.grid_history(0), iso.grid_history(0)
    assert h.shape == (1000, 9)
    assert (h == hi).all()


def test_stdp_update_weight_formula():
    """src/neuron/plasticity/mod.rs:46-65 with the defaults of :29-39."""
    L = O.lib()
    p = O.Stdp(2.0, 2.0, 4.5, 4.5, 0.1)
    w = f32(1.0)
    assert L.orc_stdp_update(C.byref(p), w, -1, 5) == 1.0           # (None, Some) -> unchanged
    assert L.orc_stdp_update(C.byref(p), w, 5, -1) == 1.0
    assert L.orc_stdp_update(C.byref(p), w, 7, 7) == 1.0            # equal times -> unchanged
    pot = L.orc_stdp_update(C.byref(p), w, 3, 10)
    dep = L.orc_stdp_update(C.byref(p), w, 10, 3)
    assert pot == pytest.approx(1.0 + 2.0 * np.exp(-0.7 / 4.5), rel=1e-6)
    assert dep == pytest.approx(1.0 - 2.0 * np.exp(-0.7 / 4.5), rel=1e-6)


def test_poisson_from_firing_rate_and_refractoriness():
    """src/neuron/spike_train/mod.rs:330-337 and :84-86 / :174-176."""
    L = O.lib()
    assert L.orc_chance_from_firing_rate(20.0, 0.1) == f32(1.0) / ((f32(1000.0) / f32(0.1)) / f32(20.0))
    # delta-dirac: a*exp(-(dt/k)*td^2)+v_rest, k = 10000
    got = L.orc_refractoriness_effect(0, 10000.0, 25, 5, 30.0, 0.0, 0.1)
    assert got == pytest.approx(30.0 * np.exp(-(0.1 / 10000.0) * 400.0), rel=1e-6)
    got = L.orc_refractoriness_effect(1, 10000.0, 25, 5, 30.0, 0.0, 0.1)
    assert got == pytest.approx(30.0 * np.exp(-(0.1 / 10000.0) * 20.0), rel=1e-6)


@pytest.mark.parametrize("synapses", [(True, False), (False, True)])
def test_poisson_to_izhikevich_behaviour(synapses):
    """tests/spike_train_neuron_interaction.rs:90-157: 1x1 Poisson (id 0) -> 1x1 Izhikevich (id 1), weight 1,
    dt = 1: <= 1 spike in 2500 steps while chance_of_firing == 0, > 2 spikes in the next 2500 with
    chance_of_firing = 0.01 * dt."""
    iterations = 2500
    be = OracleBackend(IZH, 0, 0, 0, 0)
    be.add_train_lattice(0, 1, 1)
    be.add_lattice(1, 1, 1)
    for lid in (0, 1):
        be.set_field(lid, "neurotransmitters$flags", [1, 0, 0])
    be.set_field(1, "receptors$flags", [1, 0, 0])
    be.fill_field(1, "gap_conductance", 10.0)
    be.connect_dense(0, 1, np.array([[1]], np.uint32), np.array([[1.0]], f32))
    be.set_option(0, synapses[0])
    be.set_option(1, synapses[1])
    be.set_option(4, 1, 1)
    be.set_dt(1.0)
    be.run(iterations)
    assert be.spike_history(1).sum() <= 1
    be.fill_field(0, "chance_of_firing", 0.01)
    be.run(iterations)
    after = be.spike_history(1)[iterations:].sum()
    assert after > 2, after


def test_set_dt_rescales_poisson_chance():
    """src/neuron/spike_train/mod.rs:345-349."""
    be = OracleBackend(IZH, 0, 0, 0, 0)
    be.add_train_lattice(0, 1, 2)
    be.fill_field(0, "chance_of_firing", 0.02)
    be.set_dt(0.5)
    assert np.allclose(be.get_field(0, "chance_of_firing"), 0.02 * (0.5 / 0.1))
    assert (be.get_field(0, "dt") == f32(0.5)).all()


def test_no_synapses_is_noop_and_clock_continuity():
    """neuron/mod.rs:1217 ((false,false) -> Ok(())), :964-966/:979 (last_firing_time stamped with the clock
    before the increment; the clock persists across run calls)."""
    be = OracleBackend(IZH, rows=2, cols=2)
    be.set_option(0, 0)
    v0 = be.get_field(0, "current_voltage").copy()
    be.run(10)
    assert (be.get_field(0, "current_voltage") == v0).all() and be.get_option(5) == 0
    be.set_option(0, 1)
    be.fill_field(0, "current_voltage", 29.99)
    be.run(1)
    assert be.get_option(5) == 1 and (be.get_field(0, "last_firing_time") == 0).all()
    be.run(3)
    assert be.get_option(5) == 4
