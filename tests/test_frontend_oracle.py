"""Host-side logic of the reference-shaped front end (populate / connect / apply / network assembly / error
behaviour), exercised on CPU with the oracle back end injected.  No product compute is involved."""
import numpy as np
import pytest

import scenarios as SC
import snn_b200 as S
from snn_b200 import _capi as K
from test_gpu_parity import build_network

f32 = np.float32


def test_populate_connect_apply_cell_grid(oracle_lattice_factory):
    base = S.IzhikevichNeuron(gap_conductance=10.0, c_m=25.0)
    lat = S.Lattice(S.IzhikevichNeuron, backend_factory=oracle_lattice_factory)
    lat.populate(base, 3, 4)
    lat.connect(lambda x, y: x != y, lambda x, y: 5.0)
    assert lat.get_weight((0, 0), (2, 3)) == 5.0 and lat.get_weight((1, 1), (1, 1)) is None
    with pytest.raises(S.SnnError) as ei:
        lat.get_weight((0, 0), (9, 9))
    assert ei.value.status == K.SNN_GRAPH_POSTSYNAPTIC_NOT_FOUND
    lat.apply_given_position(lambda pos, n: setattr(n, "current_voltage", -60.0 + pos[0] * 4 + pos[1]))
    assert (lat.get_field("current_voltage") == -60.0 + np.arange(12, dtype=f32)).all()
    grid = lat.cell_grid()
    assert grid[2][3].c_m == 25.0 and grid[2][3].gap_conductance == 10.0 and grid[0][0].last_firing_time is None
    with pytest.raises(S.SnnError):
        lat.set_cell_grid(grid[:2])
    lat.update_grid_history = True
    lat.run_lattice(20)
    assert lat.grid_history.history.shape == (20, 3, 4) and lat.internal_clock == 20


def test_chemistry_objects_round_trip(oracle_lattice_factory):
    T = S.IonotropicNeurotransmitterType
    base = S.HodgkinHuxleyNeuron()
    SC.add_chemistry(base, "destexhe_all")
    base.receptors[T.NMDA].mg = 0.5
    base.synaptic_neurotransmitters[T.GABA].k_p = 7.0
    lat = S.Lattice(S.HodgkinHuxleyNeuron, backend_factory=oracle_lattice_factory)
    lat.populate(base, 2, 2)
    cell = lat.cell_grid()[1][1]
    assert cell.receptors[T.NMDA].mg == 0.5 and isinstance(cell.receptors[T.AMPA].r, S.DestexheReceptor)
    assert cell.synaptic_neurotransmitters[T.GABA].k_p == 7.0 and cell.synaptic_neurotransmitters[T.AMPA].v_p == 2.0
    assert cell.na_channel_g_na == 120.0 and cell["k_channel$e_k"] == -77.0
    with pytest.raises(TypeError):
        bad = S.IzhikevichNeuron()
        bad.receptors[T.AMPA] = S.GABAReceptor()
        S.Lattice(S.IzhikevichNeuron, backend_factory=oracle_lattice_factory).populate(bad, 1, 1)


def test_network_assembly_and_errors(oracle_lattice_factory, oracle_network_factory):
    net = build_network(oracle_lattice_factory, oracle_network_factory, train="rate")
    assert net.get_all_ids() == {0, 1, 2}
    with pytest.raises(S.SnnError) as ei:
        net.connect(1, 0, lambda x, y: True)
    assert ei.value.status == K.SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN
    with pytest.raises(S.SnnError) as ei:
        net.connect(9, 1, lambda x, y: True)
    assert ei.value.status == K.SNN_NET_PRESYNAPTIC_ID_NOT_FOUND
    with pytest.raises(S.SnnError) as ei:
        net.connect(1, 9, lambda x, y: True)
    assert ei.value.status == K.SNN_NET_POSTSYNAPTIC_ID_NOT_FOUND
    dup = S.Lattice(S.IzhikevichNeuron, id=2, backend_factory=oracle_lattice_factory)
    dup.populate(S.IzhikevichNeuron(), 1, 1)
    with pytest.raises(S.SnnError) as ei:
        net.add_lattice(dup)
    assert ei.value.status == K.SNN_NET_GRAPH_ID_ALREADY_PRESENT
    c, w = net.connection_dense(2, 1)
    assert c.sum() == 20 * 19 and (w[c == 1] == -1.0).all()
    # the internal graph and per-cell state moved into the network with the lattice
    ci, wi = net.connection_dense(1, 1)
    assert ci.sum() > 0 and (wi[ci == 1] == 0.5).all()
    net.run_lattices(50)
    assert net.internal_clock == 50
    assert net.get_lattice(1).grid_history.history.shape == (50, 4, 5)
    assert net.get_spike_train_lattice(0).spike_history.history.sum() > 0
    with pytest.raises(RuntimeError):
        net.get_lattice(1).run_lattice(1)


def test_poisson_from_firing_rate_matches_reference_formula():
    p = S.PoissonNeuron.from_firing_rate(20.0, 0.1)
    assert p.chance_of_firing == float(f32(1.0) / ((f32(1000.0) / f32(0.1)) / f32(20.0)))


def test_reduced_histories_on_the_oracle(oracle_lattice_factory):
    """AverageVoltageHistory / EEGHistory restated literally (sequential f32 sums, neuron/mod.rs:266-279, 310-316)."""
    kw = dict(model="izh", rows=5, cols=6, seed=2, graph="grid")
    avg = SC.build_lattice(oracle_lattice_factory, history_type=S.AverageVoltageHistory, **kw)
    eeg = SC.build_lattice(oracle_lattice_factory, history_type=S.EEGHistory, **kw)
    grid = SC.build_lattice(oracle_lattice_factory, **kw)
    for L in (avg, eeg, grid):
        L.run_lattice(50)
    v = grid.grid_history.history.reshape(50, -1)
    want_avg, want_eeg = [], []
    for row in v:
        s, t = f32(0), f32(0)
        for x in row:
            s = f32(s + x)
            t = f32(t + f32(x - f32(0.007)))
        want_avg.append(f32(s / f32(30)))
        want_eeg.append(f32(f32(1) / f32(f32(f32(f32(4) * f32(np.pi)) * f32(251.0)) * f32(0.8))) * t)
    assert (avg.grid_history.history == np.array(want_avg, f32)).all()
    assert (eeg.grid_history.history == np.array(want_eeg, f32)).all()


def _reward_pair(oracle_lattice_factory, rows, cols, seed):
    """A C-oracle RewardModulatedLattice and the independent numpy restatement of the same lattice."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import numpy_ref as R
    lat = SC.build_lattice(oracle_lattice_factory, model="izh", rows=rows, cols=cols, seed=seed, graph="random", cls=S.RewardModulatedLattice)
    lat.reward_modulator = S.RewardModulatedSTDP(tau_c=0.05, a_plus=0.4, a_minus=0.3)
    n = rows * cols
    conn, w = lat.graph_dense()
    fields = {k: lat.get_field(k) for k in ("current_voltage", "gap_conductance", "w_value", "a", "b", "c", "d", "v_th", "tau_m", "c_m", "dt")}
    ref = R.RewardDenseLattice(R.IZH, n, conn, w, fields)
    ref.mod.update(tau_c=f32(0.05), a_plus=f32(0.4), a_minus=f32(0.3))
    return lat, ref


def test_reward_modulated_oracle_matches_independent_restatement(oracle_lattice_factory):
    """RewardModulatedLattice + RewardModulatedSTDP/TraceRSTDP (neuron/mod.rs:2717-3416, plasticity/mod.rs:114-234): the C oracle
    against the separately written numpy restatement — rasters bit-exact, weights and traces to expf rounding."""
    lat, ref = _reward_pair(oracle_lattice_factory, 3, 4, 5)
    rng = np.random.default_rng(0)
    rewards = rng.uniform(-1, 1, 120).astype(f32)
    w0 = lat.graph_csr()[2].copy()
    lat.run_lattice_with_rewards(rewards[:70])
    lat.run_lattice(30)                      # no reward signal: dopamine stays, the modulator keeps running
    lat.run_lattice_with_rewards(rewards[70:])
    ref.run_with_rewards(rewards[:70]); ref.run(30); ref.run_with_rewards(rewards[70:])
    assert np.array(ref.s_hist).sum() > 5
    assert (lat.spike_history.history.reshape(150, -1) == np.array(ref.s_hist)).all()
    rp, pre, w = lat.graph_csr()
    cnt, dw, c = lat.graph_traces()
    post = np.repeat(np.arange(12), np.diff(rp).astype(int))
    np.testing.assert_allclose(w, ref.w[pre, post], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(c, ref.c[pre, post], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(dw, ref.dw[pre, post], rtol=2e-5, atol=1e-7)
    assert (cnt == ref.counter[pre, post]).all()
    assert np.abs(w - w0).max() > 1e-3, "the modulator must have moved the weights"
    assert lat.reward_modulator.dopamine == pytest.approx(float(ref.mod["dopamine"]), rel=1e-5)


def test_bcm_plasticity_oracle_matches_independent_restatement(oracle_lattice_factory):
    """BCM rule (plasticity/mod.rs:80-112) on a BCMIzhikevichNeuron lattice: C oracle vs the numpy restatement, bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import numpy_ref as R
    lat = SC.build_lattice(oracle_lattice_factory, model="bcm_izh", rows=4, cols=5, seed=6, graph="random", stdp=True)
    lat.plasticity = S.BCM(decay=0.05, average_scalar=0.5, dt=1e-6)   # tame: activities grow with the never-reset num_spikes
    conn, w = lat.graph_dense()
    names = ("current_voltage", "gap_conductance", "w_value", "a", "b", "c", "d", "v_th", "tau_m", "c_m", "dt", "average_activity",
             "current_activity", "period", "num_spikes", "firing_rate_clock", "firing_rate_window")
    ref = R.DenseLattice(R.BCM_IZH, 20, conn, w, {k: lat.get_field(k) for k in names})
    ref.do_plasticity, ref.use_bcm = True, True
    ref.bcm = dict(decay=f32(0.05), average_scalar=f32(0.5), dt=f32(1e-6))
    lat.run_lattice(400)
    ref.run(400)
    assert np.array(ref.s_hist).sum() > 20 and ref.f["average_activity"].max() > 0 and np.isfinite(np.array(ref.v_hist)).all()
    assert (lat.spike_history.history.reshape(400, -1) == np.array(ref.s_hist)).all()
    assert (lat.grid_history.history.reshape(400, -1) == np.array(ref.v_hist)).all()
    c2, w2 = lat.graph_dense()
    assert (w2 == ref.w).all() and np.abs(w2 - w).max() > 1e-5
