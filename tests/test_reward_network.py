"""RewardModulatedLatticeNetwork (reference: backend/src/neuron/mod.rs:3455-5455).

CPU part: the front end over the oracle back end (assembly, the reference's error order, the refused configurations) and the
oracle's weight pass against a literal, independent replay of post_neuron_update_step (:5030-5062) driven by the recorded
spike rasters.  GPU part: the CUDA path against the oracle on the same networks."""
import numpy as np
import pytest

import scenarios as SC
import snn_b200 as S
from snn_b200 import _capi as K

f32 = np.float32
RMC = S.RewardModulatedConnection


def build_reward_network(lattice_factory, network_factory, mode=(True, False), variant="lsm", seed=5, plastic_liquid=False):
    """The shape of the reference's only user (examples/lsm_architecture): spike trains -> liquid (plain lattice) -> read-out
    (reward-modulated lattice).  variant "lsm": trains -> read-out RewardModulatedWeight, liquid -> read-out Weight;
    variant "swap": the two kinds exchanged."""
    rng = np.random.default_rng(seed)
    T = S.IonotropicNeurotransmitterType
    base = S.IzhikevichNeuron(gap_conductance=5.0, c_m=10.0)
    if mode[1]:
        base.synaptic_neurotransmitters[T.AMPA] = S.ApproximateNeurotransmitter()
        base.receptors[T.AMPA] = S.AMPAReceptor()
    liquid = S.Lattice(S.IzhikevichNeuron, id=1, backend_factory=lattice_factory)
    liquid.populate(base, 4, 5)
    liquid.connect(lambda x, y: x != y and abs(x[0] - y[0]) + abs(x[1] - y[1]) <= 1, lambda x, y: 0.5)
    liquid.do_plasticity = plastic_liquid
    liquid.plasticity = S.STDP(a_plus=0.002, a_minus=0.0015, tau_plus=3.5)
    readout = S.RewardModulatedLattice(S.IzhikevichNeuron, id=2, backend_factory=lattice_factory)
    readout.populate(base, 3, 3)
    readout.connect(lambda x, y: x != y, lambda x, y: 0.3)
    readout.reward_modulator = S.RewardModulatedSTDP(tau_c=0.1, a_plus=0.002, a_minus=0.0025, tau_plus=5.0, tau_minus=4.0)
    for L, n in ((liquid, 20), (readout, 9)):
        L.set_field("current_voltage", rng.uniform(-65, 20, n).astype(f32))
        L.set_field("b", rng.uniform(0.25, 0.33, n).astype(f32))
        L.update_grid_history = L.update_spike_history = True
    st_base = S.RateSpikeTrain(rate=2.0)
    if mode[1]:
        st_base.synaptic_neurotransmitters[T.AMPA] = S.ApproximateNeurotransmitter()
    st = S.SpikeTrainLattice(S.RateSpikeTrain, id=0, network_backend_factory=network_factory)
    st.populate(st_base, 3, 4)
    st.set_field("rate", rng.choice([0.0, 1.5, 2.0, 3.0], 12).astype(f32))
    st.update_spike_history = True
    net = S.RewardModulatedLatticeNetwork.generate_network([liquid], [readout], [st], backend_factory=network_factory)
    w01 = rng.uniform(0, 1, (12, 20)).astype(f32)
    w02 = rng.uniform(0, 1, (12, 9)).astype(f32)
    w12 = rng.uniform(0, 1, (20, 9)).astype(f32)
    net.connect(0, 1, lambda x, y: True, lambda x, y: float(w01[x[0] * 4 + x[1], y[0] * 5 + y[1]]))
    kind02 = RMC.RewardModulatedWeight if variant == "lsm" else RMC.Weight
    kind12 = RMC.Weight if variant == "lsm" else RMC.RewardModulatedWeight
    wrap = lambda kind, w: kind(S.TraceRSTDP(weight=w)) if kind == RMC.RewardModulatedWeight else kind(w)
    net.connect_with_reward_modulation(0, 2, lambda x, y: (x[0] + y[1]) % 3 != 0,
                                       lambda x, y: wrap(kind02, float(w02[x[0] * 4 + x[1], y[0] * 3 + y[1]])))
    net.connect_with_reward_modulation(1, 2, lambda x, y: (x[1] + y[0]) % 2 == 0,
                                       lambda x, y: wrap(kind12, float(w12[x[0] * 5 + x[1], y[0] * 3 + y[1]])))
    net.electrical_synapse, net.chemical_synapse = mode
    return net


BLOCKS = ((0, 1), (0, 2), (1, 2), (1, 1), (2, 2))


def rewards_for(n, seed=3):
    return np.random.default_rng(seed).uniform(-0.002, 0.003, n).astype(f32)   # dopamine settles around tau_d * mean / (dt / tau_d) ~ 2


# ------------------------------------------------------------------ CPU: front end + oracle
def test_assembly_and_reference_error_order(oracle_lattice_factory, oracle_network_factory):
    net = build_reward_network(oracle_lattice_factory, oracle_network_factory)
    assert net.get_all_ids() == {0, 1, 2} and net.get_reward_modulated_lattice(2).do_modulation
    always = lambda x, y: True
    rw = lambda x, y: RMC.RewardModulatedWeight()
    for call, status in (
            (lambda: net.connect(1, 2, always), K.SNN_NET_CONNECT_FUNCTION_MUST_HAVE_NON_REWARD_MODULATED_LATTICE),
            (lambda: net.connect(2, 1, always), K.SNN_NET_CONNECT_FUNCTION_MUST_HAVE_NON_REWARD_MODULATED_LATTICE),
            (lambda: net.connect(1, 0, always), K.SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN),
            (lambda: net.connect(7, 1, always), K.SNN_NET_PRESYNAPTIC_ID_NOT_FOUND),
            (lambda: net.connect(1, 7, always), K.SNN_NET_POSTSYNAPTIC_ID_NOT_FOUND),
            (lambda: net.connect_with_reward_modulation(0, 1, always, rw), K.SNN_NET_CANNOT_CONNECT_WITH_REWARD_MODULATED_CONNECTION),
            (lambda: net.connect_with_reward_modulation(2, 2, always, rw), K.SNN_NET_REWARD_MODULATED_CONNECTION_NOT_COMPATIBLE_INTERNALLY),
            (lambda: net.connect_with_reward_modulation(2, 0, always, rw), K.SNN_NET_POSTSYNAPTIC_LATTICE_CANNOT_BE_SPIKE_TRAIN),
            (lambda: net.connect_with_reward_modulation(9, 2, always, rw), K.SNN_NET_PRESYNAPTIC_ID_NOT_FOUND),
            (lambda: net.connect_reward_modulated_lattice_interally(1, always), K.SNN_NET_ID_NOT_FOUND_IN_LATTICES)):
        with pytest.raises(S.SnnError) as ei:
            call()
        assert ei.value.status == status
    mixed = lambda x, y: RMC.Weight(1.0) if x[1] else RMC.RewardModulatedWeight()
    with pytest.raises(S.SnnError):
        net.connect_with_reward_modulation(0, 2, always, mixed)
    net.run_lattices(3)
    net.run_lattices_with_reward(0.5)
    assert net.internal_clock == 4
    assert net.get_reward_modulated_lattice(2).reward_modulator.dopamine == pytest.approx(20.0 * 0.5)
    assert net.get_lattice(1).grid_history.history.shape == (4, 4, 5)


def _run_refused(net):
    with pytest.raises(Exception) as ei:
        net.run_lattices(1)
    return ei.value


def check_refused_configurations(lattice_factory, network_factory):
    # connecting edges out of a reward-modulated lattice with do_modulation (neuron/mod.rs:4931-4934)
    net = build_reward_network(lattice_factory, network_factory)
    net.connect_with_reward_modulation(2, 1, lambda x, y: x == y, lambda x, y: RMC.Weight(0.2))
    _run_refused(net)
    net.get_reward_modulated_lattice(2).do_modulation = False   # nobody walks those edges then
    net.run_lattices(2)
    # connecting edges out of a plastic plain lattice (:4760-4763)
    net = build_reward_network(lattice_factory, network_factory, plastic_liquid=True)
    _run_refused(net)
    # a RewardModulatedWeight block into a plain lattice: nothing in the reference ever advances it
    net = build_reward_network(lattice_factory, network_factory)
    net.get_reward_modulated_lattice(2).do_modulation = False
    net.connect_with_reward_modulation(2, 1, lambda x, y: x == y, lambda x, y: RMC.RewardModulatedWeight())
    _run_refused(net)


def test_refused_configurations_oracle(oracle_lattice_factory, oracle_network_factory):
    check_refused_configurations(oracle_lattice_factory, oracle_network_factory)


def _replay_reference_weight_pass(net, steps, rewards):
    """post_neuron_update_step for the reward-modulated lattices, replayed literally with dictionaries: positions in
    node order, for each the incoming connecting edges, the incoming edges of the own graph, then the outgoing ones."""
    lat = {1: net.get_lattice(1), 2: net.get_reward_modulated_lattice(2), 0: net.get_spike_train_lattice(0)}
    size = {i: lat[i].rows * lat[i].cols for i in lat}
    ras = {1: lat[1].spike_history.history.reshape(steps, -1), 2: lat[2].spike_history.history.reshape(steps, -1),
           0: lat[0].spike_history.history.reshape(steps, -1)}
    return size, ras


def _stdp_term(a_plus, a_minus, tau_plus, tau_minus, dt, t_pre, t_post):
    if t_pre is None or t_post is None:
        return f32(0)
    t_pre, t_post = f32(t_pre), f32(t_post)
    if t_pre < t_post:
        return f32(a_plus) * np.exp(f32(-1) * np.abs((t_pre - t_post) * f32(dt)) / f32(tau_plus), dtype=f32)
    if t_pre > t_post:
        return f32(-1) * f32(a_minus) * np.exp(f32(-1) * np.abs((t_post - t_pre) * f32(dt)) / f32(tau_minus), dtype=f32)
    return f32(0)


@pytest.mark.parametrize("variant", ["lsm", "swap"])
def test_oracle_weight_pass_matches_literal_replay(variant, oracle_lattice_factory, oracle_network_factory):
    steps = 120
    rewards = rewards_for(steps)
    net = build_reward_network(oracle_lattice_factory, oracle_network_factory, variant=variant)
    be = net._be
    start = {blk: be.get_connection_csr(*blk) for blk in BLOCKS}
    net.run_lattices_with_rewards(rewards)
    size, ras = _replay_reference_weight_pass(net, steps, rewards)
    m = net.get_reward_modulated_lattice(2).reward_modulator
    liquid_p = net.get_lattice(1).plasticity
    # edge dictionaries keyed by ((pre_id, pre), (post_id, post)) -> [counter, dw, weight, c]
    own, conn, kind = {}, {}, {}
    for blk in BLOCKS:
        rp, pre, w = start[blk]
        for q in range(len(rp) - 1):
            for k in range(rp[q], rp[q + 1]):
                key = ((blk[0], int(pre[k])), (blk[1], q))
                if blk == (2, 2):
                    own[key] = [0, f32(0), f32(w[k]), f32(0)]
                elif blk[1] == 2:
                    conn[key] = [0, f32(0), f32(w[k]), f32(0)]
                    kind[key] = (variant == "lsm") == (blk[0] == 0)   # True = RewardModulatedWeight
    lft = {i: [None] * size[i] for i in size}
    dopamine, decay_c = f32(0), np.exp(f32(-m.dt) / f32(m.tau_c), dtype=f32)

    def modulator_call(tr, t_pre, t_post):   # RewardModulatedSTDP::update_weight, plasticity/mod.rs:197-229
        tr[1] = f32(tr[1] + _stdp_term(m.a_plus, m.a_minus, m.tau_plus, m.tau_minus, m.dt, t_pre, t_post))
        if tr[0] == 0:
            tr[0] = 1
        else:
            tr[3] = f32(f32(tr[3] * decay_c) + f32(f32(m.tau_c) * tr[1]))
            tr[0], tr[1] = 0, f32(0)
        tr[2] = f32(tr[2] + f32(tr[3] * dopamine))

    for t in range(steps):
        dopamine = f32(f32(dopamine * np.exp(f32(-m.dt) / f32(m.tau_d), dtype=f32)) + f32(f32(m.tau_d) * rewards[t]))
        for i in (1, 2):   # neurons stamp this step; the spike trains step after the weight pass
            for j in np.flatnonzero(ras[i][t]):
                lft[i][j] = t
        for q in range(size[2]):
            pos = (2, q)
            for key, tr in conn.items():            # update_weights_from_neurons_across_reward_lattices, incoming arm
                if key[1] != pos:
                    continue
                pid, p = key[0]
                if kind[key]:
                    modulator_call(tr, lft[pid][p], lft[2][q])
                elif pid == 1:
                    tr[2] = f32(tr[2] + _stdp_term(liquid_p.a_plus, liquid_p.a_minus, liquid_p.tau_plus, liquid_p.tau_minus,
                                                   liquid_p.dt, lft[1][p], lft[2][q]))
            for key, tr in own.items():             # _within_reward_lattices: incoming edges of the own graph ...
                if key[1] == pos:
                    modulator_call(tr, lft[2][key[0][1]], lft[2][q])
            for key, tr in own.items():             # ... then the outgoing ones
                if key[0] == pos:
                    modulator_call(tr, lft[2][q], lft[2][key[1][1]])
        for j in np.flatnonzero(ras[0][t]):
            lft[0][j] = t
    assert sum(int(r.sum()) for r in ras.values()) > 20
    changed = 0
    for blk in ((0, 2), (1, 2), (2, 2)):
        rp, pre, w = be.get_connection_csr(*blk)
        cnt, dw, c = be.connection_traces(*blk)
        table = own if blk == (2, 2) else conn
        for q in range(len(rp) - 1):
            for k in range(rp[q], rp[q + 1]):
                tr = table[((blk[0], int(pre[k])), (blk[1], q))]
                assert tr[0] == cnt[k]
                np.testing.assert_allclose([tr[1], tr[2], tr[3]], [dw[k], w[k], c[k]], rtol=2e-4, atol=1e-6)
                changed += int(w[k] != start[blk][2][k])
    assert changed > 10
    # a Weight block fed by spike trains is left alone (neuron/mod.rs:4868-4884 only takes plain lattices)
    if variant == "swap":
        assert (be.get_connection_csr(0, 2)[2] == start[(0, 2)][2]).all()
    assert (be.get_connection_csr(0, 1)[2] == start[(0, 1)][2]).all() and (be.get_connection_csr(1, 1)[2] == start[(1, 1)][2]).all()


# ------------------------------------------------------------------ GPU: CUDA path against the oracle
def _sync(src, dst):
    for lid in (1,):
        SC.copy_lattice_state(src.get_lattice(lid), dst.get_lattice(lid), weights=False)
    SC.copy_lattice_state(src.get_reward_modulated_lattice(2), dst.get_reward_modulated_lattice(2), weights=False)
    SC.copy_lattice_state(src.get_spike_train_lattice(0), dst.get_spike_train_lattice(0), weights=False)
    for blk in BLOCKS:
        rp, pre, w = src._be.get_connection_csr(*blk)
        if blk[1] == 2:
            cnt, dw, c = src._be.connection_traces(*blk)
            dst._be.set_connection_traces(w, cnt, dw, c, pre_id=blk[0], post_id=blk[1])
        else:
            dst._be.set_connection_traces(w, None, None, None, pre_id=blk[0], post_id=blk[1])
    dst.get_reward_modulated_lattice(2).reward_modulator.dopamine = src.get_reward_modulated_lattice(2).reward_modulator.dopamine


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["lsm", "swap"])
@pytest.mark.parametrize("mode", [(True, False), (True, True), (False, True)])
def test_reward_network_matches_oracle(mode, variant, oracle_lattice_factory, oracle_network_factory):
    a = build_reward_network(None, None, mode=mode, variant=variant)
    b = build_reward_network(oracle_lattice_factory, oracle_network_factory, mode=mode, variant=variant)
    total, seg = 400, 40
    rewards = rewards_for(total)
    spikes = 0
    for s0 in range(0, total, seg):
        if s0 == 3 * seg:   # a stretch without a reward signal (RunNetwork::run_lattices)
            a.run_lattices(seg), b.run_lattices(seg)
        else:
            a.run_lattices_with_rewards(rewards[s0:s0 + seg]), b.run_lattices_with_rewards(rewards[s0:s0 + seg])
        for get in (lambda n: n.get_lattice(1), lambda n: n.get_reward_modulated_lattice(2)):
            ga, gb = get(a).grid_history.history, get(b).grid_history.history
            SC.assert_close_robust(ga[s0:s0 + seg], gb[s0:s0 + seg], 1e-4, 1e-3, f"voltages, steps {s0}..")
            ra, rb = get(a).spike_history.history, get(b).spike_history.history
            assert (ra[s0:s0 + seg] == rb[s0:s0 + seg]).mean() > 0.999
            spikes += int(rb[s0:s0 + seg].sum())
        assert (a.get_spike_train_lattice(0).spike_history.history == b.get_spike_train_lattice(0).spike_history.history).all()
        for blk in BLOCKS:
            wa, wb = a._be.get_connection_csr(*blk)[2], b._be.get_connection_csr(*blk)[2]
            SC.assert_close_robust(wa, wb, 2e-4, 1e-5, f"weights of block {blk}, steps {s0}..", max_abs=0.05)
            if blk[1] == 2:
                ta, tb = a._be.connection_traces(*blk), b._be.connection_traces(*blk)
                assert (ta[0] == tb[0]).all()
                SC.assert_close_robust(ta[1], tb[1], 2e-4, 1e-6, f"dw of block {blk}", max_abs=0.05)
                SC.assert_close_robust(ta[2], tb[2], 2e-4, 1e-6, f"c of block {blk}", max_abs=0.05)
        da = a.get_reward_modulated_lattice(2).reward_modulator.dopamine
        db = b.get_reward_modulated_lattice(2).reward_modulator.dopamine
        assert da == pytest.approx(db, rel=1e-5)
        _sync(b, a)
    assert spikes > (50 if mode[0] else 10)   # chemical-only input drives the Izhikevich cells weakly


@pytest.mark.gpu
def test_refused_configurations_gpu():
    check_refused_configurations(None, None)


@pytest.mark.gpu
def test_single_reward_lattice_network_and_do_modulation_off(oracle_lattice_factory, oracle_network_factory):
    """A network whose only neuron lattice is reward-modulated (the staged single-lattice step kernels stay eligible), and
    do_modulation = False freezing every TraceRSTDP."""
    def build(lf, nf, modulate):
        rng = np.random.default_rng(9)
        lat = S.RewardModulatedLattice(S.IzhikevichNeuron, id=4, backend_factory=lf)
        lat.populate(S.IzhikevichNeuron(gap_conductance=5.0, c_m=10.0), 12, 16)
        lat.connect(lambda x, y: x != y and abs(x[0] - y[0]) <= 1 and abs(x[1] - y[1]) <= 1, lambda x, y: 0.4)
        lat.set_field("current_voltage", rng.uniform(-65, 25, 192).astype(f32))
        lat.reward_modulator = S.RewardModulatedSTDP(tau_c=0.1, a_plus=0.002, a_minus=0.0025)
        lat.do_modulation = modulate
        lat.update_grid_history = True
        return S.RewardModulatedLatticeNetwork.generate_network([], [lat], [], backend_factory=nf)
    for modulate in (True, False):
        a, b = build(None, None, modulate), build(oracle_lattice_factory, oracle_network_factory, modulate)
        w0 = b._be.get_connection_csr(4, 4)[2].copy()
        r = rewards_for(60, seed=8)
        a.run_lattices_with_rewards(r), b.run_lattices_with_rewards(r)
        ga, gb = (n.get_reward_modulated_lattice(4).grid_history.history for n in (a, b))
        SC.assert_close_robust(ga, gb, 1e-4, 1e-3, "voltages")
        wa, wb = a._be.get_connection_csr(4, 4)[2], b._be.get_connection_csr(4, 4)[2]
        SC.assert_close_robust(wa, wb, 2e-4, 1e-5, "weights", max_abs=0.05)
        assert ((wb != w0).any()) == modulate
        ta, tb = a._be.connection_traces(4, 4), b._be.connection_traces(4, 4)
        assert (ta[0] == tb[0]).all() and (tb[0] == 0).all()
        SC.assert_close_robust(ta[2], tb[2], 2e-4, 1e-6, "c", max_abs=0.05)


# ------------------------------------------------------------------ several lattices of every kind
MANY_TRAINS, MANY_PLAIN, MANY_REWARD = (10, 11), (1, 5), (3, 7)
MANY_BLOCKS = ((10, 1), (11, 5), (1, 5), (10, 3), (11, 3), (1, 3), (5, 3), (7, 3), (5, 7), (11, 7), (1, 1), (5, 5), (3, 3), (7, 7))
MANY_RM = {(10, 3), (5, 3), (7, 3), (11, 7)}


def build_many(lattice_factory, network_factory, seed=17):
    """Two spike-train lattices, two plain lattices and two reward-modulated lattices with interleaved ids: every presynaptic
    class index above 0 is exercised.  Lattice 7 does not modulate, so it may feed lattice 3."""
    rng = np.random.default_rng(seed)
    base = S.IzhikevichNeuron(gap_conductance=5.0, c_m=10.0)
    shapes = {1: (3, 3), 5: (2, 4), 3: (3, 3), 7: (2, 3), 10: (2, 3), 11: (2, 2)}
    lats = {}
    for lid in MANY_PLAIN + MANY_REWARD:
        cls_ = S.RewardModulatedLattice if lid in MANY_REWARD else S.Lattice
        L = cls_(S.IzhikevichNeuron, id=lid, backend_factory=lattice_factory)
        L.populate(base, *shapes[lid])
        n = shapes[lid][0] * shapes[lid][1]
        L.connect(lambda x, y: x != y and abs(x[0] - y[0]) + abs(x[1] - y[1]) <= 2, lambda x, y: 0.3 + 0.01 * lid)
        L.set_field("current_voltage", rng.uniform(-65, 20, n).astype(f32))
        L.set_field("b", rng.uniform(0.25, 0.33, n).astype(f32))
        L.update_grid_history = L.update_spike_history = True
        lats[lid] = L
    lats[1].plasticity = S.STDP(a_plus=0.002, a_minus=0.0015, tau_plus=3.5)
    lats[5].plasticity = S.STDP(a_plus=0.001, a_minus=0.003, tau_minus=6.0)
    lats[3].reward_modulator = S.RewardModulatedSTDP(tau_c=0.1, a_plus=0.002, a_minus=0.0025, tau_plus=5.0, tau_minus=4.0)
    lats[7].reward_modulator = S.RewardModulatedSTDP(tau_c=0.2, a_plus=0.004, a_minus=0.001)
    lats[7].do_modulation = False
    trains = {}
    for tid in MANY_TRAINS:
        st = S.SpikeTrainLattice(S.RateSpikeTrain, id=tid, network_backend_factory=network_factory)
        st.populate(S.RateSpikeTrain(rate=2.0), *shapes[tid])
        st.set_field("rate", rng.choice([1.5, 2.0, 3.0], shapes[tid][0] * shapes[tid][1]).astype(f32))
        st.update_spike_history = True
        trains[tid] = st
    net = S.RewardModulatedLatticeNetwork.generate_network([lats[i] for i in MANY_PLAIN], [lats[i] for i in MANY_REWARD],
                                                           [trains[i] for i in MANY_TRAINS], backend_factory=network_factory)
    for pre, post in MANY_BLOCKS:
        if pre == post:
            continue
        wmat = rng.uniform(0.1, 1.0, (8, 8)).astype(f32)
        cond = lambda x, y, s=(pre + post): (x[0] + 2 * x[1] + y[0] + y[1] + s) % 3 != 0
        wl = lambda x, y, m=wmat: float(m[(x[0] * 3 + x[1]) % 8, (y[0] * 3 + y[1]) % 8])
        if post in MANY_REWARD or pre in MANY_REWARD:
            kind = RMC.RewardModulatedWeight if (pre, post) in MANY_RM else RMC.Weight
            net.connect_with_reward_modulation(pre, post, cond, lambda x, y, k=kind, f=wl:
                                               k(S.TraceRSTDP(weight=f(x, y))) if k == RMC.RewardModulatedWeight else k(f(x, y)))
        else:
            net.connect(pre, post, cond, wl)
    return net


def _lat(net, lid):
    return net.get_lattice(lid) or net.get_reward_modulated_lattice(lid) or net.get_spike_train_lattice(lid)


def test_many_lattices_oracle_rules(oracle_lattice_factory, oracle_network_factory):
    """Which blocks move and which stay, on the oracle: only edges INTO the modulating lattice 3 (its own graph,
    RewardModulatedWeight blocks, Weight blocks fed by plain lattices) change."""
    net = build_many(oracle_lattice_factory, oracle_network_factory)
    be = net._be
    w0 = {blk: be.get_connection_csr(*blk)[2].copy() for blk in MANY_BLOCKS}
    net.run_lattices_with_rewards(rewards_for(150))
    moved = {blk: bool((be.get_connection_csr(*blk)[2] != w0[blk]).any()) for blk in MANY_BLOCKS}
    assert moved == {blk: blk in {(10, 3), (1, 3), (5, 3), (7, 3), (3, 3)} for blk in MANY_BLOCKS}
    assert net.get_reward_modulated_lattice(7).reward_modulator.dopamine != 0.0   # every modulator takes the reward


@pytest.mark.gpu
def test_many_lattices_match_oracle(oracle_lattice_factory, oracle_network_factory):
    a, b = build_many(None, None), build_many(oracle_lattice_factory, oracle_network_factory)
    total, seg = 240, 40
    rewards = rewards_for(total, seed=4)
    for s0 in range(0, total, seg):
        a.run_lattices_with_rewards(rewards[s0:s0 + seg]), b.run_lattices_with_rewards(rewards[s0:s0 + seg])
        for lid in MANY_PLAIN + MANY_REWARD:
            ga, gb = _lat(a, lid).grid_history.history, _lat(b, lid).grid_history.history
            SC.assert_close_robust(ga[s0:s0 + seg], gb[s0:s0 + seg], 1e-4, 1e-3, f"voltages of lattice {lid}, steps {s0}..")
            assert (_lat(a, lid).spike_history.history[s0:] == _lat(b, lid).spike_history.history[s0:]).mean() > 0.999
        for tid in MANY_TRAINS:
            assert (_lat(a, tid).spike_history.history == _lat(b, tid).spike_history.history).all()
        for blk in MANY_BLOCKS:
            wa, wb = a._be.get_connection_csr(*blk)[2], b._be.get_connection_csr(*blk)[2]
            SC.assert_close_robust(wa, wb, 2e-4, 1e-5, f"weights of block {blk}, steps {s0}..", max_abs=0.05)
            if blk[1] in MANY_REWARD:
                ta, tb = a._be.connection_traces(*blk), b._be.connection_traces(*blk)
                assert (ta[0] == tb[0]).all(), blk
                SC.assert_close_robust(ta[1], tb[1], 2e-4, 1e-6, f"dw of block {blk}", max_abs=0.05)
                SC.assert_close_robust(ta[2], tb[2], 2e-4, 1e-6, f"c of block {blk}", max_abs=0.05)
        for lid in MANY_REWARD:
            assert _lat(a, lid).reward_modulator.dopamine == pytest.approx(_lat(b, lid).reward_modulator.dopamine, rel=1e-5)
        # re-synchronise a <- b
        for lid in MANY_PLAIN + MANY_REWARD + MANY_TRAINS:
            SC.copy_lattice_state(_lat(b, lid), _lat(a, lid), weights=False)
        for blk in MANY_BLOCKS:
            w = b._be.get_connection_csr(*blk)[2]
            if blk[1] in MANY_REWARD:
                cnt, dw, c = b._be.connection_traces(*blk)
                a._be.set_connection_traces(w, cnt, dw, c, pre_id=blk[0], post_id=blk[1])
            else:
                a._be.set_connection_traces(w, None, None, None, pre_id=blk[0], post_id=blk[1])
        for lid in MANY_REWARD:
            _lat(a, lid).reward_modulator.dopamine = _lat(b, lid).reward_modulator.dopamine
