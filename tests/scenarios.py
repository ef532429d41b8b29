"""Scenario builders shared by the CPU and GPU parity tests: one description, any back end."""
from __future__ import annotations

import numpy as np

import snn_b200 as S
from snn_b200 import _capi as K

f32 = np.float32

MODELS = {
    "lif": S.LeakyIntegrateAndFireNeuron,
    "qif": S.QuadraticIntegrateAndFireNeuron,
    "adlif": S.AdaptiveLeakyIntegrateAndFireNeuron,
    "adex": S.AdaptiveExpLeakyIntegrateAndFireNeuron,
    "izh": S.IzhikevichNeuron,
    "leaky_izh": S.LeakyIzhikevichNeuron,
    "simple_lif": S.SimpleLeakyIntegrateAndFire,
    "hh": S.HodgkinHuxleyNeuron,
    "bcm_izh": S.BCMIzhikevichNeuron,
}
# models whose step uses only + - * / and comparisons: rasters and voltages must match the oracle bit for bit
EXACT_MODELS = ["lif", "qif", "adlif", "izh", "leaky_izh", "simple_lif", "bcm_izh"]


def random_graph(rows, cols, seed, radius=2.0, p=0.8, weights="ones"):
    """tests/gpu_accuracy.rs:28-32 style: connect within `radius` with probability p, x != y."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    pos = [(i, j) for i in range(rows) for j in range(cols)]
    conn = np.zeros((n, n), np.uint32)
    for a, x in enumerate(pos):
        for b, y in enumerate(pos):
            if a != b and np.hypot(x[0] - y[0], x[1] - y[1]) <= radius and rng.random() <= p:
                conn[a, b] = 1
    if weights == "ones":
        w = conn.astype(f32)
    else:
        w = (rng.uniform(0.2, 2.0, (n, n)).astype(f32)) * conn
    return conn, w


def dense_to_csr(conn, w):
    n_pre, n_post = conn.shape
    row_ptr = np.zeros(n_post + 1, np.uint64)
    pre, ww = [], []
    for q in range(n_post):
        idx = np.nonzero(conn[:, q])[0]
        pre.extend(idx.tolist())
        ww.extend(w[idx, q].tolist())
        row_ptr[q + 1] = len(pre)
    return row_ptr, np.array(pre, np.uint32), np.array(ww, f32)


def add_chemistry(neuron, chem):
    """chem: None | 'approx_ampa' | 'approx_all' | 'destexhe_all' | 'expdecay_all' | 'discrete_ampa'."""
    T = S.IonotropicNeurotransmitterType
    if chem is None:
        return
    kind, which = chem.split("_")
    types = [T.AMPA] if which == "ampa" else [T.AMPA, T.NMDA, T.GABA]
    nt_cls = {"approx": S.ApproximateNeurotransmitter, "destexhe": S.DestexheNeurotransmitter,
              "expdecay": S.ExponentialDecayNeurotransmitter, "discrete": S.DiscreteSpikeNeurotransmitter}[kind]
    rc_cls = {"approx": S.ApproximateReceptor, "destexhe": S.DestexheReceptor, "expdecay": S.ExponentialDecayReceptor,
              "discrete": S.ApproximateReceptor}[kind]
    rec = {T.AMPA: S.AMPAReceptor, T.NMDA: S.NMDAReceptor, T.GABA: S.GABAReceptor}
    for ty in types:
        neuron.synaptic_neurotransmitters[ty] = nt_cls()
        neuron.receptors[ty] = rec[ty](r=rc_cls())


def build_lattice(factory, model="izh", rows=6, cols=7, seed=1, graph="grid", chem=None, stdp=False,
                  electrical=True, chemical=None, gap=10.0, hetero=True, history=True, c_m=None, history_type=None, cls=None):
    lattice_cls = cls
    cls = MODELS[model]
    base = cls(gap_conductance=gap)
    if c_m is not None:
        base.c_m = c_m
    add_chemistry(base, chem)
    lat = (lattice_cls or S.Lattice)(cls, backend_factory=factory, **({"history_type": history_type} if history_type else {}))
    lat.populate(base, rows, cols)
    n = rows * cols
    rng = np.random.default_rng(seed)
    if n:
        lo = -65.0 if model in ("izh", "leaky_izh", "hh", "bcm_izh") else -75.0
        hi = {"izh": 30.0, "leaky_izh": 30.0, "hh": -50.0, "bcm_izh": 30.0}.get(model, -55.0)
        lat.set_field("current_voltage", rng.uniform(lo, hi, n).astype(f32))
        if hetero:
            lat.set_field("gap_conductance", (gap * rng.uniform(0.5, 1.5, n)).astype(f32))
            # tonic firing without an external current: Izhikevich b > 0.27 removes the fixed point; leak reversal
            # above threshold; QIF reset above the critical voltage (same recipe as tests/golden/make_golden.py)
            if model == "bcm_izh":
                lat.set_field("firing_rate_window", rng.choice([1.5, 2.0, 3.25], n).astype(f32))
                lat.set_field("period", rng.choice([2, 3, 5], n).astype(np.uint32))
            if model in ("izh", "leaky_izh", "bcm_izh"):
                lat.set_field("a", rng.uniform(0.015, 0.03, n).astype(f32))
                lat.set_field("b", rng.uniform(0.25, 0.36, n).astype(f32))
                lat.set_field("d", rng.uniform(6.0, 9.0, n).astype(f32))
                if c_m is None:
                    lat.fill_field("c_m", 2.0)
            if model == "leaky_izh":
                lat.set_field("w_value", rng.uniform(0.0, 1.0, n).astype(f32))
            if model in ("lif", "adlif", "adex"):
                lat.set_field("e_l", rng.uniform(-62, -42, n).astype(f32))
                if c_m is None:
                    lat.fill_field("c_m", 10.0)
            if model in ("lif", "qif", "adlif", "adex"):
                lat.set_field("tref", rng.choice([0.5, 1.0, 2.0], n).astype(f32))
            if model == "qif":
                # reset above a lowered threshold: fires whenever it is not refractory
                lat.fill_field("v_th", -58.0)
                lat.set_field("v_reset", rng.uniform(-57.5, -56, n).astype(f32))
                lat.set_field("current_voltage", rng.uniform(-62, -56, n).astype(f32))
                lat.fill_field("tau_m", 10.0)
    if graph == "grid":
        lat.connect_grid(1, 1.0)
    elif graph == "grid2":
        lat.connect_grid(2, 0.5)
    elif graph == "random":
        conn, w = random_graph(rows, cols, seed + 100, weights="rand")
        lat._be.connect_dense(lat._bid, lat._bid, conn, w)
        lat._graph_spec = ("dense", conn, w)
    elif graph == "csr":
        conn, w = random_graph(rows, cols, seed + 100, weights="rand")
        lat.connect_csr(*dense_to_csr(conn, w))
    elif graph == "all":
        lat.connect(lambda x, y: x != y, lambda x, y: 2.0)
    elif graph == "none":
        pass
    else:
        raise ValueError(graph)
    lat.electrical_synapse = electrical
    lat.chemical_synapse = (chem is not None) if chemical is None else chemical
    lat.do_plasticity = stdp
    lat.update_grid_history = history
    lat.update_spike_history = history
    return lat


def drive_current(model):
    """A gap conductance that makes each model spike within a few hundred steps from random initial V."""
    return {"hh": 2.0}.get(model, 10.0)


def compare_lattices(a, b, exact=True, rtol=1e-4, atol=1e-3, fields=("current_voltage",), check_raster=True):
    """a = device under test, b = oracle."""
    if a.update_grid_history:
        ha, hb = a.grid_history.history, b.grid_history.history
        assert ha.shape == hb.shape
        if exact:
            bad = np.nonzero(ha != hb)
            assert bad[0].size == 0, f"first voltage mismatch at step {bad[0][0]} cell ({bad[1][0]},{bad[2][0]}): {ha[bad][0]!r} vs {hb[bad][0]!r}"
        else:
            np.testing.assert_allclose(ha, hb, rtol=rtol, atol=atol)
    if check_raster and a.update_spike_history:
        sa, sb = a.spike_history.history, b.spike_history.history
        assert sa.shape == sb.shape
        assert (sa == sb).all(), f"raster differs at steps {np.unique(np.nonzero(sa != sb)[0])[:5]}"
    for name in fields:
        fa, fb = a.get_field(name), b.get_field(name)
        if exact or fa.dtype.kind in "iu":
            assert (fa == fb).all(), name
        else:
            np.testing.assert_allclose(fa, fb, rtol=rtol, atol=atol, err_msg=name)


def all_field_names(lat):
    names = list(lat.neuron_type().scalar_fields())
    return names


# ------------------------------------------------------------------ lock-step comparison for chaotic scenarios
def lattice_field_names(lat):
    """Every named field that currently carries state for this lattice / spike-train lattice."""
    from snn_b200.lattice import _NT_PARAM_FIELDS, _RC_KIN_FIELDS, _TYPE_NAMES
    if hasattr(lat, "neuron_type"):
        names = list(lat.neuron_type().scalar_fields())
    else:
        names = list(lat.spike_train_type().scalar_fields()) + ["neural_refractoriness$k"]
    if getattr(lat, "_chem_touched", False):
        names += ["neurotransmitters$flags"] + [f"neurotransmitters${f}" for f in ["t", "t_max"] + _NT_PARAM_FIELDS[lat._ntk]]
    if getattr(lat, "_rc_touched", False):
        names.append("receptors$flags")
        for ty in range(3):
            names += [f"receptors${_TYPE_NAMES[ty]}_{f}" for f in S.RECEPTORS[S.IonotropicNeurotransmitterType(ty)]._defaults]
            names += [f"receptors${_TYPE_NAMES[ty]}$r$kinetics${f}" for f in _RC_KIN_FIELDS[lat._rck]]
    return names


def copy_lattice_state(src, dst, weights=True):
    """dst <- src: every field, the clock and (optionally) the graph weights."""
    for name in lattice_field_names(src):
        dst.set_field(name, src.get_field(name))
    if weights and hasattr(src, "neuron_type"):
        c, w = src.graph_dense()
        dst._be.connect_dense(dst._bid, dst._bid, c, w)


def assert_close_robust(x, y, rtol, atol, msg, max_frac=0.03, max_abs=1.0):
    """allclose, except that a few elements may deviate more: on the up-stroke of a spike dv/dv' > 1 per step, so an
    ulp-level difference in expf can reach a fraction of a millivolt for the handful of neurons that are about to
    fire, even inside one short segment.  Bounded in count (3 %) and size (1 mV); rasters are compared exactly."""
    # default-strength STDP has no weight clamp (plasticity/mod.rs:64): some scenarios run away to inf/NaN in the
    # reference algorithm itself; non-finite values must then appear at the same places with the same sign
    fin = np.isfinite(x) & np.isfinite(y)
    assert (np.isnan(x) == np.isnan(y)).all() and (np.isposinf(x) == np.isposinf(y)).all() and \
        (np.isneginf(x) == np.isneginf(y)).all(), f"{msg}: non-finite values differ"
    with np.errstate(invalid="ignore"):
        err = np.where(fin, np.abs(x - y), 0.0)
        bad = err > (atol + rtol * np.where(fin, np.abs(y), 0.0))
    assert bad.mean() <= max_frac, f"{msg}: {bad.sum()} of {bad.size} elements off (max {err.max()})"
    assert err.max() <= max_abs, f"{msg}: max deviation {err.max()} mV"


def lockstep_lattices(a, b, total, segment, rtol=1e-4, atol=1e-3, weights=False):
    """Run device `a` and oracle `b` in segments, compare each segment tightly, then re-synchronise a <- b.
    Spiking lattices are chaotic: a 1-ulp expf difference grows to millivolts within a few hundred steps (the
    oracle and the independent numpy restatement diverge the same way), so long-horizon parity is checked
    segment by segment from identical state."""
    done = 0
    spikes = 0
    while done < total:
        n = min(segment, total - done)
        a.run_lattice(n)
        b.run_lattice(n)
        ha, hb = a.grid_history.history[done:done + n], b.grid_history.history[done:done + n]
        assert_close_robust(ha, hb, rtol, atol, f"segment starting at step {done}")
        sa, sb = a.spike_history.history[done:done + n], b.spike_history.history[done:done + n]
        assert (sa == sb).all(), f"raster differs in the segment starting at step {done}"
        spikes += int(sb.sum())
        if weights:
            (ca, wa), (cb, wb) = a.graph_dense(), b.graph_dense()
            assert (ca == cb).all()
            np.testing.assert_allclose(wa, wb, rtol=1e-5, atol=1e-6, err_msg=f"weights after step {done + n}")
        assert a.internal_clock == b.internal_clock
        copy_lattice_state(b, a, weights=weights)
        done += n
    return spikes
