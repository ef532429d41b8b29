"""Step rates of the other BASELINE.json configurations (parity-test cases, not bench lines):
C1 Izhikevich 100x100 electrical, C2 LIF 1000x1000 + STDP, C3 HH 256x256 + AMPA/NMDA/GABA (Destexhe), C4 MNIST-shaped
network (784 Poisson -> 400 exc <-> 400 inh, STDP).  Device time from CUDA events on the engine stream; the CPU column is
the oracle port (single thread) on the same configuration, bounded to a few seconds.

    python tools/bench_configs.py [--cpu]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import snn_b200 as S
import scenarios as SC

f32 = np.float32


def lattice(factory, cfg):
    if cfg == "C1":
        lat = SC.build_lattice(factory, model="izh", rows=100, cols=100, seed=11, graph="grid", hetero=False, c_m=5.0, history=False)
    elif cfg == "C2":
        lat = SC.build_lattice(factory, model="lif", rows=1000, cols=1000, seed=12, graph="grid", stdp=True, history=False, gap=40.0)
    else:
        lat = SC.build_lattice(factory, model="hh", rows=256, cols=256, seed=13, graph="grid", chem="destexhe_all", history=False, gap=2.0)
    return lat


def network(lfac, nfac):
    rng = np.random.default_rng(1)
    T = S.IonotropicNeurotransmitterType
    exc_base = S.IzhikevichNeuron(gap_conductance=5.0, c_m=10.0)
    exc_base.synaptic_neurotransmitters[T.AMPA] = S.ApproximateNeurotransmitter()
    exc_base.receptors[T.AMPA] = S.AMPAReceptor(); exc_base.receptors[T.GABA] = S.GABAReceptor()
    inh_base = exc_base.clone()
    inh_base.synaptic_neurotransmitters = {T.GABA: S.ApproximateNeurotransmitter()}
    kw = {"backend_factory": lfac} if lfac else {}
    exc = S.Lattice(S.IzhikevichNeuron, id=1, **kw); exc.populate(exc_base, 20, 20)
    inh = S.Lattice(S.IzhikevichNeuron, id=2, **kw); inh.populate(inh_base, 20, 20)
    for L in (exc, inh):
        L.do_plasticity = True
    base = S.PoissonNeuron.from_firing_rate(40.0, 0.1)
    base.synaptic_neurotransmitters[T.AMPA] = S.ApproximateNeurotransmitter()
    st = S.SpikeTrainLattice(S.PoissonNeuron, id=0, **({"network_backend_factory": nfac} if nfac else {}))
    st.populate(base, 28, 28)
    st.set_field("chance_of_firing", (rng.uniform(0, 1, 784) * 0.02).astype(f32))
    net = S.LatticeNetwork.generate_network([exc, inh], [st], **({"backend_factory": nfac} if nfac else {}))
    w = rng.uniform(0, 1, (784, 400)).astype(f32)
    net.connect(0, 1, lambda x, y: True, lambda x, y: float(w[x[0] * 28 + x[1], y[0] * 20 + y[1]]))
    net.connect(1, 2, lambda x, y: x == y, lambda x, y: 1.0)
    net.connect(2, 1, lambda x, y: x != y, lambda x, y: -1.0)
    net.electrical_synapse, net.chemical_synapse = True, True
    return net


def measure(cpu=False):
    out = {}
    for cfg, steps in (("C1", 1000), ("C2", 1000), ("C3", 1000)):
        lat = lattice(None, cfg)
        lat._push_options()
        lat._be.run_timed(100)
        ms, nl = lat._be.run_timed(steps)
        n = lat.size
        out[cfg] = {"neurons": n, "us_per_step": 1e3 * ms / steps, "neuron_steps_per_s": n * steps / (ms * 1e-3), "launches": nl}
        if cfg == "C3":
            # the long run above ends NaN-saturated (the reference's own gate formulas are 0/0 at exactly V = -55 / -40 mV, the NaN
            # spreads through the gap junctions: tools/bench_hh.py); steps 20..170 from a quiet start are all-finite arithmetic
            q = lattice(None, cfg)
            q.set_field("current_voltage", np.random.default_rng(1).uniform(-64.0, -57.0, n).astype(f32))
            q._push_options()
            q._be.run_timed(20)
            ms2, _ = q._be.run_timed(150)
            out[cfg]["us_per_step_finite_window"] = 1e3 * ms2 / 150
            out[cfg]["finite_after_window"] = float(np.isfinite(q.get_field("current_voltage")).mean())
            out[cfg]["finite_after_long_run"] = float(np.isfinite(lat.get_field("current_voltage")).mean())
        if cpu:
            from oracle_api import OracleBackend
            ol = lattice(lambda m, nt, rc, rows, cols: OracleBackend(m, nt, rc, rows=rows, cols=cols), cfg)
            k = {"C1": 200, "C2": 3, "C3": 20}[cfg]
            t0 = time.perf_counter(); ol.run_lattice(k); dt = time.perf_counter() - t0
            out[cfg]["cpu_oracle_us_per_step"] = 1e6 * dt / k
    net = network(None, None)
    net.run_lattices(100)
    be = net._be
    ms, nl = be.run_timed(3500)
    out["C4"] = {"nodes": 1584, "us_per_step": 1e3 * ms / 3500, "launches": nl}
    if cpu:
        from oracle_api import OracleBackend
        onet = network(lambda m, nt, rc, rows, cols: OracleBackend(m, nt, rc, rows=rows, cols=cols), lambda *a: OracleBackend(*a))
        t0 = time.perf_counter(); onet.run_lattices(50); dt = time.perf_counter() - t0
        out["C4"]["cpu_oracle_us_per_step"] = 1e6 * dt / 50
    return out


def main():
    print(json.dumps(measure(cpu="--cpu" in sys.argv), indent=1))


if __name__ == "__main__":
    main()
