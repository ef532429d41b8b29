"""Generates the golden fixtures tests/golden/*.npz with the independent numpy float32 restatement
(oracle/numpy_ref.py).  The Rust reference cannot be built or imported in this image (no
rustc/cargo), so these vectors pin the C oracle (and the CUDA path) against a second, separately
written restatement of the same reference equations rather than against the Rust binary.

    python tests/golden/make_golden.py          # rewrites the .npz files (deterministic seeds)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy_ref as R  # noqa: E402

f32 = np.float32

DEFAULTS = {
    R.LIF: dict(current_voltage=-75, refractory_count=0, leak_constant=-1, integration_constant=1, gap_conductance=7,
                v_th=-55, v_reset=-75, tau_m=10, c_m=100, g_l=10, e_l=-75, tref=10, dt=0.1),
    R.QIF: dict(current_voltage=-75, refractory_count=0, integration_constant=1, gap_conductance=7, alpha=1, v_th=-55,
                v_reset=-75, v_c=-60, tau_m=100, c_m=100, tref=10, dt=0.1),
    R.ADLIF: dict(current_voltage=-75, refractory_count=0, leak_constant=-1, integration_constant=1, gap_conductance=7,
                  w_value=0, alpha=6, beta=10, v_th=-55, v_reset=-75, tau_m=10, c_m=100, g_l=10, e_l=-75, tref=10, dt=0.1),
    R.ADEX: dict(current_voltage=-75, refractory_count=0, leak_constant=-1, integration_constant=1, gap_conductance=7,
                 w_value=0, alpha=6, beta=10, slope_factor=1, v_th=-55, v_reset=-75, tau_m=10, c_m=100, g_l=10, e_l=-75,
                 tref=10, dt=0.1),
    R.IZH: dict(current_voltage=-65, gap_conductance=7, w_value=30, a=0.02, b=0.2, c=-55, d=8, v_th=30, tau_m=1, c_m=100,
                dt=0.1),
    R.LEAKY_IZH: dict(current_voltage=-65, gap_conductance=7, w_value=30, a=0.02, b=0.2, c=-55, d=8, v_th=30, tau_m=10,
                      c_m=100, e_l=-65, dt=0.1),
    R.BCM_IZH: dict(current_voltage=-65, gap_conductance=7, w_value=30, a=0.02, b=0.2, c=-55, d=8, v_th=30, tau_m=1, c_m=100,
                    dt=0.1, average_activity=0, current_activity=0, period=3, num_spikes=0, firing_rate_clock=0,
                    firing_rate_window=500),
    R.SIMPLE_LIF: dict(current_voltage=-75, gap_conductance=10, v_th=-55, v_reset=-75, c_m=100, g=-0.1, e=0, dt=0.1),
    R.HH: dict(current_voltage=-65, gap_conductance=7, dt=0.01, c_m=1, v_th=0, g_na=120, e_na=50, g_k=36, e_k=-77,
               g_k_leak=0.3, e_k_leak=-55, m=0, h=0, n=0),
}


def moore(rows, cols):
    n = rows * cols
    conn = np.zeros((n, n), bool)
    for i in range(rows):
        for j in range(cols):
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    a, b = i + di, j + dj
                    if (di or dj) and 0 <= a < rows and 0 <= b < cols:
                        conn[a * cols + b, i * cols + j] = True
    return conn


def random_conn(n, seed, p=0.3):
    rng = np.random.default_rng(seed)
    c = rng.random((n, n)) < p
    np.fill_diagonal(c, False)
    return c


def make(name, model, rows, cols, steps, seed, graph="moore", chem=None, stdp=False, gap=10.0, electrical=True, c_m=None, stdp_a=None):
    rng = np.random.default_rng(seed)
    n = rows * cols
    conn = moore(rows, cols) if graph == "moore" else random_conn(n, seed + 7)
    w = np.where(conn, rng.uniform(0.5, 1.5, (n, n)), 0).astype(f32)
    fields = {k: np.full(n, v, np.uint32 if k in ("period", "num_spikes") else f32) for k, v in DEFAULTS[model].items()}
    lo, hi = (-65, 30) if model in (R.IZH, R.LEAKY_IZH, R.BCM_IZH) else ((-65, -50) if model == R.HH else (-75, -55))
    fields["current_voltage"] = rng.uniform(lo, hi, n).astype(f32)
    fields["gap_conductance"] = (gap * rng.uniform(0.5, 1.5, n)).astype(f32)
    # make the lattices fire tonically without an external current (the reference drives lattices only
    # through synapses): Izhikevich b > 0.27 removes the fixed point; leak reversal above threshold; QIF reset
    # above the critical voltage
    if model == R.BCM_IZH:
        # short, per-neuron firing-rate windows so that the activity averages roll over many times in the fixture
        fields["firing_rate_window"] = rng.choice([1.5, 2.0, 3.25], n).astype(f32)
        fields["period"] = rng.choice([2, 3, 5], n).astype(np.uint32)
    if model in (R.IZH, R.LEAKY_IZH, R.BCM_IZH):
        fields["b"] = rng.uniform(0.25, 0.36, n).astype(f32)
        fields["a"] = rng.uniform(0.015, 0.03, n).astype(f32)
        fields["c_m"] = np.full(n, 2, f32)
    if model == R.LEAKY_IZH:
        fields["w_value"] = rng.uniform(0.0, 1.0, n).astype(f32)
    if model in (R.LIF, R.ADLIF, R.ADEX):
        fields["e_l"] = rng.uniform(-62, -42, n).astype(f32)
        fields["tref"] = rng.choice([0.5, 1.0, 2.0], n).astype(f32)
        fields["c_m"] = np.full(n, 10, f32)
    if model == R.QIF:
        # reset above a lowered threshold: fires whenever it is not refractory
        fields["v_th"] = np.full(n, -58, f32)
        fields["v_reset"] = rng.uniform(-57.5, -56, n).astype(f32)
        fields["current_voltage"] = rng.uniform(-62, -56, n).astype(f32)
        fields["tref"] = rng.choice([0.5, 1.0, 2.0], n).astype(f32)
        fields["tau_m"] = np.full(n, 10, f32)
    if c_m is not None:
        fields["c_m"] = np.full(n, c_m, f32)
    L = R.DenseLattice(model, n, conn, w, fields)
    L.electrical = electrical
    if chem:
        L.chemical = True
        L.ntk = L.rck = 1 if chem == "destexhe" else 0
        L.nt_flags[:] = True
        L.rc_flags[:] = True
        if chem == "ampa":
            L.nt_flags[:, 1:] = False
            L.rc_flags[:, 1:] = False
    L.do_plasticity = stdp
    if stdp_a is not None:
        L.stdp["a_plus"] = L.stdp["a_minus"] = f32(stdp_a)
    inputs = dict(model=model, rows=rows, cols=cols, steps=steps, conn=conn.astype(np.uint32), w=w.copy(),
                  chem=np.array(chem or ""), stdp=stdp, electrical=electrical,
                  stdp_a=f32(L.stdp["a_plus"]),
                  nt_flags=L.nt_flags.astype(np.uint32), rc_flags=L.rc_flags.astype(np.uint32),
                  **{"f_" + k: v.copy() for k, v in L.f.items()})
    L.run(steps)
    out = dict(v_hist=np.array(L.v_hist, f32), s_hist=np.array(L.s_hist, np.uint8), w_final=L.w, lft=L.lft.astype(np.int32),
               t_final=L.nt["t"], r_final=L.rc["r"], **{"o_" + k: v for k, v in L.f.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **inputs, **out)
    print(name, "spikes:", int(np.sum(L.s_hist)))


if __name__ == "__main__":
    make("izh_moore", R.IZH, 5, 6, 600, 1)
    make("izh_random_stdp", R.IZH, 4, 5, 600, 2, graph="random", stdp=True)
    make("lif_moore_stdp", R.LIF, 5, 5, 800, 3, stdp=True, gap=40.0)
    make("qif_random", R.QIF, 4, 4, 800, 4, graph="random", gap=30.0)
    make("adlif_moore", R.ADLIF, 4, 5, 800, 5, gap=40.0)
    make("adex_moore", R.ADEX, 4, 4, 600, 6, gap=40.0)
    make("leaky_izh_moore", R.LEAKY_IZH, 4, 4, 600, 7)
    make("simple_lif_random", R.SIMPLE_LIF, 4, 4, 600, 8, graph="random", gap=5.0)
    make("izh_chem_ampa", R.IZH, 4, 5, 600, 9, chem="ampa")
    make("izh_chem_all_stdp", R.IZH, 4, 4, 600, 10, chem="all", stdp=True, c_m=12, stdp_a=0.05)
    make("bcm_izh_moore", R.BCM_IZH, 4, 5, 600, 13)
    make("bcm_izh_chem_ampa", R.BCM_IZH, 4, 4, 600, 14, chem="ampa")
    make("hh_moore", R.HH, 3, 4, 3000, 11, gap=2.0)
    make("hh_chem_destexhe", R.HH, 3, 3, 3000, 12, chem="destexhe", gap=2.0)
