"""10^8-neuron sanity check (one B200, ~25 GB): Izhikevich 10000 x 10000, electrical + AMPA + STDP, k steps; the k-step light cone of
three 16 x 16 windows is re-computed by the CPU oracle on (16 + 2k)^2 patches and must match bit for bit (same property as
tests/test_gpu_parity.py::test_full_size_10m_izhikevich_window_property, ten times the size: 64-bit offsets, 28-bit node indices)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import snn_b200 as S
import scenarios as SC
from oracle_api import OracleBackend

f32 = np.float32
rows = cols = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
k = 6
n = rows * cols
rng = np.random.default_rng(5)
t0 = time.time()
big = SC.build_lattice(None, model="izh", rows=rows, cols=cols, seed=0, graph="grid", hetero=False, history=False, chem="approx_ampa",
                       stdp=True, c_m=2.0)
init = {"current_voltage": rng.uniform(-65, 30, n).astype(f32), "b": rng.uniform(0.25, 0.36, n).astype(f32)}
for name, arr in init.items():
    big.set_field(name, arr)
print(f"built in {time.time() - t0:.1f} s", flush=True)
ms, nl = (big._push_options(), big._be.run_timed(k))[1]
print(f"{k} steps of {n:.3g} neurons: {ms / k * 1e3:.0f} us per step ({n * k / ms / 1e6:.2f} G neuron-steps/s)", flush=True)
fac = lambda m, nt, rc, r, c: OracleBackend(m, nt, rc, rows=r, cols=c)
fields = ("current_voltage", "w_value", "last_firing_time", "neurotransmitters$t", "receptors$AMPA$r$kinetics$r")
got = {name: big.get_field(name) for name in fields}
for (r0, c0) in [(0, 0), (rows // 2 + 3, cols // 2 - 5), (rows - 16, cols - 16)]:
    ra, rb, ca, cb = max(0, r0 - k), min(rows, r0 + 16 + k), max(0, c0 - k), min(cols, c0 + 16 + k)
    patch = SC.build_lattice(fac, model="izh", rows=rb - ra, cols=cb - ca, seed=0, graph="grid", hetero=False, history=False,
                             chem="approx_ampa", stdp=True, c_m=2.0)
    for name, arr in init.items():
        patch.set_field(name, arr.reshape(rows, cols)[ra:rb, ca:cb])
    patch.run_lattice(k)
    for name in fields:
        g = got[name].reshape(rows, cols, -1)[r0:r0 + 16, c0:c0 + 16]
        w = patch.get_field(name).reshape(rb - ra, cb - ca, -1)[r0 - ra:r0 - ra + 16, c0 - ca:c0 - ca + 16]
        assert (g == w).all(), (name, r0, c0)
print("BIG_LATTICE_OK", int((got["last_firing_time"] >= 0).sum()), "neurons have spiked")
big._be.run_timed(500)
for _ in range(8):
    ms, nl = big._be.run_timed(100)
    print(f"100 steps of {n:.3g} neurons: {ms / 100 * 1e3:.0f} us per step ({n * 100 / ms / 1e6:.2f} G neuron-steps/s, "
          f"{160 * n * 100 / ms / 1e6:.0f} GB/s algorithmic)", flush=True)
