"""Hodgkin-Huxley lattice with AMPA / NMDA / GABA receptors (BASELINE.json configs[2] model) at a size that fills the GPU:
neuron-steps/s and the fraction of the 277 B/neuron-step HBM accounting it corresponds to.

    python tools/bench_hh.py [side] [quiet|default]

`quiet` (default): V starts in [-64, -57] mV; the first timed window (steps 20..170) is all-finite arithmetic.  Later windows and  `default`: V starts in
[-65, -50] as in the parity scenarios — then the reference's own arithmetic hits its singular points (n_alpha = 0/0 at exactly
V = -55 mV, m_alpha at -40 mV; hodgkin_huxley gates, ion_channels/mod.rs:219-281) within ~100 steps somewhere in a lattice of
this size, the NaN spreads one cell per step through the gap junctions, and from then on every IEEE division takes its slow
path: that state is what round 1's 476 us figure was measured on.  The CPU oracle reproduces the NaNs bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import scenarios as SC
rows = cols = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mode = sys.argv[2] if len(sys.argv) > 2 else "quiet"
lat = SC.build_lattice(None, model="hh", rows=rows, cols=cols, seed=13, graph="grid", chem="destexhe_all", history=False, gap=2.0)
n = rows * cols
if mode == "quiet":
    lat.set_field("current_voltage", np.random.default_rng(1).uniform(-64.0, -57.0, n).astype(np.float32))
lat._push_options()
lat._be.run_timed(20 if mode == "quiet" else 300)
for k in ((150,) if mode == "quiet" else ()) + (200, 200, 200):
    ms, nl = lat._be.run_timed(k)
    v = lat.get_field("current_voltage")
    print(f"HH {rows}x{cols} + 3 receptors [{mode}]: {ms / k * 1e3:.1f} us per step over {k} steps, {n * k / ms / 1e6:.2f} G neuron-steps/s, "
          f"{277 * n * k / ms / 1e6:.0f} GB/s of the 277 B accounting; finite after: {np.isfinite(v).mean():.4f}, V in [{np.nanmin(v):.1f}, {np.nanmax(v):.1f}]",
          flush=True)
