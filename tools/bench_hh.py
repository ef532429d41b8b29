"""Hodgkin-Huxley lattice with AMPA / NMDA / GABA receptors (BASELINE.json configs[2] model) at a size that fills the GPU:
neuron-steps/s and the fraction of the 277 B/neuron-step HBM accounting it corresponds to."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import scenarios as SC
rows = cols = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
lat = SC.build_lattice(None, model="hh", rows=rows, cols=cols, seed=13, graph="grid", chem="destexhe_all", history=False, gap=2.0)
lat._push_options()
lat._be.run_timed(300)
n = rows * cols
for _ in range(3):
    ms, nl = lat._be.run_timed(200)
    print(f"HH {rows}x{cols} + 3 receptors: {ms / 200 * 1e3:.1f} us per step, {n * 200 / ms / 1e6:.2f} G neuron-steps/s, {277 * n * 200 / ms / 1e6:.0f} GB/s of the 277 B accounting", flush=True)
