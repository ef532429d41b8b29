"""RewardModulatedLattice at the bench shape: Izhikevich 3163x3163, electrical synapses, 8-neighbour grid,
RewardModulatedSTDP over TraceRSTDP weights, modulation on.  Device time of the step loop (step kernel + per-edge modulator kernel)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200")):
    sys.path.insert(0, p)
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
rows = cols = int(os.environ.get("ROWS", "3163"))
n = rows * cols
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
f = bench.init_fields(np, n, 0x5EED)
for name, arr in f.items():
    if "$" not in name:
        be.set_field(0, name, arr)
be.connect_grid(0, 1, 1.0)
be.set_option(0, 1); be.set_option(1, 0)
be.set_reward_modulator(True, True, dopamine=0.5, tau_d=20.0, tau_c=0.05, a_plus=0.02, a_minus=0.02, tau_plus=4.5, tau_minus=4.5, dt=0.1)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
be.run_timed(600)
res = []
for _ in range(3):
    ms, nl = be.run_timed(steps)
    res.append(ms / steps * 1e3)
us = min(res)
cnt, dw, c = (None, None, None)
out = {"us_per_timestep": us, "neuron_steps_per_s": n / (us * 1e-6), "launches_per_step": nl / steps,
       # full TraceRSTDP traffic (col 4, weight RW 8, counter RW 2, dw RW 8, c RW 8 per edge) and what is left of it while the
       # traces are in the canonical state (counter == 0 and dw == 0 between timesteps: neither array is touched)
       "algorithmic_bytes_per_neuron_step": {"general": 120 + 8 + 8 * 30, "canonical": 120 + 8 + 8 * 20},
       "achieved_GBps_canonical": (120 + 8 + 8 * 20) * n / (us * 1e-6) / 1e9}
print(json.dumps(out))
