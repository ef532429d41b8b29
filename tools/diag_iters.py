"""Per-timestep device time of the bench workload as a function of timesteps per run (sustained vs isolated launches)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
rows = cols = int(os.environ.get("ROWS", "3163"))
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
bench.configure(be, bench.init_fields(np, rows * cols, 0x5EED))
be.run_timed(50)
for iters in (1, 2, 5, 10, 50, 200, 1000, 1, 5, 200):
    reps = max(1, 400 // iters)
    tot = 0.0
    for _ in range(reps):
        ms, nl = be.run_timed(iters)
        tot += ms
    print(f"iters {iters:5d} reps {reps:4d}: {1e3 * tot / (reps * iters):8.1f} us per timestep (incl. flush_stdp per run)", flush=True)
