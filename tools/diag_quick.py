"""Steady-state step time of the bench workload (after 1000 steps), for A/B runs of library variants."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
rows = cols = int(os.environ.get("ROWS", "3163"))
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
bench.configure(be, bench.init_fields(np, rows * cols, 0x5EED))
be.run_timed(100)
a, _ = be.run_timed(100)
be.run_timed(800)
b = [be.run_timed(500)[0] / 500 * 1e3 for _ in range(3)]
print(f"{os.environ.get('SNN_B200_LIB', 'default'):>40s}: early {a / 100 * 1e3:6.1f} us/step, steady {min(b):6.1f} .. {max(b):6.1f} us/step", flush=True)
