"""Step time vs spiking activity along the bench trajectory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
rows = cols = 3163
n = rows * cols
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
bench.configure(be, bench.init_fields(np, n, 0x5EED))
clock = 0
CH = 50
for c in range(80):
    ms, nl = be.run_timed(CH)
    clock += CH
    if c % 4 == 0 or c < 8:
        lft = be.get_field(0, "last_firing_time")
        recent = int((lft >= clock - CH).sum())
        print(f"steps {clock - CH:5d}..{clock:5d}: {1e3 * ms / CH:7.1f} us/step  spikes/neuron/step {recent / n / CH:.5f}", flush=True)
