import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
rows = cols = 3163
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
bench.configure(be, bench.init_fields(np, rows * cols, 0x5EED))
be.run_timed(300)
for it in (1, 2, 3, 12, 30, 30):
    be.run_timed(it)
