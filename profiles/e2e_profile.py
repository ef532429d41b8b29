import sys, time, os
sys.path.insert(0, "spiking-neural-networks_b200"); sys.path.insert(0, ".")
import numpy as np, torch
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
n = 3163 * 3163
f = bench.init_fields(np, n, 1)
pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in f.items()}
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, 3163, 3163)
bench.configure(be, pinned); be.run(5)
for rep in range(2):
    tot = 0
    for name, arr in pinned.items():
        torch.cuda.synchronize(); t = time.perf_counter(); be.set_field(0, name, arr); dt = time.perf_counter() - t; tot += dt
        print(f"set {name:40s} {arr.nbytes/1e6:8.1f} MB {dt*1e3:8.2f} ms  {arr.nbytes/dt/1e9:6.1f} GB/s")
    t = time.perf_counter(); be.connect_grid(0, 1, 1.0); be.set_option(0,1); be.set_option(1,1); be.set_option(2,1,0); print("graph+opts", (time.perf_counter()-t)*1e3)
    t = time.perf_counter(); be.run(1); print("run(1) incl finalize_graph", (time.perf_counter()-t)*1e3)
    t = time.perf_counter(); be.run(10); print("run(10)", (time.perf_counter()-t)*1e3)
    for name in bench.STATE_FIELDS:
        t = time.perf_counter(); a = be.get_field(0, name); dt = time.perf_counter() - t
        print(f"get {name:40s} {a.nbytes/1e6:8.1f} MB {dt*1e3:8.2f} ms  {a.nbytes/dt/1e9:6.1f} GB/s")
    print("total set", tot*1e3)
