"""General graphs at the bench size: Izhikevich 3163 x 3163, electrical synapses, 8 random in-edges per neuron — drawn from a
(2r+1)^2 neighbourhood ("local", the random-radius graphs of tests/gpu_accuracy.rs grown) or from the whole lattice ("uniform").
These take the general step kernel (kernels.cu: every operand straight from HBM), not the TMA-staged stencil kernels."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200")):
    sys.path.insert(0, p)
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend

rows = cols = int(os.environ.get("ROWS", "3163"))
n, fan = rows * cols, 8
rng = np.random.default_rng(3)
out = {}
for name, radius in (("local_r4", 4), ("local_r32", 32), ("uniform", 0)):
    i, j = np.divmod(np.arange(n, dtype=np.int64), cols)
    if radius:
        a = np.clip(i[:, None] + rng.integers(-radius, radius + 1, (n, fan)), 0, rows - 1)
        b = np.clip(j[:, None] + rng.integers(-radius, radius + 1, (n, fan)), 0, cols - 1)
        pre = a * cols + b
    else:
        pre = rng.integers(0, n, (n, fan))
    pre = np.sort(pre, 1)
    keep = np.ones_like(pre, bool)
    keep[:, 1:] = pre[:, 1:] != pre[:, :-1]
    keep &= pre != np.arange(n)[:, None]
    rp = np.zeros(n + 1, np.uint64)
    rp[1:] = np.cumsum(keep.sum(1))
    be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
    f = bench.init_fields(np, n, 0x5EED)
    for fname, arr in f.items():
        if "$" not in fname:
            be.set_field(0, fname, arr)
    be.connect_csr(0, 0, rp, pre[keep].astype(np.uint32), rng.uniform(0.5, 1.5, int(keep.sum())).astype(np.float32))
    be.set_option(K.OPT_ELECTRICAL_SYNAPSE, 1); be.set_option(K.OPT_CHEMICAL_SYNAPSE, 0)
    be.run_timed(50)
    res = []
    for _ in range(3):
        ms, nl = be.run_timed(100)
        res.append(ms / 100 * 1e3)
    us = min(res)
    # DESIGN.md / SURVEY 8(d) accounting for Izhikevich electrical K = 8: state R+W 16 B, 9 parameters 36 B, CSR 4 + 8 * 8 = 68 B;
    # the gathered voltages are not counted (4 B per edge if every 32 B sector were shared, 32 B per edge when none is)
    algo = 120 * n
    out[name] = {"us_per_timestep": us, "neuron_steps_per_s": n / (us * 1e-6), "edges": int(keep.sum()),
                 "algorithmic_GBps": algo / (us * 1e-6) / 1e9, "launches_per_step": nl / 100}
    be.close()
    del pre, keep, rp
print(json.dumps(out))
