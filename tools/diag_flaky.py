"""Repeat the 4-rank general-graph partition on one device many times; report time-outs (hardware-queue aliasing hypothesis)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import scenarios as SC
from snn_b200 import _capi as K
from snn_b200.dist import LocalStrips
import snn_b200 as S
rows, cols, world = 12, 9, int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = rows * cols
conn, w = SC.random_graph(rows, cols, 41, radius=2.5, p=0.6, weights="rand")
rp, pre, ww = SC.dense_to_csr(conn, w)
bad = 0
keep = []
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    parts = LocalStrips(K.MODEL_IZH, rows, cols, world)
    parts.each(lambda be: be.set_option(K.OPT_HALO_TIMEOUT_MS, 1500))
    parts.set_field("current_voltage", np.random.default_rng(rep).uniform(-65, 30, n).astype(np.float32))
    parts.each(lambda be: be.set_option(K.OPT_ELECTRICAL_SYNAPSE, 1))
    parts.set_graph_csr(rp, pre, ww)
    parts.attach_general()
    t0 = time.time()
    try:
        for k in (21, 1, 20):
            parts.run(k)
    except S.SnnError as e:
        bad += 1
        print("rep", rep, "FAILED after", round(time.time() - t0, 2), "s:", str(e)[:80], flush=True)
    if rep % 3 == 0:
        keep.append(parts)   # keep some handles (and their streams) alive, as a long pytest session does
print("time-outs:", bad, "connections:", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS"))
