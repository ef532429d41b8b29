"""Diagnostic: multi-step launch vs one launch per step on the same lattice; reports the first differing step / cells."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import scenarios as SC
from snn_b200 import _capi as K

rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100, 100)
chem = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "none" else None
out = []
for spg in (1, 0):
    lat = SC.build_lattice(None, model="izh", rows=rows, cols=cols, seed=3, graph="grid2", chem=chem, history=True)
    lat._be.set_option(K.OPT_STEPS_PER_GRAPH, spg)
    lat.run_lattice(40)
    out.append(lat.grid_history.history.copy())
a, b = out
for s in range(a.shape[0]):
    bad = np.argwhere(a[s] != b[s])
    if bad.size:
        print(f"first difference at step {s}: {bad.shape[0]} cells, e.g. {bad[:5].tolist()}, {a[s][tuple(bad[0])]} vs {b[s][tuple(bad[0])]}")
        break
else:
    print("identical over", a.shape[0], "steps")
