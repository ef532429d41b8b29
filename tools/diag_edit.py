import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import scenarios as SC
from oracle_api import OracleBackend
f32 = np.float32
rows, cols = 264, 256
fac = lambda m, nt, rc, r, c: OracleBackend(m, nt, rc, rows=r, cols=c)
def make(f):
    lat = SC.build_lattice(f, model="izh", rows=rows, cols=cols, seed=4, graph="grid", hetero=False, history=False, chem="approx_ampa", stdp=False, c_m=2.0)
    rng = np.random.default_rng(9)
    lat.set_field("current_voltage", rng.uniform(-65, 30, rows * cols).astype(f32))
    lat.set_field("b", rng.uniform(0.25, 0.36, rows * cols).astype(f32))
    return lat
a, b = make(None), make(fac)
for lat in (a, b):
    lat.edit_weight((100, 100), (100, 101), 3.0)
    lat.edit_weight((263, 255), (262, 254), 0.25)
a.run_lattice(12); b.run_lattice(12)
print("phase 1 equal:", (a.get_field("current_voltage") == b.get_field("current_voltage")).all())
for lat in (a, b):
    lat.edit_weight((10, 10), (10, 11), None)
    lat.edit_weight((10, 12), (10, 10), 2.0)
ga, gb = a.graph_csr(), b.graph_csr()
print("graphs equal:", [bool((x == y).all()) if x.shape == y.shape else (x.shape, y.shape) for x, y in zip(ga, gb)])
for name in ("current_voltage", "neurotransmitters$t", "w_value", "receptors$AMPA$r$kinetics$r"):
    print("before phase 2", name, (a.get_field(name) == b.get_field(name)).all())
for step in range(12):
    a.run_lattice(1); b.run_lattice(1)
    va, vb = a.get_field("current_voltage").reshape(rows, cols), b.get_field("current_voltage").reshape(rows, cols)
    bad = np.argwhere(va != vb)
    print("step", step, "diff cells:", bad.shape[0], bad[:6].tolist(), [(float(va[tuple(x)]), float(vb[tuple(x)])) for x in bad[:3]])
    if bad.shape[0]:
        break
