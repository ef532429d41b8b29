import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import scenarios as SC
from oracle_api import OracleBackend
rows = cols = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = rows * cols
np.set_printoptions(precision=9)
lat = SC.build_lattice(None, model="hh", rows=rows, cols=cols, seed=13, graph="grid", chem="destexhe_all", history=False, gap=2.0)
v0 = np.random.default_rng(1).uniform(-64.0, -57.0, n).astype(np.float32)
lat.set_field("current_voltage", v0)
names = lat and SC.lattice_field_names(lat)
step = 0
while step < 1200:
    snap = {nm: lat.get_field(nm).copy() for nm in names}
    lat.run_lattice(10); step += 10
    v = lat.get_field("current_voltage")
    if not np.isfinite(v).all():
        bad = np.argwhere(~np.isfinite(v)).ravel()
        print("non-finite after step", step, "count", bad.size, "first cells", bad[:5], "rows/cols", [(b // cols, b % cols) for b in bad[:5]])
        b = int(bad[0]); r0, c0 = b // cols, b % cols
        # oracle on the light cone of that cell from the snapshot 10 steps earlier
        k = 10
        ra, rb, ca, cb = max(0, r0 - k - 1), min(rows, r0 + k + 2), max(0, c0 - k - 1), min(cols, c0 + k + 2)
        fac = lambda m, nt, rc, r, c: OracleBackend(m, nt, rc, rows=r, cols=c)
        patch = SC.build_lattice(fac, model="hh", rows=rb - ra, cols=cb - ca, seed=13, graph="grid", chem="destexhe_all", history=False, gap=2.0)
        for nm in names:
            a = snap[nm]; per = a.size // n
            patch.set_field(nm, a.reshape(rows, cols, per)[ra:rb, ca:cb].reshape(-1))
        patch._be.set_option(5, step - 10)
        for s in range(10):
            patch.run_lattice(1)
            pv = patch.get_field("current_voltage").reshape(rb - ra, cb - ca)
            print(" oracle step", s, "cell V", pv[r0 - ra, c0 - ca], "finite in patch", np.isfinite(pv).mean())
        print(" snapshot V around cell:", snap["current_voltage"].reshape(rows, cols)[max(0,r0-1):r0+2, max(0,c0-1):c0+2])
        break
    if step % 100 == 0:
        print("step", step, "V range", v.min(), v.max(), flush=True)
else:
    print("finite throughout")
