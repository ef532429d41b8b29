import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import snn_b200 as S
import scenarios as SC
from oracle_api import OracleBackend
fac = lambda m, nt, rc, rows, cols: OracleBackend(m, nt, rc, rows=rows, cols=cols)
kw = dict(model="izh", rows=9, cols=11, seed=9, graph="grid", cls=S.RewardModulatedLattice)
a, b = SC.build_lattice(None, **kw), SC.build_lattice(fac, **kw)
for L in (a, b):
    L.reward_modulator = S.RewardModulatedSTDP(tau_c=0.05, a_plus=0.4, a_minus=0.3)
rng = np.random.default_rng(1)
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for step in range(0, 60, chunk):
    r = rng.uniform(-1, 1, chunk).astype(np.float32)
    a.run_lattice_with_rewards(r); b.run_lattice_with_rewards(r)
    wa, wb = a.graph_csr()[2], b.graph_csr()[2]
    (ca, da, cca), (cb, db, ccb) = a.graph_traces(), b.graph_traces()
    va, vb = a.get_field("current_voltage"), b.get_field("current_voltage")
    la, lb = a.get_field("last_firing_time"), b.get_field("last_firing_time")
    print(step, "dv", np.abs(va - vb).max(), "lft", (la != lb).sum(), "w", np.abs(wa - wb).max(), "dw", np.abs(da - db).max(), "c", np.abs(cca - ccb).max(),
          "cnt", (ca != cb).sum(), "dop", a.reward_modulator.dopamine, b.reward_modulator.dopamine, "spikes", (lb >= 0).sum())
