import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scenarios as SC
rows = cols = int(sys.argv[1])
n = rows * cols
rng = np.random.default_rng(2024)
big = SC.build_lattice(None, model="izh", rows=rows, cols=cols, seed=0, graph="grid", hetero=False, history=False)
big.fill_field("c_m", 2.0)
big.set_field("current_voltage", rng.uniform(-65, 30, n).astype(np.float32))
big.run_lattice(4)
print("ok", rows)
