import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
sys.path.insert(0, os.path.join(ROOT, "tools"))
for side in (100, 64, 32):
    import scenarios as SC
    lat = SC.build_lattice(None, model="izh", rows=side, cols=side, seed=11, graph="grid", hetero=False, c_m=5.0, history=False)
    lat._push_options()
    lat._be.run_timed(100)
    ms, nl = lat._be.run_timed(2000)
    print(f"izh {side}x{side}: {1e3 * ms / 2000:.2f} us per step, launches {nl}", flush=True)
