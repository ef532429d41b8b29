"""BASELINE.json configs[3] is too small to shard ("replicas only"): R independent networks (one per image) on ONE GPU,
each handle with its own stream, stepped concurrently from R host threads (ctypes releases the GIL inside snn_network_run)."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as B

steps = 2000
for R in (1, 4, 8, 16):
    nets = [B.network(None, None) for _ in range(R)]
    for n in nets:
        n.run_lattices(50)
    th = [threading.Thread(target=n.run_lattices, args=(steps,)) for n in nets]
    t0 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    print(f"{R:2d} replicas: {1e6 * dt / steps:7.1f} us per step of all replicas, {R * steps / dt:9.0f} network-steps/s, "
          f"{633600 * R * steps / dt / 1e9:.2f} G synaptic events/s", flush=True)
    for n in nets:
        n._be.close()
