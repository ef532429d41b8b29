"""Sample clocks / power while the bench workload runs for a few seconds."""
import os, sys, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaLatticeBackend
rows = cols = 3163
be = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=0)
bench.configure(be, bench.init_fields(np, rows * cols, 0x5EED))
be.run_timed(50)
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active,utilization.memory",
                      "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
time.sleep(0.5)
be.run_timed(3000)
for _ in range(3):
    ms, nl = be.run_timed(2000)
    print(f"2000 steps: {ms / 2000 * 1e3:.1f} us/step", flush=True)
time.sleep(0.3)
p.terminate()
out = p.stdout.read().splitlines()
print("\n".join(out[-8::3]))
