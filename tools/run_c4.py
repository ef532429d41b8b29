import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as B
net = B.network(None, None)
net.run_lattices(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
