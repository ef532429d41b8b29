"""RewardModulatedLatticeNetwork at scale (the shape of examples/lsm_architecture, grown): spike trains -> liquid (plain lattice)
-> read-out (reward-modulated lattice).  Device time per timestep with and without do_modulation; the difference is the
post-step weight pass (rstdp_net_edge_kernel), reported against its algorithmic bytes."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spiking-neural-networks_b200")):
    sys.path.insert(0, p)
import numpy as np
import bench
from snn_b200 import _capi as K
from snn_b200.backend import CudaNetworkBackend

R = int(os.environ.get("READOUT", "1024"))      # read-out lattice R x R
LQ = int(os.environ.get("LIQUID", "256"))       # liquid lattice LQ x LQ
TR = int(os.environ.get("TRAINS", "32"))        # spike trains TR x TR
FAN = 8
rng = np.random.default_rng(11)


def stencil_csr(rows, cols):
    i, j = np.divmod(np.arange(rows * cols, dtype=np.int64), cols)
    pre, ok = [], []
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            if di == 0 and dj == 0:
                continue
            a, b = i + di, j + dj
            ok.append((a >= 0) & (a < rows) & (b >= 0) & (b < cols))
            pre.append(a * cols + b)
    pre, ok = np.stack(pre, 1), np.stack(ok, 1)
    rp = np.zeros(rows * cols + 1, np.uint64)
    rp[1:] = np.cumsum(ok.sum(1))
    return rp, pre[ok].astype(np.uint32)


def random_csr(n_pre, n_post, fan):
    pre = np.sort(rng.integers(0, n_pre, (n_post, fan)), 1)
    keep = np.ones_like(pre, bool)
    keep[:, 1:] = pre[:, 1:] != pre[:, :-1]   # no duplicate edges inside a row
    rp = np.zeros(n_post + 1, np.uint64)
    rp[1:] = np.cumsum(keep.sum(1))
    return rp, pre[keep].astype(np.uint32)


def build(modulate):
    be = CudaNetworkBackend(K.MODEL_IZH, train_kind=K.TRAIN_RATE, device=0)
    be.add_train_lattice(0, TR, TR)
    be.add_lattice(1, LQ, LQ)
    be.add_reward_lattice(2, R, R)
    for lid, n in ((1, LQ * LQ), (2, R * R)):
        f = bench.init_fields(np, n, 0x5EED + lid)
        for name, arr in f.items():
            if "$" not in name:
                be.set_field(lid, name, arr)
    be.set_field(0, "rate", rng.choice([0.0, 15.0, 20.0, 30.0], TR * TR).astype(np.float32))
    edges = {}
    for blk, (rp, pre) in {(1, 1): stencil_csr(LQ, LQ), (2, 2): stencil_csr(R, R), (0, 1): random_csr(TR * TR, LQ * LQ, FAN),
                           (0, 2): random_csr(TR * TR, R * R, FAN), (1, 2): random_csr(LQ * LQ, R * R, FAN)}.items():
        be.connect_csr(blk[0], blk[1], rp, pre, rng.uniform(0.2, 1.0, pre.size).astype(np.float32))
        edges[blk] = int(pre.size)
    be.mark_connection_reward(0, 2, True)
    be.set_lattice_reward_modulator(2, modulate, dopamine=0.5, tau_d=20.0, tau_c=0.05, a_plus=0.002, a_minus=0.002,
                                    tau_plus=4.5, tau_minus=4.5, dt=0.1)
    be.set_plasticity(1, 0.002, 0.002, 4.5, 4.5, 0.1)
    be.set_option(K.OPT_ELECTRICAL_SYNAPSE, 1); be.set_option(K.OPT_CHEMICAL_SYNAPSE, 0)
    return be, edges


steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
for modulate in (True, False):
    be, edges = build(modulate)
    be.run_timed(300)
    res = []
    for _ in range(3):
        ms, nl = be.run_timed(steps)
        res.append(ms / steps * 1e3)
    out["modulated" if modulate else "frozen"] = {"us_per_timestep": min(res), "launches_per_step": nl / steps}
    be.close()
# TraceRSTDP edges (own graph, RewardModulatedWeight block): col 4 + weight RW 8 + counter RW 2 + dw RW 8 + c RW 8 = 30 B; Weight
# edges fed by a plain lattice: col 4 + weight RW 8 = 12 B; per read-out neuron its own spike time, per edge one gathered spike time
trace_edges, stdp_edges = edges[(2, 2)] + edges[(0, 2)], edges[(1, 2)]
algo = 30 * trace_edges + 12 * stdp_edges + 4 * (trace_edges + stdp_edges) + 4 * R * R
dt_us = out["modulated"]["us_per_timestep"] - out["frozen"]["us_per_timestep"]
out.update({"neurons": LQ * LQ + R * R, "trains": TR * TR, "edges": {f"{a}->{b}": n for (a, b), n in edges.items()},
            "weight_pass_us": dt_us, "weight_pass_algorithmic_bytes": algo, "weight_pass_GBps": algo / (dt_us * 1e-6) / 1e9})
print(json.dumps(out))
