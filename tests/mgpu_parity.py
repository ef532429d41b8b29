"""Multi-GPU parity driver (launched with torchrun, one rank per GPU; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_parity.py --rows 203 --cols 97 --steps 150

Every rank steps its row strip (halos pushed GPU-to-GPU inside the step kernel); rank 0 also steps the whole lattice on
its own GPU and the strips must reproduce it bit for bit (same canonical summation order => identical arithmetic).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=203)
    ap.add_argument("--cols", type=int, default=97)
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--radius", type=int, default=1)
    ap.add_argument("--no-chem", action="store_true")
    ap.add_argument("--no-stdp", action="store_true")
    ap.add_argument("--runs", type=int, default=2, help="split the steps over this many run() calls")
    ap.add_argument("--graph", choices=["grid", "random"], default="grid",
                    help="random: tests/gpu_accuracy.rs-style random-radius graph whose edges cross rank boundaries anywhere (general-graph partition)")
    ap.add_argument("--graph-radius", type=float, default=6.0)
    ap.add_argument("--reward", action="store_true", help="RewardModulatedLattice: RewardModulatedSTDP over TraceRSTDP weights, a reward per step")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from snn_b200 import _capi as K
    from snn_b200.backend import CudaLatticeBackend
    from snn_b200.dist import StripLattice, attach_general

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows, cols = args.rows, args.cols
    n = rows * cols
    rng = np.random.default_rng(77)
    f32 = np.float32
    init = {"current_voltage": rng.uniform(-65, 30, n).astype(f32), "b": rng.uniform(0.25, 0.36, n).astype(f32),
            "gap_conductance": (10 * rng.uniform(0.5, 1.5, n)).astype(f32), "c_m": np.full(n, 4.0, f32)}
    chem = not args.no_chem
    flags = np.zeros((n, 3), np.uint32)
    flags[:, 0] = 1

    csr = None
    if args.graph == "random":
        # vectorised random-radius graph: for every offset (dr, dc) within the radius, connect with probability 0.6
        g = np.random.default_rng(123)
        R = int(np.ceil(args.graph_radius))
        rr, cc = np.divmod(np.arange(n), cols)
        pres, posts, ws = [], [], []
        for dr in range(-R, R + 1):
            for dc in range(-R, R + 1):
                if (dr == 0 and dc == 0) or np.hypot(dr, dc) > args.graph_radius:
                    continue
                ok = (rr + dr >= 0) & (rr + dr < rows) & (cc + dc >= 0) & (cc + dc < cols) & (g.random(n) <= 0.6)
                posts.append(np.nonzero(ok)[0]); pres.append((rr[ok] + dr) * cols + cc[ok] + dc)
                ws.append(g.uniform(0.2, 1.5, int(ok.sum())).astype(f32))
        posts, pres, ws = np.concatenate(posts), np.concatenate(pres), np.concatenate(ws)
        order = np.lexsort((pres, posts))
        posts, pres, ws = posts[order], pres[order].astype(np.uint32), ws[order]
        rp = np.zeros(n + 1, np.uint64)
        np.add.at(rp, posts + 1, 1)
        csr = (np.cumsum(rp).astype(np.uint64), pres, ws)

    def configure(be, sl):
        for name, arr in init.items():
            be.set_field(0, name, arr[sl])
        if chem:
            be.set_field(0, "neurotransmitters$flags", flags[sl])
            be.set_field(0, "receptors$flags", flags[sl])
        if csr is None:
            be.connect_grid(0, args.radius, 0.8)
        else:
            rp, pre, w = csr
            q0, q1 = sl.start, sl.stop
            s, t = int(rp[q0]), int(rp[q1])
            if sl.stop - sl.start != n:
                be.set_option(K.OPT_GENERAL_PARTITION, 1)
            be.connect_csr(0, 0, rp[q0:q1 + 1] - rp[q0], pre[s:t], w[s:t])
        be.set_option(K.OPT_ELECTRICAL_SYNAPSE, 1)
        be.set_option(K.OPT_CHEMICAL_SYNAPSE, int(chem))
        be.set_option(K.OPT_DO_PLASTICITY, int(not args.no_stdp and not args.reward))
        be.set_plasticity(0, 0.05, 0.04, 4.5, 3.0, 0.1)
        if args.reward:
            be.set_reward_modulator(True, True, dopamine=0.0, tau_d=20.0, tau_c=0.05, a_plus=0.1, a_minus=0.08, tau_plus=4.5,
                                    tau_minus=3.0, dt=0.1)

    strip = StripLattice(K.MODEL_IZH, rows, cols, rank, world, device=local)
    sl = slice(strip.row_begin * cols, strip.row_end * cols)
    configure(strip.be, sl)
    if csr is None:
        strip.attach()
    else:
        peers = attach_general(strip.be, rank, world)
        print(f"[rank {rank}] general-graph partition: exchanges with ranks {peers}", flush=True)
    per = [args.steps // args.runs] * args.runs
    per[-1] += args.steps - sum(per)
    rewards = np.random.default_rng(5).uniform(-0.3, 0.3, args.steps).astype(f32)

    def advance(be):
        done = 0
        for i, k in enumerate(per):
            if args.reward and i % 2 == 0:
                be.run_with_rewards(rewards[done:done + k])   # run_lattice_with_reward per step
            else:
                be.run(k)                                      # RunLattice::run_lattice (no reward signal; modulation stays on)
            done += k

    advance(strip.be)
    names = ["current_voltage", "w_value", "last_firing_time", "is_spiking"] + (["neurotransmitters$t", "receptors$AMPA$r$kinetics$r"] if chem else [])
    mine = {nm: strip.be.get_field(0, nm) for nm in names}
    rp, pre, w = strip.be.get_connection_csr()
    mine["weights"] = w
    mine["pre"] = pre
    if args.reward:
        mine["traces"] = strip.be.connection_traces()
        mine["dopamine"] = strip.be.get_reward_modulator()["dopamine"]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        full = CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=local)
        configure(full, slice(0, n))
        advance(full)
        for nm in names:
            want = full.get_field(0, nm)
            got = np.concatenate([g[nm] for g in gathered])
            same = (got == want).all()
            print(f"{nm:32s} {'OK' if same else 'MISMATCH'} ({int((got != want).sum())} of {want.size})")
            ok &= bool(same)
        rp2, pre2, w2 = full.get_connection_csr()
        gw = np.concatenate([g["weights"] for g in gathered])
        gp = np.concatenate([g["pre"] for g in gathered])
        same = (gp == pre2).all() and (gw == w2).all()
        print(f"{'graph (pre, weights)':32s} {'OK' if same else 'MISMATCH'}; max |dw| from the initial weights: {np.abs(w2 - (0.8 if csr is None else csr[2])).max():.4f}")
        ok &= bool(same)
        if args.reward:
            want_tr = full.connection_traces()
            for i, nm in enumerate(("counter", "dw", "c")):
                got = np.concatenate([g["traces"][i] for g in gathered])
                same = (got == want_tr[i]).all()
                print(f"{'trace ' + nm:32s} {'OK' if same else 'MISMATCH'} ({int((got != want_tr[i]).sum())} of {got.size})")
                ok &= bool(same)
            dop = full.get_reward_modulator()["dopamine"]
            same = all(g["dopamine"] == dop for g in gathered)
            print(f"{'dopamine':32s} {'OK' if same else 'MISMATCH'} ({dop})")
            ok &= bool(same) and np.abs(w2 - 0.8).max() > 1e-3
        spikes = int((full.get_field(0, "last_firing_time") >= 0).sum())
        print(f"world={world} rows={rows} cols={cols} steps={args.steps} neurons that spiked: {spikes}")
        print("MGPU_PARITY", "PASS" if ok and spikes > 0 else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
