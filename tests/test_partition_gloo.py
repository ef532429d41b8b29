"""world_size-2 `gloo` test of the N > 1 host logic on CPU: the library's own strip plan (snn_partition_begin),
the rank-to-rank plumbing of snn_b200.dist, and the halo-exchange schedule (boundary rows of step s feed step s+1),
with the per-strip arithmetic done by the oracle.  The strips must reproduce the single-domain oracle bit for bit."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path[:0] = [os.path.join(ROOT, "spiking-neural-networks_b200"), os.path.join(ROOT, "oracle")]
    from snn_b200.dist import partition_rows, exchange_blobs
    from oracle_api import OracleBackend

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rows, cols, steps = 13, 7, 60
    r0, r1 = partition_rows(rows, world, rank)
    # every rank learns its neighbours through the same all-gather the CUDA path uses for its IPC blobs
    lo, hi = exchange_blobs(("rows", r0, r1), rank, world)
    assert (lo is None) == (rank == 0) and (hi is None) == (rank == world - 1)
    if lo is not None: assert lo[2] == r0
    if hi is not None: assert hi[1] == r1
    rng = np.random.default_rng(5)
    n = rows * cols
    V = rng.uniform(-65, 30, n).astype(np.float32)
    B = rng.uniform(0.25, 0.36, n).astype(np.float32)
    g0, g1 = max(0, r0 - 1), min(rows, r1 + 1)            # strip plus one halo row on each side
    be = OracleBackend(4, 0, 0, rows=g1 - g0, cols=cols)
    sl = slice(g0 * cols, g1 * cols)
    be.set_field(0, "current_voltage", V[sl]); be.set_field(0, "b", B[sl]); be.fill_field(0, "c_m", 4.0)
    be.fill_field(0, "gap_conductance", 10.0)
    be.connect_grid(0, 1, 1.0)
    state = ["current_voltage", "w_value", "last_firing_time", "is_spiking"]
    own = slice((r0 - g0) * cols, (r1 - g0) * cols)
    for s in range(steps):
        be.run(1)
        # halo exchange: my first/last owned row -> neighbour's ghost row (ghost results of this step are garbage)
        for name in state:
            a = be.get_field(0, name)
            first, last = a[own][:cols].copy(), a[own][-cols:].copy()
            reqs, recv_lo, recv_hi = [], None, None
            if rank > 0:
                recv_lo = torch.empty(cols, dtype=torch.from_numpy(first).dtype)
                reqs += [dist.isend(torch.from_numpy(first), rank - 1), dist.irecv(recv_lo, rank - 1)]
            if rank < world - 1:
                recv_hi = torch.empty(cols, dtype=torch.from_numpy(last).dtype)
                reqs += [dist.isend(torch.from_numpy(last), rank + 1), dist.irecv(recv_hi, rank + 1)]
            for r in reqs: r.wait()
            if recv_lo is not None: a[:cols] = recv_lo.numpy()
            if recv_hi is not None: a[-cols:] = recv_hi.numpy()
            be.set_field(0, name, a)
    out = {name: be.get_field(0, name)[own] for name in state}
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        full = OracleBackend(4, 0, 0, rows=rows, cols=cols)
        full.set_field(0, "current_voltage", V); full.set_field(0, "b", B); full.fill_field(0, "c_m", 4.0)
        full.fill_field(0, "gap_conductance", 10.0)
        full.connect_grid(0, 1, 1.0)
        full.run(steps)
        for name in state:
            got = np.concatenate([g[name] for g in gathered])
            assert (got == full.get_field(0, name)).all(), name
        assert (full.get_field(0, "last_firing_time") >= 0).sum() > 0
        print("GLOO_PARTITION_OK")
    dist.barrier()
    dist.destroy_process_group()
''')


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_strip_partition_matches_single_domain(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "GLOO_PARTITION_OK" in res.stdout


def test_partition_rows_cover():
    sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))
    from snn_b200.dist import partition_rows
    for rows in (1, 2, 13, 3163):
        for world in (1, 2, 4, 8):
            spans = [partition_rows(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
