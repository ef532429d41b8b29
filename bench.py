#!/usr/bin/env python
"""bench.py — neuron-steps/s of the lattice stepping hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (default)
    python bench.py --impl reference ...                           # the reference algorithm's CPU path (oracle port)
    python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...   # one rank per GPU

Workload (config.workload): BASELINE.json configs[4] shape on ONE GPU — Izhikevich 3163x3163 (10^7 neurons), 8-neighbour
Moore grid, electrical + chemical (AMPA, ApproximateNeurotransmitter/ApproximateReceptor) synapses, STDP on, dt = 0.1,
default_impl() neurons with gap_conductance = 10 and V ~ U[v_init, v_th].  For N > 1 the lattice is
(3163*N) x 3163, one 3163-row strip per rank (weak scaling), halos pushed GPU-to-GPU inside the step kernel.

A bench "step" is one `run_lattice(iters)` call = `iters` simulation timesteps over the whole lattice.
`value` times the step loop on the device (CUDA events on the engine's stream, state resident in HBM);
`e2e` times the drop-in call with HOST buffers: every SoA field uploaded from pinned memory, run, state read back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "spiking-neural-networks_b200"))

BYTES_PER_NEURON_STEP = 160.0  # SURVEY.md §8(d), Izhikevich electro-chemical T=1 + STDP, K=8 (derivation in DESIGN.md)
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rows", type=int, default=3163)
    ap.add_argument("--cols", type=int, default=3163)
    ap.add_argument("--iters", type=int, default=200, help="simulation timesteps per bench step (one run_lattice call)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the step rates of the other BASELINE.json configurations")
    ap.add_argument("--cpu-rows", type=int, default=512, help="side of the bounded CPU sample lattice")
    ap.add_argument("--ref-budget-s", type=float, default=60.0, help="CPU seconds the whole --impl reference run may take")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
def init_fields(np, n, seed):
    """Synthetic inputs of SURVEY.md §8(d): type defaults, gap_conductance = 10, V ~ U[v_init, v_th]."""
    rng = np.random.default_rng(seed)
    f = {
        "current_voltage": rng.uniform(-65.0, 30.0, n).astype(np.float32),
        "gap_conductance": np.full(n, 10.0, np.float32),
        "w_value": np.full(n, 30.0, np.float32),
        "a": np.full(n, 0.02, np.float32), "b": np.full(n, 0.2, np.float32), "c": np.full(n, -55.0, np.float32),
        "d": np.full(n, 8.0, np.float32), "v_th": np.full(n, 30.0, np.float32), "tau_m": np.full(n, 1.0, np.float32),
        "c_m": np.full(n, 100.0, np.float32), "dt": np.full(n, 0.1, np.float32),
        "v_init": np.full(n, -65.0, np.float32), "w_init": np.full(n, 30.0, np.float32),
        "is_spiking": np.zeros(n, np.uint32), "last_firing_time": np.full(n, -1, np.int32),
    }
    flags = np.zeros((n, 3), np.uint32)
    flags[:, 0] = 1  # AMPA only
    chem = {
        "neurotransmitters$flags": flags.reshape(-1), "receptors$flags": flags.reshape(-1).copy(),
        "neurotransmitters$t": np.zeros(n * 3, np.float32), "neurotransmitters$t_max": np.ones(n * 3, np.float32),
        "neurotransmitters$clearance_constant": np.full(n * 3, 0.01, np.float32),
        "receptors$AMPA_g": np.ones(n, np.float32), "receptors$AMPA_e": np.zeros(n, np.float32),
        "receptors$AMPA$r$kinetics$r": np.zeros(n, np.float32),
    }
    f.update(chem)
    return f


STATE_FIELDS = ["current_voltage", "w_value", "is_spiking", "last_firing_time", "neurotransmitters$t",
                "receptors$AMPA$r$kinetics$r"]


def configure(be, fields):
    for name, arr in fields.items():
        be.set_field(0, name, arr)
    be.connect_grid(0, 1, 1.0)
    be.set_option(0, 1)      # electrical_synapse
    be.set_option(1, 1)      # chemical_synapse
    be.set_option(2, 1, 0)   # do_plasticity (STDP defaults 2, 2, 4.5, 4.5, dt 0.1)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU path
def cpu_oracle_run(rows, cols, iters, steps, warmup, parallel, threads):
    """The reference algorithm's CPU path (oracle port: C restatement with array storage — an upper bound on the Rust
    path's speed) on a bounded sample of the workload.  Returns neuron-steps/s."""
    import numpy as np
    os.environ["OMP_NUM_THREADS"] = str(threads)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle_api import OracleBackend
    n = rows * cols
    ob = OracleBackend(4, 0, 0, rows=rows, cols=cols)
    f = init_fields(np, n, 0x5EED)
    for name, arr in f.items():
        if name in ("v_init", "w_init"):
            continue
        ob.set_field(0, name, arr)
    ob.connect_grid(0, 1, 1.0)
    ob.set_option(0, 1); ob.set_option(1, 1); ob.set_option(2, 1, 0); ob.set_option(6, int(parallel))
    for _ in range(warmup):
        ob.run(iters)
    t0 = time.perf_counter()
    for _ in range(steps):
        ob.run(iters)
    dt = time.perf_counter() - t0
    ob.close()
    return n * iters * steps / dt, dt / steps


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    rows = cols = args.cpu_rows
    # calibrate so that the whole --steps/--warmup run stays within a couple of minutes
    rate, _ = cpu_oracle_run(rows, cols, 2, 1, 0, True, cores)
    budget = args.ref_budget_s / max(1, args.steps + args.warmup)
    iters = max(1, min(args.iters, int(rate * budget / (rows * cols))))
    value, per_step = cpu_oracle_run(rows, cols, iters, args.steps, args.warmup, True, cores)
    sample = f"{rows}x{cols} Izhikevich lattice (same synapses/STDP), {iters} timesteps per step, OpenMP gather on {cores} threads"
    line = {
        "impl": "reference", "metric": "neuron-steps/s (Izhikevich lattice)", "value": value, "unit": "neuron-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus, iters_override=iters, sample=sample),
        "cpu_baseline": {"value": value, "unit": "neuron-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "neuron-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Rust reference cannot be built here (no rustc/cargo): this is the C oracle port of its CPU algorithm, "
                "gather phase parallel like `parallel = true`, neuron/STDP phase serial (neuron/mod.rs:775-808, 954-982)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, iters_override=None, sample=None):
    cfg = {
        "workload": "Izhikevich lattice, 8-neighbour Moore grid, electrical + chemical(AMPA) synapses, STDP, dt=0.1 "
                    "(BASELINE.json configs[4] shape)",
        "rows_per_gpu": args.rows, "cols": args.cols, "neurons_per_gpu": args.rows * args.cols,
        "global_rows": args.rows * world, "timesteps_per_step": iters_override or args.iters,
        "partition": "row strips, halo pushed over NVLink peer memory inside the step kernel" if world > 1 else "single GPU",
        "l2": "per-step working set ~1.7 GB per GPU >> 126 MB L2 (no flush needed)",
    }
    if sample:
        cfg["sample"] = sample
    return cfg


# ------------------------------------------------------------------------------------------------ our arm
def main():
    args = parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from snn_b200 import _capi as K
    from snn_b200.backend import CudaLatticeBackend
    from snn_b200.dist import StripLattice

    # stdout carries the one JSON line and nothing else: whatever libraries print while we work (NCCL's "NCCL version ..."
    # banner at the first collective, for one) is sent to stderr at the file-descriptor level; fd 1 comes back for the line
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rows, cols, iters = args.rows, args.cols, args.iters
    n_local = rows * cols

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def make():
        if world > 1:
            sl = StripLattice(K.MODEL_IZH, rows * world, cols, rank, world, device=local_rank)
            assert sl.n_local == n_local
            return sl.be, sl
        return CudaLatticeBackend(K.MODEL_IZH, 0, 0, rows, cols, device=local_rank), None

    fields = init_fields(np, n_local, 0x5EED + rank)
    be, strip = make()
    configure(be, fields)
    if strip is not None:
        strip.attach()

    for _ in range(args.warmup):
        be.run_timed(iters)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, launches = 0.0, 0
    for _ in range(args.steps):
        ms, nl = be.run_timed(iters)
        dev_ms += ms
        launches += nl
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()

    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = t.tolist()
    total_neurons = n_local * world
    neuron_steps = total_neurons * iters * args.steps
    value = neuron_steps / (dev_ms_max * 1e-3)
    spikes = int((be.get_field(0, "last_firing_time") >= 0).sum())
    nnz = be.connection_nnz() if n_local <= 4_000_000 else None
    edges_local = nnz if nnz is not None else (8 * n_local - 6 * (rows + cols) + 4)

    # ---- e2e: the drop-in call with host buffers (upload every field from pinned memory, run, read the state back)
    e2e = None
    if not args.no_e2e:
        pinned = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in fields.items()}
        outs = {name: torch.from_numpy(np.empty_like(fields[name])).pin_memory().numpy() for name in STATE_FIELDS}
        h2d = sum(v.nbytes for v in pinned.values())
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.e2e_steps):
            ta = time.perf_counter()
            configure(be, pinned)
            tb = time.perf_counter()
            be.run(iters)
            tc = time.perf_counter()
            for name in STATE_FIELDS:
                be.get_field(0, name, out=outs[name])
            td = time.perf_counter()
            if os.environ.get("SNN_BENCH_VERBOSE"):
                print(f"[e2e rank {rank}] upload {1e3 * (tb - ta):.1f} ms, run {1e3 * (tc - tb):.1f} ms, download {1e3 * (td - tc):.1f} ms",
                      file=sys.stderr, flush=True)
        barrier()
        e2e_wall = time.perf_counter() - t1
        tt = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = sum(v.nbytes for v in outs.values())
        e2e = {"value": total_neurons * iters * args.e2e_steps / tt.item(), "unit": "neuron-steps/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "what": "snn_lattice_set_field(all fields, pinned host) + set_graph_grid + run(iters) + get_field(state)"}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = FALLBACK_HBM_GBS, "B200_PROFILING.md fallback (of fallback)"
        achieved = BYTES_PER_NEURON_STEP * n_local * iters * args.steps / (dev_ms_max * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("step_kernel_dram_bytes_per_launch")
        line = {
            "metric": "neuron-steps/s (Izhikevich lattice)", "value": value, "unit": "neuron-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "synaptic_events_per_s": edges_local * world * iters * args.steps / (dev_ms_max * 1e-3),
            "us_per_timestep": dev_ms_max * 1e3 / (iters * args.steps),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "spiked_fraction": spikes / n_local,
            "clocks": clocks, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "snn::step_win_kernel<IZHIKEVICH, CHEMG=1, NTREL, STDP, G=3> (csrc/step_win.cu)",
                         "bytes_per_neuron_step": BYTES_PER_NEURON_STEP,
                         "avg_launch_us": dev_ms_max * 1e3 / max(1, launches)},
        }
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            r = args.cpu_rows
            v1, _ = cpu_oracle_run(r, r, 8, 1, 0, False, 1)
            it = max(2, min(200, int(v1 * 6.0 / (r * r))))
            single, _ = cpu_oracle_run(r, r, it, 1, 0, False, 1)
            par, _ = cpu_oracle_run(r, r, it, 2, 0, True, cores)
            line["cpu_baseline"] = {
                "value": par, "unit": "neuron-steps/s", "cores": cores, "kind": "port",
                "single_thread_value": single,
                "sample": f"{r}x{r} lattice, same model/synapses/STDP, {it} timesteps x 2 (gather phase on {cores} OpenMP threads, "
                          f"update+STDP serial, mirroring parallel=true); oracle port = upper bound on the Rust path",
            }
        if world == 1 and not args.no_configs:
            # BASELINE.json configs[0..3] are parity-test cases, not bench lines; their device step rates ride along for the
            # record (C1 / C4 are latency-bound, report us per step).  Never allowed to cost the headline line.
            try:
                be.close()
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_configs
                line["other_configs"] = bench_configs.measure(cpu=False)
            except Exception as exc:  # noqa: BLE001
                line["other_configs"] = {"error": repr(exc)}
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
