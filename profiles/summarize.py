"""Turn an .ncu-rep (brought back in gpurun_out/) into the small text summaries committed under profiles/.

    python profiles/summarize.py gpurun_out/prof_tma_r1d.ncu-rep profiles/r1_step_tma_full.txt
"""
import csv
import subprocess
import sys
from collections import Counter

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
       "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct",
       "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(rep, out):
    lines = []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        lines.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for w in RAW:
            if w in hdr:
                lines.append(f"  {w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1
        lines.append("  warp stall samples: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:8]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) > 6 and r[0] != "Address"]
    i_inst, i_src, i_smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    ops = Counter()
    for r in data:
        t = r[i_src].split()
        ops[t[1] if t[0].startswith("@") else t[0]] += int(r[i_inst])
    total = sum(ops.values())
    lines.append(f"  warp-level instructions executed: {total}")
    lines.append("  opcode mix: " + ", ".join(f"{op} {100 * v / total:.1f}%" for op, v in ops.most_common(16)))
    tma = {op: v for op, v in ops.items() if op.startswith(("UBLKCP", "UTMA", "SYNCS"))}
    lines.append(f"  TMA / mbarrier SASS: {tma}")
    lines.append("  hottest SASS lines (stall samples, executions, instruction):")
    for r in sorted(data, key=lambda r: -int(r[i_smp]))[:8]:
        lines.append(f"    {r[i_smp]:>6} {r[i_inst]:>9} {r[i_src][:90]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
