"""Summarise an .ncu-rep (read here, without a GPU) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x.txt "note"
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum.per_second", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__sass_average_data_bytes_per_sector_mem_local_op_ld.ratio", "sass__inst_executed_local_loads",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warp_latency_per_inst_issued.ratio",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = [f"# ncu summary of {rep}", f"# {note}", ""]
    for row in raw[2:]:
        d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
        lines.append(f"## kernel: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d and d[k] != "":
                lines.append(f"{k:95s} {d[k]:>18s} {u[k]}")
        lines.append("")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h = None
    agg = collections.Counter(); samp = collections.Counter()
    for r in src:
        if "Source" in r and "Instructions Executed" in r:
            h = r; ci, ei, si = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"); continue
        if h is None or len(r) <= max(ci, ei, si):
            continue
        try:
            n, s = int(r[ei]), int(r[si])
        except ValueError:
            continue
        t = r[ci].split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?"))
        agg[op] += n; samp[op] += s
    tot, stot = sum(agg.values()) or 1, sum(samp.values()) or 1
    lines.append("## executed warp-instructions by opcode (first kernel in the report) and share of stall samples")
    for op, n in agg.most_common(24):
        lines.append(f"{op:28s} {n:16d} {100 * n / tot:6.2f}%   samples {100 * samp[op] / stot:6.2f}%")
    lines.append(f"{'total':28s} {tot:16d}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
