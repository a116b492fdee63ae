#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.txt [--top-sass 25]
"""
import csv
import io
import subprocess
import sys

KEYS = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__sass_average_branch_targets_threads_uniform.pct",
)


def run(args):
    return subprocess.run(["ncu"] + args, check=True, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top-sass") + 1]) if "--top-sass" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    lines = [f"# ncu summary of {rep}", "# command: ncu --set full --clock-control none --import-source on (see profiles/README.md)", ""]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        lines.append(f"## kernel {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:90s} {d[k]:>18s} {units[hdr.index(k)]}")
        lines.append("")
    try:
        src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
        h = None
        body = []
        for r in src:
            if r and r[0] == "Address":
                h = r
                continue
            if h and len(r) == len(h):
                body.append(r)
        if h:
            ci = {n: h.index(n) for n in ("Source", "Warp Stall Sampling (All Samples)", "Instructions Executed",
                                          "Avg. Threads Executed") if n in h}
            body.sort(key=lambda r: -int(r[ci["Warp Stall Sampling (All Samples)"]] or 0))
            tot = sum(int(r[ci["Warp Stall Sampling (All Samples)"]] or 0) for r in body) or 1
            lines.append(f"## top {top} SASS instructions by warp-stall samples (share of {tot} samples)")
            for r in body[:top]:
                s = int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
                lines.append(f"{100.0 * s / tot:6.2f}%  execs={r[ci['Instructions Executed']]:>12s}  "
                             f"avg_threads={r[ci['Avg. Threads Executed']]:>5s}  {r[ci['Source']].strip()}")
    except Exception as e:  # source page is optional
        lines.append(f"(no source page: {e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
