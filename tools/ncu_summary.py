#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_conv_tc.md "title"
"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (reg limit)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none), one column per captured launch", ""]
    launches = rows[2:]
    names = []
    for r in launches:
        n = r[col["Kernel Name"]]
        n = n.replace("void ", "").replace("afldm::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        names.append(n.split("(")[0][:40] + " grid " + r[col["Grid Size"]] + " block " + r[col["Block Size"]])
    lines.append("| metric | " + " | ".join(names) + " |")
    lines.append("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in col:
            continue
        i = col[key]
        vals = [f"{r[i]} {units[i]}" for r in launches]
        lines.append(f"| {label} (`{key}`) | " + " | ".join(vals) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
