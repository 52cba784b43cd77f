#!/usr/bin/env python3
"""Summarise an .ncu-rep (one `ncu --set full --import-source on` capture) into a small JSON + text file for profiles/.

usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx [edges_per_launch]
Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    edges = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    summary = {"report": rep, "launches": []}
    for r in data:
        d = {"kernel": r[ki]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": r[i], "unit": units[i]}
        summary["launches"].append(d)
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    starts = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if starts:
        s, e = starts[0], (starts[1] - 1 if len(starts) > 1 else len(src))
        h = src[s]
        ci, si = h.index("Instructions Executed"), h.index("Source")
        stall_cols = {x: i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x}
        ops, stalls, tot = collections.Counter(), collections.Counter(), 0
        for r in src[s + 1:e]:
            try:
                n = int(r[ci])
            except Exception:
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.x]+)", r[si])
            ops[m.group(2).split(".")[0] if m else "?"] += n
            tot += n
            for x, i in stall_cols.items():
                try:
                    stalls[x] += int(r[i])
                except Exception:
                    pass
        summary["warp_instructions"] = tot
        summary["opcode_mix"] = {k: v for k, v in ops.most_common(24)}
        ts = sum(stalls.values()) or 1
        summary["stall_samples_pct"] = {k: round(100.0 * v / ts, 1) for k, v in stalls.most_common(12)}
        if edges:
            summary["warp_instructions_per_warp_edge"] = tot / edges
            summary["opcode_per_warp_edge"] = {k: round(v / edges, 2) for k, v in ops.most_common(16)}
    json.dump(summary, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        for L in summary["launches"]:
            f.write(L["kernel"] + "\n")
            for k in KEYS:
                if k in L:
                    f.write(f"  {k:72s} {L[k]['value']:>16s} {L[k]['unit']}\n")
        if "opcode_mix" in summary:
            f.write(f"warp instructions (first launch): {summary['warp_instructions']}\n")
            if edges:
                f.write(f"per warp-edge-update: {summary['warp_instructions_per_warp_edge']:.2f}  {summary['opcode_per_warp_edge']}\n")
            f.write(f"stall samples %: {summary['stall_samples_pct']}\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
