#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into the text kept under profiles/: headline metrics of the first kernel in the
report plus the dynamic SASS opcode mix per problem/rollout.  Usage: tools/ncu_summary.py REPORT UNITS_PER_LAUNCH > profiles/x.md"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]
STALLS = "smsp__average_warps_issue_stalled_"


def ncu(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep, units = sys.argv[1], float(sys.argv[2])
    rows = list(csv.reader(io.StringIO(ncu([rep, "--page", "raw", "--csv"]))))
    head, unit_row, val = rows[0], rows[1], rows[2]
    d = dict(zip(head, val))
    u = dict(zip(head, unit_row))
    print(f"# ncu --set full summary: {d.get('Kernel Name', '?')}\n")
    print(f"report: `{rep}` (kept in gpurun_out/, not tracked); units per launch: {units:g}; grid {d.get('Grid Size')} x block {d.get('Block Size')}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k]} | {u.get(k, '')} |")
    print("\n| warp stall reason (warps per issue-active cycle) | value |\n|---|---|")
    st = sorted(((float(d[k]), k[len(STALLS):-len('_per_issue_active.ratio')]) for k in d if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio")), reverse=True)
    for v, k in st[:10]:
        print(f"| {k} | {v:.3f} |")
    src = list(csv.reader(io.StringIO(ncu([rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hrow = next(i for i, r in enumerate(src) if "Instructions Executed" in r)
    h = src[hrow]
    ix, sx, smp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    ops, samples, tot = collections.Counter(), collections.Counter(), 0
    for r in src[hrow + 1:]:
        if len(r) <= ix or not r[ix].isdigit():
            continue
        s = r[sx].split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0].rstrip(";")
        ops[op] += int(r[ix]); samples[op] += int(r[smp]); tot += int(r[ix])
    print(f"\nDynamic SASS mix: {tot / units:.1f} warp-instructions per unit\n\n| opcode | warp-instr per unit | stall samples |\n|---|---|---|")
    for k, v in ops.most_common(24):
        print(f"| {k} | {v / units:.2f} | {samples[k]} |")


if __name__ == "__main__":
    main()
