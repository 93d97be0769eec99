#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) or an ncu launch-list csv into markdown for profiles/.

    python tools/ncu_summary.py rep  gpurun_out/prof.ncu-rep      > profiles/xxx.md     (or the report's raw-page csv)
    python tools/ncu_summary.py list gpurun_out/launches.csv      > profiles/yyy.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "idc__requests.sum", "idc__requests_lookup_miss.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def rep(path):
    # a .ncu-rep is read with ncu itself; a .csv is the `--page raw --csv` dump of a report made on the GPU box
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full summary: `%s`\n" % path.split("/")[-1])
    print("(captured with `--clock-control none`; per-launch values; read offline with `ncu -i ... --page raw --csv`)\n")
    for r in rows[2:]:
        d = OrderedDict(zip(hdr, r))
        print("## %s  (launch id %s)\n" % (d.get("Kernel Name", "?")[:160], d.get("ID", "?")))
        print("| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in d and d[k] != "":
                print("| %s | %s | %s |" % (k, d[k], units[hdr.index(k)]))
        print()


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0][:110]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    unit = rows[1][hdr.index("Metric Unit")]
    print("# ncu launch list: `%s`\n" % path.split("/")[-1])
    print("(`ncu --metrics gpu__time_duration.sum --clock-control none`; durations are serialised and cold-cache: compare SHARES)\n")
    print("| kernel | launches | total %s | share |\n|---|---|---|---|" % unit)
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.0f | %.2f%% |" % (name, n, t, 100 * t / total))


if __name__ == "__main__":
    (rep if sys.argv[1] == "rep" else launch_list)(sys.argv[2])
