"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`): key raw metrics + hottest source lines.
usage: ncu_summary.py report.ncu-rep [top_lines] > profiles/<name>.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__thread_inst_executed_per_inst_executed.pct", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "sm__cycles_elapsed.avg.per_second"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
print(f"# {rep.split('/')[-1]}  (ncu --set full --clock-control none --import-source on; values are per launch)")
for r in rows[2:]:
    print("kernel:", r[h.index("Kernel Name")], " grid", r[h.index("Grid Size")], " block", r[h.index("Block Size")])
    for k in KEYS:
        if k in h:
            print(f"  {k:90s} {r[h.index(k)]:>18s} {u[h.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, hdr, lines = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r; ix = {n: i for i, n in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        lines[(cur, int(r[0]))] = (float(r[ix["Instructions Executed"]]), float(r[ix["Thread Instructions Executed"]]), float(r[ix["# Samples"]] or 0), r[1].strip()[:110])
    except (ValueError, KeyError):
        pass
tot = sum(v[0] for v in lines.values()) or 1
tots = sum(v[2] for v in lines.values()) or 1
print(f"\nsource lines by warp instructions executed (total {tot:.4g} warp instructions, {tots:.0f} PC samples)")
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"  {f}:{l:<4d} inst {v[0] / tot * 100:5.2f}%  samples {v[2] / tots * 100:5.2f}%  active lanes {v[1] / max(v[0], 1):4.1f}  | {v[3]}")
