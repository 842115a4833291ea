"""Write profiles/traffic.json (per-launch DRAM bytes and pipe figures of the dominant kernel) from ONE
`ncu --set full --clock-control none` capture.  usage: ncu_traffic.py report.ncu-rep key capture_note [extra_copy.json]
bench.py copies the entry `key` (e.g. mixed1024_4k_strict) into roofline.traffic / roofline.ncu."""
import csv
import json
import os
import subprocess
import sys

rep, key, note = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, r = rows[0], rows[1], rows[2]


def val(name):
    v = float(r[h.index(name)].replace(",", ""))
    unit = u[h.index(name)].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3}.get(unit, 1.0)


rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
entry = {
    "dram_bytes_per_launch": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
    "kernel_ms_under_ncu": round(val("gpu__time_duration.sum") / (1e6 if u[h.index("gpu__time_duration.sum")] == "ns" else 1.0), 3)
    if u[h.index("gpu__time_duration.sum")] in ("ns", "ms") else val("gpu__time_duration.sum"),
    "sm__pipe_fma_cycles_active_pct": round(val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), 2),
    "smsp__issue_active_pct": round(val("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
    "smsp__inst_executed": int(val("smsp__inst_executed.sum")),
    "active_lanes_per_warp_instruction": round(val("smsp__thread_inst_executed_per_inst_executed.ratio"), 2),
    "registers_per_thread": int(val("launch__registers_per_thread")),
    "capture": note,
}
# instruction mix by opcode (source page, SASS view): an FFMA2 holds the issue port of a sub-partition for two cycles
# (tools/micro/fma_mix_probe.cu), so the issue-bound time of the kernel is (scalar + 2 * FFMA2 + everything else) cycles
import re
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
hdr = next(x for x in srows if x and x[0] == "Address")
ie = hdr.index("Instructions Executed")
te = hdr.index("Thread Instructions Executed")
n_p = n_s = n_all = 0.0
# executed fp32 flops = what the FMA pipe really did (thread level): FFMA 2, FMUL / FADD 1, packed forms twice that.  Reported next to
# the ALGORITHMIC flops of roofline.achieved so that strength reduction cannot pass as utilisation (and vice versa).
FLOPS = {"FFMA": 2, "FMUL": 1, "FADD": 1, "FFMA2": 4, "FMUL2": 2, "FADD2": 2}
executed_flops = 0.0
for x in srows:
    if len(x) != len(hdr) or x[0] == "Address":
        continue
    m = re.match(r"(?:@!?U?P\d\s+)?([A-Z0-9_]+)", x[1].strip())
    if not m:
        continue
    n = float(x[ie] or 0)
    n_all += n
    executed_flops += FLOPS.get(m.group(1), 0) * float(x[te] or 0)
    if m.group(1) in ("FFMA2", "FMUL2", "FADD2"):
        n_p += n
    elif m.group(1) in ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2"):
        n_s += n
cycles = val("smsp__cycles_active.sum")
entry.update({"executed_fp32_flops": executed_flops, "warp_inst_ffma2": int(n_p), "warp_inst_fma_pipe_scalar": int(n_s), "warp_inst_other": int(n_all - n_p - n_s),
              "smsp_cycles_active": int(cycles), "fma_pipe_cycles_frac": round((n_s + 2 * n_p) / cycles, 4),
              "issue_cycles_frac": round((n_all + n_p) / cycles, 4)})
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "traffic.json")
try:
    doc = json.load(open(path))
except Exception:
    doc = {}
doc["_comment"] = "per-launch figures of the dominant kernel from ONE `ncu --set full --clock-control none` capture; bench.py copies them into roofline.traffic / roofline.ncu"
doc[key] = entry
json.dump(doc, open(path, "w"), indent=1)
for extra in sys.argv[4:]:
    json.dump(doc, open(extra, "w"), indent=1)
print(json.dumps(entry))
