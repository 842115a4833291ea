"""Aggregate an ncu report's per-instruction counters by CUDA source line.  usage: ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if len(r) > 4 and r[0] == "Line No":
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":       # a source line row (aggregate of its SASS rows)
        try:
            ie = float(r[ix["Instructions Executed"]])
            te = float(r[ix["Thread Instructions Executed"]])
            ns = float(r[ix["# Samples"]] or 0)
        except ValueError:
            continue
        lines[(cur_file, int(r[0]))] = (ie, te, ns, r[1].strip()[:100])
tot = sum(v[0] for v in lines.values())
tots = sum(v[2] for v in lines.values())
print(f"total warp instructions {tot:.4g}, samples {tots:.0f}")
byfile = {}
for (f, l), v in lines.items():
    byfile[f] = byfile.get(f, 0) + v[0]
print({k: round(v / tot * 100, 1) for k, v in byfile.items()})
for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{f}:{l:4d} inst {v[0] / tot * 100:5.2f}%  smp {v[2] / max(tots, 1) * 100:5.2f}%  thr/inst {v[1] / max(v[0], 1):5.1f}  {v[3]}")
