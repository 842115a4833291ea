"""Loop-level instruction statistics of one kernel's SASS (development; CPU only).
The kernels are issue bound and an FFMA2 holds the issue port for two cycles (profiles/README.md), so the cost of a loop body is
    issue cycles = scalar FP32 (FFMA/FMUL/FADD) + 2 x FFMA2 + every other instruction  (+ ~1 per scalar->packed switch).
usage: python tools/sass_loops.py [object-or-library] [mangled kernel name]
default: raytracing-opengl_b200/csrc/rt_kernels_strict.o, the strict persistent kernel."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "raytracing-opengl_b200", "csrc", "rt_kernels_strict.o")
fun = sys.argv[2] if len(sys.argv) > 2 else "_ZN10rtb_strict17persistent_kernelILb0ELi640EEEv11FrameParams"
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for line in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        text = m.group(2).strip()
        op = [t for t in text.split() if not t.startswith("@")][0].split(".")[0]
        ins.append((int(m.group(1), 16), op, text))
index = {a: i for i, (a, _, _) in enumerate(ins)}
SCALAR = {"FFMA", "FMUL", "FADD"}
print(f"{fun}: {len(ins)} instructions")
print(f"{'loop':>17s} {'instr':>6s} {'FFMA2':>6s} {'scalar':>7s} {'other':>6s} {'calls':>6s} {'issue cycles':>13s} {'S->P switches':>14s}")
for i, (a, op, text) in enumerate(ins):
    m = re.search(r"\bBRA\S*\s+(?:!?U?P\d,\s*)?(?:UR\d+,\s*)?(0x[0-9a-f]+)", text)
    if not m:
        continue
    t = int(m.group(1), 16)
    if t >= a or t not in index:
        continue
    body = ins[index[t]:i + 1]
    if len(body) > 1200:
        continue
    p = sum(o == "FFMA2" for _, o, _ in body)
    s = sum(o in SCALAR for _, o, _ in body)
    calls = sum(o == "CALL" for _, o, _ in body)
    seq = "".join("P" if o == "FFMA2" else "S" if o in SCALAR else "" for _, o, _ in body)
    print(f"{t:#8x}-{a:#8x} {len(body):6d} {p:6d} {s:7d} {len(body) - p - s:6d} {calls:6d} {s + 2 * p + (len(body) - p - s):13d} {len(re.findall('SP', seq)):14d}")
