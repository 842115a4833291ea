"""Development: why do some ranks of an N-way split take longer?  Every rank's share of mixed1024@4K is rendered one after the other
on ONE GPU: kernel ms, exact work counters, algorithmic flops, and (second pass, RTB_DEBUG_TIMES=1) the per-CTA drain / end stamps.
usage: rank_spread.py [world] [block_rows] [strict]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
br = int(sys.argv[2]) if len(sys.argv) > 2 else 4
strict = int(sys.argv[3]) if len(sys.argv) > 3 else 1
debug = os.environ.get("RTB_DEBUG_TIMES") is not None
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=256)
sc = scenes.build_config("mixed1024_4k")
rows = []
for r in range(world):
    gl = rtb200.GLWrapper(3840, 2160); gl.init_window(); gl.set_partition(r, world, br)
    rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube)); gl.set_option("strict", strict)
    ms = []
    for _ in range(4):
        gl.draw(); gl.sync(); ms.append(gl.stats().kernel_ms)
    row = {"rank": r, "ms": [round(m, 2) for m in ms]}
    if not debug:
        c = gl.draw_counted()
        row.update(rays=c.rays, dk=c.dk_iterations, flops=c.flops, ms_per_Tflop=round(min(ms) / (c.flops / 1e12), 2))
    rows.append(row)
    print(json.dumps(row), flush=True)
    gl.stop()
t = np.array([min(r["ms"]) for r in rows])
print(json.dumps({"world": world, "block_rows": br, "max_over_mean": round(float(t.max() / t.mean()), 4), "sum_ms": round(float(t.sum()), 1)}))
