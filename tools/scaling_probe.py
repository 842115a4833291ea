"""Development: kernel ms of rank 0's share for world = 1,2,4,...: intercept = per-launch fixed cost (drain)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=256)
sc = scenes.build_config("mixed1024_4k")
rows = []
for world in (1, 2, 4, 8, 16, 32, 64):
    gl = rtb200.GLWrapper(3840, 2160); gl.init_window(); gl.set_partition(0, world, 16)
    rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube)); gl.set_option("strict", 1)
    ms = []
    for _ in range(3):
        gl.draw(); gl.sync(); ms.append(gl.stats().kernel_ms)
    c = gl.draw_counted(); gl.stop()
    rows.append({"world": world, "ms": round(min(ms), 2), "rays": c.rays, "flops": c.flops, "ms_per_Mray": round(min(ms) / (c.rays / 1e6), 3)})
    print(json.dumps(rows[-1]), flush=True)
x = np.array([r["flops"] for r in rows]); y = np.array([r["ms"] for r in rows])
A = np.vstack([x, np.ones_like(x)]).T
k, b = np.linalg.lstsq(A, y, rcond=None)[0]
print(json.dumps({"fit_ms_per_Tflop": k * 1e12, "intercept_ms": b}))
