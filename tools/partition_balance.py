"""Development: kernel time of every rank's share of an N-way split, run one after the other on ONE GPU
(load balance of the row-block partition).  usage: partition_balance.py [world] [block_rows...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
from rtb200 import scenes, textures
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
brs = [int(x) for x in sys.argv[2:]] or [16]
ts = textures.procedural_textures(cube_size=256)
sc = scenes.build_config("mixed1024_4k")
for br in brs:
    out = []
    for r in range(world):
        gl = rtb200.GLWrapper(3840, 2160); gl.init_window(); gl.set_partition(r, world, br)
        rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube)); gl.set_option("strict", 1)
        ms = []
        for _ in range(3):
            gl.draw(); gl.sync(); ms.append(gl.stats().kernel_ms)
        c = gl.draw_counted()
        out.append((round(min(ms), 2), c.rays, c.dk_iterations))
        gl.stop()
    t = np.array([o[0] for o in out]); dk = np.array([o[2] for o in out], dtype=float)
    print(json.dumps({"world": world, "block_rows": br, "ms": t.tolist(), "sum_ms": round(float(t.sum()), 1), "max_over_mean": round(float(t.max() / t.mean()), 3),
                      "dk_max_over_mean": round(float(dk.max() / dk.mean()), 3)}), flush=True)
