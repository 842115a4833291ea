"""Development (GPU): A/B of the cost-ordered tile hand-out (option "lpt") — kernel ms of whole frames and of single shares of an
8-way split, rendered one after the other on ONE GPU.  usage: python tools/lpt_ab.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=256)
cube = textures.TextureSet(cube=ts.cube)
for name, strict, part in (("mixed1024_4k", 0, None), ("mixed1024_4k", 1, None), ("spheres4k", 0, None), ("spheres4k", 1, None), ("tori1080", 0, None),
                           ("mixed1024_4k", 0, (0, 8)), ("mixed1024_4k", 0, (3, 8)), ("mixed1024_4k", 1, (2, 8)), ("mixed1024_8k", 0, (5, 8))):
    sc = scenes.build_config(name)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    row = {"config": name, "strict": strict, "share": part}
    for lpt in (0, 1):
        gl = rtb200.GLWrapper(w, h); gl.init_window()
        if part:
            gl.set_partition(part[0], part[1], 4)
        rtb200.setup_scene(gl, sc, cube); gl.set_option("strict", strict); gl.set_option("lpt", lpt)
        ms = []
        for _ in range(5):
            gl.draw(); gl.sync(); ms.append(round(gl.stats().kernel_ms, 3))
        row[f"lpt{lpt}_ms"] = ms
        gl.stop()
    row["gain"] = round(min(row["lpt0_ms"][1:]) / min(row["lpt1_ms"][1:]), 4)
    print(json.dumps(row), flush=True)
