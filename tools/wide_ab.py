"""Development (GPU): A/B of the 24-warp persistent-kernel variant (option "wide") on the torus-free BASELINE scene, both builds."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=256); cube = textures.TextureSet(cube=ts.cube)
for name, strict in (("spheres4k", 0), ("spheres4k", 1)):
    sc = scenes.build_config(name); w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    row = {"config": name, "strict": strict}
    for wide in (0, 1):
        gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, cube); gl.set_option("strict", strict); gl.set_option("wide", wide)
        ms = []
        for _ in range(6):
            gl.draw(); gl.sync(); ms.append(round(gl.stats().kernel_ms, 3))
        row[f"wide{wide}_ms"] = ms; row[f"wide{wide}_block"] = gl.stats().block
        gl.stop()
    print(json.dumps(row), flush=True)
