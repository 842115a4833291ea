"""Development: parity (strict vs oracle, incl. DK trip counts) + full-size timing of the library named by RTB200_LIB."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import rtb200
from rtb200 import scenes, textures
from oracle.binding import Oracle, Stats
ts = textures.procedural_textures(cube_size=256)
tag = os.environ.get("RTB200_LIB", "default").split("/")[-2] if os.environ.get("RTB200_LIB") else "default"
res = {"variant": tag}
if "--noparity" not in sys.argv:
    for name, scale in (("tori1080", 1 / 8), ("mixed1024_4k", 1 / 16)):
        sc = scenes.build_config(name, scale)
        ost = Stats(); want = Oracle(sc, ts).render(stats=ost)
        w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
        gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts)
        gl.set_option("strict", 1); gl.set_option("kernel", 2)
        st = gl.draw_counted(); img = gl.read_pixels(); gl.stop()
        res[name + "_par"] = {"maxerr": float(np.abs(img - want).max()), "dk_equal": bool(st.dk_iterations == ost.as_dict()["dk_iterations"])}
for name, builds in (("mixed1024_4k", ("strict", "fast")), ("tori1080", ("strict",)), ("spheres4k", ("strict",))):
    sc = scenes.build_config(name)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube))
    for b in builds:
        gl.set_option("strict", 1 if b == "strict" else 0)
        ms = []
        for _ in range(6):
            gl.draw(); gl.sync(); ms.append(round(gl.stats().kernel_ms, 2))
        st = gl.stats()
        res[f"{name}_{b}"] = {"ms": ms, "grid": st.grid}
    gl.stop()
# one rank's share of an 8-way split of the 4K frame (what strong scaling at N=8 times)
sc = scenes.build_config("mixed1024_4k")
gl = rtb200.GLWrapper(3840, 2160); gl.init_window(); gl.set_partition(3, 8, 16); rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube))
gl.set_option("strict", 1)
ms = []
for _ in range(5):
    gl.draw(); gl.sync(); ms.append(round(gl.stats().kernel_ms, 2))
res["mixed1024_4k_strict_rank3of8"] = ms
gl.stop()
print(json.dumps(res), flush=True)
