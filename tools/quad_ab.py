"""Development: timings of the textured default scene (quad kernel), both builds, of the library named by RTB200_LIB."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=512)
tag = os.environ["RTB200_LIB"].split("/")[-2] if os.environ.get("RTB200_LIB") else "default"
res = {"variant": tag}
for name in ("default1080", "default256"):
    sc = scenes.build_config(name)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts)
    for build, strict in (("fused", 0), ("strict", 1)):
        gl.set_option("strict", strict)
        ms = []
        for _ in range(8):
            gl.draw(); gl.sync(); ms.append(round(gl.stats().kernel_ms, 4))
        res[f"{name}_{build}"] = sorted(ms)[:4]
    gl.stop()
print(json.dumps(res), flush=True)
