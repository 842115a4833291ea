"""Development: full-size timings of the STRICT build of the library named by RTB200_LIB (A/B runs of build variants)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=256)
tag = os.environ["RTB200_LIB"].split("/")[-2] if os.environ.get("RTB200_LIB") else "default"
res = {"variant": tag}
for name in sys.argv[1:] or ("mixed1024_4k", "spheres4k"):
    sc = scenes.build_config(name)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube))
    gl.set_option("strict", 1)
    ms = []
    for _ in range(4):
        gl.draw(); gl.sync(); ms.append(round(gl.stats().kernel_ms, 2))
    res[name] = ms
    gl.stop()
print(json.dumps(res), flush=True)
