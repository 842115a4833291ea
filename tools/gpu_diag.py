import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures()
def render(sc, kernel, strict):
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts)
    gl.set_option("kernel", kernel); gl.set_option("strict", strict); gl.draw(); img = gl.read_pixels(); ms = gl.stats().kernel_ms; gl.stop(); return img, ms
for cfg, scale in (("mixed1024_4k", 0.1), ("tori1080", 0.2), ("spheres4k", 0.1)):
    sc = scenes.build_config(cfg, scale)
    for strict in (1, 0):
        a, _ = render(sc, 1, strict); b, _ = render(sc, 2, strict)
        d = np.abs(a - b).max(axis=2)
        print(cfg, "strict" if strict else "fast", "quad vs persistent: max", d.max(), "n diff", int((d > 0).sum()), "of", d.size, flush=True)
BIN = "raytracing-opengl_b200/host/build/rt_headless"
if os.path.isfile(BIN):
    td = "gpurun_out/dropin"; os.makedirs(td, exist_ok=True)
    env = dict(os.environ, RT_WIDTH="256", RT_HEIGHT="256", RT_ITERATIONS="1", RT_FRAMES="1", RT_DUMP_DIR=td)
    r = subprocess.run([BIN], capture_output=True, text=True, env=env); print(r.stdout[-300:], r.stderr[-300:])
    for f in os.listdir(td):
        if f.startswith(("cube_", "tex_")): os.remove(os.path.join(td, f))
