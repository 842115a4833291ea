"""Render one workload a few times (for ncu captures and quick timings).  Development tool."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
from rtb200 import scenes
from rtb200.textures import TextureSet, procedural_textures

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="mixed1024_4k")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--build", default="strict")
ap.add_argument("--kernel", type=int, default=0)
ap.add_argument("--cull", type=int, default=0)
ap.add_argument("--ctas", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
sc = scenes.build_config(a.workload, a.scale)
w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
gl = rtb200.GLWrapper(w, h)
gl.init_window()
ts = procedural_textures(cube_size=256) if sc.uses_textures() else TextureSet(cube=procedural_textures(cube_size=256).cube)
rtb200.setup_scene(gl, sc, ts)
gl.set_option("strict", 1 if a.build == "strict" else 0)
gl.set_option("kernel", a.kernel)
gl.set_option("cull", a.cull)
if a.ctas:
    gl.set_option("ctas_per_sm", a.ctas)
for _ in range(a.reps):
    gl.draw()
    gl.sync()
    st = gl.stats()
    print(f"{a.workload} {w}x{h} build={a.build} kernel={st.kernel_used} grid={st.grid}x{st.block} smem={st.smem_bytes}: {st.kernel_ms:.3f} ms", flush=True)
c = gl.draw_counted()
print(f"rays={c.rays} flops={c.flops:.4g} -> {c.rays / st.kernel_ms / 1e3:.1f} Mrays/s, {c.flops / st.kernel_ms / 1e9:.2f} TFLOP/s algorithmic", flush=True)
gl.stop()
