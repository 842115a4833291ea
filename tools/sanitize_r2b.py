"""Development (under compute-sanitizer): the kernels added in the second half of round 2 — per-tile cost atomics + the three tile-order
kernels + the permuted hand-out (three frames of a 1/5.5-scale mixed1024 frame: 8 400 tiles, option "lpt" automatic), the 24-warp
variant (spheres4k at 1/10 scale) and the four SMAA kernels (ULTRA, compacted pass 2) on the rendered frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=64); cube = textures.TextureSet(cube=ts.cube)
tabs = textures.smaa_tables()
for name, scale, frames in (("mixed1024_4k", 0.18, 3), ("spheres4k", 0.1, 2)):
    sc = scenes.build_config(name, scale); w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, cube); gl.set_option("strict", 0)
    if tabs is not None:
        gl.smaa_set_tables(*tabs); gl.enable_SMAA(3)
    for _ in range(frames):
        gl.draw(); gl.sync()
    st = gl.stats()
    img = gl.read_pixels_u8()
    print(name, w, h, "block", st.block, "kernel", st.kernel_used, "checksum", int(img.astype(np.int64).sum()), flush=True)
    gl.stop()
