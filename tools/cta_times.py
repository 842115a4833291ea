"""Development: per-CTA drain/end stamps for each rank's share (RTB_DEBUG_TIMES=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RTB_DEBUG_TIMES"] = "1"
import rtb200
from rtb200 import scenes, textures
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ts = textures.procedural_textures(cube_size=256)
sc = scenes.build_config("mixed1024_4k")
for r in range(world):
    gl = rtb200.GLWrapper(3840, 2160); gl.init_window(); gl.set_partition(r, world, 16)
    rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube)); gl.set_option("strict", 1)
    for _ in range(2):
        gl.draw(); gl.sync()
        print(f"rank {r}: kernel {gl.stats().kernel_ms:.2f} ms", file=sys.stderr, flush=True)
    gl.stop()
