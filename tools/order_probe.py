"""Development (GPU): cost of the tile-order kernel = frame time through CUDA events around draw() minus the kernel's own time."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtb200
from rtb200 import scenes, textures
ts = textures.procedural_textures(cube_size=256); cube = textures.TextureSet(cube=ts.cube)
for name in ("spheres4k", "mixed1024_4k"):
    sc = scenes.build_config(name); w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    row = {"config": name}
    for lpt in (0, 1):
        gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, cube); gl.set_option("strict", 0); gl.set_option("lpt", lpt)
        buf = torch.empty((h, w, 4), dtype=torch.float32, device="cuda"); st = torch.cuda.Stream(); torch.cuda.set_stream(st)
        out = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); gl.draw_to(buf.data_ptr(), st.cuda_stream); b.record(st); torch.cuda.synchronize()
            out.append((round(a.elapsed_time(b), 3), round(gl.stats().kernel_ms, 3)))
        row[f"lpt{lpt}_frame_vs_kernel_ms"] = out
        gl.stop()
    print(json.dumps(row), flush=True)
