"""Development (multi-GPU box): ONE process driving all GPUs through rtb_create_multi — per-rank kernel times and device-side
frame time (kernels + gather) for both gather modes and both builds, against the same frame on one GPU.
usage: python tools/multi_probe.py [workload] > profiles/r2_multi_probe_nN.jsonl"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rtb200
from rtb200 import scenes, textures
name = sys.argv[1] if len(sys.argv) > 1 else "mixed1024_4k"
n = torch.cuda.device_count()
ts = textures.procedural_textures(cube_size=512)
sc = scenes.build_config(name)
w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
ref = {}
one = rtb200.GLWrapper(w, h); one.init_window(); rtb200.setup_scene(one, sc, textures.TextureSet(cube=ts.cube))
for build, strict in (("fused", 0), ("strict", 1)):
    one.set_option("strict", strict)
    ms = []
    for _ in range(4):
        one.draw(); one.sync(); ms.append(one.stats().kernel_ms)
    ref[build] = (min(ms), one.read_pixels())
one.stop()
gl = rtb200.GLWrapper(w, h, n_gpus=n, block_rows=4); gl.init_window(); rtb200.setup_scene(gl, sc, textures.TextureSet(cube=ts.cube))
for build, strict in (("fused", 0), ("strict", 1)):
    gl.set_option("strict", strict)
    for gname, g in (("nccl", 0), ("p2p", 1)):
        gl.set_option("gather", g)
        frames, ranks = [], None
        for _ in range(6):
            gl.draw()
            k, f = gl.rank_times()
            frames.append(f); ranks = k
        same = bool(np.array_equal(gl.read_pixels().view(np.uint32), ref[build][1].view(np.uint32)))
        print(json.dumps({"workload": name, "n_gpus": n, "build": build, "gather": gname, "frame_ms": [round(x, 3) for x in frames],
                          "rank_kernel_ms": [round(x, 3) for x in ranks], "one_gpu_ms": round(ref[build][0], 3),
                          "speedup": round(ref[build][0] / min(frames[1:]), 3), "bit_identical_to_one_gpu": same}), flush=True)
gl.stop()
