"""Development probe (run under gpurun): parity of both kernels / both builds against the oracle on small
canvases, then full-size timings.  Not part of the test suite."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rtb200
from rtb200 import scenes, textures
from oracle.binding import Oracle, Stats

out = {}
ts = textures.procedural_textures()


def run(gl, kernel, strict, cull=0):
    gl.set_option("kernel", kernel); gl.set_option("strict", strict); gl.set_option("cull", cull)
    gl.draw(); img = gl.read_pixels(); st = gl.stats()
    return img, st.kernel_ms


def cmp(a, b):
    d = np.abs(a - b)
    d = np.where(np.isnan(d), np.inf, d)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan, 0, d)
    px = d.max(axis=2)
    return float(px.max()), float((px > 1e-4).mean()), float((px > 1e-6).mean())


cases = [("mini1", scenes.synthetic_scene("mini1", 160, 96, 4), True),
         ("default_tex", scenes.default_scene(192, 108, 4), True),
         ("default_notex", scenes.default_scene(192, 108, 4, textured=False), True),
         ("spheres4k@1/16", scenes.build_config("spheres4k", 1 / 16), True),
         ("tori1080@1/8", scenes.build_config("tori1080", 1 / 8), True),
         ("mixed1024@1/20", scenes.build_config("mixed1024_4k", 1 / 20), True)]
for name, sc, _ in cases:
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    t0 = time.time(); ost = Stats(); want = Oracle(sc, ts).render(stats=ost); t_or = time.time() - t0
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts)
    res = {"size": [w, h], "oracle_s": round(t_or, 3)}
    kernels = [1] if sc.uses_textures() else [1, 2]
    for k in kernels:
        for strict in (1, 0):
            img, ms = run(gl, k, strict)
            res[f"k{k}_{'strict' if strict else 'fast'}"] = {"cmp(max,>1e-4,>1e-6)": cmp(img, want), "ms": round(ms, 3)}
    if not sc.uses_textures():
        img, ms = run(gl, 2, 0, cull=1)
        res["k2_fast_cull"] = {"cmp": cmp(img, want), "ms": round(ms, 3)}
    cst = gl.draw_counted()
    o, c = ost.as_dict(), cst.as_dict()
    res["counters_equal"] = all(o[k] == c[k] for k in ("pixels", "rays_nearest", "rays_shadow", "tests", "dk_iterations", "shaded_hits", "light_evals"))
    if not res["counters_equal"]:
        res["counters"] = {"oracle": o, "cuda": c}
    gl.stop()
    out[name] = res
    print(name, json.dumps(res), flush=True)

print("fp32 peak TFLOP/s:", rtb200.measure_fp32_peak(0), flush=True)

# full-size timings
for cfg in ("spheres4k", "tori1080", "mixed1024_4k"):
    sc = scenes.build_config(cfg)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts)
    res = {}
    for k, strict, cull in ((2, 0, 0), (2, 1, 0), (1, 0, 0), (2, 0, 1)):
        _, ms = run(gl, k, strict, cull)
        _, ms2 = run(gl, k, strict, cull)
        res[f"k{k}_s{strict}_c{cull}"] = [round(ms, 2), round(ms2, 2)]
        print(cfg, k, strict, cull, ms, ms2, flush=True)
    st = gl.draw_counted()
    d = st.as_dict()
    res["stats"] = d
    best = min(v[1] for kk, v in res.items() if kk.startswith("k2_s0_c0"))
    res["Mrays/s"] = st.rays / best / 1e3
    res["TFLOP/s"] = st.flops / best / 1e9
    print(cfg, json.dumps(res), flush=True)
    out[cfg] = res
    gl.stop()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
