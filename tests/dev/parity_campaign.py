"""Development (GPU + oracle): a parity campaign beyond the seeds of tests/test_gpu_parity.py — N fresh hostile random scenes
(tests/test_gpu_parity._random_scene), strict build in the quad kernel, the 20-warp and the 24-warp persistent kernel against the CPU
oracle: every pixel within 1e-4 (NaN where the oracle has NaN), work counters equal; the fused build's fraction of pixels beyond 1e-4 is
reported next to it (its gate is the envelope criterion).  usage: python tests/dev/parity_campaign.py [first_seed] [count]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rtb200
from rtb200 import textures
from oracle.binding import Oracle, Stats
import test_gpu_parity as T

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ts = textures.procedural_textures()
worst, fails, fused_bad, nan_px = 0.0, [], [], 0
for seed in range(first, first + count):
    sc = T._random_scene(seed)
    ost = Stats()
    want = Oracle(sc, ts).render(stats=ost)
    o = ost.as_dict()
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    nan_px += int(np.isnan(want).any(axis=2).sum())
    for kernel, wide in ((T.KERNEL_QUAD, 0), (T.KERNEL_PERSISTENT, 0), (T.KERNEL_PERSISTENT, 1)):
        if kernel == T.KERNEL_PERSISTENT and sc.uses_textures():
            continue
        gl = rtb200.GLWrapper(w, h); gl.init_window()
        try:
            rtb200.setup_scene(gl, sc, ts); gl.set_option("kernel", kernel); gl.set_option("strict", 1); gl.set_option("wide", wide)
            st = gl.draw_counted(); got = gl.read_pixels()
            both_nan = np.isnan(got) & np.isnan(want)
            err = np.where(both_nan, 0.0, np.abs(got - want))
            bad = bool(np.isnan(err).any() or np.nanmax(err) > 1e-4)
            c = st.as_dict()
            bad |= any(o[k] != c[k] for k in ("rays_nearest", "rays_shadow", "dk_iterations", "shaded_hits", "light_evals"))
            if bad:
                fails.append((seed, kernel, wide))
            worst = max(worst, float(np.nanmax(err)))
            if kernel == T.KERNEL_QUAD:
                gl.set_option("strict", 0); gl.draw(); f = gl.read_pixels()
                fe = np.where(np.isnan(f) & np.isnan(want), 0.0, np.abs(f - want))
                fused_bad.append(float((np.nan_to_num(fe, nan=1.0).max(axis=2) > 1e-4).mean()))
        finally:
            gl.stop()
print(json.dumps({"seeds": [first, first + count - 1], "scenes": count, "strict_failures": fails, "strict_worst_abs_err": worst,
                  "oracle_nan_pixels": nan_px, "fused_px_beyond_1e-4_mean": float(np.mean(fused_bad)), "fused_px_beyond_1e-4_max": float(np.max(fused_bad))}))
