"""Development (GPU): the SMAA passes on real frames — total CUDA-event time per preset; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split.  usage: python tools/smaa_probe.py [config ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtb200
from rtb200 import scenes, textures
tabs = textures.smaa_tables()
assert tabs is not None, "host/build/assets/smaa is filled by `make -C raytracing-opengl_b200/host` where /root/reference exists"
once = "--once" in sys.argv                 # one ULTRA run of the default (compacted) passes: the ncu --set full target
names = [a for a in sys.argv[1:] if not a.startswith("--")]
for name in names or ("mixed1024_4k", "default1080"):
    sc = scenes.build_config(name)
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    ts = textures.procedural_textures(cube_size=512)
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts if sc.uses_textures() else textures.TextureSet(cube=ts.cube))
    gl.set_option("strict", 0); gl.smaa_set_tables(*tabs)
    gl.draw(); frame8 = gl.read_pixels_u8()
    out = {"config": name, "size": [w, h]}
    if once:
        gl.enable_SMAA(3)
        gl.smaa_apply(frame8)
        gl.stop()
        continue
    for preset, pname in ((0, "LOW"), (3, "ULTRA")):
        gl.enable_SMAA(preset)
        ref = None
        for compact in (0, 1):
            gl.set_option("smaa_compact", compact)
            ms = []
            for _ in range(5):
                o, e, b, t = gl.smaa_apply(frame8); ms.append(round(t, 4))
            if ref is None:
                ref = (o, e, b)
            same = all(np.array_equal(p, q) for p, q in zip(ref, (o, e, b)))
            out[f"{pname}_compact{compact}"] = {"ms": sorted(ms)[:3], "edge_px_frac": round(float(e.any(axis=2).mean()), 4), "GBs": round(24.0 * w * h / (min(ms) * 1e-3) / 1e9, 1),
                                                "identical_to_compact0": same}
    print(json.dumps(out), flush=True)
    gl.stop()
