"""Development (GPU): the fused build judged by the envelope criterion on every committed fixture -> one JSON line per case.
usage: python tools/envelope_report.py [--save DIR] > profiles/r2_envelope_report.jsonl      (--save: also write the fused frames as .npy)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rtb200
from rtb200 import scenes, textures
from oracle.binding import Oracle
import envelope as env
ts = textures.procedural_textures()
save = sys.argv[sys.argv.index("--save") + 1] if "--save" in sys.argv else None
for name in env.CASES:
    if not os.path.isfile(env.fixture_path(name)):
        continue
    cfg, scale, spread, pathdiff, digest = env.load_fixture(name)
    sc = scenes.build_config(cfg, scale)
    o32 = Oracle(sc, ts).render()
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h); gl.init_window(); rtb200.setup_scene(gl, sc, ts)
    out = {"case": name, "config": cfg, "size": [w, h], "independent_conformant_samples_avoidable": env.fixture_calibration(name)}
    for build, strict in (("fused", 0), ("strict", 1)):
        gl.set_option("strict", strict); gl.draw()
        img = gl.read_pixels()
        if save and build == "fused":
            os.makedirs(save, exist_ok=True)
            np.save(os.path.join(save, name + "_fused.npy"), img)
        v = env.judge(img, o32, spread, pathdiff)
        out[build] = {k: (round(x, 6) if isinstance(x, float) else x) for k, x in v.items()}
    gl.stop()
    print(json.dumps(out), flush=True)
