#!/usr/bin/env python3
"""Generate the envelope fixtures under tests/golden/envelope/ (see tests/envelope.py for the criterion).

For each BASELINE.json config at its test size: render the oracle restatement in fp32, in fp64, K1 times with stochastic
rounding and K2 times with exaggerated rounding noise (oracle/real_types.h), and store per pixel the ensemble's spread around
the fp32 image and whether any member took a different discrete path.  The fixture also carries the scene's digest: the scenes are generated (rtb200.scenes, PCG32 seeds),
and a fixture only applies to the exact bytes it was computed from.  Nothing under /root/reference is needed.

    python tests/golden/make_envelope.py [case ...]          # up to ~15 minutes of CPU per case on 8 cores

Also prints, per case, the verdict of the criterion on INDEPENDENT stochastic-rounding samples (calibration: a conformant
evaluation must have no avoidable outliers).
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import rtb200  # noqa: E402,F401
from rtb200 import scenes, textures  # noqa: E402
from oracle.binding import Oracle, build  # noqa: E402
import envelope as env  # noqa: E402



def main():
    build()
    ts = textures.procedural_textures()
    for name in sys.argv[1:] or list(env.CASES):
        cfg, scale = env.CASES[name]
        sc = scenes.build_config(cfg, scale)
        t0 = time.time()
        o32, spread, pathdiff = env.ensemble(sc, ts)
        env.save_fixture(name, sc, spread, pathdiff)
        _, _, spread_s, pathdiff_s, _ = env.load_fixture(name)
        cals = []
        for k in (100, 101, 102):
            probe, _, _ = Oracle(sc, ts, precision="sr").render_ex(sample=k)
            cals.append(env.judge(probe, o32, spread_s, pathdiff_s))
        env.save_fixture(name, sc, spread, pathdiff, [c["avoidable_outliers"] for c in cals])
        print(json.dumps({"case": name, "size": list(o32.shape[:2]), "seconds": round(time.time() - t0, 1), "undetermined": round(cals[0]["frac_undetermined"], 4),
                          "calibration_avoidable": [c["avoidable_outliers"] for c in cals], "calibration_within_tol": [round(c["frac_within_tol"], 4) for c in cals],
                          "bytes": os.path.getsize(env.fixture_path(name))}), flush=True)


if __name__ == "__main__":
    main()
