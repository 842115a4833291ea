"""Development study (CPU only): what would a per-lane ("flattened") Durand-Kerner schedule gain over the lock-step loop?

Trip counts come from the oracle (TEST INFRASTRUCTURE — this script lives under tests/ for that reason): primary rays of random
pixels of mixed1024@4K plus, for those that hit, the two shadow rays and a mirror ray from the hit point — the mix the persistent
kernel traces.  A warp is a random group of 32 such rays (the refill mixes unrelated rays).

lock-step (the kernel today)   cost = sum over tori (C_TORUS + C_TRIP * max over lanes of k)
flattened, ring of depth D     every lane walks its own chain of solves; the setup of torus p (all lanes, lock-step) may run at most D
                               tori ahead of the slowest lane; per trip OVH extra issue cycles for "store roots, load next setup"
Cycle constants from the round-1b SASS: C_TRIP = 300 (264 FP + 36 other), C_TORUS = 261.

usage: python tests/dev/dk_schedule_sim.py [n_primary_rays]      (about 1 minute for 1200)
Result of 2026-10 (3483 rays): lock-step lane utilisation 0.74 in this sample (ncu on the real frame: 0.82 — real warps are more
coherent); flattened D=2..8, OVH 20: 0.87 of the lock-step cycles, OVH 35: 0.91; a ring of depth 2 is as good as an unbounded one.
The implementation that was then written needed OVH ~ 95 in its SASS (floor by hand count ~ 80), i.e. break-even at the measured 0.82:
the schedule was dropped.  Set OVH to what a new idea costs before writing it.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rtb200  # noqa: F401,E402
from oracle.binding import Oracle  # noqa: E402
from rtb200 import scenes, textures  # noqa: E402

C_TRIP, C_TORUS = 300.0, 261.0
TYPE_TORUS = 4


def collect(n_primary, rng):
    sc = scenes.build_config("mixed1024_4k", 1.0)
    orc = Oracle(sc, textures.procedural_textures(cube_size=64))
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    cam = np.array(sc.scene["camera_pos"][:3], dtype=np.float32)
    lp = np.array([3, 5, 0], np.float32)
    ld = -np.array([3, -1, 1], np.float32)
    ld /= np.linalg.norm(ld)
    rays = []
    for _ in range(n_primary):
        rd = np.asarray(orc.ray_dir(int(rng.integers(w)), int(rng.integers(h))), dtype=np.float32)
        rays.append((cam, rd))
        t = orc.calc_inter(cam, rd)[0]
        if t < 1e6:
            pt = (cam + rd * np.float32(t)).astype(np.float32)
            d1 = lp - pt
            rays.append((pt, (d1 / np.linalg.norm(d1)).astype(np.float32)))
            rays.append((pt, ld.astype(np.float32)))
            n = rng.normal(size=3).astype(np.float32)
            n /= np.linalg.norm(n)                                   # a stand-in normal for the mirror ray
            rays.append((pt, (rd - 2 * np.dot(rd, n) * n).astype(np.float32)))
    n_tori = len(sc.toruses)
    k = np.zeros((len(rays), n_tori), dtype=np.int32)
    for i, (ro, rd) in enumerate(rays):
        for j in range(n_tori):
            k[i, j] = orc.intersect(TYPE_TORUS, j, ro, rd)[2]
    return k


def lockstep(g):
    return g.shape[1] * C_TORUS + C_TRIP * g.max(axis=0).sum()


def flat_ring(g, depth, ovh, c_torus):
    n_t = g.shape[1]
    cur = np.zeros(32, int)
    left = g[:, 0].copy()
    trips = 0
    while True:
        live = cur < n_t
        if not live.any():
            break
        can = live & (cur < cur[live].min() + depth)
        trips += 1
        left[can] -= 1
        done = can & (left <= 0)
        cur[done] += 1
        nxt = done & (cur < n_t)
        left[nxt] = g[nxt, cur[nxt]]
    return n_t * c_torus + (C_TRIP + ovh) * trips


def main():
    rng = np.random.default_rng(1)
    k = collect(int(sys.argv[1]) if len(sys.argv) > 1 else 1200, rng)
    print(f"rays {len(k)}, mean trips {k.mean():.2f}, std {k.std():.2f}, per-torus std over rays {k.std(axis=0).mean():.2f}")
    res, util = {}, []
    for _ in range(300):
        g = k[rng.choice(len(k), 32, replace=False)]
        util.append(g.sum() / (32 * g.max(axis=0).sum()))
        res.setdefault("lock-step", []).append(lockstep(g))
        for depth, ovh in ((2, 20), (4, 20), (8, 20), (4, 35)):
            res.setdefault(f"ring depth {depth}, +{ovh} cycles per trip", []).append(flat_ring(g, depth, ovh, C_TORUS + 14))
    base = np.mean(res["lock-step"])
    print(f"lock-step lane utilisation in the trips: {np.mean(util):.3f}")
    for name, v in res.items():
        print(f"{name:40s} {np.mean(v):10.0f} cycles per ray scan (tori part)   {np.mean(v) / base:.3f} of lock-step")


if __name__ == "__main__":
    main()
