"""Helpers shared by the tests: golden-fixture I/O and image comparison."""
import os

import numpy as np

from rtb200 import scene as S
from rtb200.scene import SceneContainer

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ARRAYS = ("spheres", "planes", "surfaces", "boxes", "toruses", "rings", "lights_point", "lights_direct")
DTYPES = dict(SceneContainer._ARRAYS)


def scene_to_npz_dict(sc: SceneContainer) -> dict:
    d = {"scene": np.frombuffer(np.ascontiguousarray(sc.scene).tobytes(), dtype=np.uint8),
         "ambient_color": np.asarray(sc.ambient_color, dtype=np.float32),
         "shadow_ambient": np.asarray(sc.shadow_ambient, dtype=np.float32)}
    for n in ARRAYS:
        d[n] = np.frombuffer(sc.array(n).tobytes(), dtype=np.uint8)
    return d


def scene_from_npz(z) -> SceneContainer:
    sc = SceneContainer()
    sc.scene = np.frombuffer(z["scene"].tobytes(), dtype=S.rt_scene)[0].copy()
    sc.ambient_color = tuple(float(x) for x in z["ambient_color"])
    sc.shadow_ambient = tuple(float(x) for x in z["shadow_ambient"])
    for n in ARRAYS:
        setattr(sc, n, np.frombuffer(z[n].tobytes(), dtype=DTYPES[n]).copy())
    return sc


def pixel_err(a, b):
    """Per-pixel max-abs difference over RGBA; NaN in both = 0, NaN in one = inf."""
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    both = np.isnan(a) & np.isnan(b)
    d = np.where(both, 0.0, np.where(np.isnan(d), np.inf, d))
    return d.max(axis=-1)


def golden_files():
    """Image fixtures produced by oracle/_ref (default_scene_t0.npz is a scene dump, not an image fixture)."""
    return sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and f != "default_scene_t0.npz")
