#!/usr/bin/env python3
"""Generate the golden input/output vectors under tests/golden/ from oracle/_ref.

oracle/_ref is the reference's OWN shader (assets/shaders/rt.frag) compiled as C++ where it lies under
/root/reference (oracle/build_ref.py), so these images are outputs of the reference's code, not of the
restatement.  Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
Each .npz holds the scene's uniform-buffer bytes (inputs) and the RGBA32F image (output, row 0 = bottom).
Textures are the deterministic procedural set (rtb200.textures.procedural_textures); derivative pairing
at diverged quads is the k-th-call rule oracle/_ref can observe (ORC_PAIR_ORDINAL).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import rtb200  # noqa: E402,F401
from rtb200 import scenes, textures  # noqa: E402
from rtb200.scene import SceneManager, SurfaceFactory  # noqa: E402
from oracle.binding import Oracle, build, have_ref  # noqa: E402
from util import scene_to_npz_dict  # noqa: E402


def quadric_zoo(w, h, it):
    """All nine SurfaceFactory presets, a hollow glass sphere, a ring, a plane: exercises every intersector/quirk."""
    sc = scenes._base(w, h, it)
    cm = SceneManager.create_material
    F = SurfaceFactory
    presets = [F.GetEllipsoid(1.0, 0.6, 0.8, cm((0.9, 0.2, 0.2), 50, 0.2)), F.GetEllipticParaboloid(0.7, 0.9, cm((0.2, 0.9, 0.2), 10, 0.0)),
               F.GetHyperbolicParaboloid(0.8, 0.8, cm((0.2, 0.2, 0.9), 100, 0.3)), F.GetEllipticHyperboloidOneSheet(0.5, 0.5, 0.9, cm((0.9, 0.9, 0.2), 200, 0.1)),
               F.GetEllipticHyperboloidTwoSheets(0.5, 0.6, 0.4, cm((0.9, 0.2, 0.9), 0, 0.0)), F.GetEllipticCone(0.4, 0.4, 1.0, cm((0.2, 0.9, 0.9), 50, 0.25)),
               F.GetEllipticCylinder(0.5, 0.7, cm((0.7, 0.7, 0.7), 100, 0.5)), F.GetHyperbolicCylinder(0.6, 0.6, cm((0.9, 0.5, 0.1), 10, 0.0)),
               F.GetParabolicCylinder(0.5, cm((0.4, 0.3, 0.8), 200, 0.2))]
    rng = scenes.PCG32(77)
    for i, s in enumerate(presets):
        c = (-8 + 2.0 * i, 1.5 + (i % 3), 4 + (i % 2) * 3)
        s["pos"] = c
        s["quat_rotation"] = rng.quat()
        s["v_min"] = tuple(np.float32(x) - np.float32(1.5) for x in c)
        s["v_max"] = tuple(np.float32(x) + np.float32(1.5) for x in c)
        sc.surfaces.append(s)
    sc.spheres.append(SceneManager.create_sphere((0, 2, 0), 1.2, cm((1, 1, 1), 200, 0.1, 1.125, (1, 0, 2), 1), True))
    sc.spheres.append(SceneManager.create_sphere((3, 1, 2), 1.0, cm((1, 1, 1), 50, 0.0, 1.5, (0.2, 0.5, 0.1), 1), False))
    ring = SceneManager.create_ring((-3, 3, 6), 1.0, 2.5, cm((0.8, 0.8, 0.3), 10, 0.0))
    ring["quat_rotation"] = rng.quat()
    sc.rings.append(ring)
    sc.planes.append(SceneManager.create_plane((0, 1, 0), (0, 0, 0), cm((0.5, 0.5, 0.5), 10, 0.3)))
    return sc


CASES = {
    "default_tex_96x64_it3": lambda: scenes.default_scene(96, 64, 3),
    "default_notex_96x64_it5": lambda: scenes.default_scene(96, 64, 5, textured=False),
    "mini1_64x48_it4": lambda: scenes.synthetic_scene("mini1", 64, 48, 4),
    "mini7_64x48_it8": lambda: scenes.synthetic_scene("mini7", 64, 48, 8),
    "quadric_zoo_96x64_it4": lambda: quadric_zoo(96, 64, 4),
    "tori1080_48x28_it4": lambda: scenes.synthetic_scene("tori1080", 48, 28, 4),
    "spheres4k_64x36_it8": lambda: scenes.synthetic_scene("spheres4k", 64, 36, 8),
    "mixed1024_32x18_it8": lambda: scenes.synthetic_scene("mixed1024", 32, 18, 8),
}


def main():
    build(ref=True)
    assert have_ref(), "oracle/_ref could not be built (needs /root/reference)"
    ts = textures.procedural_textures()
    for name, mk in CASES.items():
        sc = mk()
        img = Oracle(sc, ts, impl="ref").render()
        d = scene_to_npz_dict(sc)
        d["image"] = img
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, img.shape, float(img[..., :3].mean()))


if __name__ == "__main__":
    main()
