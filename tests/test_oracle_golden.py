"""The restated oracle against (a) the golden vectors produced by the reference's own shader (oracle/_ref)
and (b) oracle/_ref itself when it is built here.  rt.frag has no tests of its own, so these
fixtures ARE the reference outputs for this path."""
import os

import numpy as np
import pytest

from oracle.binding import Oracle, have_ref
from util import GOLDEN, golden_files, pixel_err, scene_from_npz

# + - * / sqrt are IEEE on every host; powf/expf/atan2f/asinf/log2f may differ by an ulp between libm builds
LIBM_TOL = 2e-6


@pytest.mark.parametrize("fname", golden_files())
def test_oracle_reproduces_reference_shader_outputs(fname, procedural):
    z = np.load(os.path.join(GOLDEN, fname))
    sc = scene_from_npz(z)
    o = Oracle(sc, procedural)
    o.set_pairing(1)                     # the k-th-call pairing rule oracle/_ref uses at diverged quads
    img = o.render()
    err = pixel_err(img, z["image"])
    assert err.max() <= LIBM_TOL, f"{fname}: {int((err > LIBM_TOL).sum())} pixels differ, max {err.max()}"


def test_golden_set_covers_every_primitive_class():
    seen = set()
    for f in golden_files():
        z = np.load(os.path.join(GOLDEN, f))
        for n in ("spheres", "planes", "surfaces", "boxes", "toruses", "rings", "lights_point", "lights_direct"):
            if z[n].size:
                seen.add(n)
    assert len(seen) == 8


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref needs /root/reference (build container only)")
@pytest.mark.parametrize("case", ["default_tex", "default_notex", "mini3", "tori", "mixed"])
def test_restatement_is_bit_identical_to_compiled_reference_shader(case, procedural):
    from rtb200 import scenes
    sc = {"default_tex": lambda: scenes.default_scene(160, 90, 5), "default_notex": lambda: scenes.default_scene(128, 72, 4, textured=False),
          "mini3": lambda: scenes.synthetic_scene("mini3", 96, 64, 6), "tori": lambda: scenes.synthetic_scene("tori1080", 40, 24, 4),
          "mixed": lambda: scenes.synthetic_scene("mixed1024", 24, 14, 8)}[case]()
    o = Oracle(sc, procedural)
    o.set_pairing(1)
    a = o.render()
    b = Oracle(sc, procedural, impl="ref").render()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"max diff {np.abs(a - b).max()}"


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref needs /root/reference (build container only)")
def test_scan_functions_match_compiled_reference_shader(procedural):
    """calcInter / inShadow on random rays: same t, same winner, same shadow term."""
    from rtb200 import scenes
    sc = scenes.synthetic_scene("mini5", 64, 64, 2)
    o, r = Oracle(sc, procedural), Oracle(sc, procedural, impl="ref")
    rng = np.random.default_rng(1)
    for _ in range(300):
        ro = rng.uniform([-15, 0.5, -12], [15, 8, 30]).astype(np.float32)
        rd = rng.normal(size=3).astype(np.float32)
        rd /= np.linalg.norm(rd)
        assert o.calc_inter(ro, rd) == r.calc_inter(ro, rd)
        assert o.in_shadow(ro, rd, 25.0) == r.in_shadow(ro, rd, 25.0)


def test_program_order_pairing_differs_only_at_textured_diverged_quads(procedural):
    from rtb200 import scenes
    sc = scenes.default_scene(128, 72, 3)
    o = Oracle(sc, procedural)
    a = o.render()
    o.set_pairing(1)
    b = o.render()
    assert (pixel_err(a, b) > 0).mean() < 0.02
    sc2 = scenes.default_scene(128, 72, 3, textured=False)
    o2 = Oracle(sc2, procedural)
    a2 = o2.render()
    o2.set_pairing(1)
    assert np.array_equal(a2, o2.render())
