"""The envelope criterion (tests/envelope.py): the oracle's precision variants, the committed fixtures, and — on the GPU —
the parity gate of the FUSED build: no determined pixel farther than 1e-4 from the fp32 oracle (beyond the calibrated allowance)."""
import numpy as np
import pytest

import envelope as env
from oracle.binding import Oracle
from rtb200 import scenes
from rtb200.scene import SceneManager as SM


def _simple_scene():
    """one diffuse sphere over a diffuse plane, one bounce: every pixel away from the silhouettes is determined"""
    sc = scenes._base(64, 40, 1)
    cm = SM.create_material
    sc.spheres.append(SM.create_sphere((0, 2, 0), 2.0, cm((0.8, 0.3, 0.2), 50, 0.0)))
    sc.planes.append(SM.create_plane((0, 1, 0), (0, 0, 0), cm((0.4, 0.5, 0.6), 10, 0.0)))
    return sc


def test_precision_variants_agree_where_the_arithmetic_is_benign(procedural):
    sc = _simple_scene()
    o32, p32, _ = Oracle(sc, procedural).render_ex()
    o64, p64, _ = Oracle(sc, procedural, precision="f64").render_ex()
    osr, psr, _ = Oracle(sc, procedural, precision="sr").render_ex(sample=3)
    same = (p32 == p64) & (p32 == psr)
    assert same.mean() > 0.97                                   # only silhouette / shadow-edge pixels may change their path
    assert env.pix_err(o64, o32)[same].max() < env.TOL           # (the specular highlight, pow(x, 50), is the worst: ~4e-5)
    assert env.pix_err(osr, o32)[same].max() < env.TOL
    assert not np.array_equal(osr, o32)                         # ... but the stochastic variant does round differently


def test_stochastic_rounding_is_a_deterministic_function_of_the_sample(procedural):
    sc = scenes.synthetic_scene("mini1", 32, 24, 3)
    o = Oracle(sc, procedural, precision="sr")
    a, pa, da = o.render_ex(sample=1, threads=1)
    b, pb, db = o.render_ex(sample=1, threads=4)                # independent of the thread that renders a pixel
    c, _, _ = o.render_ex(sample=2, threads=4)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and np.array_equal(pa, pb) and np.array_equal(da, db)
    assert not np.array_equal(a.view(np.uint32), c.view(np.uint32))
    # every operation stays within one ulp of the exact result: the image stays close to fp32 on determined pixels
    o32 = Oracle(sc, procedural).render()
    assert np.median(env.pix_err(a, o32)) < 1e-6


def test_fp64_variant_takes_fp32_inputs(procedural):
    """same scene bytes, same fp32 literals: a ray that misses everything samples the same sky texel in every precision"""
    sc = scenes._base(16, 8, 2)
    a = Oracle(sc, procedural).render()
    b = Oracle(sc, procedural, precision="f64").render()
    assert env.pix_err(a, b).max() < 1e-6


@pytest.mark.parametrize("name", list(env.CASES))
def test_fixture_belongs_to_the_generated_scene(name):
    cfg, scale, spread, pathdiff, digest = env.load_fixture(name)
    sc = scenes.build_config(cfg, scale)
    assert env.scene_digest(sc) == digest, "scene generator drifted: regenerate with tests/golden/make_envelope.py"
    assert spread.shape == (int(sc.scene["canvas_height"]), int(sc.scene["canvas_width"]))
    assert max(env.fixture_calibration(name)) <= env.ALLOWANCE   # what independent conformant samples scored when the fixture was made


def test_criterion_does_not_blame_a_conformant_evaluation(procedural):
    """calibration on a window: an independent stochastic-rounding sample is a conformant evaluation of the shader"""
    cfg, scale, spread, pathdiff, _ = env.load_fixture("default1080")
    sc = scenes.build_config(cfg, scale)
    x0, y0, w, h = 64, 40, 96, 48
    o32 = Oracle(sc, procedural).render(x0, y0, w, h)
    probe, _, _ = Oracle(sc, procedural, precision="sr").render_ex(x0, y0, w, h, sample=777)
    v = env.judge(probe, o32, spread[y0:y0 + h, x0:x0 + w], pathdiff[y0:y0 + h, x0:x0 + w])
    assert v["avoidable_outliers"] <= env.ALLOWANCE, v
    assert v["frac_undetermined"] < 0.5


def test_criterion_has_teeth(procedural):
    """an image that is wrong by 1e-3 on determined pixels is blamed for every one of them"""
    cfg, scale, spread, pathdiff, _ = env.load_fixture("default256")
    sc = scenes.build_config(cfg, scale)
    o32 = Oracle(sc, procedural).render()
    wrong = o32.copy()
    wrong[..., 1] += 1e-3
    v = env.judge(wrong, o32, spread, pathdiff)
    assert v["avoidable_outliers"] == int(((spread <= env.TOL / 4) & ~pathdiff).sum()) > 0.9 * spread.size


# ---------------------------------------------------------------- GPU: the parity gate of the fused build
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(env.CASES))
def test_fused_build_has_no_avoidable_outliers(name, procedural):
    import rtb200
    cfg, scale, spread, pathdiff, digest = env.load_fixture(name)
    sc = scenes.build_config(cfg, scale)
    assert env.scene_digest(sc) == digest
    o32 = Oracle(sc, procedural).render()
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h)
    gl.init_window()
    try:
        rtb200.setup_scene(gl, sc, procedural)
        gl.set_option("strict", 0)
        gl.draw()
        fused = gl.read_pixels()
        gl.set_option("strict", 1)
        gl.draw()
        strict = gl.read_pixels()
    finally:
        gl.stop()
    vs = env.judge(strict, o32, spread, pathdiff)
    vf = env.judge(fused, o32, spread, pathdiff)
    print(f"\n{name} {w}x{h}: fused within 1e-4 of oracle32 on {vf['frac_within_tol']:.4%} of pixels, undetermined {vf['frac_undetermined']:.4%}, "
          f"avoidable outliers {vf['avoidable_outliers']} (independent conformant samples: {env.fixture_calibration(name)}), "
          f"max error on determined pixels {vf['max_err_determined']:.3g}")
    assert vs["frac_within_tol"] == 1.0 and vs["avoidable_outliers"] == 0          # the strict build needs no envelope
    assert vf["avoidable_outliers"] <= env.ALLOWANCE, vf
    assert vf["frac_within_tol"] >= 0.85, vf                                       # (gross breakage would hide behind nothing)
