"""GPU parity tests proper: the CUDA path, called through the C-ABI (librtb200.so), against the CPU oracle.

Tolerance (BASELINE.json north_star): per-channel |CUDA - reference| <= 1e-4 on the RGBA32F framebuffer.
  * strict build: EVERY pixel must satisfy it (observed <= 5e-7: + - * / sqrt are bit-identical to the oracle,
    only libm's pow/exp/atan/asin/log2 differ by an ulp).
  * fused build (FMA contraction, MUFU reciprocals, rotation matrices): judged by the envelope criterion of
    tests/envelope.py (tests/test_envelope.py); here only a first-hit sanity bound.
"""
import os

import numpy as np
import pytest

import rtb200
from oracle.binding import Oracle, Stats
from rtb200 import scenes
from rtb200.api import KERNEL_PERSISTENT, KERNEL_QUAD
from rtb200.scene import SceneManager as SM
from util import GOLDEN, golden_files, pixel_err, scene_from_npz

pytestmark = pytest.mark.gpu
TOL = 1e-4


def gpu_render(sc, ts, kernel=0, strict=1, cull=0, counted=False, coop=1):
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h)
    gl.init_window()
    try:
        rtb200.setup_scene(gl, sc, ts)
        gl.set_option("kernel", kernel)
        gl.set_option("strict", strict)
        gl.set_option("cull", cull)
        gl.set_option("coop", coop)
        if counted:
            st = gl.draw_counted()
            return gl.read_pixels(), st
        gl.draw()
        return gl.read_pixels(), gl.stats()
    finally:
        gl.stop()


CASES = {
    "default256": lambda: scenes.build_config("default256"),                       # BASELINE configs[0], full size
    "default1080/8": lambda: scenes.build_config("default1080", 1 / 8),
    "spheres4k/12": lambda: scenes.build_config("spheres4k", 1 / 12),
    "tori1080/8": lambda: scenes.build_config("tori1080", 1 / 8),
    "mixed1024/16": lambda: scenes.build_config("mixed1024_4k", 1 / 16),
    "mini4": lambda: scenes.synthetic_scene("mini4", 250, 130, 6),                 # canvas not a multiple of the 8x4 tile
}


@pytest.mark.parametrize("case", list(CASES))
def test_strict_build_matches_oracle_on_every_pixel(case, procedural):
    sc = CASES[case]()
    ost = Stats()
    want = Oracle(sc, procedural).render(stats=ost)
    kernels = [KERNEL_QUAD] if sc.uses_textures() else [KERNEL_QUAD, KERNEL_PERSISTENT]
    for k in kernels:
        got, st = gpu_render(sc, procedural, kernel=k, strict=1, counted=True)
        err = pixel_err(got, want)
        assert err.max() <= TOL, f"{case} kernel {k}: {int((err > TOL).sum())} px beyond {TOL}, max {err.max()}"
        assert st.kernel_used == k
        o = ost.as_dict()
        c = st.as_dict()
        for key in ("pixels", "rays_nearest", "rays_shadow", "tests", "dk_iterations", "shaded_hits", "light_evals"):
            assert o[key] == c[key], (case, k, key, o[key], c[key])       # same rays, same tests, same solver iterations


@pytest.mark.parametrize("fname", [f for f in golden_files() if "default_tex" not in f])
def test_strict_build_matches_reference_shader_golden_vectors(fname, procedural):
    """tests/golden/*.npz were produced by the reference's own rt.frag (oracle/_ref)."""
    z = np.load(os.path.join(GOLDEN, fname))
    sc = scene_from_npz(z)
    for k in (KERNEL_QUAD, KERNEL_PERSISTENT):
        got, _ = gpu_render(sc, procedural, kernel=k)
        err = pixel_err(got, z["image"])
        assert err.max() <= TOL, f"{fname} kernel {k}: max {err.max()}"


def test_textured_golden_vector_outside_diverged_quads(procedural):
    """The textured fixture pairs derivatives by call ordinal (all oracle/_ref can observe); the quad kernel pairs by
    program position.  They agree wherever the 2x2 quad did not diverge: compare through the oracle's two rules."""
    z = np.load(os.path.join(GOLDEN, "default_tex_96x64_it3.npz"))
    sc = scene_from_npz(z)
    o = Oracle(sc, procedural)
    prog = o.render()
    got, _ = gpu_render(sc, procedural, kernel=KERNEL_QUAD)
    assert pixel_err(got, prog).max() <= TOL
    same_rule = pixel_err(prog, z["image"]) <= 2e-6
    assert same_rule.mean() > 0.97
    assert pixel_err(got, z["image"])[same_rule].max() <= TOL


@pytest.mark.parametrize("case", ["spheres4k/12", "tori1080/8", "mixed1024/16", "default1080/8"])
def test_fused_build_first_hit(case, procedural):
    """Paths cut after the first hit (no chaotic amplification through reflections): the fused build may differ from the fp32
    oracle only where a silhouette, a shadow edge or a solver trip count flips.  The full-depth gate is the envelope
    criterion (tests/test_envelope.py::test_fused_build_has_no_avoidable_outliers)."""
    sc = CASES[case]()
    sc.scene["reflect_depth"] = 1
    want1 = Oracle(sc, procedural).render()
    got1, _ = gpu_render(sc, procedural, strict=0)
    err = pixel_err(got1, want1)
    frac1 = float((err > TOL).mean())
    assert frac1 <= 0.05, f"{case}: {frac1:.3%} of first-hit pixels beyond {TOL}"      # tori: the fp32 Durand-Kerner root itself is only good to ~1e-4 (3 % observed)
    assert float(np.median(err)) <= 1e-6


def test_absorb_distance_quirk_q5_on_the_gpu():
    """rt.frag:816,859 — absorbDistance accumulates over the whole path (tests/test_oracle_kat.py has the closed form)."""
    from test_oracle_kat import _two_glass_spheres
    sc, sky = _two_glass_spheres(0.4)
    want = Oracle(sc, sky).render()
    for k in (KERNEL_QUAD, KERNEL_PERSISTENT):
        got, _ = gpu_render(sc, sky, kernel=k)
        assert pixel_err(got, want).max() <= TOL
    fused, _ = gpu_render(sc, sky, strict=0)
    assert float((pixel_err(fused, want) > TOL).mean()) < 0.002          # silhouette pixels only


def test_quad_and_persistent_kernels_are_bit_identical(procedural):
    sc = scenes.build_config("mixed1024_4k", 1 / 10)
    a, _ = gpu_render(sc, procedural, kernel=KERNEL_QUAD)
    b, _ = gpu_render(sc, procedural, kernel=KERNEL_PERSISTENT)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # (the fast build is NOT bit-stable across kernels: nvcc contracts a*b+c differently in different inlining contexts)


def test_torus_cull_option_preserves_results(procedural):
    sc = scenes.build_config("tori1080", 1 / 6)
    a, _ = gpu_render(sc, procedural, cull=0)
    b, _ = gpu_render(sc, procedural, cull=1)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


# ---------------------------------------------------------------- edge cases
def test_empty_scene_is_sky_only(procedural):
    sc = scenes._base(64, 40, 4)
    sc.lights_point.clear()
    sc.lights_direct.clear()
    want = Oracle(sc, procedural).render()
    for k in (KERNEL_QUAD, KERNEL_PERSISTENT):
        got, _ = gpu_render(sc, procedural, kernel=k)
        assert pixel_err(got, want).max() <= TOL
    got, _ = gpu_render(sc, None)
    assert np.array_equal(got[..., :3], np.zeros_like(got[..., :3])) and (got[..., 3] == 1).all()


def test_zero_iterations_and_no_lights(procedural):
    sc = scenes.synthetic_scene("mini1", 32, 20, 0)
    got, _ = gpu_render(sc, procedural)
    assert np.array_equal(got[..., :3], np.zeros_like(got[..., :3]))
    sc = scenes.synthetic_scene("mini1", 32, 20, 3)
    sc.lights_point.clear()
    sc.lights_direct.clear()
    want = Oracle(sc, procedural).render()
    for k in (KERNEL_QUAD, KERNEL_PERSISTENT):
        got, _ = gpu_render(sc, procedural, kernel=k)
        assert pixel_err(got, want).max() <= TOL


def test_glass_hollow_sphere_ring_and_rotated_camera(procedural):
    sc = scenes.synthetic_scene("mini6", 96, 64, 5)
    cm = SM.create_material
    sc.spheres.append(SM.create_sphere((0, 2, -4), 1.5, cm((1, 1, 1), 200, 0.1, 1.125, (1, 0, 2), 1), True))     # glass, hollow
    sc.spheres.append(SM.create_sphere((3, 2, -2), 1.0, cm((1, 1, 1), 50, 0.0, 1.5, (0.2, 0.5, 0.1), 1), False))  # glass, no mirror term
    ring = SM.create_ring((-3, 3, 0), 1.0, 2.5, cm((0.8, 0.8, 0.3), 10, 0.0))
    sc.rings.append(ring)
    sc.lights_point.append(SM.create_light_point((-4, 6, -6, 0.5), (1, 0.8, 0.6), 12))
    from rtb200.scene import quat_from_euler
    sc.scene["quat_camera_rotation"] = quat_from_euler(0.15, -0.2, 0)
    want = Oracle(sc, procedural).render()
    for k in (KERNEL_QUAD, KERNEL_PERSISTENT):
        got, _ = gpu_render(sc, procedural, kernel=k)
        err = pixel_err(got, want)
        assert err.max() <= TOL, (k, err.max(), int((err > TOL).sum()))


def test_textured_ring_shadows_and_box_texture(procedural):
    """fwidth / implicit-LOD sites, including alpha-accumulating ring shadows (rt.frag:644-651)."""
    sc = scenes.default_scene(160, 90, 3)
    sc.scene["camera_pos"] = (6, 2, -3)
    ring = SM.create_ring((8, 3.5, 6), 0.5, 3.0, SM.create_material((0, 0, 0), 0, 0))
    ring["textureNum"] = 4
    from rtb200.scene import quat_angle_axis
    ring["quat_rotation"] = quat_angle_axis(1.3, (1, 0, 0))
    sc.rings.append(ring)
    want = Oracle(sc, procedural).render()
    got, st = gpu_render(sc, procedural)
    assert st.kernel_used == KERNEL_QUAD
    err = pixel_err(got, want)
    assert err.max() <= TOL, (err.max(), int((err > TOL).sum()))


def test_maximum_uniform_block_population(procedural):
    """585 spheres = the most a 64 KB GL uniform block holds."""
    sc = scenes._base(64, 36, 2)
    rng = scenes.PCG32(9)
    scenes._add_spheres(sc, rng, 585)
    want = Oracle(sc, procedural).render()
    got, _ = gpu_render(sc, procedural)
    assert pixel_err(got, want).max() <= TOL


# ---------------------------------------------------------------- error behaviour
def test_call_order_and_argument_errors(procedural):
    gl = rtb200.GLWrapper(32, 32)
    gl.init_window()
    with pytest.raises(rtb200.RtbError, match="init_shaders|set_defines"):
        gl.draw()
    sc = scenes.default_scene(32, 32, 1)
    rtb200.setup_scene(gl, sc, procedural)
    gl.set_option("kernel", KERNEL_PERSISTENT)
    with pytest.raises(rtb200.RtbError, match="2-D textures"):
        gl.draw()
    with pytest.raises(rtb200.RtbError, match="unknown option"):
        gl.set_option("bogus", 1)
    with pytest.raises(rtb200.RtbError, match="partition"):
        gl.set_partition(3, 2, 16)
    gl.stop()
    gl = rtb200.GLWrapper(32, 32)
    gl.init_window()
    sc = scenes.synthetic_scene("mini1", 64, 64, 1)          # canvas size differs from the context's
    rtb200.setup_scene(gl, sc, procedural)
    with pytest.raises(rtb200.RtbError, match="canvas"):
        gl.draw()
    gl.stop()


# ---------------------------------------------------------------- full BASELINE sizes: size-independent properties
def _render_partitioned(sc, ts, world, strict=1):
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    parts = []
    for r in range(world):
        gl = rtb200.GLWrapper(w, h)
        gl.init_window()
        gl.set_partition(r, world, 16)
        rtb200.setup_scene(gl, sc, ts)
        gl.set_option("strict", strict)
        gl.draw()
        parts.append(gl.read_pixels())
        gl.stop()
    return rtb200.gather_rows(parts, h, world, 16)


def test_full_size_4k_mixed1024_properties(procedural):
    """BASELINE configs[4] scene at the headline 3840x2160 / 8 bounces."""
    sc = scenes.build_config("mixed1024_4k")
    full, st = gpu_render(sc, procedural, kernel=KERNEL_PERSISTENT)
    assert full.shape == (2160, 3840, 4) and np.isfinite(full).all() and (full[..., 3] == 1).all()
    again, _ = gpu_render(sc, procedural, kernel=KERNEL_PERSISTENT)
    assert np.array_equal(full.view(np.uint32), again.view(np.uint32))              # idempotent / deterministic
    tiled = _render_partitioned(sc, procedural, 8)
    assert np.array_equal(full.view(np.uint32), tiled.view(np.uint32))              # N-GPU tiling is bit-invariant
    rng = np.random.default_rng(0)                                                  # sampled quads against the oracle
    qx = (rng.integers(0, 3840 // 2, 1500) * 2).astype(np.int32)
    qy = (rng.integers(0, 2160 // 2, 1500) * 2).astype(np.int32)
    want = Oracle(sc, procedural).render_quads(qx, qy)
    got = np.stack([full[qy, qx], full[qy, qx + 1], full[qy + 1, qx], full[qy + 1, qx + 1]], axis=1)
    err = pixel_err(got, want)
    assert err.max() <= TOL, (err.max(), int((err > TOL).sum()))


def test_full_size_default1080_and_spheres4k_sampled(procedural):
    for cfg, n in (("default1080", 3000), ("spheres4k", 3000), ("tori1080", 1500)):
        sc = scenes.build_config(cfg)
        w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
        full, _ = gpu_render(sc, procedural)
        rng = np.random.default_rng(1)
        qx = (rng.integers(0, w // 2, n) * 2).astype(np.int32)
        qy = (rng.integers(0, h // 2, n) * 2).astype(np.int32)
        want = Oracle(sc, procedural).render_quads(qx, qy)
        got = np.stack([full[qy, qx], full[qy, qx + 1], full[qy + 1, qx], full[qy + 1, qx + 1]], axis=1)
        err = pixel_err(got, want)
        assert err.max() <= TOL, (cfg, err.max(), int((err > TOL).sum()))


def test_full_size_8k_partition_invariance(procedural):
    """BASELINE configs[4]: 7680x4320 tiled over 8 ranks equals the single-GPU frame bit for bit."""
    sc = scenes.build_config("mixed1024_8k")
    sc.scene["reflect_depth"] = 2                    # keeps the test short; tiling does not depend on depth
    full, _ = gpu_render(sc, procedural)
    tiled = _render_partitioned(sc, procedural, 8)
    assert np.array_equal(full.view(np.uint32), tiled.view(np.uint32))


def test_partition_with_empty_ranks_and_tiny_canvases(procedural):
    """More ranks than 4-scanline blocks (some contexts own no pixels at all) and canvases far smaller than the machine
    (every CTA of the persistent kernel goes straight to its cooperative drain): frames stay bit-identical."""
    for (w, h) in ((30, 10), (64, 48)):
        sc = scenes.synthetic_scene("mini4", w, h, 5)
        want = Oracle(sc, procedural).render()
        full, _ = gpu_render(sc, procedural, kernel=KERNEL_PERSISTENT)
        assert pixel_err(full, want).max() <= TOL
        world, parts = 8, []
        for r in range(world):
            gl = rtb200.GLWrapper(w, h)
            gl.init_window()
            try:
                gl.set_partition(r, world, 4)
                rtb200.setup_scene(gl, sc, procedural)
                gl.set_option("kernel", KERNEL_PERSISTENT)
                gl.set_option("strict", 1)
                gl.draw()
                parts.append(gl.read_pixels())
            finally:
                gl.stop()
        assert sum(p.shape[0] for p in parts) == h
        assert any(p.shape[0] == 0 for p in parts) == ((h + 3) // 4 < world)
        assert np.array_equal(rtb200.gather_rows(parts, h, world, 4), full)


def test_rgba8_readback_is_the_gl_unorm_conversion(procedural):
    """rtb_read_rgba8 = what glReadPixels returns from the reference's RGBA8 colour buffer (GLWrapper.cpp:127,216):
    clamp to [0,1], x255, round to nearest — converted on the device, compared with numpy on the float frame."""
    sc = scenes.synthetic_scene("mini2", 90, 50, 3)
    gl = rtb200.GLWrapper(90, 50)
    gl.init_window()
    try:
        rtb200.setup_scene(gl, sc, procedural)
        gl.draw()
        f = gl.read_pixels()
        u = gl.read_pixels_u8()
    finally:
        gl.stop()
    want = (np.clip(np.nan_to_num(f, nan=0.0), 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    assert u.shape == (50, 90, 4) and np.array_equal(u, want)


def _random_scene(seed, w=72, h=40):
    """A hostile little scene: every primitive class, glass / hollow / diffuse-with-alpha materials, identity AND random
    rotations (axis-aligned boxes give exact zeros, quadrics with the ray along the axis hit the degenerate branch),
    objects thousands of units away (Durand-Kerner overflows to inf / NaN there) and a rotated, displaced camera."""
    from rtb200.scene import SurfaceFactory, quat_from_euler
    rng = scenes.PCG32(1000 + seed)
    sc = scenes._base(w, h, 2 + seed % 5)
    cm = SM.create_material

    def mat():
        kind = rng.index(4)
        col = (rng.range(0, 1), rng.range(0, 1), rng.range(0, 1))
        spec = scenes.SPECULARS[rng.index(5)]
        if kind == 0:
            return cm(col, spec, 0.0)
        if kind == 1:
            return cm(col, spec, rng.range(0.05, 0.9))
        if kind == 2:
            return cm(col, spec, rng.range(0.0, 0.3), rng.range(1.05, 1.6), (rng.range(0, 2), rng.range(0, 2), rng.range(0, 2)), 1)
        return cm(col, spec, 0.0, 0.0, (0, 0, 0), rng.range(0.2, 1.0))

    def centre():
        far = 3000.0 if rng.index(6) == 0 else 1.0
        return (rng.range(-8, 8) * far, rng.range(0.0, 6) * far, rng.range(-2, 14) * far)

    def quat():
        return (0, 0, 0, 1) if rng.index(3) == 0 else tuple(rng.quat())

    for _ in range(3 + rng.index(4)):
        c = centre()
        far = abs(c[0]) > 50 or abs(c[2]) > 50
        sc.spheres.append(SM.create_sphere(c, rng.range(0.3, 1.5) * (800 if far else 1), mat(), bool(rng.index(2))))
    for _ in range(1 + rng.index(3)):
        b = SM.create_box(centre(), (rng.range(0.3, 2), rng.range(0.3, 2), rng.range(0.3, 2)), mat())
        b["quat_rotation"] = quat()
        sc.boxes.append(b)
    for _ in range(1 + rng.index(3)):
        t = SM.create_torus(centre(), (rng.range(0.6, 1.5), rng.range(0.15, 0.5)), mat())
        t["quat_rotation"] = quat()
        sc.toruses.append(t)
    for _ in range(rng.index(3)):
        r = SM.create_ring(centre(), rng.range(0.3, 1.0), rng.range(1.2, 3.0), mat())
        r["quat_rotation"] = quat()
        sc.rings.append(r)
    makers = (lambda m: SurfaceFactory.GetEllipsoid(rng.range(0.4, 1.2), rng.range(0.4, 1.2), rng.range(0.4, 1.2), m),
              lambda m: SurfaceFactory.GetEllipticCone(rng.range(0.3, 1), rng.range(0.3, 1), rng.range(0.5, 1.2), m),
              lambda m: SurfaceFactory.GetEllipticCylinder(rng.range(0.3, 1), rng.range(0.3, 1), m),
              lambda m: SurfaceFactory.GetEllipticParaboloid(rng.range(0.4, 1), rng.range(0.4, 1), m),
              lambda m: SurfaceFactory.GetHyperbolicParaboloid(rng.range(0.4, 1), rng.range(0.4, 1), m),
              lambda m: SurfaceFactory.GetParabolicCylinder(rng.range(0.4, 1), m))
    for _ in range(1 + rng.index(3)):
        q = makers[rng.index(len(makers))](mat())
        c = centre()
        q["pos"] = c
        q["quat_rotation"] = quat()
        q["v_min"] = tuple(np.float32(x) - np.float32(2.5) for x in c)
        q["v_max"] = tuple(np.float32(x) + np.float32(2.5) for x in c)
        sc.surfaces.append(q)
    if rng.index(2):
        sc.planes.append(SM.create_plane((0, 1, 0), (0, -0.5, 0), mat()))
    if rng.index(2):
        sc.lights_point.append(SM.create_light_point((rng.range(-5, 5), rng.range(3, 8), rng.range(-4, 8), 0.3), (1, 0.9, 0.8), rng.range(5, 30)))
    sc.scene["camera_pos"] = (rng.range(-3, 3), rng.range(0.5, 4), rng.range(-10, -4))
    if seed % 3:
        sc.scene["quat_camera_rotation"] = quat_from_euler(rng.range(-0.2, 0.2), rng.range(-0.3, 0.3), rng.range(-0.1, 0.1))
    return sc


@pytest.mark.parametrize("seed", range(40))
def test_randomized_scenes_match_the_oracle_in_both_kernels(seed, procedural):
    sc = _random_scene(seed)
    ost = Stats()
    want = Oracle(sc, procedural).render(stats=ost)
    o = ost.as_dict()
    for k in (KERNEL_QUAD, KERNEL_PERSISTENT):
        got, st = gpu_render(sc, procedural, kernel=k, strict=1, counted=True)
        both_nan = np.isnan(got) & np.isnan(want)
        err = np.where(both_nan, 0.0, np.abs(got - want))
        assert np.nanmax(err) <= TOL and not np.isnan(err).any(), (seed, k, float(np.nanmax(err)), int((err > TOL).sum()))
        c = st.as_dict()
        for key in ("rays_nearest", "rays_shadow", "dk_iterations", "shaded_hits", "light_evals"):
            assert o[key] == c[key], (seed, k, key, o[key], c[key])


@pytest.mark.parametrize("seed", (1, 7, 11, 13, 21, 34))
def test_cooperative_drain_is_bit_identical_to_the_serial_drain(seed, procedural):
    """On a canvas this small every scan of the persistent kernel runs in the cooperative drain (coop_scan: one ray per
    warp, primitives spread over the lanes, order-dependent cases replayed in index order); with "coop" = 0 every warp
    scans serially.  Same bits either way — including NaN hit distances (seed 1: a box whose slab test is 0 * inf)."""
    sc = _random_scene(seed)
    a, _ = gpu_render(sc, procedural, kernel=KERNEL_PERSISTENT, coop=1)
    b, _ = gpu_render(sc, procedural, kernel=KERNEL_PERSISTENT, coop=0)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_cost_ordered_tiles_of_one_share_of_a_partition(procedural):
    """The tile order lives in local tile numbers: rank 3 of an 8-way split (4-scanline blocks) renders the same rows with the
    option on (second and third frame: ordered) and off, and moving the context to another rank drops the old order."""
    sc = scenes.build_config("mixed1024_4k", 1 / 3)                # 1280x720
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    rows = {}
    for lpt in (0, 1):
        gl = rtb200.GLWrapper(w, h)
        gl.init_window()
        try:
            rtb200.setup_scene(gl, sc, procedural)
            gl.set_option("strict", 0)
            gl.set_option("lpt", lpt)
            out = []
            for rank in (3, 5):
                gl.set_partition(rank, 8, 4)
                for _ in range(3):
                    gl.draw()
                out.append(gl.read_pixels().copy())
            rows[lpt] = out
        finally:
            gl.stop()
    for a, b in zip(rows[0], rows[1]):
        assert a.shape[0] in (88, 92) and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert not np.array_equal(rows[1][0][:88], rows[1][1][:88])


@pytest.mark.parametrize("strict", (1, 0))
def test_cost_ordered_tiles_render_the_same_frames(strict, procedural):
    """Option "lpt": from the second frame on the persistent kernel hands its tiles out costliest first (the order is a
    counting sort of the previous frame's per-tile path lengths, rtb_api.cu tile_hist / tile_plan / tile_scatter_kernel).  Pixels are independent, so
    every frame — the first in scan order, the later ones in cost order, also after the camera moved and the order is stale —
    carries the bits of a frame rendered with the option off, and the work counters stay equal."""
    sc = scenes.build_config("mixed1024_4k", 1 / 6)                # 640x360: 3600 tiles -> switch the option on explicitly
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    frames = {}
    for lpt in (0, 1):
        gl = rtb200.GLWrapper(w, h)
        gl.init_window()
        try:
            handles = rtb200.setup_scene(gl, sc, procedural)
            gl.set_option("kernel", KERNEL_PERSISTENT)
            gl.set_option("strict", strict)
            gl.set_option("lpt", lpt)
            out = []
            for k in range(3):
                gl.draw()
                out.append(gl.read_pixels())
            scene2 = sc.scene.copy()
            scene2["camera_pos"][0] += 0.75                        # frame 4: the camera moved, the order is the old frame's
            gl.update_buffer(handles["scene_buf"], np.ascontiguousarray(scene2).reshape(1))
            gl.draw()
            out.append(gl.read_pixels())
            out.append(gl.draw_counted().as_dict())
            frames[lpt] = out
        finally:
            gl.stop()
    for k in range(4):
        assert np.array_equal(frames[0][k].view(np.uint32), frames[1][k].view(np.uint32)), k
    assert np.array_equal(frames[1][0].view(np.uint32), frames[1][2].view(np.uint32))
    assert not np.array_equal(frames[1][2], frames[1][3])         # the camera did move
    work = [{k: v for k, v in frames[lpt][4].items() if not k.endswith("_ms")} for lpt in (0, 1)]
    assert work[0] == work[1] and work[0]["dk_iterations"] > 0


@pytest.mark.parametrize("case", ("spheres4k/12", "mixed1024/16"))
def test_the_24_warp_variant_of_the_persistent_kernel_renders_the_same_bits(case, procedural):
    """Option "wide": scenes without tori run the persistent kernel with 24 warps and 80 registers per thread instead of 20 and 96
    (automatic; forced on here for a scene with tori as well).  Same code, another register budget: both builds give the bits
    and the work counters of the 20-warp variant."""
    sc = CASES[case]()
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    for strict in (1, 0):
        out = {}
        for wide in (0, 1):
            gl = rtb200.GLWrapper(w, h)
            gl.init_window()
            try:
                rtb200.setup_scene(gl, sc, procedural)
                gl.set_option("kernel", KERNEL_PERSISTENT)
                gl.set_option("strict", strict)
                gl.set_option("wide", wide)
                st = gl.draw_counted()
                assert st.block == (768 if wide else 640)
                out[wide] = (gl.read_pixels(), {k: v for k, v in st.as_dict().items() if k not in ("kernel_ms", "block", "smem_bytes", "grid")})
            finally:
                gl.stop()
        assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32)), (case, strict)
        assert out[0][1] == out[1][1], (case, strict)


def test_the_tile_order_is_a_permutation_with_the_cheapest_tiles_last(procedural):
    """rtb_tile_order: after a frame the hand-out order of the next one is a permutation of the tiles whose tail holds the
    cheapest tiles (3 x the tiles in flight, at most half the frame; by cost bucket, costliest first) and whose head keeps
    the scan order."""
    sc = scenes.build_config("mixed1024_4k", 1 / 3)                # 1280x720: 14 400 tiles, the option is on by itself
    w, h = int(sc.scene["canvas_width"]), int(sc.scene["canvas_height"])
    gl = rtb200.GLWrapper(w, h)
    gl.init_window()
    try:
        rtb200.setup_scene(gl, sc, procedural)
        gl.set_option("strict", 0)
        assert gl.tile_order() is None
        gl.draw()
        cost, order = gl.tile_order()
        st = gl.stats()
    finally:
        gl.stop()
    n = ((w + 7) // 8) * ((h + 3) // 4)
    assert len(order) == n and np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    assert cost.min() >= 32 and int(cost.sum()) >= 32 * n         # every pixel's path has at least length 1
    tail = min(3 * (st.grid * st.block // 32), n // 2)
    head_t, tail_t = order[: n - tail], order[n - tail:]
    bucket = np.minimum(cost >> 2, 255)
    assert np.all(np.diff(head_t.astype(np.int64)) > 0)             # scan order
    assert bucket[tail_t].max() <= bucket[head_t].min()
    assert np.all(np.diff(bucket[tail_t].astype(np.int64)) <= 0)    # costliest bucket of the tail first
    assert bucket[tail_t][-1] == bucket.min()
