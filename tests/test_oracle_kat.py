"""Known-answer tests of single shader functions with closed-form answers, and one test per quirk listed in
SURVEY.md 8a (they are easy to 'fix' by accident)."""
import math

import numpy as np
import pytest

from oracle.binding import Oracle
from rtb200 import scenes
from rtb200.scene import SceneManager as SM, SurfaceFactory as SF, quat_angle_axis

T_SPHERE, T_PLANE, T_SURFACE, T_BOX, T_TORUS, T_RING, T_LIGHT = range(7)
mat = SM.create_material((1, 1, 1), 0, 0.0)


def empty(w=64, h=64, it=1):
    sc = scenes._base(w, h, it)
    sc.lights_point.clear()
    sc.lights_direct.clear()
    return sc


def test_sphere_closed_form():
    sc = empty()
    sc.spheres.append(SM.create_sphere((0, 0, 6), 1, mat))
    hit, t, _ = Oracle(sc).intersect(T_SPHERE, 0, (0, 0, -5), (0, 0, 1))
    assert hit and t == 10.0                      # camera (0,0,-5), sphere (0,0,6) r=1
    hit, _, _ = Oracle(sc).intersect(T_SPHERE, 0, (0, 0, -5), (0, 0, 1), tmin=9.0)
    assert not hit                                # accept only 0 < t < tmin


def test_hollow_sphere_returns_far_root_from_inside():
    sc = empty()
    sc.spheres.append(SM.create_sphere((0, 0, 0), 2, mat, hollow=True))
    sc.spheres.append(SM.create_sphere((0, 0, 0), 2, mat, hollow=False))
    o = Oracle(sc)
    assert o.intersect(T_SPHERE, 0, (0, 0, 0), (0, 0, 1)) == (True, 2.0, 0)
    assert o.intersect(T_SPHERE, 1, (0, 0, 0), (0, 0, 1))[0] is False
    assert o.in_shadow((0, 0, 0), (0, 0, 1), 100.0) == 0.0      # inShadow passes hollow=false (rt.frag:636)


def test_plane_is_one_sided():
    sc = empty()
    sc.planes.append(SM.create_plane((0, 1, 0), (0, 0, 0), mat))
    o = Oracle(sc)
    hit, t, _ = o.intersect(T_PLANE, 0, (0, 3, 0), (0, -1, 0))
    assert hit and t == 3.0
    assert not o.intersect(T_PLANE, 0, (0, -3, 0), (0, 1, 0))[0]      # back face: PLANE_ONESIDE (rt.frag:21,358)
    assert o.in_shadow((0, 3, 0), (0, -1, 0), 100.0) == 0.0           # planes never occlude (rt.frag:652)


def test_box_slab_and_inside_quirk():
    sc = empty()
    sc.boxes.append(SM.create_box((0, 0, 5), (1, 2, 3), mat))
    o = Oracle(sc)
    hit, t, _ = o.intersect(T_BOX, 0, (0.25, 0.5, -4), (0, 0, 1))
    assert hit and t == 6.0
    hit, t, _ = o.intersect(T_BOX, 0, (0.25, 0.5, 5), (0, 0, 1))      # origin inside: negative tN is returned as a hit
    assert hit and t == -3.0                                          # quirk 3, rt.frag:417-423


def test_box_rotated_matches_axis_aligned_after_counter_rotation():
    sc = empty()
    q = quat_angle_axis(math.radians(90), (0, 1, 0))
    b = SM.create_box((0, 0, 0), (1, 2, 3), mat)
    b["quat_rotation"] = q
    sc.boxes.append(b)
    d = np.array([1, 0.01, 0.02]) / np.linalg.norm([1, 0.01, 0.02])     # (an axis-exact ray gives inf*0 = NaN in the slab test)
    hit, t, _ = Oracle(sc).intersect(T_BOX, 0, (-10, 0, 0), d)
    assert hit and abs(t - 7.0 / d[0]) < 1e-4     # 90 deg about y: the half-extent 3 now lies along x


def test_ring_uses_squared_radii():
    sc = empty()
    sc.rings.append(SM.create_ring((0, 0, 4), 1.0, 2.0, mat))
    assert float(sc.rings[0]["r1"]) == 1.0 and float(sc.rings[0]["r2"]) == 4.0
    o = Oracle(sc)
    assert o.intersect(T_RING, 0, (1.5, 0, 0), (0, 0, 1)) == (True, 4.0, 0)
    assert not o.intersect(T_RING, 0, (0.5, 0, 0), (0, 0, 1))[0]      # inside the hole
    assert not o.intersect(T_RING, 0, (2.5, 0, 0), (0, 0, 1))[0]


def test_quadric_ellipsoid_equals_sphere():
    sc = empty()
    s = SF.GetEllipsoid(2, 2, 2, mat)
    s["pos"] = (0, 0, 10)
    sc.surfaces.append(s)
    hit, t, _ = Oracle(sc).intersect(T_SURFACE, 0, (0, 0, 0), (0, 0, 1))
    assert hit and abs(t - 8.0) < 1e-5


def test_quadric_clip_box_falls_back_to_far_root():
    sc = empty()
    s = SF.GetEllipticCylinder(1, 1, mat)          # x^2 + y^2 = 1, infinite along z
    s["v_min"] = (-10, -10, 0.5)                   # world-space clip keeps only z > 0.5
    sc.surfaces.append(s)
    hit, t, _ = Oracle(sc).intersect(T_SURFACE, 0, (-5, 0, 1), (1, 0, 0))
    assert hit and abs(t - 4.0) < 1e-5
    hit, t, _ = Oracle(sc).intersect(T_SURFACE, 0, (-5, 0, 0), (1, 0, 0))      # z = 0 is clipped away at both roots
    assert not hit


def test_quadric_degenerate_branch_accepts_t_greater_than_tmin():
    """quirk 2, rt.frag:541-545: when |p2| < 1e-6 the shader returns t = -p3/p1 and accepts it if t > tmin."""
    sc = empty()
    s = SF.GetParabolicCylinder(0.5, mat)          # x^2 + y = 0: a ray with d1 = 0 makes p2 = 0
    sc.surfaces.append(s)
    o = Oracle(sc)
    hit, t, _ = o.intersect(T_SURFACE, 0, (1, 5, 0), (0, -1, 0), tmin=1e6)
    assert abs(t - 6.0) < 1e-5 and not hit         # 6 > 1e6 is false
    hit, t, _ = o.intersect(T_SURFACE, 0, (1, 5, 0), (0, -1, 0), tmin=2.0)
    assert hit                                     # ... but 6 > 2 is "a hit" (sic)


def quartic_roots_f64(ro, rd, R, r):
    ro, rd = np.asarray(ro, np.float64), np.asarray(rd, np.float64)
    a, b, c = rd @ rd, ro @ rd, ro @ ro + R * R - r * r
    axy, bxy, cxy = rd[:2] @ rd[:2], ro[:2] @ rd[:2], ro[:2] @ ro[:2]
    p = np.polymul([a, 2 * b, c], [a, 2 * b, c])
    p = np.polysub(p, 4 * R * R * np.array([axy, 2 * bxy, cxy]))
    roots = np.roots(p)
    real = roots[(np.abs(roots.imag) < 1e-6) & (roots.real > 0)].real
    return np.sort(real)


def test_torus_durand_kerner_finds_the_nearest_real_root():
    sc = empty()
    sc.toruses.append(SM.create_torus((0, 0, 0), (1.0, 0.4), mat))
    o = Oracle(sc)
    rng = np.random.default_rng(3)
    n_hit = 0
    for _ in range(200):
        ro = rng.uniform(-3, 3, 3).astype(np.float32)
        if np.linalg.norm(ro) < 1.6:
            continue
        rd = (-ro + rng.normal(scale=0.5, size=3)).astype(np.float32)
        rd = (rd / np.linalg.norm(rd)).astype(np.float32)
        hit, t, k = o.intersect(T_TORUS, 0, ro, rd)
        want = quartic_roots_f64(ro, rd, 1.0, 0.4)
        assert 1 <= k <= 60
        if hit:
            n_hit += 1
            assert len(want) and abs(t - want[0]) < 5e-3            # eps = 1e-3 in the solver (rt.frag:463)
        elif len(want) >= 2 and want[1] - want[0] > 0.05:
            pytest.fail(f"missed a clean root {want} (k={k})")
    assert n_hit > 30


def test_torus_hit_range_is_capped_at_100():
    sc = empty()
    sc.toruses.append(SM.create_torus((0, 0, 150), (1.0, 0.4), mat))
    assert not Oracle(sc).intersect(T_TORUS, 0, (1, 0, 0), (0, 0, 1))[0]       # t ~ 149.6 > 100 (rt.frag:486)


def test_calc_inter_order_and_ties():
    """Strict t<tmin: a later primitive at the same distance does not replace an earlier one (rt.frag:587-628)."""
    sc = empty()
    sc.spheres.append(SM.create_sphere((0, 0, 6), 1, mat))
    sc.spheres.append(SM.create_sphere((0, 0, 6), 1, mat))
    sc.lights_point.append(SM.create_light_point((0, 0, 6, 1), (1, 1, 1), 1))
    t, num, typ = Oracle(sc).calc_inter((0, 0, -5), (0, 0, 1))
    assert (t, num, typ) == (10.0, 0, T_SPHERE)
    t, num, typ = Oracle(sc).calc_inter((0, 0, -5), (0, 1, 0), num=7, type_=3)
    assert t == 1e6 and (num, typ) == (7, 3)       # quirk 1: num/type untouched on a miss


def test_ray_direction_convention():
    sc = empty(200, 100)
    o = Oracle(sc)
    c = o.ray_dir(99, 49)                          # pixel centre (99.5, 49.5): just left/below the optical axis
    assert c[0] < 0 and c[1] < 0 and c[2] > 0.99
    top_right = o.ray_dir(199, 99)
    assert top_right[0] > 0 and top_right[1] > 0   # row 0 is the BOTTOM scanline (GL window coordinates)
    v = np.array([(199.5 - 100) / 100, (99.5 - 50) / 100, 1.0])      # both axes divided by H (rt.frag:315)
    assert np.allclose(top_right, v / np.linalg.norm(v), atol=1e-6)


def test_light_sphere_is_visible_and_ends_the_path():
    sc = empty(16, 16, 3)
    sc.scene["camera_pos"] = (0, 0, -5)
    sc.lights_point.append(SM.create_light_point((0, 0, 5, 3.0), (0.25, 0.5, 0.75), 10))
    img = Oracle(sc).render()
    assert np.allclose(img[8, 8], [0.25, 0.5, 0.75, 1.0])               # rt.frag:829-832: colour * mask, no shading


def test_sky_ignores_bg_color_and_alpha_is_one(procedural):
    sc = empty(16, 16, 2)
    sc.scene["bg_color"] = (1, 0, 0)
    img = Oracle(sc, None).render()
    assert np.array_equal(img[..., :3], np.zeros_like(img[..., :3])) and (img[..., 3] == 1).all()    # unbound cubemap samples 0
    img = Oracle(sc, procedural).render()
    assert img[..., :3].max() > 0.5


def test_zero_iterations_renders_black():
    sc = scenes.synthetic_scene("mini1", 16, 16, 0)
    img = Oracle(sc).render()
    assert np.array_equal(img[..., :3], np.zeros((16, 16, 3), np.float32))


def test_ambient_passes_through_percent_f():
    """quirk 7: AMBIENT_COLOR reaches the shader as std::to_string(float), i.e. rounded to 6 decimals."""
    def render(amb):
        sc = empty(8, 8, 1)
        sc.scene["camera_pos"] = (0, 0, -5)
        sc.ambient_color = (amb,) * 3
        sc.spheres.append(SM.create_sphere((0, 0, 5), 3, SM.create_material((1, 1, 1), 0, 0.0)))
        return Oracle(sc).render()[4, 4, 0]
    assert render(0.1234564) == render(0.123456)
    assert render(0.1234566) == render(0.123457)


def test_textured_ring_shadow_accumulates_alpha(procedural):
    """quirk 8, rt.frag:644-651: a textured ring adds its alpha to the shadow term; an untextured one sets it to 1."""
    sc = empty()
    ring = SM.create_ring((0, 0, 5), 0.5, 3.0, mat)
    ring["textureNum"] = 4
    sc.rings.append(ring)
    sc.rings.append(ring.copy())
    s2 = Oracle(sc, procedural).in_shadow((1.5, 0.3, 0), (0, 0, 1), 100.0)
    sc.rings.pop()
    s1 = Oracle(sc, procedural).in_shadow((1.5, 0.3, 0), (0, 0, 1), 100.0)
    assert 0 < s1 <= 1 and s2 == pytest.approx(min(1.0, 2 * s1), abs=1e-6)
    sc.rings[0]["textureNum"] = 0
    assert Oracle(sc, procedural).in_shadow((1.5, 0.3, 0), (0, 0, 1), 100.0) == 1.0


def test_glass_sphere_path_terminates():
    """quirk 4: refractive hits do not consume iterations (i--); the pinned cap keeps the loop finite."""
    sc = empty(32, 32, 1)
    sc.scene["camera_pos"] = (0, 0, -5)
    sc.spheres.append(SM.create_sphere((0, 0, 2), 1.5, SM.create_material((1, 1, 1), 200, 0.1, 1.125, (1, 0, 2), 1), True))
    img = Oracle(sc).render()
    assert np.isfinite(img).all()


def _two_glass_spheres(absorb=0.4):
    """camera on the z axis, two absorbing glass spheres of index 1 (no bending, no Fresnel reflection on the axis), uniform sky"""
    from rtb200.textures import TextureSet
    sc = empty(512, 512, 4)
    glass = SM.create_material((1, 1, 1), 0, 0.0, refract=1.0, absorb=(absorb, absorb, absorb))
    sc.spheres.append(SM.create_sphere((0, 2, 0), 1.0, glass, hollow=True))     # hollow: the far root is returned from inside (rt.frag:350)
    sc.spheres.append(SM.create_sphere((0, 2, 6), 1.5, glass, hollow=True))
    sky = TextureSet(cube=[np.full((8, 8, 3), 200, dtype=np.uint8) for _ in range(6)])
    return sc, sky


def test_absorb_distance_accumulates_over_the_whole_path():
    """Quirk Q5 (rt.frag:816,859): absorbDistance is initialised once per pixel and only ever grows — the second glass body
    attenuates by exp(-absorb * (d1 + d2)), not by exp(-absorb * d2)."""
    a = 0.4
    sc, sky = _two_glass_spheres(a)
    px = Oracle(sc, sky).render_quads([256], [256])[0, 0, :3]          # the pixel next to the optical axis
    # chords of that pixel's ray through the two spheres, in fp64
    rd = np.array([0.5 / 512, 0.5 / 512, 1.0]); rd /= np.linalg.norm(rd)
    ro = np.array([0.0, 2.0, -12.0])

    def chord(c, r):
        oc = ro - np.array(c, dtype=float)
        b = oc @ rd
        return 2.0 * math.sqrt(b * b - (oc @ oc - r * r))

    d1, d2 = chord((0, 2, 0), 1.0), chord((0, 2, 6), 1.5)
    sky_c = 200 / 255
    cumulative = sky_c * math.exp(-a * d1) * math.exp(-a * (d1 + d2))
    reset = sky_c * math.exp(-a * d1) * math.exp(-a * d2)
    assert abs(reset - cumulative) > 0.05                               # the two readings are far apart ...
    assert np.allclose(px, cumulative, atol=3e-3), (px, cumulative, reset)   # ... and the shader's is the cumulative one (bias offsets ~1e-3)


def test_update_buffers_never_resends_the_directional_lights(procedural):
    """Quirk Q10 (SceneManager.cpp:254 vs :266-276): lights_direct_buf is written by init_buffers only; update_buffers, which
    runs every frame, sends the other eight blocks.  A drop-in must keep the init-time bytes — checked here on the host logic
    of the Python mirror (the C++ host inherits it from the reference's unchanged SceneManager.cpp)."""
    import rtb200
    sent = []

    class Recorder(rtb200.GLWrapper):
        def init_window(self):
            return True

        def init_shaders(self, d):
            pass

        def load_cubemap(self, faces, genMipmap=False):
            return 1

        def load_texture(self, *a, **k):
            return 1

        def init_buffer(self, name, bindingPoint, data):
            sent.append(("init", name))
            return len(sent)

        def update_buffer(self, ubo, data):
            sent.append(("update", ubo))

    sc = scenes.default_scene(64, 48, 2)
    gl = Recorder(64, 48)
    handles = rtb200.setup_scene(gl, sc, procedural)
    assert ("init", "lights_direct_buf") in sent
    sent.clear()
    rtb200.update_buffers(gl, sc, handles)                             # SceneManager::update_buffers, SceneManager.cpp:266-276
    updated = {h for kind, h in sent if kind == "update"}
    assert handles["lights_direct_buf"] not in updated
    assert updated == {handles[n] for n in handles if n != "lights_direct_buf" and (n == "scene_buf" or len(sc.array(n[:-4])))}
