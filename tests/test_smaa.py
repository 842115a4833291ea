"""SMAA post-pass: the CUDA passes (csrc/smaa.cu) against the reference's OWN SMAA.h compiled as C++
(oracle/_ref/libsmaa_ref.so, oracle/build_smaa_ref.py).  The lookup tables (src/AreaTex.h, src/SearchTex.h) are not in this
repository; the tests read them out of the compiled reference library, so everything here needs it (it is built where
/root/reference exists and travels to the GPU box)."""
import numpy as np
import pytest

from oracle.smaa_binding import PRESETS, have_smaa_ref, smaa_ref, tables

pytestmark = pytest.mark.skipif(not have_smaa_ref(), reason="oracle/_ref/libsmaa_ref.so is built where /root/reference exists")


def _shapes(w, h, seed=1):
    """axis-aligned boxes, thin lines, diagonals at several slopes, a disc and noise: every SMAA pattern class"""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w, 4), np.uint8)
    img[..., :3] = 30
    img[..., 3] = 255
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(6):
        x0, y0 = rng.integers(0, w - 20), rng.integers(0, h - 20)
        img[y0:y0 + rng.integers(6, 40), x0:x0 + rng.integers(6, 60), :3] = rng.integers(60, 255, 3)
    for k, slope in enumerate((1.0, -1.0, 0.5, 2.0, 0.2, -3.0)):
        m = np.abs((yy - h / 2) - slope * (xx - w / 2) - 7 * k) < 2.5
        img[m, :3] = rng.integers(80, 255, 3)
    m = (xx - w * 0.7) ** 2 + (yy - h * 0.3) ** 2 < (min(w, h) * 0.2) ** 2
    img[m, :3] = (220, 40, 90)
    noise = rng.integers(0, 255, (h // 4, w // 4, 3)).astype(np.uint8)
    img[: h // 4, : w // 4, :3] = noise
    return img


def test_reference_smaa_leaves_a_flat_image_alone_and_blends_a_step():
    flat = np.full((32, 48, 4), 128, np.uint8)
    out, edges, blend = smaa_ref(flat, PRESETS["ULTRA"])
    assert np.array_equal(out, flat) and not edges.any() and not blend.any()
    step = flat.copy()
    step[:, 24:, :3] = 250
    step[10:, 24:26, :3] = 128                       # a one-pixel-high jog in the vertical edge: something to smooth
    out, edges, blend = smaa_ref(step, PRESETS["ULTRA"])
    assert edges[:, 24, 0].any() and not edges[:, :20].any()         # red channel = edge at the LEFT of the pixel
    assert blend.any() and (out != step).any()
    assert np.array_equal(out[:, :20], step[:, :20])                 # nothing changes away from the edge


def test_presets_differ_in_what_they_detect():
    img = _shapes(160, 96)
    n = {name: int(smaa_ref(img, p)[1].any(axis=2).sum()) for name, p in PRESETS.items()}
    assert n["LOW"] < n["MEDIUM"] <= n["HIGH"] < n["ULTRA"]          # thresholds 0.15 / 0.1 / 0.1 / 0.05 (SMAA.h:304-324)


def test_tables_have_the_documented_shapes():
    area, search = tables()
    assert area.shape == (560, 160, 2) and search.shape == (16, 64)   # AreaTex.h:33-34, SearchTex.h:33-34


def test_unorm8_decode_is_the_ieee_quotient():
    """csrc/smaa.cu u2f(): q = v * (1/255); q + fma(-q, 255, v) * (1/255) equals the IEEE quotient v / 255.0f for every byte
    (the kernels decode texels with it instead of a division; the compiled reference divides)."""
    import ctypes, ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.fmaf.restype = ctypes.c_float
    libm.fmaf.argtypes = [ctypes.c_float] * 3
    rcp = np.float32(1.0) / np.float32(255.0)
    plain_differs = 0
    for v in range(256):
        x = np.float32(v)
        q = np.float32(x * rcp)
        got = np.float32(libm.fmaf(libm.fmaf(-q, 255.0, x), rcp, q))
        assert got == np.float32(x / np.float32(255.0)), v
        plain_differs += int(q != np.float32(x / np.float32(255.0)))
    assert plain_differs > 0                                        # the correction step is needed


# ---------------------------------------------------------------- GPU
def _gl(w, h):
    import rtb200
    gl = rtb200.GLWrapper(w, h)
    gl.init_window()
    gl.smaa_set_tables(*tables())
    return gl


@pytest.mark.gpu
@pytest.mark.parametrize("preset", list(PRESETS))
@pytest.mark.parametrize("size", [(320, 180), (250, 130), (64, 40)])
def test_cuda_smaa_matches_the_reference_shader(preset, size):
    w, h = size
    gl = _gl(w, h)
    try:
        gl.enable_SMAA(PRESETS[preset])
        for seed in (1, 2):
            img = _shapes(w, h, seed)
            want_out, want_edges, want_blend = smaa_ref(img, PRESETS[preset])
            for compact in (1, 0):                                  # pass 2 over the compacted edge pixels (default) / over every pixel
                gl.set_option("smaa_compact", compact)
                out, edges, blend, ms = gl.smaa_apply(img)
                assert np.array_equal(edges, want_edges), f"{preset} {size} compact={compact}: {int((edges != want_edges).any(axis=2).sum())} edge pixels differ"
                assert np.array_equal(blend, want_blend), f"{preset} {size} compact={compact}: {int((blend != want_blend).any(axis=2).sum())} weight pixels differ"
                assert np.array_equal(out, want_out), f"{preset} {size} compact={compact}: {int((out != want_out).any(axis=2).sum())} output pixels differ"
                assert ms > 0
    finally:
        gl.stop()


@pytest.mark.gpu
def test_draw_with_smaa_enabled_post_processes_the_ray_traced_frame(procedural):
    """GLWrapper::draw() with SMAA on (GLWrapper.cpp:155-204): ray trace -> RGBA8 colour target -> three passes -> screen."""
    import rtb200
    from rtb200 import scenes
    sc = scenes.default_scene(256, 144, 3)
    gl = _gl(256, 144)
    try:
        rtb200.setup_scene(gl, sc, procedural)
        gl.draw()
        raw8 = gl.read_pixels_u8()                                  # SMAA off: the quantised ray-traced frame
        frame = gl.read_pixels()
        assert np.array_equal(raw8, (np.clip(frame, 0, 1) * 255.0 + 0.5).astype(np.uint8))
        gl.enable_SMAA(PRESETS["ULTRA"])                            # main.cpp:32
        gl.draw()
        post = gl.read_pixels_u8()
        want, _, _ = smaa_ref(raw8, PRESETS["ULTRA"])
        assert np.array_equal(post, want)
        assert (post != raw8).any()                                 # it did something
        assert np.array_equal(gl.read_pixels(), frame)              # the float frame stays the ray-traced one
        assert gl.smaa_last_ms() > 0
        gl.enable_SMAA(None)
        gl.draw()
        assert np.array_equal(gl.read_pixels_u8(), raw8)
    finally:
        gl.stop()


@pytest.mark.gpu
def test_smaa_needs_its_tables():
    import rtb200
    gl = rtb200.GLWrapper(64, 40)
    gl.init_window()
    try:
        gl.enable_SMAA(PRESETS["ULTRA"])
        with pytest.raises(rtb200.RtbError, match="tables were never set"):
            gl.smaa_apply(np.zeros((40, 64, 4), np.uint8))
    finally:
        gl.stop()
