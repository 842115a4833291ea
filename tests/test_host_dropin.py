"""The C++ drop-in: the reference's UNCHANGED main.cpp + SceneManager.cpp, built against host/GLWrapper.{h,cpp}
(raytracing-opengl_b200/host/Makefile), rendering through librtb200.so."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "raytracing-opengl_b200", "host")
BIN = os.path.join(HOST, "build", "rt_headless")


def test_replacement_glwrapper_declares_every_public_member_of_the_reference():
    text = open(os.path.join(HOST, "GLWrapper.h")).read()
    for decl in ("GLWrapper(int width, int height, bool fullScreen);", "GLWrapper(bool fullScreen);", "int getWidth();", "int getHeight();",
                 "GLuint getProgramId();", "bool init_window();", "void init_shaders(rt_defines& defines);", "void set_skybox(unsigned int textureId);",
                 "void stop();", "void enable_SMAA(SMAA_PRESET preset);", "GLFWwindow* window;", "void draw();",
                 "static GLuint load_cubemap(std::vector<std::string> faces, bool genMipmap = false);",
                 "GLuint load_texture(int texNum, const char* name, const char* uniformName, GLuint wrapMode = GL_REPEAT);",
                 "void init_buffer(GLuint* ubo, const char* name, int bindingPoint, size_t size, void* data) const;",
                 "static void update_buffer(GLuint ubo, size_t size, void* data);"):
        assert decl in text, decl                      # src/GLWrapper.h:17-38, verbatim signatures


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources only exist in the build container")
def test_reference_sources_are_compiled_where_they_lie_not_copied():
    for f in ("main.cpp", "SceneManager.cpp", "SceneManager.h", "scene.h", "Surface.h"):
        p = os.path.join(HOST, "build", "refsrc", f)
        if os.path.exists(p):
            assert os.path.islink(p) and os.path.realpath(p).startswith("/root/reference/"), p
    tracked = subprocess.check_output(["git", "ls-files"], cwd=ROOT).decode().split()
    assert not any(t.endswith(("main.cpp", "SceneManager.cpp", "scene.h", "Surface.h", "rt.frag", "stb_image.h")) for t in tracked)


@pytest.mark.skipif(not os.path.isfile(BIN), reason="rt_headless is built only where /root/reference exists")
def test_headless_binary_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(BIN), reason="rt_headless did not travel (built only where /root/reference exists)")
def test_unchanged_main_cpp_renders_the_default_scene_like_the_oracle():
    """BASELINE configs[0]: default main.cpp scene, 256x256, 1 bounce.  The frame written by the C++ drop-in must
    match the oracle fed with the SAME uniform-buffer bytes and the SAME decoded textures (both dumped by the run)."""
    import rtb200  # noqa: F401
    from oracle.binding import Oracle
    from rtb200 import scene as S
    from rtb200.scene import SceneContainer
    from rtb200.textures import TextureSet
    from util import pixel_err
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, RT_WIDTH="256", RT_HEIGHT="256", RT_ITERATIONS="1", RT_FRAMES="1", RT_DUMP_DIR=td, RT_STRICT="1")
        r = subprocess.run([BIN], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        frame = np.load(os.path.join(td, "frame_0000.npy"))
        assert frame.shape == (256, 256, 4)
        sc = SceneContainer()
        for name, attr, dt in (("spheres_buf", "spheres", S.rt_sphere), ("surfaces_buf", "surfaces", S.rt_surface), ("boxes_buf", "boxes", S.rt_box),
                               ("toruses_buf", "toruses", S.rt_torus), ("rings_buf", "rings", S.rt_ring), ("lights_point_buf", "lights_point", S.rt_light_point),
                               ("lights_direct_buf", "lights_direct", S.rt_light_direct)):
            raw = np.load(os.path.join(td, name + ".npy"))           # the bytes the unchanged SceneManager uploaded for this frame
            setattr(sc, attr, np.frombuffer(raw.tobytes(), dtype=dt).copy())
        assert len(sc.spheres) == 6 and len(sc.boxes) == 2 and len(sc.toruses) == 1 and len(sc.rings) == 1 and len(sc.surfaces) == 2
        sc.scene = np.frombuffer(np.load(os.path.join(td, "scene_buf.npy")).tobytes(), dtype=S.rt_scene)[0].copy()
        assert int(sc.scene["canvas_width"]) == 256 and int(sc.scene["reflect_depth"]) == 5       # SceneManager.cpp:233
        sc.scene["reflect_depth"] = 1                                # RT_ITERATIONS=1 overrides the {ITERATIONS} token
        sc.ambient_color, sc.shadow_ambient = (0.025,) * 3, (0.1,) * 3                            # main.cpp:47-48
        ts = TextureSet(cube=[np.load(os.path.join(td, f"cube_{f}.npy")) for f in range(6)],
                        tex2d={u: np.load(os.path.join(td, f"tex_{u}.npy")) for u in range(1, 6)})
        assert ts.tex2d[4].shape == (500, 8192, 4)
        want = Oracle(sc, ts).render()
        err = pixel_err(frame, want)
        assert err.max() <= 1e-4, (float(err.max()), int((err > 1e-4).sum()))
        # and the restated Python scene script agrees with what main.cpp built
        from rtb200 import scenes
        mine = scenes.default_scene(256, 256)
        def fields(arr):            # every named field, recursively (padding members may hold garbage and are never read)
            out = []
            for n in arr.dtype.names:
                out += fields(arr[n]) if arr[n].dtype.names else [np.asarray(arr[n], dtype=np.float64).ravel()]
            return out
        for attr in ("spheres", "surfaces", "boxes", "toruses", "rings", "lights_point", "lights_direct"):
            a, b = np.concatenate(fields(mine.array(attr))), np.concatenate(fields(sc.array(attr)))
            assert a.shape == b.shape and np.allclose(a, b, rtol=3e-7, atol=1e-30), attr
        out = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out):                                       # keep the dump: tests/golden/default_scene_t0.npz is made from it
            np.savez_compressed(os.path.join(out, "default_scene_t0.npz"), width=256, height=256,
                                **{a: np.frombuffer(sc.array(a).tobytes(), dtype=np.uint8) for a in
                                   ("spheres", "surfaces", "boxes", "toruses", "rings", "lights_point", "lights_direct")})


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(BIN), reason="rt_headless did not travel (built only where /root/reference exists)")
def test_unchanged_frame_loop_animates_and_reuploads_every_frame():
    """SURVEY.md 8f-2, the per-frame update path: main.cpp's loop (update_scene :197-246 -> SceneManager::update ->
    update_buffers :266-276 -> draw) runs three frames on the deterministic clock; the LAST frame must match the oracle
    fed with the uniform-buffer bytes of the last upload, and must differ from the first (the scene moved)."""
    import rtb200  # noqa: F401
    from oracle.binding import Oracle
    from rtb200 import scene as S
    from rtb200.scene import SceneContainer
    from rtb200.textures import TextureSet
    from util import pixel_err
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, RT_WIDTH="192", RT_HEIGHT="108", RT_ITERATIONS="3", RT_FRAMES="3", RT_DUMP_DIR=td, RT_STRICT="1")
        r = subprocess.run([BIN], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        frames = [np.load(os.path.join(td, f"frame_{i:04d}.npy")) for i in range(3)]
        assert frames[2].shape == (108, 192, 4)
        assert np.abs(frames[2] - frames[0]).max() > 1e-3            # the box and the torus spin (main.cpp:234-245)
        sc = SceneContainer()
        for name, attr, dt in (("spheres_buf", "spheres", S.rt_sphere), ("surfaces_buf", "surfaces", S.rt_surface), ("boxes_buf", "boxes", S.rt_box),
                               ("toruses_buf", "toruses", S.rt_torus), ("rings_buf", "rings", S.rt_ring), ("lights_point_buf", "lights_point", S.rt_light_point),
                               ("lights_direct_buf", "lights_direct", S.rt_light_direct)):
            setattr(sc, attr, np.frombuffer(np.load(os.path.join(td, name + ".npy")).tobytes(), dtype=dt).copy())
        sc.scene = np.frombuffer(np.load(os.path.join(td, "scene_buf.npy")).tobytes(), dtype=S.rt_scene)[0].copy()
        sc.scene["reflect_depth"] = 3
        sc.ambient_color, sc.shadow_ambient = (0.025,) * 3, (0.1,) * 3
        ts = TextureSet(cube=[np.load(os.path.join(td, f"cube_{f}.npy")) for f in range(6)],
                        tex2d={u: np.load(os.path.join(td, f"tex_{u}.npy")) for u in range(1, 6)})
        err = pixel_err(frames[2], Oracle(sc, ts).render())
        assert err.max() <= 1e-4, (float(err.max()), int((err > 1e-4).sum()))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isfile(BIN), reason="rt_headless did not travel (built only where /root/reference exists)")
def test_unchanged_main_cpp_gets_its_smaa_post_pass():
    """main.cpp:32 calls enable_SMAA(ULTRA) before init_window(): what the drop-in puts "on screen" is the ray-traced frame after
    the reference's three SMAA passes (GLWrapper.cpp:173-204) — checked against the reference's own SMAA.h compiled as C++."""
    from oracle.smaa_binding import PRESETS, have_smaa_ref, smaa_ref
    if not have_smaa_ref():
        pytest.skip("oracle/_ref/libsmaa_ref.so did not travel")
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, RT_WIDTH="320", RT_HEIGHT="180", RT_ITERATIONS="3", RT_FRAMES="1", RT_DUMP_DIR=td, RT_STRICT="1")
        r = subprocess.run([BIN], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        frame = np.load(os.path.join(td, "frame_0000.npy"))
        screen = np.load(os.path.join(td, "screen_0000.npy"))
        raw8 = (np.clip(frame, 0, 1) * 255.0 + 0.5).astype(np.uint8)
        want, _, _ = smaa_ref(raw8, PRESETS["ULTRA"])
        assert np.array_equal(screen, want)
        assert (screen != raw8).any()
        env["RT_SMAA"] = "0"                                          # the headless switch
        r = subprocess.run([BIN], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0 and not os.path.isfile(os.path.join(td, "screen_0001.npy"))
