"""the reference's OWN GLWrapper.cpp (with main.cpp and SceneManager.cpp, all three unchanged, compiled against
the reference's own glad / GLFW / shader headers) running on the fake OpenGL driver raytracing-opengl_b200/host/fakegl/fakegl.cpp,
which renders with librtb200.so (build/rt_fakegl, raytracing-opengl_b200/host/Makefile)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "raytracing-opengl_b200", "host")
FAKEGL = os.path.join(HOST, "build", "rt_fakegl")
HEADLESS = os.path.join(HOST, "build", "rt_headless")
BLOCKS = {"spheres": "spheres_buf", "surfaces": "surfaces_buf", "boxes": "boxes_buf", "toruses": "toruses_buf", "rings": "rings_buf",
          "lights_point": "lights_point_buf", "lights_direct": "lights_direct_buf", "scene": "scene_buf"}


def test_fake_driver_defines_every_gl_and_glfw_symbol_the_unchanged_sources_need():
    """The 45 glad pointers and 22 GLFW functions of SURVEY.md 8b/8f (nm of the unchanged objects) are all defined in fakegl.cpp."""
    text = open(os.path.join(HOST, "fakegl", "fakegl.cpp")).read()
    gl = ("ActiveTexture AttachShader BindBuffer BindBufferBase BindFramebuffer BindTexture BindVertexArray BufferData BufferSubData "
          "CheckFramebufferStatus Clear ClearColor CompileShader CreateProgram CreateShader DeleteBuffers DeleteFramebuffers DeleteProgram "
          "DeleteShader DeleteTextures DeleteVertexArrays DrawArrays EnableVertexAttribArray FramebufferTexture2D GenBuffers GenFramebuffers "
          "GenTextures GenVertexArrays GenerateMipmap GetError GetProgramInfoLog GetProgramiv GetShaderInfoLog GetShaderiv GetUniformBlockIndex "
          "GetUniformLocation LinkProgram ShaderSource TexImage2D TexParameteri Uniform1i UniformBlockBinding UseProgram VertexAttribPointer "
          "Viewport").split()
    assert len(gl) == 45
    for name in gl:
        assert f" glad_gl{name} = fk_{name};" in text, name
    for name in ("gladLoadGL", "glfwInit", "glfwTerminate", "glfwSetErrorCallback", "glfwGetPrimaryMonitor", "glfwGetVideoMode", "glfwWindowHint",
                 "glfwCreateWindow", "glfwDestroyWindow", "glfwGetWindowSize", "glfwMakeContextCurrent", "glfwGetTime", "glfwPollEvents",
                 "glfwSwapInterval", "glfwSwapBuffers", "glfwWindowShouldClose", "glfwSetWindowShouldClose", "glfwSetWindowUserPointer",
                 "glfwGetWindowUserPointer", "glfwSetCursorPosCallback", "glfwSetKeyCallback", "glfwSetFramebufferSizeCallback", "glfwSetInputMode"):
        assert f" {name}(" in text, name
    assert "struct gladGLversionStruct GLVersion" in text


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources only exist in the build container")
def test_no_reference_source_is_copied_for_the_fake_driver_build():
    tracked = subprocess.check_output(["git", "ls-files"], cwd=ROOT).decode().split()
    assert not any(t.endswith(("GLWrapper.cpp", "shader.h", "SMAA_Builder.h", "glad.h", "glfw3.h", "quad.vert")) and "/host/GLWrapper.cpp" not in t and
                   "shim/" not in t for t in tracked), tracked
    mk = open(os.path.join(HOST, "Makefile")).read()
    assert "build/fgl_%.o: $(REFERENCE)/src/%.cpp" in mk            # compiled where they lie


@pytest.mark.skipif(not os.path.isfile(FAKEGL), reason="rt_fakegl is built only where /root/reference exists")
def test_state_captured_behind_the_unchanged_glwrapper_equals_the_uploaded_scene():
    """No GPU needed: RT_FAKEGL_CAPTURE_ONLY stops at the first ray-trace glDrawArrays and dumps what the driver would hand to the
    C-ABI.  It must be, byte for byte, what the replacement GLWrapper uploaded for the same unchanged main.cpp
    (tests/golden/default_scene_t0.npz), with the specialisation parsed back out of the substituted shader text."""
    golden = np.load(os.path.join(ROOT, "tests", "golden", "default_scene_t0.npz"))
    with tempfile.TemporaryDirectory() as td:
        env = dict(os.environ, RT_WIDTH="256", RT_HEIGHT="256", RT_FRAMES="1", RT_DUMP_DIR=td, RT_FAKEGL_CAPTURE_ONLY="1")
        r = subprocess.run([FAKEGL], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "OpenGL 3.3" in r.stdout and "captured the state of the first ray-trace draw (256x256)" in r.stdout
        for key, fname in BLOCKS.items():
            got = np.load(os.path.join(td, fname + ".npy"))
            assert got.shape == golden[key].shape and (got == golden[key]).all(), key
        assert not os.path.exists(os.path.join(td, "planes_buf.npy"))            # the default scene has no planes: a zero-size block
        d = np.load(os.path.join(td, "defines.npy"))
        assert d[:36].view(np.int32).tolist() == [6, 0, 2, 2, 1, 1, 1, 1, 5]      # scene.h:9-17 order; reflect_depth 5 (SceneManager.cpp:233)
        assert np.allclose(d[36:].view(np.float32), [0.025, 0.025, 0.025, 0.1, 0.1, 0.1], rtol=0, atol=1e-7)
        shapes = {1: (2048, 4096, 3), 2: (2048, 4096, 3), 3: (1024, 2048, 3), 4: (500, 8192, 4), 5: (512, 512, 4)}   # main.cpp:149-153
        for role, shape in shapes.items():
            assert np.load(os.path.join(td, f"tex_{role}.npy"), mmap_mode="r").shape == shape, role
        for f in range(6):
            assert np.load(os.path.join(td, f"cube_{f}.npy"), mmap_mode="r").shape == (2048, 2048, 3)


@pytest.mark.skipif(not os.path.isfile(FAKEGL), reason="rt_fakegl is built only where /root/reference exists")
def test_fake_driver_binary_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([FAKEGL], capture_output=True, text=True, timeout=300, env=dict(os.environ, RT_WIDTH="64", RT_HEIGHT="64"))
    assert r.returncode != 0 and "no CUDA device" in (r.stdout + r.stderr)


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.isfile(FAKEGL) and os.path.isfile(HEADLESS)), reason="the host binaries did not travel (built only where /root/reference exists)")
def test_unchanged_glwrapper_on_the_fake_driver_renders_the_same_frames_as_the_replacement_glwrapper():
    """Three animated frames of the unchanged main.cpp loop (per-frame glBufferSubData updates, texture re-binds), 256x256, 1 bounce:
    the frames presented behind the reference's own GLWrapper.cpp are bit-identical to those of the replacement GLWrapper, whose
    frames are checked against the oracle in test_host_dropin.py."""
    frames = {}
    for name, binary in (("fakegl", FAKEGL), ("headless", HEADLESS)):
        with tempfile.TemporaryDirectory() as td:
            env = dict(os.environ, RT_WIDTH="256", RT_HEIGHT="256", RT_ITERATIONS="1", RT_FRAMES="3", RT_DUMP_DIR=td, RT_STRICT="1")
            r = subprocess.run([binary], capture_output=True, text=True, timeout=300, env=env)
            assert r.returncode == 0, name + ": " + r.stdout[-1500:] + r.stderr[-1500:]
            frames[name] = [np.load(os.path.join(td, f"frame_{i:04d}.npy")) for i in range(3)]
    for i in range(3):
        a, b = frames["fakegl"][i], frames["headless"][i]
        assert a.shape == b.shape == (256, 256, 4)
        assert np.array_equal(a, b), f"frame {i}: max |d| = {np.abs(a - b).max()}"
    assert not np.array_equal(frames["fakegl"][0], frames["fakegl"][2])         # the scene is animated (main.cpp:197-246)
