"""Synthetic workloads (BASELINE.json configs) and the restated default scene."""
import os

import numpy as np
import pytest

from rtb200 import scenes
from rtb200.scene import rt_sphere

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pcg32_reference_vector():
    """O'Neill's pcg32 demo: seed 42, stream 54 -> 0xa15c02b7 0x7b47f409 0xba1d3330 ..."""
    r = scenes.PCG32(42, 54)
    assert [r.next_u32() for _ in range(6)] == [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]


@pytest.mark.parametrize("cfg,counts", [("spheres4k", (256, 1, 0, 64, 0, 0)), ("tori1080", (0, 0, 0, 0, 128, 0)),
                                        ("mixed1024_4k", (512, 0, 192, 256, 64, 0)), ("mixed1024_8k", (512, 0, 192, 256, 64, 0))])
def test_baseline_configs_have_the_named_primitive_counts(cfg, counts):
    sc = scenes.build_config(cfg, 1 / 64)
    d = sc.get_defines()
    got = tuple(int(d[k]) for k in ("sphere_size", "plane_size", "surface_size", "box_size", "torus_size", "ring_size"))
    assert got == counts
    assert int(d["iterations"]) == scenes.CONFIGS[cfg][3]
    assert int(d["light_point_size"]) == 1 and int(d["light_direct_size"]) == 1
    assert not sc.uses_textures()


def test_mixed1024_is_1024_primitives_and_fits_uniform_block_limits():
    sc = scenes.build_config("mixed1024_4k", 1 / 64)
    assert len(sc.spheres) + len(sc.boxes) + len(sc.surfaces) + len(sc.toruses) == 1024
    for name in ("spheres", "boxes", "surfaces", "toruses"):
        assert sc.array(name).nbytes <= 65536          # a typical GL uniform-block limit


def test_generator_is_deterministic_and_unit_quaternions():
    a = scenes.synthetic_scene("mixed1024", 64, 36, 8)
    b = scenes.synthetic_scene("mixed1024", 64, 36, 8)
    for n in ("spheres", "boxes", "surfaces", "toruses"):
        assert a.array(n).tobytes() == b.array(n).tobytes()
    q = a.array("boxes")["quat_rotation"].astype(np.float64)
    assert np.allclose((q * q).sum(axis=1), 1.0, atol=1e-6)


def test_full_size_configs():
    assert scenes.CONFIGS["mixed1024_4k"][1:] == (3840, 2160, 8)
    assert scenes.CONFIGS["mixed1024_8k"][1:] == (7680, 4320, 8)
    assert scenes.CONFIGS["default256"][1:] == (256, 256, 1)
    assert scenes.CONFIGS["default1080"][1:] == (1920, 1080, 4)


def test_default_scene_matches_main_cpp():
    sc = scenes.default_scene(1280, 720)
    d = sc.get_defines()
    assert [int(d[k]) for k in ("sphere_size", "plane_size", "surface_size", "box_size", "torus_size", "ring_size",
                                "light_point_size", "light_direct_size", "iterations")] == [6, 0, 2, 2, 1, 1, 1, 1, 5]
    sp = sc.array("spheres")
    assert tuple(sp["textureNum"]) == (0, 0, 0, 1, 2, 3) and tuple(sp["hollow"]) == (0, 1, 1, 0, 0, 0)
    assert np.allclose(sp["obj"][3], (20000, 0, 0, 5000))                 # jupiter at t = 0 (main.cpp:199-203)
    assert float(sc.rings[0]["r1"]) == pytest.approx((4150 * 1.1166) ** 2, rel=1e-6)
    assert sc.uses_textures() and not scenes.default_scene(64, 64, textured=False).uses_textures()
    assert scenes.default_scene(63, 31).scene["canvas_width"] == 64       # odd sizes are bumped (main.cpp:40-41)


def _fields(arr):
    out = []
    for n in arr.dtype.names:
        out += _fields(arr[n]) if arr[n].dtype.names else [np.asarray(arr[n], dtype=np.float64).ravel()]
    return out


def test_default_scene_equals_the_dump_of_the_unchanged_main_cpp():
    """tests/golden/default_scene_t0.npz = the uniform buffers the reference's OWN main.cpp + SceneManager.cpp uploaded for
    their first frame when run, unchanged, through host/GLWrapper (rt_headless on the B200 box, RT_DUMP_DIR)."""
    from rtb200 import scene as S
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "default_scene_t0.npz"))
    sc = scenes.default_scene(int(z["width"]), int(z["height"]))
    for name, dt in (("spheres", S.rt_sphere), ("surfaces", S.rt_surface), ("boxes", S.rt_box), ("toruses", S.rt_torus), ("rings", S.rt_ring),
                     ("lights_point", S.rt_light_point), ("lights_direct", S.rt_light_direct)):
        want = np.frombuffer(z[name].tobytes(), dtype=dt)
        a, b = np.concatenate(_fields(sc.array(name))), np.concatenate(_fields(want))
        assert a.shape == b.shape, name
        assert np.allclose(a, b, rtol=3e-7, atol=1e-30), name
    scene = np.frombuffer(z["scene"].tobytes(), dtype=S.rt_scene)[0]
    assert np.allclose(scene["quat_camera_rotation"], sc.scene["quat_camera_rotation"]) and np.allclose(scene["camera_pos"], sc.scene["camera_pos"])


def test_cpp_scene_generator_produces_the_same_bytes(tmp_path):
    """the workload generator lives in Python (scenes.py) and in C++ (host/scene_gen.cpp); the two must agree byte
    for byte on every uniform block, at full size, for every synthetic BASELINE config."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    src = os.path.join(ROOT, "raytracing-opengl_b200", "host", "scene_gen.cpp")
    exe = tmp_path / "rt_scene_gen"
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-ffp-contract=off", "-DRTB_SCENE_GEN_MAIN", "-o", str(exe), src])
    for config in ("spheres4k", "tori1080", "mixed1024_4k", "mixed1024_8k"):
        d = tmp_path / config
        d.mkdir()
        subprocess.check_call([str(exe), config, "0", "0", "0", str(d)])
        sc = scenes.build_config(config)
        assert (d / "scene_buf.bin").read_bytes() == np.ascontiguousarray(sc.scene).tobytes(), config
        assert (d / "defines.bin").read_bytes() == sc.get_defines().tobytes(), config
        for blk, attr in (("spheres_buf", "spheres"), ("planes_buf", "planes"), ("surfaces_buf", "surfaces"), ("boxes_buf", "boxes"), ("toruses_buf", "toruses"),
                          ("rings_buf", "rings"), ("lights_point_buf", "lights_point"), ("lights_direct_buf", "lights_direct")):
            assert (d / (blk + ".bin")).read_bytes() == sc.array(attr).tobytes(), (config, blk)
