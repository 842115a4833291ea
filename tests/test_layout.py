"""Layout contract: numpy dtypes == include/rtb200_types.h == (when present) the reference's src/scene.h."""
import os
import subprocess
import tempfile

import pytest

from rtb200 import scene as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = {"material": S.rt_material, "sphere": S.rt_sphere, "plane": S.rt_plane, "box": S.rt_box, "torus": S.rt_torus,
           "ring": S.rt_ring, "surface": S.rt_surface, "light_direct": S.rt_light_direct, "light_point": S.rt_light_point,
           "scene": S.rt_scene}
# numpy field -> C member where the names differ
ALIASES = {("surface", "v_min"): "xMin", ("surface", "v_max"): "xMax"}
EXPECTED_SIZE = {"material": 64, "sphere": 112, "plane": 96, "box": 112, "torus": 112, "ring": 112, "surface": 160,
                 "light_direct": 32, "light_point": 48, "scene": 64}


def _probe(prefix, include_flags, header, use_alias):
    lines = ["#include <cstdio>", "#include <cstddef>", f'#include "{header}"', "int main(){"]
    for name, dt in STRUCTS.items():
        lines.append(f'printf("{name} size %zu\\n", sizeof({prefix}{name}));')
        for f in dt.names:
            member = ALIASES.get((name, f), f) if use_alias else f
            lines.append(f'printf("{name} {f} %zu\\n", offsetof({prefix}{name}, {member}));')
    lines.append("return 0;}")
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "probe.cpp")
        open(src, "w").write("\n".join(lines))
        exe = os.path.join(td, "probe")
        subprocess.check_call(["g++", "-std=c++11", "-w", "-Wno-invalid-offsetof"] + include_flags + [src, "-o", exe])
        out = subprocess.check_output([exe]).decode()
    res = {}
    for ln in out.splitlines():
        a, b, c = ln.split()
        res[(a, b)] = int(c)
    return res


def _check(res):
    for name, dt in STRUCTS.items():
        assert res[(name, "size")] == dt.itemsize == EXPECTED_SIZE[name], name
        for f in dt.names:
            assert res[(name, f)] == dt.fields[f][1], (name, f)


def test_numpy_dtypes_match_the_c_abi_header():
    _check(_probe("rtb_", ["-I", os.path.join(ROOT, "include")], "rtb200_types.h", use_alias=False))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference sources only exist in the build container")
def test_numpy_dtypes_match_the_reference_scene_h():
    _check(_probe("rt_", ["-I", "/root/reference/src", "-I", "/root/reference/external_sources/glm"], "scene.h", use_alias=True))


def test_defines_struct_is_60_bytes_in_reference_field_order():
    assert S.rt_defines.itemsize == 60
    assert S.rt_defines.names[:9] == ("sphere_size", "plane_size", "surface_size", "box_size", "torus_size", "ring_size",
                                      "light_point_size", "light_direct_size", "iterations")


def test_quaternion_memory_order_is_xyzw():
    s = S.SceneManager.create_sphere((0, 0, 0), 1, S.SceneManager.create_material((1, 1, 1), 0, 0))
    assert tuple(s["quat_rotation"]) == (0, 0, 0, 1)       # glm::quat(1,0,0,0) in memory; rt.frag:320 tests vec4(0,0,0,1)
