#!/usr/bin/env python3
"""Build oracle/_ref/libref.so: the reference's OWN fragment shader, executed on the CPU.

TEST INFRASTRUCTURE.  The reference's GL render cannot run in this image (no
libGL/EGL/OSMesa), but its shader is ~900 lines of C-like GLSL.
This script reads /root/reference/assets/shaders/rt.frag WHERE IT LIES, applies
the purely mechanical rewrites listed below so that the text is valid C++, and
compiles it (ref_harness.cpp + glsl_prelude.h, built-ins delegated to the glm
that the reference vendors).  Outputs go ONLY to oracle/_ref/ (git-ignored):
the generated include never enters the repository's history.

Rewrites (none changes what a statement computes):
  R1  drop `#version`; strip // comments
  R2  float literals get an `f` suffix (GLSL literals are fp32, C++'s are double)
  R3  the {TOKEN} specialisation constants (GLWrapper.cpp:237-247) become reads of
      run-time values U.<name>, so one build serves every scene
  R4  each std140 uniform block (rt.frag:155-230) becomes a pointer / value
      member filled by the harness from the same bytes the UBO would hold
  R5  `uniform` / `out vec4 FragColor` storage qualifiers dropped;
      in/out/inout parameters become values / references
  R6  swizzles .xyz .xy .yzx .zxy .zy .zx .rgb become accessor calls
  R7  hit_record(...) constructor calls get a matching C++ constructor
  R8  pins for undefined behaviour, identical to rt_oracle.cpp's header:
      uninitialised `num,type` / `color` zero-initialised (Q1), the unbounded
      `i--` loop capped at 64 refractive events (Q4)
  R9  `void main()` -> `void main_()`
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

TOKENS = {
    "{SPHERE_SIZE}": "(U.sphere_size)", "{PLANE_SIZE}": "(U.plane_size)", "{SURFACE_SIZE}": "(U.surface_size)",
    "{BOX_SIZE}": "(U.box_size)", "{TORUS_SIZE}": "(U.torus_size)", "{RING_SIZE}": "(U.ring_size)",
    "{LIGHT_DIRECT_SIZE}": "(U.light_direct_size)", "{LIGHT_POINT_SIZE}": "(U.light_point_size)",
    "{AMBIENT_COLOR}": "(U.ambient_color)", "{SHADOW_AMBIENT}": "(U.shadow_ambient)", "{ITERATIONS}": "(U.iterations)",
}

BLOCKS = {  # uniform block -> member declaration (R4)
    "scene_buf": "rt_scene scene;",
    "spheres_buf": "const rt_sphere* spheres = nullptr;",
    "planes_buf": "const rt_plane* planes = nullptr;",
    "surfaces_buf": "const rt_surface* surfaces = nullptr;",
    "boxes_buf": "const rt_box* boxes = nullptr;",
    "toruses_buf": "const rt_torus* toruses = nullptr;",
    "rings_buf": "const rt_ring* rings = nullptr;",
    "lights_point_buf": "const rt_light_point* lights_point = nullptr;",
    "lights_direct_buf": "const rt_light_direct* lights_direct = nullptr;",
}

FLOAT_LIT = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def transform(src: str) -> str:
    src = src.replace("\r\n", "\n")
    src = re.sub(r"^#version.*$", "", src, flags=re.M)                                   # R1
    src = re.sub(r"//[^\n]*", "", src)
    src = FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)                                 # R2
    for k, v in TOKENS.items():                                                          # R3
        assert k in src, k
        src = src.replace(k, v)

    def block(m):                                                                        # R4
        name = m.group(1)
        assert name in BLOCKS, name
        return BLOCKS[name]
    src, n = re.subn(r"layout\(\s*std140\s*\)\s*uniform\s+(\w+)\s*\{.*?\n\};", block, src, flags=re.S)
    assert n == len(BLOCKS), n
    src = src.replace("out vec4 FragColor;", "vec4 FragColor;")                          # R5
    src = re.sub(r"^\s*uniform\s+", "", src, flags=re.M)
    src = re.sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", src)
    src = re.sub(r"\bin\s+(vec\d|float|int)\s+(\w+)", r"\1 \2", src)
    src = re.sub(r"\.(xyz|xy|yzx|zxy|zy|zx|rgb)\b", r".\1()", src)                       # R6
    src, n = re.subn(r"struct hit_record \{",                                            # R7
                     "struct hit_record { hit_record() : mat(), normal(0), bias_mult(0), alpha(0) {} "
                     "hit_record(rt_material m_, vec3 n_, float b_, float a_) : mat(m_), normal(n_), bias_mult(b_), alpha(a_) {}",
                     src)
    assert n == 1
    for old, new in (("int num, type;", "int num = 0, type = 0;"),                       # R8
                     ("int num;", "int num = 0;"),
                     ("vec4 color;", "vec4 color = vec4(0);"),
                     ("i--;", "i--; if (++glass_events__ >= 64) break;")):
        assert src.count(old) == 1, (old, src.count(old))
        src = src.replace(old, new)
    assert src.count("void main()") == 1
    src = src.replace("void main()", "void main_()")                                     # R9
    # FLT_MIN / FLT_MAX are re-#defined by the shader (rt.frag:3-4)
    src = "#undef FLT_MIN\n#undef FLT_MAX\n" + src
    return src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--cxx", default=os.environ.get("CXX", "g++"))
    args = ap.parse_args()
    frag = os.path.join(args.reference, "assets", "shaders", "rt.frag")
    glm = os.path.join(args.reference, "external_sources", "glm")
    if not os.path.isfile(frag):
        print(f"build_ref: {frag} not found; oracle/_ref cannot be built here", file=sys.stderr)
        return 2
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    with open(frag, encoding="latin-1") as f:
        gen = transform(f.read())
    inc = os.path.join(out_dir, "rt_frag_gen.inc")
    with open(inc, "w", encoding="latin-1") as f:
        f.write(gen)
    cmd = [args.cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread",
           "-w", "-I", glm, "-I", out_dir, "-o", os.path.join(out_dir, "libref.so"), os.path.join(HERE, "ref_harness.cpp")]
    print(" ".join(cmd))
    subprocess.check_call(cmd)
    # the TIMING copy (bench.py's cpu_baseline / --impl reference): same source, same IEEE semantics (no contraction, no fast-math),
    # optimised the way BASELINE.md section 3 states.  -march=x86-64-v3 (AVX2) stands in for -march=native: the library is built in
    # the container where /root/reference exists and must run on the GPU box's host CPU, whatever that is.  Golden vectors and parity
    # tests keep using libref.so above, so their bits never depend on the optimiser.
    fast = [a for a in cmd]
    fast[fast.index("-O2")] = "-O3"
    fast.insert(2, "-march=x86-64-v3")
    fast[fast.index(os.path.join(out_dir, "libref.so"))] = os.path.join(out_dir, "libref_fast.so")
    print(" ".join(fast))
    subprocess.check_call(fast)
    with open(os.path.join(out_dir, "libref_fast.flags"), "w") as f:
        f.write("g++ -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
