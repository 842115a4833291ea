#!/usr/bin/env python3
"""Build oracle/_ref/libsmaa_ref.so: the reference's OWN SMAA shader code, executed on the CPU.

TEST INFRASTRUCTURE.  Reads /root/reference/assets/shaders/SMAA.h WHERE IT LIES, applies the mechanical rewrites below so that
the text is valid C++ against oracle/smaa_prelude.h, and compiles it with smaa_ref_harness.cpp and the reference's lookup
tables (src/AreaTex.h, src/SearchTex.h, included from where they lie).  Outputs go ONLY to oracle/_ref/ (git-ignored).

Rewrites (none changes what a statement computes):
  S1  comments stripped
  S2  float literals get an `f` suffix (shader literals are fp32, C++'s are double)
  S3  parameter qualifiers: `inout float2 x` -> `Ref2 x`, `inout float4 x` -> `Ref4 x` (references to the components of the lvalue
      passed, which may be a swizzle), `out T x` -> `T& x`, `out float4 x[3]` -> `float4 x[3]`
  S4  the configuration / porting section (SMAA.h up to "// Misc functions") is kept as is: the presets and the derived
      constants come from the reference; only the shading-language macros come from the harness (SMAA_CUSTOM_SL)
A second generated file, smaa_undef.inc, #undefs every macro the body defines, so the body can be included once per preset.
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FLOAT_LIT = re.compile(r"(?<![\w.])(\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def transform(src: str) -> str:
    src = src.replace("\r\n", "\n")
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)      # S1
    src = re.sub(r"//[^\n]*", "", src)
    src = FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)                                          # S2
    src = re.sub(r"\binout\s+float2\s+(\w+)", r"Ref2 \1", src)                                     # S3
    src = re.sub(r"\binout\s+float4\s+(\w+)", r"Ref4 \1", src)
    src = re.sub(r"\bout\s+float4\s+(\w+)\[3\]", r"float4 \1[3]", src)
    src = re.sub(r"\bout\s+(float2|float4)\s+(\w+)", r"\1& \2", src)
    assert "inout" not in src and not re.search(r"\bout\s", src), "unhandled parameter qualifier"
    return src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--cxx", default=os.environ.get("CXX", "g++"))
    args = ap.parse_args()
    smaa = os.path.join(args.reference, "assets", "shaders", "SMAA.h")
    if not os.path.isfile(smaa):
        print(f"build_smaa_ref: {smaa} not found; oracle/_ref/libsmaa_ref.so cannot be built here", file=sys.stderr)
        return 2
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    with open(smaa, encoding="latin-1") as f:
        gen = transform(f.read())
    with open(os.path.join(out_dir, "smaa_gen.inc"), "w", encoding="latin-1") as f:
        f.write(gen)
    macros = sorted(set(re.findall(r"^\s*#\s*define\s+(\w+)", gen, flags=re.M)))
    # the porting layer is the harness's (smaa_ref_harness.cpp, smaa_prelude.h): those names survive between the four inclusions
    keep = {"SMAA_CUSTOM_SL", "SMAATexture2D", "SMAATexturePass2D", "SMAASampleLevelZero", "SMAASampleLevelZeroPoint", "SMAASampleLevelZeroOffset",
            "SMAASample", "SMAASamplePoint", "SMAASampleOffset", "SMAA_FLATTEN", "SMAA_BRANCH", "mad", "lerp", "saturate", "float2", "float3", "float4",
            "int2", "int3", "int4", "bool2", "bool3", "bool4", "SMAATexture2DMS2", "SMAALoad", "SMAAGather"}
    with open(os.path.join(out_dir, "smaa_undef.inc"), "w") as f:
        f.write("".join(f"#undef {m}\n" for m in macros if m not in keep))
    cmd = [args.cxx, "-O2", "-std=gnu++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-w",
           "-I", os.path.join(args.reference, "src"), "-I", out_dir, "-I", HERE,
           "-o", os.path.join(out_dir, "libsmaa_ref.so"), os.path.join(HERE, "smaa_ref_harness.cpp")]
    print(" ".join(cmd))
    subprocess.check_call(cmd)
    return 0


if __name__ == "__main__":
    sys.exit(main())
