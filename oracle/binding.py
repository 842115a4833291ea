"""ctypes binding of the CPU oracle (liboracle.so) and of oracle/_ref (libref.so).

TEST INFRASTRUCTURE: import this only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs.  The product package
(raytracing-opengl_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF = os.path.join(HERE, "_ref", "libref.so")
LIB_REF_FAST = os.path.join(HERE, "_ref", "libref_fast.so")          # timing copy (-O3 -march=x86-64-v3), bench.py only
LIB_ORACLE_NATIVE = os.path.join(HERE, "liboracle_native.so")        # timing copy of the restatement, built on the timing box


class _Image(C.Structure):
    _fields_ = [("px", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("ch", C.c_int32)]


class _Defines(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("sphere_size", "plane_size", "surface_size", "box_size", "torus_size", "ring_size",
                                         "light_point_size", "light_direct_size", "iterations")] + \
               [("ambient_color", C.c_float * 3), ("shadow_ambient", C.c_float * 3)]


class _Desc(C.Structure):
    _fields_ = [("defines", _Defines)] + [(n, C.c_void_p) for n in (
        "scene", "spheres", "planes", "surfaces", "boxes", "toruses", "rings", "lights_point", "lights_direct")] + \
        [("cube", _Image * 6), ("tex2d", _Image * 6)]


class Stats(C.Structure):
    _fields_ = [("pixels", C.c_uint64), ("rays_nearest", C.c_uint64), ("rays_shadow", C.c_uint64), ("tests", C.c_uint64 * 7),
                ("dk_iterations", C.c_uint64), ("shaded_hits", C.c_uint64 * 7), ("light_evals", C.c_uint64), ("dk_hist", C.c_uint64 * 61)]

    def as_dict(self):
        return {"pixels": self.pixels, "rays_nearest": self.rays_nearest, "rays_shadow": self.rays_shadow,
                "tests": list(self.tests), "dk_iterations": self.dk_iterations, "shaded_hits": list(self.shaded_hits),
                "light_evals": self.light_evals}


def build(ref: bool = False, quiet: bool = True):
    """Compile liboracle.so (and, where /root/reference exists, _ref/libref.so)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=out)
    if ref and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out)


def have_ref() -> bool:
    return os.path.isfile(LIB_REF)


def timing_flags(impl: str) -> str | None:
    """Compiler flags of the timing copy of `impl` ('ref' / 'oracle'), building the restatement's copy on this box if needed; None = absent."""
    if impl == "ref":
        f = os.path.join(HERE, "_ref", "libref_fast.flags")
        return open(f).read().strip() if os.path.isfile(LIB_REF_FAST) and os.path.isfile(f) else None
    try:
        subprocess.check_call(["make", "-C", HERE, "liboracle_native.so"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return None
    return "g++ -O3 -march=native -ffp-contract=off -fno-fast-math"


def _make_desc(scene, textures, keep):
    """scene: SceneContainer; textures: TextureSet or None.  `keep` collects arrays that must outlive the call."""
    d = _Desc()
    defs = scene.get_defines()
    for n in ("sphere_size", "plane_size", "surface_size", "box_size", "torus_size", "ring_size", "light_point_size",
              "light_direct_size", "iterations"):
        setattr(d.defines, n, int(defs[n]))
    d.defines.ambient_color[:] = [float(x) for x in defs["ambient_color"]]
    d.defines.shadow_ambient[:] = [float(x) for x in defs["shadow_ambient"]]
    sc = np.ascontiguousarray(scene.scene).reshape(1)
    keep.append(sc)
    d.scene = sc.ctypes.data
    for n in ("spheres", "planes", "surfaces", "boxes", "toruses", "rings", "lights_point", "lights_direct"):
        a = scene.array(n)
        keep.append(a)
        setattr(d, n, a.ctypes.data if len(a) else None)
    if textures is not None:
        if textures.cube is not None:
            for f in range(6):
                a = np.ascontiguousarray(textures.cube[f])
                keep.append(a)
                d.cube[f] = _Image(a.ctypes.data, a.shape[1], a.shape[0], a.shape[2])
        for u, a in textures.tex2d.items():
            a = np.ascontiguousarray(a)
            keep.append(a)
            d.tex2d[u] = _Image(a.ctypes.data, a.shape[1], a.shape[0], a.shape[2])
    return d


PRECISIONS = {"f32": "orc", "f64": "orc64", "sr": "orcsr"}


class Oracle:
    """One scene loaded into the restated oracle (impl='oracle') or into the compiled reference shader (impl='ref').

    precision (restatement only, oracle/real_types.h): "f32" = the pinned fp32 restatement, "f64" = the same control flow in
    double, "sr" = fp32 with stochastic rounding (render_ex(sample=k) selects the k-th random-rounding stream)."""

    def __init__(self, scene, textures=None, impl: str = "oracle", precision: str = "f32", timing_build: bool = False):
        """timing_build: load the -O3 copy of the library (bench.py's CPU arm); results are the same bits, only faster."""
        self.impl = impl
        self.precision = precision
        if impl != "oracle" and precision != "f32":
            raise ValueError("oracle/_ref exists in fp32 only")
        self.p = PRECISIONS[precision] if impl == "oracle" else "ref"
        path = LIB_ORACLE if impl == "oracle" else LIB_REF
        if timing_build:
            path = LIB_ORACLE_NATIVE if impl == "oracle" else LIB_REF_FAST
        if not os.path.isfile(path):
            raise FileNotFoundError(f"{path} not built (make -C oracle{' ref' if impl == 'ref' else ''})")
        self.lib = C.CDLL(path)
        L, p = self.lib, self.p
        getattr(L, p + "_create").restype = C.c_void_p
        getattr(L, p + "_create").argtypes = [C.POINTER(_Desc)]
        getattr(L, p + "_destroy").argtypes = [C.c_void_p]
        getattr(L, p + "_calc_inter").restype = C.c_float
        getattr(L, p + "_calc_inter").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        getattr(L, p + "_in_shadow").restype = C.c_float
        getattr(L, p + "_in_shadow").argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        if impl == "oracle":
            getattr(L, p + "_render").argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.POINTER(Stats), C.c_int]
            getattr(L, p + "_render_quads").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Stats), C.c_int]
            getattr(L, p + "_render_ex").argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(Stats), C.c_int]
            getattr(L, p + "_render_quads_ex").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                                           C.POINTER(Stats), C.c_int]
            getattr(L, p + "_set_pairing").argtypes = [C.c_void_p, C.c_int]
            getattr(L, p + "_intersect").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
            getattr(L, p + "_ray_dir").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            L.orc_sample_cube.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
            L.orc_sample_2d.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p]
            L.orc_mip_levels.argtypes = [C.c_void_p, C.c_int]
            L.orc_mip_level.restype = C.POINTER(C.c_uint8)
            L.orc_mip_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        else:
            L.ref_render.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p, C.c_int]
            L.ref_render_quads.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        keep = []
        desc = _make_desc(scene, textures, keep)
        self.h = getattr(L, p + "_create")(C.byref(desc))
        if not self.h:
            raise RuntimeError("oracle create failed")
        self.width = int(scene.scene["canvas_width"])
        self.height = int(scene.scene["canvas_height"])

    def close(self):
        if self.h:
            getattr(self.lib, self.p + "_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_pairing(self, rule: int):
        if self.impl == "oracle":
            getattr(self.lib, self.p + "_set_pairing")(self.h, rule)

    def render(self, x0=0, y0=0, w=None, h=None, threads=0, stats: Stats | None = None) -> np.ndarray:
        """RGBA32F [h, w, 4]; row 0 = bottom scanline (GL window coordinates)."""
        w = self.width if w is None else w
        h = self.height if h is None else h
        out = np.empty((h, w, 4), dtype=np.float32)
        if self.impl == "oracle":
            rc = getattr(self.lib, self.p + "_render")(self.h, x0, y0, w, h, out.ctypes.data, C.byref(stats) if stats is not None else None, threads)
        else:
            rc = self.lib.ref_render(self.h, x0, y0, w, h, out.ctypes.data, threads)
        if rc != 0:
            raise ValueError("render window must be even-aligned and non-empty")
        return out

    def render_quads(self, qx, qy, threads=0, stats: Stats | None = None) -> np.ndarray:
        """[n, 4, 4]: the four pixels (x,y) (x+1,y) (x,y+1) (x+1,y+1) of each quad."""
        qx = np.ascontiguousarray(qx, dtype=np.int32)
        qy = np.ascontiguousarray(qy, dtype=np.int32)
        out = np.empty((len(qx), 4, 4), dtype=np.float32)
        if self.impl == "oracle":
            getattr(self.lib, self.p + "_render_quads")(self.h, len(qx), qx.ctypes.data, qy.ctypes.data, out.ctypes.data,
                                                        C.byref(stats) if stats is not None else None, threads)
        else:
            self.lib.ref_render_quads(self.h, len(qx), qx.ctypes.data, qy.ctypes.data, out.ctypes.data, threads)
        return out

    def render_ex(self, x0=0, y0=0, w=None, h=None, sample=0, threads=0, stats: Stats | None = None):
        """(RGBA32F [h, w, 4], path hash uint64 [h, w], Durand-Kerner trips uint32 [h, w]) — restatement only."""
        w = self.width if w is None else w
        h = self.height if h is None else h
        out = np.empty((h, w, 4), dtype=np.float32)
        path = np.empty((h, w), dtype=np.uint64)
        dk = np.empty((h, w), dtype=np.uint32)
        rc = getattr(self.lib, self.p + "_render_ex")(self.h, x0, y0, w, h, out.ctypes.data, path.ctypes.data, dk.ctypes.data, sample,
                                                     C.byref(stats) if stats is not None else None, threads)
        if rc != 0:
            raise ValueError("render window must be even-aligned and non-empty")
        return out, path, dk

    def render_quads_ex(self, qx, qy, sample=0, threads=0, stats: Stats | None = None):
        """([n, 4, 4] colours, [n, 4] path hashes, [n, 4] Durand-Kerner trips) — restatement only."""
        qx = np.ascontiguousarray(qx, dtype=np.int32)
        qy = np.ascontiguousarray(qy, dtype=np.int32)
        out = np.empty((len(qx), 4, 4), dtype=np.float32)
        path = np.empty((len(qx), 4), dtype=np.uint64)
        dk = np.empty((len(qx), 4), dtype=np.uint32)
        getattr(self.lib, self.p + "_render_quads_ex")(self.h, len(qx), qx.ctypes.data, qy.ctypes.data, out.ctypes.data, path.ctypes.data,
                                                       dk.ctypes.data, sample, C.byref(stats) if stats is not None else None, threads)
        return out, path, dk

    def calc_inter(self, ro, rd, num=0, type_=0):
        ro = np.ascontiguousarray(ro, dtype=np.float32)
        rd = np.ascontiguousarray(rd, dtype=np.float32)
        n, t = C.c_int32(num), C.c_int32(type_)
        tm = getattr(self.lib, self.p + "_calc_inter")(self.h, ro.ctypes.data, rd.ctypes.data, C.byref(n), C.byref(t))
        return float(tm), n.value, t.value

    def in_shadow(self, ro, rd, dist):
        ro = np.ascontiguousarray(ro, dtype=np.float32)
        rd = np.ascontiguousarray(rd, dtype=np.float32)
        return float(getattr(self.lib, self.p + "_in_shadow")(self.h, ro.ctypes.data, rd.ctypes.data, dist))

    # ---- restatement-only probes ----
    def intersect(self, type_, index, ro, rd, tmin=1e6):
        ro = np.ascontiguousarray(ro, dtype=np.float32)
        rd = np.ascontiguousarray(rd, dtype=np.float32)
        t, k = C.c_float(0), C.c_int32(0)
        hit = getattr(self.lib, self.p + "_intersect")(self.h, type_, index, ro.ctypes.data, rd.ctypes.data, tmin, C.byref(t), C.byref(k))
        return bool(hit), float(t.value), int(k.value)

    def ray_dir(self, x, y):
        out = np.empty(3, dtype=np.float32)
        getattr(self.lib, self.p + "_ray_dir")(self.h, x, y, out.ctypes.data)
        return out

    def sample_cube(self, d):
        d = np.ascontiguousarray(d, dtype=np.float32)
        out = np.empty(4, dtype=np.float32)
        self.lib.orc_sample_cube(self.h, d.ctypes.data, out.ctypes.data)
        return out

    def sample_2d(self, unit, u, v, lod):
        out = np.empty(4, dtype=np.float32)
        self.lib.orc_sample_2d(self.h, unit, u, v, lod, out.ctypes.data)
        return out

    def mip_chain(self, unit):
        levels = []
        for l in range(self.lib.orc_mip_levels(self.h, unit)):
            w, h = C.c_int32(), C.c_int32()
            p = self.lib.orc_mip_level(self.h, unit, l, C.byref(w), C.byref(h))
            levels.append(np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy())
        return levels
