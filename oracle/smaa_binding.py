"""ctypes binding of oracle/_ref/libsmaa_ref.so — the reference's own SMAA.h compiled as C++ (oracle/build_smaa_ref.py).

TEST INFRASTRUCTURE: the checker of the CUDA SMAA passes; import only from tests/, tools/ and bench.py's CPU legs."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsmaa_ref.so")
PRESETS = {"LOW": 0, "MEDIUM": 1, "HIGH": 2, "ULTRA": 3}


def have_smaa_ref() -> bool:
    return os.path.isfile(LIB)


def _lib():
    L = C.CDLL(LIB)
    L.smaa_ref_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.smaa_ref_area_tex.restype = C.POINTER(C.c_uint8)
    L.smaa_ref_area_tex.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.smaa_ref_search_tex.restype = C.POINTER(C.c_uint8)
    L.smaa_ref_search_tex.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return L


def smaa_ref(rgba8, preset=3, threads=0):
    """(out [H,W,4], edges [H,W,2], blend [H,W,4]) of the reference's three passes on an RGBA8 image (row 0 = texture row 0)."""
    img = np.ascontiguousarray(rgba8, dtype=np.uint8)
    h, w = img.shape[:2]
    out, edges, blend = np.zeros((h, w, 4), np.uint8), np.zeros((h, w, 2), np.uint8), np.zeros((h, w, 4), np.uint8)
    rc = _lib().smaa_ref_run(img.ctypes.data, w, h, int(preset), edges.ctypes.data, blend.ctypes.data, out.ctypes.data, threads)
    if rc != 0:
        raise ValueError("smaa_ref_run failed")
    return out, edges, blend


def tables():
    """(AreaTex [560,160,2], SearchTex [16,64]) — the reference's lookup tables, read out of the compiled reference library."""
    L = _lib()
    w, h = C.c_int(), C.c_int()
    p = L.smaa_ref_area_tex(C.byref(w), C.byref(h))
    area = np.ctypeslib.as_array(p, shape=(h.value, w.value, 2)).copy()
    p = L.smaa_ref_search_tex(C.byref(w), C.byref(h))
    search = np.ctypeslib.as_array(p, shape=(h.value, w.value)).copy()
    return area, search
