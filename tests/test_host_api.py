"""Host-side mirror of GLWrapper and the C-ABI library: loads, exports every declared symbol, fails loudly."""
import ctypes
import os
import re

import numpy as np
import pytest

import rtb200
from rtb200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_symbol_the_header_declares():
    lib = api.load_library()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"librtb200.so does not export {n}"


def test_every_entry_point_cites_the_reference_interface_it_replaces():
    text = open(os.path.join(ROOT, "include", "rtb200.h")).read()
    for method in ("init_window", "init_shaders", "init_buffer", "update_buffer", "load_cubemap", "load_texture", "draw()"):
        assert method in text
    assert len(re.findall(r"GLWrapper\.(h|cpp):\d+", text)) >= 8


def test_version_string():
    assert b"sm_100a" in api.load_library().rtb_version()


def test_python_glwrapper_has_the_reference_method_names():
    for m in ("getWidth", "getHeight", "getProgramId", "init_window", "init_shaders", "set_skybox", "stop", "enable_SMAA", "draw",
              "load_cubemap", "load_texture", "init_buffer", "update_buffer"):
        assert callable(getattr(rtb200.GLWrapper, m)), m       # src/GLWrapper.h:17-38


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    gl = rtb200.GLWrapper(64, 64)
    with pytest.raises(rtb200.RtbError, match="no CUDA device|CPU fallback"):
        gl.init_window()
    out = ctypes.c_double()
    assert api.load_library().rtb_measure_fp32_peak(0, ctypes.byref(out)) != 0


def test_unknown_ubo_block_name_is_an_error():
    gl = rtb200.GLWrapper(64, 64)
    with pytest.raises(rtb200.RtbError, match="Invalid ubo block name"):       # GLWrapper.cpp:371-375
        gl.init_buffer("bogus_buf", 3, None)


def test_null_context_calls_return_errors_not_crashes():
    L = api.load_library()
    assert L.rtb_render(None) < 0 and L.rtb_sync(None) < 0 and L.rtb_upload(None, 1, None, 0) < 0
    assert L.rtb_local_rows(None) == 0
    L.rtb_destroy(None)


def test_gather_rows_inverts_the_row_block_partition():
    h, w, world, br = 100, 8, 3, 16
    full = np.arange(h * w * 4, dtype=np.float32).reshape(h, w, 4)
    parts = []
    for r in range(world):
        rows = [y for y in range(h) if (y // br) % world == r]
        parts.append(full[rows])
    assert np.array_equal(api.gather_rows(parts, h, world, br), full)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "raytracing-opengl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dp, f), errors="ignore").read()
                code = re.sub(r"/\*.*?\*/|//[^\n]*|#[^\n]*|\"\"\".*?\"\"\"", "", text, flags=re.S)
                assert "liboracle" not in code and "rt_oracle" not in code and "import oracle" not in code and "from oracle" not in code, f
