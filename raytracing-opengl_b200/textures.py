"""Decoded sampler inputs: the cubemap faces and the five 2-D textures.

The reference decodes files with stb_image inside GLWrapper::load_cubemap /
load_texture (src/GLWrapper.cpp:284-363) and hands raw RGB8/RGBA8 rows to GL.
Python callers decode with PIL (`load_assets`) or, where the reference's asset
files are not present (the GPU box), use `procedural_textures` — deterministic
stand-ins of the same shapes/channel counts, good for parity tests because the
CUDA path and the oracle are given identical bytes.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from .scenes import DEFAULT_CUBEMAP, DEFAULT_TEXTURES


@dataclass
class TextureSet:
    cube: list | None = None                       # 6 x uint8 [h, w, ch], order +X,-X,+Y,-Y,+Z,-Z, row 0 = t 0
    tex2d: dict = field(default_factory=dict)      # unit (1..5) -> uint8 [h, w, ch]


def load_assets(assets_dir: str, cubemap=DEFAULT_CUBEMAP, textures=DEFAULT_TEXTURES) -> TextureSet:
    """Decode the reference's asset files (ASSETS_DIR/textures/...), rows in file order (stb_image default)."""
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None

    def load(rel):
        img = Image.open(os.path.join(assets_dir, "textures", rel))
        if img.mode not in ("RGB", "RGBA", "L"):
            img = img.convert("RGBA" if "A" in img.mode else "RGB")
        a = np.asarray(img, dtype=np.uint8)
        return np.ascontiguousarray(a if a.ndim == 3 else a[:, :, None])

    ts = TextureSet()
    ts.cube = [load(p) for p in cubemap]
    ts.tex2d = {u: load(p) for u, p in textures.items()}
    return ts


def _hash2(x, y, seed):
    m = np.uint64(0xFFFFFFFF)
    h = (x.astype(np.uint64) * np.uint64(374761393) + y.astype(np.uint64) * np.uint64(668265263) + np.uint64(seed * 2246822519 & 0xFFFFFFFF)) & m
    h = ((h ^ (h >> np.uint64(13))) * np.uint64(1274126177)) & m
    return (h ^ (h >> np.uint64(16))).astype(np.uint32)


def procedural_textures(cube_size=64, small=True) -> TextureSet:
    """Deterministic textures: smooth gradients + hashed detail so that filtering and LOD errors show."""
    ts = TextureSet()
    faces = []
    for f in range(6):
        y, x = np.mgrid[0:cube_size, 0:cube_size]
        r = (x * 255 // max(1, cube_size - 1)).astype(np.uint8)
        g = (y * 255 // max(1, cube_size - 1)).astype(np.uint8)
        b = ((_hash2(x // 4, y // 4, f) & 0xFF)).astype(np.uint8)
        faces.append(np.ascontiguousarray(np.stack([r, g, np.maximum(b, 40 * f)], axis=-1).astype(np.uint8)))
    ts.cube = faces
    shapes = {1: (128, 256, 3), 2: (128, 256, 3), 3: (64, 128, 3), 4: (25, 512, 4), 5: (64, 64, 4)} if small else \
             {1: (2048, 4096, 3), 2: (2048, 4096, 3), 3: (1024, 2048, 3), 4: (500, 8192, 4), 5: (512, 512, 4)}
    for unit, (h, w, ch) in shapes.items():
        y, x = np.mgrid[0:h, 0:w]
        chans = []
        for c in range(ch):
            base = ((x * (c + 2) + y * (5 - c)) * 255 // (w + h)).astype(np.uint32)
            noise = _hash2(x // 2, y // 2, unit * 8 + c) & 0x3F
            chans.append(((base + noise) & 0xFF).astype(np.uint8))
        a = np.stack(chans, axis=-1)
        if ch == 4:
            a[..., 3] = np.where((x // max(1, w // 16)) % 2 == 0, 255, (x * 200 // w + 30)).astype(np.uint8)
        ts.tex2d[unit] = np.ascontiguousarray(a)
    return ts


def smaa_tables():
    """(AreaTex [560,160,2] uint8, SearchTex [16,64] uint8) from the asset mirror the host Makefile fills from the reference's
    src/AreaTex.h / src/SearchTex.h (host/build/assets/smaa, git-ignored; the repository ships no copy), or None when absent."""
    import os
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "build", "assets", "smaa")
    a, s = os.path.join(d, "area_rg8_160x560.bin"), os.path.join(d, "search_r8_64x16.bin")
    if not (os.path.isfile(a) and os.path.isfile(s)):
        return None
    return np.fromfile(a, dtype=np.uint8).reshape(560, 160, 2), np.fromfile(s, dtype=np.uint8).reshape(16, 64)
