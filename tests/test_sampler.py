"""The GL sampler model (oracle/gl_sampler.h): cube face selection, bilinear, mip chain, LOD clamping."""
import numpy as np
import pytest

from oracle.binding import Oracle
from rtb200 import scenes
from rtb200.textures import TextureSet


def _scene():
    return scenes._base(8, 8, 1)


def solid_cube(size=4):
    faces = []
    for f in range(6):
        a = np.zeros((size, size, 3), np.uint8)
        a[...] = (40 * f + 10, 255 - 40 * f, 7 * f)
        faces.append(a)
    return faces


@pytest.mark.parametrize("d,face", [((1, .2, .3), 0), ((-1, .2, .3), 1), ((.2, 1, .3), 2), ((.2, -1, .3), 3), ((.2, .3, 1), 4), ((.2, .3, -1), 5)])
def test_cube_face_selection(d, face):
    ts = TextureSet(cube=solid_cube())
    c = Oracle(_scene(), ts).sample_cube(d)
    assert np.allclose(c[:3] * 255, (40 * face + 10, 255 - 40 * face, 7 * face), atol=1e-4) and c[3] == 1.0


def test_cube_orientation_and_clamp():
    faces = solid_cube(2)
    faces[4] = np.array([[[0, 0, 0], [255, 0, 0]], [[0, 255, 0], [255, 255, 0]]], np.uint8)      # +Z: row 0 is t = 0
    o = Oracle(_scene(), TextureSet(cube=faces))
    # +Z: s = (x/z+1)/2, t = (-y/z+1)/2: looking up (+y) reads row 0 (file top row), right (+x) reads column 1
    assert np.allclose(o.sample_cube((-0.9, 0.9, 1))[:3], (0, 0, 0), atol=1e-6)
    assert np.allclose(o.sample_cube((0.9, 0.9, 1))[:3], (1, 0, 0), atol=1e-6)
    assert np.allclose(o.sample_cube((-0.9, -0.9, 1))[:3], (0, 1, 0), atol=1e-6)
    assert np.allclose(o.sample_cube((0, 0, 1))[:3], (0.5, 0.5, 0), atol=1e-6)         # centre: equal bilinear weights


def test_mip_chain_shape_of_the_saturn_ring_texture():
    a = (np.arange(500 * 8192 * 4) % 251).astype(np.uint8).reshape(500, 8192, 4)
    chain = Oracle(_scene(), TextureSet(tex2d={4: a})).mip_chain(4)
    dims = [(l.shape[1], l.shape[0]) for l in chain]
    assert dims[:4] == [(8192, 500), (4096, 250), (2048, 125), (1024, 62)] and dims[-1] == (1, 1) and len(chain) == 14
    lvl1 = (a[0::2, 0::2].astype(int) + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2
    assert np.array_equal(chain[1], lvl1.astype(np.uint8))


def test_texture_lod_repeat_and_clamping():
    a = np.zeros((4, 4, 3), np.uint8)
    a[:, :2] = 200
    o = Oracle(_scene(), TextureSet(tex2d={1: a}))
    texel_centre = o.sample_2d(1, 0.125, 0.125, 0.0)
    assert np.allclose(texel_centre, (200 / 255,) * 3 + (1,))
    assert np.allclose(o.sample_2d(1, 1.125, -0.875, 0.0), texel_centre)               # GL_REPEAT
    assert np.allclose(o.sample_2d(1, 0.125, 0.125, -5.0), texel_centre)               # lod < 0 -> magnification = level 0
    assert np.allclose(o.sample_2d(1, 0.125, 0.125, float("-inf")), texel_centre)      # log2(0) from fwidth == 0
    top = o.sample_2d(1, 0.3, 0.3, 99.0)                                               # clamped to the 1x1 level
    assert np.allclose(top[:3], 100 / 255, atol=1e-6)
    mid = o.sample_2d(1, 0.125, 0.125, 0.5)                                            # trilinear between level 0 and 1
    l1 = o.sample_2d(1, 0.125, 0.125, 1.0)
    assert np.allclose(mid, 0.5 * texel_centre + 0.5 * l1, atol=1e-6)


def test_unbound_samplers_return_black_with_alpha_one():
    o = Oracle(_scene(), None)
    assert tuple(o.sample_2d(2, 0.3, 0.3, 0)) == (0, 0, 0, 1)
