"""texture::value on the device (tpt_debug_texture) against the reference's texture classes:
perlin turbulence marble (src/texture.cc:18-25, src/utils.cc:160-225) with the golden tables,
checker (src/texture.cc:4-16), nearest-texel image lookup (src/texture.cc:27-42); Philox KAT."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def test_philox_known_answers(T, gpu):
    for row in common.golden("philox_kat")["kat"]:
        assert T.philox([int(x) for x in row[0:4]], [int(x) for x in row[4:6]]) == [int(x) for x in row[6:10]]


def tex_index(desc, kind):
    d = desc.contents
    return [i for i in range(d.n_textures) if d.textures[i].kind == kind]


def test_perlin_marble(T, gpu):
    g = common.golden("textures")
    hs = common.host_scene(T, "two_perlin_spheres", perlin=common.perlin_struct(T, g))
    sc = T.Scene(hs)
    (ti,) = tex_index(hs.desc, 2)
    uvp = np.zeros((len(g["pts"]), 5), np.float32)
    uvp[:, 2:5] = g["pts"]
    got = sc.texture_value(ti, uvp, mode=T.MODE_PARITY)
    # sinf(scale*z + 10*turb): glibc sinf vs correctly-rounded sin differ by <= 1 ulp
    assert np.abs(got - g["turb_scale2"]).max() <= 2e-7
    assert (got == g["turb_scale2"]).mean() > 0.97
    fast = sc.texture_value(ti, uvp, mode=T.MODE_FAST)
    small = np.abs(g["pts"]).max(axis=1) < 50
    assert np.abs(fast - g["turb_scale2"])[small].max() < 2e-3


def test_checker(T, gpu):
    g = common.golden("textures")
    hs = common.host_scene(T, "random_scene")
    sc = T.Scene(hs)
    (ti,) = tex_index(hs.desc, 1)
    uvp = np.zeros((len(g["pts"]), 5), np.float32)
    uvp[:, 2:5] = g["pts"]
    got = sc.texture_value(ti, uvp, mode=T.MODE_PARITY)
    # a sign flip needs sin(10x) within an ulp of zero: none of the fixture points is that close
    assert np.array_equal(got, g["checker"])


def test_image_lookup(T, gpu):
    g = common.golden("textures")
    hs = common.host_scene(T, "earth")
    sc = T.Scene(hs)
    (ti,) = tex_index(hs.desc, 3)
    uvp = np.zeros((len(g["uv"]), 5), np.float32)
    uvp[:, 0:2] = g["uv"]
    for mode in (T.MODE_PARITY, T.MODE_FAST):
        got = sc.texture_value(ti, uvp, mode=mode)
        assert np.array_equal(got, g["image"]) if mode == T.MODE_PARITY else np.abs(got - g["image"]).max() < 1e-6
