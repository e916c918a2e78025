"""Pins the plain-C restatement (oracle/tpt_oracle.c): bit for bit against the reference itself
(oracle/_ref) and against the committed golden fixtures -- hit records for every scene and ray
class, per-SAMPLE radiance under the injected Philox stream for every render case, texture
values, ray and draw counts. The restatement runs on the FLATTENED scene produced by the host
front end, so these tests also pin the flattener's semantics (tie rules, transform chains)
without a GPU."""
import ctypes as C

import numpy as np
import pytest

import common


@pytest.fixture(scope="module")
def P(T):
    import oracle_port
    if not oracle_port.available():
        pytest.skip("oracle/_build/libtptoracle.so not built (run __graft_entry__.build())")
    return oracle_port


def test_philox_known_answers(P):
    for row in common.golden("philox_kat")["kat"]:
        ctr = (C.c_uint32 * 4)(*[int(x) for x in row[0:4]])
        key = (C.c_uint32 * 2)(*[int(x) for x in row[4:6]])
        out = (C.c_uint32 * 4)()
        P.lib().tpto_philox4x32_10(ctr, key, out)
        assert [int(x) for x in out] == [int(x) for x in row[6:10]]


@pytest.mark.parametrize("scene", common.HIT_SCENES)
def test_port_hits_equal_golden(T, P, scene):
    g = common.golden("hits_" + scene)
    hs = common.host_scene(T, scene)
    got = P.hit_batch(T, hs, g["rays"])
    exp = g["hits"]
    for f in ("hit", "prim", "mat"):
        assert np.array_equal(got[f], exp[f]), f
    ok = exp["hit"] == 1
    for f in ("t", "u", "v", "p", "n"):
        assert common.same_float(got[f][ok], exp[f][ok]).all(), f


@pytest.mark.parametrize("tmin,tmax", [(0.001, 1.0), (0.5, 2.5), (0.0, 0.75)])
def test_port_hits_equal_reference_on_windows(T, O, P, tmin, tmax):
    import raygen
    rays = raygen.primary_batch("cornell_box", 3000, 3000, seed=19)
    exp = O.RefScene("cornell_box").hit_batch(rays, tmin, tmax)
    got = P.hit_batch(T, common.host_scene(T, "cornell_box"), rays, tmin, tmax)
    assert np.array_equal(got["hit"], exp["hit"]) and np.array_equal(got["prim"], exp["prim"])
    ok = exp["hit"] == 1
    assert common.same_float(got["t"][ok], exp["t"][ok]).all()


@pytest.mark.parametrize("case", list(common.RENDER_CASES))
def test_port_radiance_equals_golden_per_sample(T, P, case):
    c = common.RENDER_CASES[case]
    g = common.golden("render_" + case)
    hs = common.host_scene(T, c["scene"], perlin=common.perlin_struct(T, g), lights=c.get("lights"))
    cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
    p = T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], slices=c.get("slices", 1), seed=c["seed"])
    out, samples, st = P.render(T, hs, cam, p, threads=4, per_sample=True)
    assert common.same_float(samples, g["samples"]).all()
    assert common.same_float(out, g["sum_rgb"]).all()
    assert st["rays"] == int(g["rays"][0]) and st["draws"] == int(g["draws"][0])


def test_port_radiance_equals_live_reference(T, O, P):
    """fresh seeds / sizes against the live injected reference, variant B (depth 50)."""
    cam = dict(common.CORNELL_CAM, vfov=61.93)
    ref, rsamples, rst = O.RefScene("cornell_box").render(cam, 40, 28, 6, 50, seed=4242, per_sample=True)
    hs = common.host_scene(T, "cornell_box")
    out, samples, st = P.render(T, hs, common.product_camera(T, cam, 40, 28), T.make_params(40, 28, 6, 50, seed=4242),
                                threads=3, per_sample=True)
    assert common.same_float(samples, rsamples).all() and common.same_float(out, ref).all()
    assert st["rays"] == rst["rays"] and st["draws"] == rst["draws"]


def test_port_textures_equal_golden(T, P):
    g = common.golden("textures")
    hs = common.host_scene(T, "two_perlin_spheres", perlin=common.perlin_struct(T, g))
    uvp = np.zeros((len(g["pts"]), 5), np.float32)
    uvp[:, 2:5] = g["pts"]
    assert common.same_float(P.texture_value(T, hs, 0, uvp), g["turb_scale2"]).all()
    hs = common.host_scene(T, "random_scene")
    d = hs.desc.contents
    chk = [i for i in range(d.n_textures) if d.textures[i].kind == 1][0]
    assert common.same_float(P.texture_value(T, hs, chk, uvp), g["checker"]).all()
    hs = common.host_scene(T, "earth")
    uv = np.zeros((len(g["uv"]), 5), np.float32)
    uv[:, 0:2] = g["uv"]
    assert common.same_float(P.texture_value(T, hs, 0, uv), g["image"]).all()


def test_port_quantiser(P):
    sums = np.array([0.0, 1.0, 4.0, 8.0, 2000.0, np.nan, 0.25], np.float32)
    out = np.zeros(len(sums), np.uint8)
    P.lib().tpto_quantise(C.c_void_p(sums.ctypes.data), C.c_size_t(len(sums)), C.c_float(8.0), C.c_void_p(out.ctypes.data))
    exp = [0, int(255.99 * np.sqrt(1 / 8)), int(255.99 * np.sqrt(0.5)), 255, 255, 0, int(255.99 * np.sqrt(0.25 / 8))]
    assert list(out) == exp
