"""Pins the oracle: the live checker (oracle/_ref = the reference compiled unmodified + harness)
must reproduce every committed golden fixture bit for bit, the harness's injected stream must be
deterministic and independent of threading, and its statistics must match the survey's
measurements of the reference (rays/path, draws/path, NaN share)."""
import numpy as np
import pytest

import common
import raygen


def test_cases_in_sync_with_generator():
    src = open(common.GOLDEN + "/make_golden.py").read()
    for name, c in common.RENDER_CASES.items():
        assert f'"{name}": dict(scene="{c["scene"]}"' in src
        assert f'nx={c["nx"]}, ny={c["ny"]}, ns={c["ns"]}, depth={c["depth"]}, seed={c["seed"]}' in src


@pytest.mark.parametrize("scene", common.HIT_SCENES)
def test_ref_reproduces_golden_hits(O, scene):
    g = common.golden("hits_" + scene)
    img = common.earth_small() if scene in ("earth", "textured_lit") else None
    rs = O.RefScene(scene, image=img)
    assert rs.n_leaves == int(g["n_leaves"][0])
    hits = rs.hit_batch(g["rays"])
    exp = g["hits"]
    for f in ("hit", "prim", "mat"):
        assert np.array_equal(hits[f], exp[f]), f
    ok = exp["hit"] == 1
    for f in ("t", "u", "v", "p", "n"):
        assert common.same_float(hits[f][ok], exp[f][ok]).all(), f


def test_ray_generator_is_deterministic():
    a = raygen.primary_batch("cornell_box", 100, 100, seed=7)
    b = raygen.primary_batch("cornell_box", 100, 100, seed=7)
    assert common.same_float(a, b).all()
    g = common.golden("hits_cornell_box")["rays"]
    full = raygen.primary_batch("cornell_box", 1500, 1500, seed=7)
    assert common.same_float(g[: len(full)], full).all()


@pytest.mark.parametrize("case", ["cornell_A", "cornell_slices", "light_spheres", "textured_lit", "cornell_smoke",
                                  "oneweek_final"])
def test_ref_reproduces_golden_radiance(O, case, monkeypatch):
    c = common.RENDER_CASES[case]
    g = common.golden("render_" + case)
    img = common.earth_small() if c["scene"] in ("earth", "textured_lit") else None
    monkeypatch.chdir(common.GOLDEN)  # oneweek_final() loads ./earthmap.jpg (src/utils.cc:400)
    rs = O.RefScene(c["scene"], image=img)
    # the perlin tables are wall-clock seeded (src/perlin_noise.cc:15-17): put the golden ones back
    tabs = [np.ascontiguousarray(g[k]) for k in ("ranvec", "perm_x", "perm_y", "perm_z")]  # keep alive
    rs.lib.ref_set_perlin(*[t.ctypes.data for t in tabs])
    out, samples, st = rs.render(c["cam"], c["nx"], c["ny"], c["ns"], c["depth"], slices=c.get("slices", 1),
                                 lights=c.get("lights", O.REFERENCE_LIGHTS), seed=c["seed"], per_sample=True)
    assert common.same_float(out, g["sum_rgb"]).all()
    assert common.same_float(samples, g["samples"]).all()
    assert st["rays"] == int(g["rays"][0]) and st["draws"] == int(g["draws"][0])


def test_injected_stream_is_thread_independent(O):
    rs = O.RefScene("cornell_box")
    a, _, _ = rs.render(O.CORNELL_CAM, 40, 40, 6, 15, seed=99, threads=1)
    b, _, _ = rs.render(O.CORNELL_CAM, 40, 40, 6, 15, seed=99, threads=5)
    c, _, _ = rs.render(O.CORNELL_CAM, 40, 40, 6, 15, seed=100, threads=5)
    assert common.same_float(a, b).all()
    assert not np.array_equal(a, c)


def test_window_render_equals_full_render(O):
    rs = O.RefScene("cornell_box")
    full, _, _ = rs.render(O.CORNELL_CAM, 32, 32, 4, 15, seed=1)
    win, _, _ = rs.render(O.CORNELL_CAM, 32, 32, 4, 15, seed=1, window=(8, 4, 20, 30))
    assert common.same_float(win[0, 4:30, 8:20], full[0, 4:30, 8:20]).all()
    assert win[0, :4].sum() == 0


def test_reference_statistics_match_survey(O):
    """SURVEY 6.2 / 8a (variant A = fov 90, depth 15): 2.41 rays/path, 10.2 draws/path."""
    rs = O.RefScene("cornell_box")
    out, samples, st = rs.render(O.CORNELL_CAM, 96, 96, 16, 15, seed=42, per_sample=True)
    assert abs(st["rays"] / st["paths"] - 2.41) < 0.08
    assert abs(st["draws"] / st["paths"] - 10.2) < 0.4
    assert np.isfinite(out).all()


def test_harness_philox_known_answers(O):
    import ctypes as C
    L = O.ref_lib(True)
    for row in common.golden("philox_kat")["kat"]:
        ctr = (C.c_uint32 * 4)(*[int(x) for x in row[0:4]])
        key = (C.c_uint32 * 2)(*[int(x) for x in row[4:6]])
        out = (C.c_uint32 * 4)()
        L.ref_philox4x32_10(ctr, key, out)
        assert [int(x) for x in out] == [int(x) for x in row[6:10]]


def test_plain_reference_build_has_no_injection(O):
    rs = O.RefScene("cornell_box", det=False)
    with pytest.raises(RuntimeError):
        rs.render(O.CORNELL_CAM, 8, 8, 2, 5, deterministic=True)
    out, _, st = rs.render(O.CORNELL_CAM, 16, 16, 4, 15, deterministic=False, threads=2)
    assert st["paths"] == 16 * 16 * 4 and np.isfinite(out).all()
