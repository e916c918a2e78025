"""Gate 1 (north_star): ray-primitive and BVH hit records on identical ray batches.
CUDA path (tpt_intersect_batch through the C-ABI) vs the reference's hit functions:
  * hit-or-miss and closest-object ids: BIT-EXACT (parity mode)
  * t / p / normal within 1e-5 relative (observed: bit-exact in parity mode)
Golden fixtures were produced by the reference itself (tests/golden/make_golden.py); when
oracle/_ref travelled with the snapshot the same comparison is repeated live on fresh rays."""
import numpy as np
import pytest

import common
import raygen

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5  # north_star gate 1


def compare_hits(got, exp, exact_ids=True):
    stats = {}
    stats["hit_mismatch"] = int((got["hit"] != exp["hit"]).sum())
    both = (got["hit"] == 1) & (exp["hit"] == 1)
    stats["prim_mismatch"] = int((got["prim"][both] != exp["prim"][both]).sum())
    stats["mat_mismatch"] = int((got["mat"][both] != exp["mat"][both]).sum())
    same = both & (got["prim"] == exp["prim"])
    finite = same & np.isfinite(exp["t"])
    for f in ("t", "p", "n"):
        g, e = got[f][finite].astype(np.float64), exp[f][finite].astype(np.float64)
        if f != "t":  # records with a non-finite component (overflowing adversarial rays): bitwise check only
            keep = np.isfinite(e).all(axis=1) & np.isfinite(g).all(axis=1)
            g, e = g[keep], e[keep]
        scale = np.maximum(np.abs(e), 1e-3) if f != "p" else np.maximum(np.linalg.norm(e, axis=1, keepdims=True), 1e-3)
        stats["rel_" + f] = float((np.abs(g - e) / scale).max()) if len(e) else 0.0
        stats["exact_" + f] = bool(common.same_float(got[f][same], exp[f][same]).all())
    du = np.abs(got["u"][finite] - exp["u"][finite])
    dv = np.abs(got["v"][finite] - exp["v"][finite])
    # asin() of an argument an ulp above 1 is NaN on both sides: NaN == NaN here
    du[np.isnan(got["u"][finite]) & np.isnan(exp["u"][finite])] = 0
    dv[np.isnan(got["v"][finite]) & np.isnan(exp["v"][finite])] = 0
    stats["rel_uv"] = float(max(du.max(initial=0), dv.max(initial=0)))
    return stats


@pytest.mark.parametrize("scene", common.HIT_SCENES)
def test_parity_hits_vs_golden(T, gpu, scene):
    g = common.golden("hits_" + scene)
    sc = T.Scene(common.host_scene(T, scene))
    got = sc.intersect(g["rays"], mode=T.MODE_PARITY)
    st = compare_hits(got, g["hits"])
    assert st["hit_mismatch"] == 0 and st["prim_mismatch"] == 0 and st["mat_mismatch"] == 0, st
    assert st["rel_t"] <= REL_TOL and st["rel_p"] <= REL_TOL and st["rel_n"] <= REL_TOL, st
    assert st["exact_t"] and st["exact_p"] and st["exact_n"], st  # stronger than the gate asks
    assert st["rel_uv"] <= 2e-6, st  # atan2f/asinf: glibc vs correctly-rounded, <= 1 ulp of [0,1]


N_PRIMARY = {"cornell_box": 3000, "sphere_cornell_box": 1200, "random_scene": 3000, "random_scene_list": 800,
             "two_perlin_spheres": 600, "light_spheres": 600, "earth": 600, "textured_lit": 800}
# OBSERVED on B200 (profiles/r02_fast_mismatch.json, tools/gpu_mismatch.py): fast mode picks the reference's
# closest object on EVERY well-posed ray of every scene -- 0 hit/miss and 0 object-id mismatches. The budget is
# therefore 0, not a percentage. "Well-posed" excludes (a) the hand-made adversarial block (rays lying IN a
# rectangle's plane, zero direction components that turn 0 * inf into NaN, exact corner hits: ties the reference
# resolves by evaluation order) and (b) rays with a non-finite component (secondary rays spawned from a NaN hit
# record of that block).
FAST_ID_BUDGET = {scene: 0 for scene in common.HIT_SCENES}


def well_posed(scene, rays):
    n_adv = len(raygen.adversarial_rays(*raygen.SCENE_INFO[scene][:2]))
    ok = np.isfinite(rays).all(axis=1)
    ok[N_PRIMARY[scene]:N_PRIMARY[scene] + n_adv] = False
    return ok


@pytest.mark.parametrize("scene", common.HIT_SCENES)
def test_fast_hits_vs_golden(T, gpu, scene):
    """fast mode: fp32 + FMA + its own acceleration structures (uniform brute force / SAH BVH, slab-folded
    boxes). Gate 1 asks for bit-exact hit/miss and closest-object ids: measured and asserted with budget 0 on
    every well-posed ray; records within the stated tolerances of the reference's."""
    g = common.golden("hits_" + scene)
    rays, exp = g["rays"], g["hits"]
    ok = well_posed(scene, rays)
    sc = T.Scene(common.host_scene(T, scene))
    assert len(sc.intersect(rays, mode=T.MODE_FAST)) == len(rays)  # degenerate rays: a record for every ray, no crash
    got = sc.intersect(rays[ok], mode=T.MODE_FAST)
    exp = exp[ok]
    st = compare_hits(got, exp)
    print(f"\nfast-mode gate 1 [{scene}]: {int(ok.sum())} rays, hit/miss mismatches {st['hit_mismatch']}, "
          f"object-id mismatches {st['prim_mismatch']}, max rel t error {st['rel_t']:.3g}")
    assert st["hit_mismatch"] <= FAST_ID_BUDGET[scene], st
    assert st["prim_mismatch"] <= FAST_ID_BUDGET[scene], st
    # hit points: within 1e-3 of the scene extent, except for a handful of ill-conditioned rays
    # (grazing hits on the 1000-radius ground sphere: the fp32 discriminant is a catastrophic cancellation
    # there and differs with and without FMA -- the reference's own answer is one of several equally
    # valid ones)
    same = (got["hit"] == 1) & (exp["hit"] == 1) & (got["prim"] == exp["prim"])
    fin = same & np.isfinite(exp["p"]).all(axis=1) & np.isfinite(got["p"]).all(axis=1)
    dp = np.linalg.norm(got["p"][fin].astype(np.float64) - exp["p"][fin], axis=1)
    extent = raygen.SCENE_INFO[scene][0]
    assert (dp > 1e-3 * extent).mean() < 2e-3, (scene, float(dp.max()))
    assert np.percentile(dp, 99) < 1e-4 * extent


@pytest.mark.parametrize("scene", ["cornell_box", "random_scene", "textured_lit"])
def test_parity_hits_live_reference(T, O, gpu, scene):
    """fresh, larger batches against the live reference (oracle/_ref)."""
    img = common.earth_small() if scene == "textured_lit" else None
    rs = O.RefScene(scene, image=img)
    sc = T.Scene(common.host_scene(T, scene))
    rays = raygen.primary_batch(scene, 40000, 40000, seed=1234)
    exp = rs.hit_batch(rays)
    sec = raygen.secondary_rays(exp, np.random.default_rng(5))
    rays = np.concatenate([rays, sec])
    exp = rs.hit_batch(rays)
    got = sc.intersect(rays, mode=T.MODE_PARITY)
    st = compare_hits(got, exp)
    assert st["hit_mismatch"] == 0 and st["prim_mismatch"] == 0 and st["mat_mismatch"] == 0, st
    assert st["exact_t"] and st["exact_p"] and st["exact_n"], st


def test_tmin_tmax_window(T, O, gpu):
    """arbitrary [t_min, t_max] windows (hitable_list shrinks t_max, bvh does not)."""
    rs = O.RefScene("cornell_box")
    sc = T.Scene(common.host_scene(T, "cornell_box"))
    rays = raygen.primary_batch("cornell_box", 5000, 5000, seed=3)
    for tmin, tmax in [(0.001, 1.0), (0.5, 2.5), (0.0, 0.75), (1.0, common.FLT_MAX)]:
        exp = rs.hit_batch(rays, tmin, tmax)
        got = sc.intersect(rays, tmin, tmax, mode=T.MODE_PARITY)
        st = compare_hits(got, exp)
        assert st["hit_mismatch"] == 0 and st["prim_mismatch"] == 0, (tmin, tmax, st)
        assert st["exact_t"], (tmin, tmax, st)


def test_empty_batch(T, gpu):
    sc = T.Scene(common.host_scene(T, "cornell_box"))
    assert len(sc.intersect(np.zeros((0, 7), np.float32))) == 0


def test_open_block_equals_separate_rects(T, gpu, monkeypatch):
    """fast mode folds the five Cornell walls into one open block (one slab test); with
    TPT_SMALL_OPEN_BLOCKS=0 they stay five rectangle tests. Same closest object on every ray (a
    tie at a shared edge aside), t within an ulp ((k - o) / d against (k - o) * (1 / d))."""
    hs = common.host_scene(T, "cornell_box")
    extent, eye, lookat = raygen.SCENE_INFO["cornell_box"]
    rng = np.random.default_rng(77)  # camera + interior rays; no hand-made ties (fast mode does not promise those)
    rays = np.concatenate([raygen.camera_rays(60000, eye, lookat, 90.0, rng), raygen.interior_rays(60000, extent, rng)])
    folded = T.Scene(hs)
    a = folded.intersect(rays, mode=T.MODE_FAST)
    sec = raygen.secondary_rays(a, np.random.default_rng(9))
    rays = np.concatenate([rays, sec])
    a = folded.intersect(rays, mode=T.MODE_FAST)
    monkeypatch.setenv("TPT_SMALL_OPEN_BLOCKS", "0")
    b = T.Scene(hs).intersect(rays, mode=T.MODE_FAST)
    assert (a["hit"] == 1).sum() > len(rays) // 3
    differ = (a["hit"] != b["hit"]) | ((a["hit"] == 1) & (a["prim"] != b["prim"]))
    assert differ.sum() <= max(2, len(rays) // 20000), int(differ.sum())
    same = ~differ & (a["hit"] == 1)
    assert np.allclose(a["t"][same], b["t"][same], rtol=4e-7, atol=0)


@pytest.mark.parametrize("scene", ["random_scene", "random_scene_list"])
def test_parity_walk_cull_changes_nothing(T, gpu, monkeypatch, scene):
    """The parity walk of large trees (closest_hit_skip) does not enter a box that lies wholly behind the best
    hit so far, plus a margin; the reference enters it (bvh_node::hit never shrinks t_max, src/hitable.cc:63-90)
    and finds nothing closer there. With TPT_PARITY_SKIP_CULL=0 the walk tests every box against the caller's
    t_max like the reference: the two must return the same record on every ray, bit for bit -- camera rays,
    rays from inside the scene, the hand-made adversarial ones and two generations of secondary rays."""
    hs = common.host_scene(T, scene)
    rays = raygen.primary_batch(scene, 150000, 50000, seed=2024)
    rays[::3, 6] = np.random.default_rng(5).uniform(-1.0, 2.0, size=len(rays[::3]))  # shutter times inside and OUTSIDE the moving spheres' [0, 1]: outside it they leave their boxes
    culled = T.Scene(hs)
    monkeypatch.setenv("TPT_PARITY_SKIP_CULL", "0")
    plain = T.Scene(hs)
    for gen in range(3):
        a = culled.intersect(rays, mode=T.MODE_PARITY)
        b = plain.intersect(rays, mode=T.MODE_PARITY)
        assert (a["hit"] == 1).sum() > len(rays) // 10
        assert a.tobytes() == b.tobytes(), f"generation {gen}: {int((a != b).sum())} records differ"
        rays = raygen.secondary_rays(a, np.random.default_rng(100 + gen))
        if len(rays) == 0:
            break


@pytest.mark.parametrize("scene", ["cornell_box", "sphere_cornell_box", "random_scene", "light_spheres"])
def test_parity_hits_on_perturbed_scenes_vs_port(T, gpu, scene):
    """Differential check away from the fixed fixtures: the flattened scene gets a few geometry / transform floats
    scaled, shifted or snapped onto another record's value (coincident planes and centres: the tie cases), node
    boxes left as they were -- both sides consume the same description, consistent or not. Parity-mode records
    must equal the plain-C restatement's bit for bit on camera, interior, adversarial and secondary rays.
    (tools/gpu_diff_fuzz.py is the long-running form of this test: profiles/r02_fuzz.txt.)"""
    import ctypes as C

    import oracle_port as P
    import test_abi_fuzz as F
    if not P.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    hs = common.host_scene(T, scene)
    src = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
    rng = np.random.default_rng(31 + len(scene))
    checked = 0
    for trial in range(8):
        d, keep = F._clone(T, src)
        muts = F.perturb_geometry(rng, keep, int(rng.integers(1, 6)))

        class Holder:
            desc = C.pointer(d)
        try:
            sc = T.Scene(Holder.desc)
        except Exception:
            continue  # a perturbation the validator refuses
        rays = raygen.primary_batch(scene, 3000, 1500, seed=500 + trial)
        for gen in range(2):
            got = sc.intersect(rays, mode=T.MODE_PARITY)
            exp = P.hit_batch(T, Holder, rays)
            both = (got["hit"] == 1) & (exp["hit"] == 1)
            differ = (got["hit"] != exp["hit"]) | (both & ((got["prim"] != exp["prim"]) | ~common.same_float(got["t"], exp["t"])
                                                           | ~common.same_float(got["p"], exp["p"]).all(axis=1)
                                                           | ~common.same_float(got["n"], exp["n"]).all(axis=1)))
            assert differ.sum() == 0, f"trial {trial} generation {gen}: {int(differ.sum())} of {len(rays)} records differ after {muts!r}"
            checked += len(rays)
            rays = raygen.secondary_rays(got, np.random.default_rng(trial))
            if len(rays) == 0:
                break
        sc.close()
    assert checked > 20000


def test_parity_honours_boxes_that_do_not_contain_their_primitive(T, gpu):
    """The reference tests a primitive whenever the boxes of the bvh_nodes above it are crossed -- wherever the
    primitive actually is. A hand-made description can put a sphere OUTSIDE those boxes (here: a small sphere of
    random_scene moved in front of everything else, between the camera and the scene, its node boxes left where
    they were): rays that cross its old boxes hit it first, the others never test it. Parity mode must follow the
    boxes literally -- its walk's cull ("nothing in a box behind the best hit can be closer") does not hold for such
    a description and has to stay off (found by tools/gpu_diff_fuzz.py, profiles/r02_fuzz.txt)."""
    import ctypes as C

    import oracle_port as P
    import test_abi_fuzz as F
    if not P.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    hs = common.host_scene(T, "random_scene")
    src = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
    extent, eye, lookat = raygen.SCENE_INFO["random_scene"]
    eye = np.array(eye, np.float32)
    rng = np.random.default_rng(3)
    wide = raygen.camera_rays(100000, tuple(eye), lookat, 25.0, rng)
    total_displaced_hits = 0
    for victim in (40, 200, 333):
        d, keep = F._clone(T, src)
        prims = keep["prims"]
        assert prims[victim].kind in (0, 1)  # a sphere / moving sphere
        old_centre = np.array([prims[victim].p[k] for k in range(3)], np.float32)
        new_centre = eye + 0.5 * (old_centre - eye)  # half way to the camera, on the line of sight of its old place
        for k in range(3):
            prims[victim].p[k] = float(new_centre[k])
            if prims[victim].kind == 1:
                prims[victim].p[4 + k] = float(new_centre[k])
        prims[victim].p[3] = 0.3
        # rays from the camera towards the sphere's OLD place: through the displaced sphere first, then through its old boxes
        aimed = np.zeros((50000, 7), np.float32)
        aimed[:, 0:3] = eye
        aimed[:, 3:6] = old_centre + rng.normal(0, 0.12, size=(len(aimed), 3)).astype(np.float32) - eye
        aimed[:, 6] = rng.uniform(0, 1, len(aimed)).astype(np.float32)
        rays = np.concatenate([aimed, wide])

        class Holder:
            desc = C.pointer(d)
        got = T.Scene(Holder.desc).intersect(rays, mode=T.MODE_PARITY)
        exp = P.hit_batch(T, Holder, rays)
        both = (got["hit"] == 1) & (exp["hit"] == 1)
        differ = (got["hit"] != exp["hit"]) | (both & ((got["prim"] != exp["prim"]) | ~common.same_float(got["t"], exp["t"])))
        assert differ.sum() == 0, f"victim {victim}: {int(differ.sum())} of {len(rays)} records differ"
        total_displaced_hits += int(((exp["hit"] == 1) & (exp["prim"] == victim)).sum())
    print("displaced-sphere hits:", total_displaced_hits)
    assert total_displaced_hits > 1000, total_displaced_hits  # the displaced spheres are really hit through their old boxes
