"""Away from the handful of fixed scenes: random "scene programs" (tiny-path-tracer_b200/host/tpt_scene_programs.h).
One generator header is compiled against the reference's own classes (oracle/ref_harness.cc) and against the
product's front end, so both build the same hitable tree from a seed: spheres, moving spheres, rects, flip_normal,
boxes, translate / rotate_y around primitives and around whole groups, hitable_lists and bvh_nodes nested in each
other (one-element bvh_nodes, bvh_nodes under lists), all surface materials, checker textures.

CPU part (here): the REFERENCE ITSELF (world->hit, color()) against the plain-C restatement running on the flattened
description -- this pins the flattener and the restatement on tree shapes no fixed scene has. GPU part
(test_gpu_scene_programs.py): the CUDA path against the restatement on the same programs."""
import numpy as np
import pytest

import common
import raygen

SEEDS = list(range(1, 81))
# light-sampling list that matches the programs' lamp (the reference's hard-coded one, main.cpp:99-106, points elsewhere and
# leaves these rooms nearly black), plus a sphere shape in mid-room; the programs use the book's 0..555 room, the reference's own
# cornell_box() is centred on the origin and seen from +z (common.CORNELL_CAM)
PROGRAM_LIGHTS = [(0, (213.0, 343.0, 227.0, 332.0, 554.0)), (1, (278.0, 278.0, 278.0, 60.0, 0.0))]
PROGRAM_CAM = dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vup=(0, 1, 0), vfov=40.0, aperture=0.1, focus_dist=10.0)  # the book's view into the 0..555 room


@pytest.fixture(scope="module")
def P(T):
    import oracle_port
    if not oracle_port.available():
        pytest.skip("oracle/_build/libtptoracle.so not built (run __graft_entry__.build())")
    return oracle_port


def program_rays(seed, n=2500):
    """camera / interior / adversarial rays of the Cornell room plus rays between random points of the room (the objects
    sit in [60, 495]^3), shutter times in [0, 1]"""
    rng = np.random.default_rng(2000 + seed)
    through = np.zeros((4 * n, 7), np.float32)
    through[:, 0:3] = rng.uniform(-80, 635, size=(4 * n, 3))
    through[:, 3:6] = rng.uniform(60, 495, size=(4 * n, 3)) - through[:, 0:3]
    through[:, 6] = rng.uniform(0, 1, size=4 * n)
    view = raygen.camera_rays(n, PROGRAM_CAM["lookfrom"], PROGRAM_CAM["lookat"], 40.0, rng)
    return np.concatenate([raygen.primary_batch("cornell_box", n, n, seed=1000 + seed), through, view])


def test_programs_are_diverse(T):
    """the family really covers the node kinds and wrappers it claims to"""
    kinds, chains, dups, depth, prim_kinds = set(), 0, 0, 0, set()
    for seed in SEEDS:
        hs = T.HostScene(f"program:{seed}")
        d = hs.desc.contents
        open_ends = []
        for i in range(d.n_nodes):
            while open_ends and open_ends[-1] == i:
                open_ends.pop()
            k = d.nodes[i].kind
            kinds.add(k & 0xff)
            dups += bool(k & 0x100)
            if (k & 0xff) != 2:
                open_ends.append(d.nodes[i].end_or_prim)
                depth = max(depth, len(open_ends))
        chains = max(chains, d.n_chains)
        prim_kinds |= {d.prims[i].kind for i in range(d.n_prims)}
    assert kinds == {0, 1, 2} and dups > 0 and chains > 3 and depth >= 4 and prim_kinds >= {0, 1, 2, 3, 4}


@pytest.mark.parametrize("seed", SEEDS)
def test_reference_and_restatement_agree_on_hit_records(T, O, P, seed):
    hs = T.HostScene(f"program:{seed}")
    ref = O.RefScene(f"program:{seed}")
    rays = program_rays(seed)
    for gen in range(2):
        exp = ref.hit_batch(rays)
        got = P.hit_batch(T, hs, rays)
        assert gen > 0 or (exp["hit"] == 1).sum() > 500
        for f in ("hit", "prim", "mat"):
            assert np.array_equal(got[f], exp[f]), (seed, gen, f, int((got[f] != exp[f]).sum()))
        ok = exp["hit"] == 1
        for f in ("t", "u", "v", "p", "n"):
            assert common.same_float(got[f][ok], exp[f][ok]).all(), (seed, gen, f)
        rays = raygen.secondary_rays(exp, np.random.default_rng(seed))
        if len(rays) == 0:
            break


@pytest.mark.parametrize("seed", SEEDS[::2])
def test_reference_and_restatement_agree_on_radiance_per_sample(T, O, P, seed):
    nx, ny, ns, depth = 20, 20, 3, 12
    ref, rsamples, rst = O.RefScene(f"program:{seed}").render(PROGRAM_CAM, nx, ny, ns, depth, seed=900 + seed, per_sample=True, lights=PROGRAM_LIGHTS)
    hs = T.HostScene(f"program:{seed}", lights=PROGRAM_LIGHTS)
    out, samples, st = P.render(T, hs, common.product_camera(T, PROGRAM_CAM, nx, ny),
                                T.make_params(nx, ny, ns, depth, seed=900 + seed), threads=4, per_sample=True)
    assert common.same_float(samples, rsamples).all(), int((~common.same_float(samples, rsamples)).sum())
    assert common.same_float(out, ref).all()
    assert st["rays"] == rst["rays"] and st["draws"] == rst["draws"]
    assert (ref.sum(axis=-1) > 0).mean() > 0.05, "the frame should not be black"


@pytest.mark.parametrize("seed", list(range(1, 41)))
def test_reference_and_restatement_agree_with_participating_media(T, O, P, seed):
    """"programm:<seed>": the same programs plus one or two constant_medium objects (sphere / box / moved, rotated box
    as boundary; src/hitable.cc:92-128 draws from the stream INSIDE hit): per-sample radiance bit for bit, equal ray
    and draw counts."""
    nx, ny, ns, depth = 20, 20, 3, 12
    name = f"programm:{seed}"
    ref, rsamples, rst = O.RefScene(name).render(PROGRAM_CAM, nx, ny, ns, depth, seed=900 + seed, per_sample=True, lights=PROGRAM_LIGHTS)
    hs = T.HostScene(name, lights=PROGRAM_LIGHTS)
    out, samples, st = P.render(T, hs, common.product_camera(T, PROGRAM_CAM, nx, ny),
                                T.make_params(nx, ny, ns, depth, seed=900 + seed), threads=4, per_sample=True)
    assert common.same_float(samples, rsamples).all(), int((~common.same_float(samples, rsamples)).sum())
    assert st["rays"] == rst["rays"] and st["draws"] == rst["draws"]


@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_reference_and_restatement_agree_on_large_programs(T, O, P, seed):
    """"programL:<seed>" / "programLm:<seed>": 150-850 primitives, the size class of random_scene and oneweek_final
    (SAH BVH, skip-pointer and replay walks on the CUDA side): hit records, and per-sample radiance with media."""
    hs = T.HostScene(f"programL:{seed}")
    assert hs.desc.contents.n_prims > 100
    rays = program_rays(seed, 1500)
    exp = O.RefScene(f"programL:{seed}").hit_batch(rays)
    got = P.hit_batch(T, hs, rays)
    for f in ("hit", "prim", "mat"):
        assert np.array_equal(got[f], exp[f]), (seed, f)
    ok = exp["hit"] == 1
    for f in ("t", "u", "v", "p", "n"):
        assert common.same_float(got[f][ok], exp[f][ok]).all(), (seed, f)
    if seed % 3 == 0:
        nx, ny, ns, depth = 16, 16, 2, 10
        name = f"programLm:{seed}"
        ref, rsamples, rst = O.RefScene(name).render(PROGRAM_CAM, nx, ny, ns, depth, seed=900 + seed, per_sample=True, lights=PROGRAM_LIGHTS)
        out, samples, st = P.render(T, T.HostScene(name, lights=PROGRAM_LIGHTS), common.product_camera(T, PROGRAM_CAM, nx, ny),
                                    T.make_params(nx, ny, ns, depth, seed=900 + seed), threads=4, per_sample=True)
        assert common.same_float(samples, rsamples).all()
        assert st["rays"] == rst["rays"] and st["draws"] == rst["draws"]


def test_every_program_passes_the_library_validator_and_the_folding(T):
    """host-only parts of tpt_scene_create on all four families: validate_desc accepts the description (the call then ends
    with TPT_ERR_NO_DEVICE here, or succeeds on a GPU box), and the small-scene folding (tpt_debug_small_scene) runs and
    reports a layout that is consistent with the primitive count."""
    import ctypes as C
    lib = T.lib()
    lib.tpt_debug_small_scene.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(C.c_int32)]
    lib.tpt_debug_small_scene.restype = C.c_int
    small = large = 0
    for fam, seeds in (("program", SEEDS), ("programm", range(1, 41)), ("programL", range(1, 13)), ("programLm", range(3, 13, 3))):
        for seed in seeds:
            hs = T.HostScene(f"{fam}:{seed}", lights=PROGRAM_LIGHTS)
            out = C.c_void_p()
            rc = lib.tpt_scene_create(hs.desc, 0, C.byref(out))
            assert rc in (0, -3), (fam, seed, rc, lib.tpt_last_error())
            if rc == 0:
                lib.tpt_scene_destroy(out)
            info = (C.c_int32 * 64)()
            assert lib.tpt_debug_small_scene(hs.desc, info) == 0, (fam, seed, lib.tpt_last_error())
            if info[0]:  # folded into the constant-bank layout: at most 48 primitives by construction
                small += 1
                assert 0 < info[1] <= 8 and hs.desc.contents.n_prims <= 600
            else:
                large += 1
    assert small > 20 and large > 10, (small, large)


@pytest.mark.parametrize("seed", list(range(1, 21)))
def test_reference_and_restatement_agree_with_perlin_textures(T, O, P, seed):
    """"programp:<seed>": every third texture is the Perlin marble (src/texture.cc:18-25, src/utils.cc:160-225) on moved and
    rotated objects at world coordinates of several hundred. The static noise tables of the reference-side build are read
    back and forced on the other side (every perlin_noise ctor re-randomises them from the wall clock)."""
    name = f"programp:{seed}"
    rs = O.RefScene(name)
    rv, px, py, pz = rs.perlin_tables()
    hs = T.HostScene(name, perlin=common.perlin_struct(T, dict(ranvec=rv, perm_x=px, perm_y=py, perm_z=pz)), lights=PROGRAM_LIGHTS)
    nx, ny, ns, depth = 16, 16, 2, 10
    ref, rsamples, rst = rs.render(PROGRAM_CAM, nx, ny, ns, depth, seed=900 + seed, per_sample=True, lights=PROGRAM_LIGHTS)
    out, samples, st = P.render(T, hs, common.product_camera(T, PROGRAM_CAM, nx, ny), T.make_params(nx, ny, ns, depth, seed=900 + seed),
                                threads=4, per_sample=True)
    assert common.same_float(samples, rsamples).all(), int((~common.same_float(samples, rsamples)).sum())
    assert st["rays"] == rst["rays"] and st["draws"] == rst["draws"]
