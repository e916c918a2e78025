"""GPU part of the random scene programs (tests/test_scene_programs.py pins the restatement to the reference on them):
the CUDA path through the C-ABI against the restatement on tree shapes no fixed scene has.
  * parity mode: hit records bit for bit; radiance under the injected stream within 1e-4 on every pixel, both kernels
  * fast mode: same closest object on the well-posed rays (budget written below), t within 1e-4 relative"""
import numpy as np
import pytest

import common
import raygen
from test_scene_programs import PROGRAM_CAM, PROGRAM_LIGHTS, program_rays

pytestmark = pytest.mark.gpu
SEEDS = list(range(1, 41))
REL_TOL = 1e-4
FAST_ID_BUDGET = 2e-4  # share of rays whose closest object may differ in FAST mode (ties at shared edges, grazing hits); OBSERVED on B200: 0 of 700 000


@pytest.fixture(scope="module")
def P(T):
    import oracle_port
    if not oracle_port.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    return oracle_port


def make_scene(T, hs):
    try:
        return T.Scene(hs)
    except T.TptError as e:  # a program nested deeper than the library's frame limit: refused at upload, by design
        if e.code == -4:
            pytest.skip(f"refused as unsupported: {e}")
        raise


@pytest.mark.parametrize("seed", SEEDS)
def test_parity_hit_records_equal_restatement(T, P, gpu, seed):
    hs = T.HostScene(f"program:{seed}")
    sc = make_scene(T, hs)
    rays = program_rays(seed)
    for gen in range(2):
        got = sc.intersect(rays, mode=T.MODE_PARITY)
        exp = P.hit_batch(T, hs, rays)
        for f in ("hit", "prim", "mat"):
            assert np.array_equal(got[f], exp[f]), (seed, gen, f, int((got[f] != exp[f]).sum()))
        ok = exp["hit"] == 1
        for f in ("t", "p", "n"):
            assert common.same_float(got[f][ok], exp[f][ok]).all(), (seed, gen, f)
        rays = raygen.secondary_rays(exp, np.random.default_rng(seed))
        if len(rays) == 0:
            break


@pytest.mark.parametrize("seed", SEEDS[::2])
def test_parity_radiance_equals_restatement(T, P, gpu, seed):
    nx, ny, ns, depth = 24, 24, 4, 12
    hs = T.HostScene(f"program:{seed}", lights=PROGRAM_LIGHTS, background=T.BG_SKY if seed % 4 == 1 else T.BG_BLACK)
    sc = make_scene(T, hs)
    cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        p = T.make_params(nx, ny, ns, depth, mode=T.MODE_PARITY, seed=900 + seed, kernel=kernel)
        ref, _, _ = P.render(T, hs, cam, p, threads=4)
        res = sc.render(cam, p)
        rel = common.rel_err(res.sum_rgb, ref, 1e-3 * ns)
        bad = int((rel > REL_TOL).any(axis=-1).sum())
        assert bad == 0, f"seed {seed} kernel {kernel}: {bad} pixels beyond {REL_TOL}, worst {float(rel.max())}"
        assert (ref.sum(axis=-1) > 0).mean() > 0.3, "the frame should not be black"


def test_fast_mode_finds_the_same_objects(T, P, gpu):
    """FAST mode's own structures (constant-bank groups with folded boxes / SAH BVH) against the restatement."""
    n = bad = 0
    worst_t = 0.0
    for seed in SEEDS:
        hs = T.HostScene(f"program:{seed}")
        try:
            sc = T.Scene(hs)
        except T.TptError:
            continue
        rays = program_rays(seed)
        n_adv = len(raygen.adversarial_rays(*raygen.SCENE_INFO["cornell_box"][:2]))
        keep = np.isfinite(rays).all(axis=1)
        keep[5000:5000 + n_adv] = False  # the hand-made tie / in-plane block (FAST mode does not promise those)
        got = sc.intersect(rays, mode=T.MODE_FAST)
        exp = P.hit_batch(T, hs, rays)
        both = (got["hit"] == 1) & (exp["hit"] == 1)
        differ = ((got["hit"] != exp["hit"]) | (both & (got["prim"] != exp["prim"]))) & keep
        same = both & (got["prim"] == exp["prim"]) & keep & np.isfinite(exp["t"])
        if same.any():
            worst_t = max(worst_t, float((np.abs(got["t"][same].astype(np.float64) - exp["t"][same]) / np.maximum(np.abs(exp["t"][same]), 1e-3)).max()))
        n += int(keep.sum())
        bad += int(differ.sum())
    print(f"\nfast mode on {len(SEEDS)} scene programs: {bad} of {n} rays with another closest object, worst relative t difference {worst_t:.3g}")
    assert n > 300000
    assert bad <= FAST_ID_BUDGET * n, (bad, n)
    assert worst_t < 1e-3


@pytest.mark.parametrize("seed", list(range(1, 21)))
def test_parity_radiance_with_participating_media(T, P, gpu, seed):
    """"programm:<seed>": programs with constant_medium objects (the stream is consumed inside world->hit)."""
    nx, ny, ns, depth = 24, 24, 4, 12
    hs = T.HostScene(f"programm:{seed}", lights=PROGRAM_LIGHTS, background=T.BG_SKY if seed % 4 == 1 else T.BG_BLACK)
    sc = make_scene(T, hs)
    cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        p = T.make_params(nx, ny, ns, depth, mode=T.MODE_PARITY, seed=900 + seed, kernel=kernel)
        ref, _, _ = P.render(T, hs, cam, p, threads=4)
        res = sc.render(cam, p)
        rel = common.rel_err(res.sum_rgb, ref, 1e-3 * ns)
        bad = int((rel > REL_TOL).any(axis=-1).sum())
        assert bad == 0, f"seed {seed} kernel {kernel}: {bad} pixels beyond {REL_TOL}, worst {float(rel.max())}"


@pytest.mark.parametrize("seed", list(range(1, 13)))
def test_large_programs(T, P, gpu, seed):
    """"programL:<seed>" (150-850 primitives: SAH BVH in fast mode, skip-pointer / replay walks in parity mode) and, every
    third seed, "programLm:<seed>" with participating media (the "trace" wavefront variant in fast mode)."""
    hs = T.HostScene(f"programL:{seed}")
    sc = make_scene(T, hs)
    rays = program_rays(seed, 1500)
    exp = P.hit_batch(T, hs, rays)
    got = sc.intersect(rays, mode=T.MODE_PARITY)
    for f in ("hit", "prim", "mat"):
        assert np.array_equal(got[f], exp[f]), (seed, f, int((got[f] != exp[f]).sum()))
    ok = exp["hit"] == 1
    for f in ("t", "p", "n"):
        assert common.same_float(got[f][ok], exp[f][ok]).all(), (seed, f)
    # fast mode: same closest object on the well-posed rays
    n_adv = len(raygen.adversarial_rays(*raygen.SCENE_INFO["cornell_box"][:2]))
    keep = np.isfinite(rays).all(axis=1)
    keep[3000:3000 + n_adv] = False
    fast = sc.intersect(rays, mode=T.MODE_FAST)
    both = (fast["hit"] == 1) & (exp["hit"] == 1)
    differ = ((fast["hit"] != exp["hit"]) | (both & (fast["prim"] != exp["prim"]))) & keep
    assert differ.sum() <= 3, f"seed {seed}: {int(differ.sum())} of {int(keep.sum())} rays on another object in fast mode"
    nx, ny, ns, depth = 24, 24, 4, 10
    cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
    for name in ([f"programL:{seed}"] + ([f"programLm:{seed}"] if seed % 3 == 0 else [])):
        hs2 = T.HostScene(name, lights=PROGRAM_LIGHTS)
        sc2 = make_scene(T, hs2)
        for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
            p = T.make_params(nx, ny, ns, depth, mode=T.MODE_PARITY, seed=900 + seed, kernel=kernel)
            ref, _, _ = P.render(T, hs2, cam, p, threads=4)
            res = sc2.render(cam, p)
            rel = common.rel_err(res.sum_rgb, ref, 1e-3 * ns)
            bad = int((rel > REL_TOL).any(axis=-1).sum())
            assert bad == 0, f"{name} kernel {kernel}: {bad} pixels beyond {REL_TOL}, worst {float(rel.max())}"
            # fast mode runs the same frame without an error and with a comparable amount of light
            pf = T.make_params(nx, ny, 16, depth, mode=T.MODE_FAST, seed=900 + seed, kernel=kernel)
            fres = sc2.render(cam, pf)
            assert np.isfinite(fres.sum_rgb).all()


@pytest.mark.parametrize("name", ["program:2", "program:3", "program:5", "program:21", "programm:2", "programL:2"])
def test_fast_mode_is_unbiased_on_programs(T, gpu, name):
    """Converged frames (48 x 48, 2048 spp) in FAST and in PARITY mode: the image means agree far inside 1 % (two PARITY
    frames with different seeds differ by 0.05-0.3 %). This is the check that found FAST mode 4.4 % dark on rooms whose x = 0
    wall carried a checker_texture -- not a defect of its sampling but the checker's sin(10 x) evaluated AT its zero, where
    the reference's answer is the sign of a rounding residue; the generator now keeps checkers off
    that plane."""
    nx = ny = 48
    ns, depth = 2048, 12
    cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
    sc = make_scene(T, T.HostScene(name, lights=PROGRAM_LIGHTS))

    def mean(mode, seed):
        r = sc.render(cam, T.make_params(nx, ny, ns, depth, mode=mode, seed=seed, kernel=T.KERNEL_WAVEFRONT))
        return float(np.minimum(np.nan_to_num(r.sum_rgb[0] / ns), 10.0).mean())  # fireflies clamped for the comparison
    a, a2, b = mean(T.MODE_PARITY, 11), mean(T.MODE_PARITY, 12), mean(T.MODE_FAST, 11)
    print(f"\n{name}: parity {a:.5f} / {a2:.5f} (other seed), fast {b:.5f}: fast - parity = {(b - a) / a:+.2e} of the mean")
    assert a > 0.02  # a lit frame (OBSERVED on B200: fast - parity between -1.3e-4 and +2.1e-4 of the mean on these six)
    assert abs(b - a) <= 0.01 * a, (a, b)


def test_parity_radiance_with_perlin_textures(T, P, gpu):
    """"programp:<seed>": Perlin marble on moved / rotated objects. The marble is 0.5 (1 + sin(...)): where sin -> -1 one ulp of
    sinf (glibc's against the double evaluation rounded once) is a large RELATIVE error of an albedo near 0, so a few dark-vein
    pixels may exceed 1e-4 (test_gpu_radiance.py documents the same for two_perlin_spheres); budget 0.5 % of the pixels."""
    nx, ny, ns, depth = 24, 24, 4, 12
    cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
    perlin = common.perlin_struct(T, common.golden("textures"))
    n = bad = 0
    for seed in range(1, 13):
        hs = T.HostScene(f"programp:{seed}", perlin=perlin, lights=PROGRAM_LIGHTS)
        sc = make_scene(T, hs)
        for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
            p = T.make_params(nx, ny, ns, depth, mode=T.MODE_PARITY, seed=900 + seed, kernel=kernel)
            ref, _, _ = P.render(T, hs, cam, p, threads=4)
            res = sc.render(cam, p)
            rel = common.rel_err(res.sum_rgb, ref, 1e-3 * ns)
            n += nx * ny
            bad += int((rel > REL_TOL).any(axis=-1).sum())
    print(f"\nPerlin programs: {bad} of {n} pixels beyond {REL_TOL}")
    assert bad <= 0.005 * n, (bad, n)
