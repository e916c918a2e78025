"""Size-independent properties of the CUDA sample loop, checked at and near BASELINE sizes:
determinism, partition invariance (tiles / ranks), sample-range invariance, linearity in the
emitted radiance, the quantiser, statistics, error behaviour."""
import ctypes as C

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cornell(T, gpu):
    return T.Scene(common.host_scene(T, "cornell_box"))


def test_deterministic_and_seeded(T, cornell):
    cam = T.cornell_camera(200, 200)
    a = cornell.render(cam, T.make_params(200, 200, 32, 15, seed=1))
    b = cornell.render(cam, T.make_params(200, 200, 32, 15, seed=1))
    c = cornell.render(cam, T.make_params(200, 200, 32, 15, seed=2))
    assert np.array_equal(a.sum_rgb, b.sum_rgb) and np.array_equal(a.rgb8, b.rgb8)
    assert not np.array_equal(a.sum_rgb, c.sum_rgb)
    assert a.stats["rays"] == b.stats["rays"]


@pytest.mark.parametrize("mode", ["parity", "fast"])
def test_partition_invariance(T, cornell, mode):
    """Philox is keyed on (pixel, sample): the union of the parts of a 3-way tile split is
    bit-identical to the undivided render (what the multi-GPU gather relies on)."""
    m = T.MODE_PARITY if mode == "parity" else T.MODE_FAST
    nx, ny, ns = 150, 90, 16  # not multiples of the tile size
    cam = T.cornell_camera(nx, ny)
    whole = cornell.render(cam, T.make_params(nx, ny, ns, 15, mode=m, seed=5, subs=1))
    acc = np.zeros_like(whole.sum_rgb)
    owner = np.zeros((ny, nx), np.int32)
    paths = 0
    for part in range(3):
        r = cornell.render(cam, T.make_params(nx, ny, ns, 15, mode=m, seed=5, part_index=part, part_count=3, subs=1))
        owner += (np.abs(r.sum_rgb[0]).sum(axis=-1) > 0)
        acc += r.sum_rgb
        paths += r.stats["paths"]
    assert paths == nx * ny * ns
    assert owner.max() <= 1
    assert np.array_equal(acc, whole.sum_rgb)


@pytest.mark.parametrize("mode", ["parity", "fast"])
def test_wavefront_equals_megakernel(T, cornell, mode):
    """the two scheduling variants run the same device functions on the same Philox stream with
    the same per-pixel summation order: bit-identical sums, identical ray / NaN statistics."""
    m = T.MODE_PARITY if mode == "parity" else T.MODE_FAST
    nx, ny, ns = 200, 136, 24
    for fov, depth in ((90.0, 15), (61.93, 50)):
        cam = T.cornell_camera(nx, ny, fov=fov)
        a = cornell.render(cam, T.make_params(nx, ny, ns, depth, mode=m, seed=21, kernel=T.KERNEL_MEGA, slices=3), want_slices=True)
        b = cornell.render(cam, T.make_params(nx, ny, ns, depth, mode=m, seed=21, kernel=T.KERNEL_WAVEFRONT, slices=3), want_slices=True)
        if mode == "parity":  # no FMA contraction: every operation is the same in both kernels
            assert np.array_equal(a.sum_rgb, b.sum_rgb)
            assert np.array_equal(a.rgb8, b.rgb8) and np.array_equal(a.rgb8_slices, b.rgb8_slices)
            for k in ("paths", "rays", "nan_samples"):
                assert a.stats[k] == b.stats[k], k
        else:  # fast: the compiler contracts different mul/add pairs in the two kernels (ulp-level)
            rel = common.rel_err(b.sum_rgb, a.sum_rgb, 1e-3 * ns)
            assert (rel > 1e-4).any(axis=-1).mean() < 0.02
            assert a.stats["paths"] == b.stats["paths"]
            assert abs(a.stats["rays"] - b.stats["rays"]) <= 2e-3 * a.stats["rays"]


def test_wavefront_partition_and_tail(T, cornell):
    """wavefront variant: 3-way tile split reproduces the whole frame; tiny frames (fewer bins than
    slots) and a frame whose size is not a multiple of the tile work."""
    nx, ny, ns = 150, 90, 8
    cam = T.cornell_camera(nx, ny)
    whole = cornell.render(cam, T.make_params(nx, ny, ns, 15, seed=5, kernel=T.KERNEL_WAVEFRONT, subs=1))
    acc = np.zeros_like(whole.sum_rgb)
    for part in range(3):
        acc += cornell.render(cam, T.make_params(nx, ny, ns, 15, seed=5, kernel=T.KERNEL_WAVEFRONT, part_index=part,
                                                 part_count=3, subs=1)).sum_rgb
    assert np.array_equal(acc, whole.sum_rgb)
    tiny = cornell.render(T.cornell_camera(3, 2), T.make_params(3, 2, 5, 15, seed=5, kernel=T.KERNEL_WAVEFRONT))
    ref = cornell.render(T.cornell_camera(3, 2), T.make_params(3, 2, 5, 15, seed=5, kernel=T.KERNEL_MEGA))
    assert np.array_equal(tiny.sum_rgb, ref.sum_rgb) and tiny.stats["paths"] == 30


def test_wavefront_on_a_large_scene(T, gpu):
    """random_scene (92 KB of tables, 485 leaves) does not fit the shared-memory budget: the
    wavefront variant then reads the scene through L1 and must still agree with the megakernel
    bit for bit in parity mode (sky background so the image is not black)."""
    hs = common.host_scene(T, "random_scene", background=T.BG_SKY)
    sc = T.Scene(hs)
    cam = T.book_camera(96, 64, t0=0.0, t1=1.0)
    a = sc.render(cam, T.make_params(96, 64, 6, 15, mode=T.MODE_PARITY, seed=2, kernel=T.KERNEL_MEGA))
    b = sc.render(cam, T.make_params(96, 64, 6, 15, mode=T.MODE_PARITY, seed=2, kernel=T.KERNEL_WAVEFRONT))
    assert a.sum_rgb.mean() > 0.01 and np.array_equal(a.sum_rgb, b.sum_rgb)
    f = sc.render(cam, T.make_params(96, 64, 64, 15, mode=T.MODE_FAST, seed=2, kernel=T.KERNEL_WAVEFRONT))
    g = sc.render(cam, T.make_params(96, 64, 64, 15, mode=T.MODE_PARITY, seed=3, kernel=T.KERNEL_MEGA))
    assert abs(f.sum_rgb.mean() - g.sum_rgb.mean()) < 0.02 * g.sum_rgb.mean()  # SAH BVH vs reference tree: same image


def test_sub_range_invariance(T, cornell):
    """cutting a pixel's samples into sub-ranges only changes the fp32 summation order."""
    nx = ny = 128
    cam = T.cornell_camera(nx, ny)
    a = cornell.render(cam, T.make_params(nx, ny, 64, 15, seed=9, subs=1))
    b = cornell.render(cam, T.make_params(nx, ny, 64, 15, seed=9, subs=4))
    assert a.stats["rays"] == b.stats["rays"]
    assert common.rel_err(b.sum_rgb, a.sum_rgb, 1e-3).max() < 1e-5


def test_linearity_in_emission(T, gpu):
    """radiance is linear in the lamp's emission; scaling it by 2 is exact in floating point."""
    hs = common.host_scene(T, "cornell_box")
    d = hs.desc.contents
    lamp = [i for i in range(d.n_materials) if d.materials[i].kind == 3]
    assert len(lamp) == 1
    tex = d.materials[lamp[0]].texture
    cam = T.cornell_camera(96, 96)
    p = T.make_params(96, 96, 16, 15, seed=4)
    base = T.Scene(hs).render(cam, p)
    for c in range(3):
        d.textures[tex].color[c] *= 2.0
    twice = T.Scene(hs).render(cam, p)
    assert base.sum_rgb.max() > 0
    assert np.array_equal(twice.sum_rgb, 2.0 * base.sum_rgb)


def test_quantiser_matches_reference_formula(T, cornell):
    """main.cpp:135-139: col/=ns; sqrt; int(255.99f*c); clamp on write (:176-182)."""
    cam = T.cornell_camera(256, 256, fov=61.93)
    r = cornell.render(cam, T.make_params(256, 256, 8, 15, seed=3))
    col = r.sum_rgb[0] / np.float32(8)
    q = np.clip((np.float32(255.99) * np.sqrt(col)).astype(np.int32), 0, 255).astype(np.uint8)
    assert np.array_equal(q, r.rgb8)
    assert (r.rgb8 == 255).any() and (r.rgb8 == 0).any()  # the lamp saturates, shadows are black


def test_statistics_headline_frame(T, cornell):
    """full 1200x1200 frame (the headline resolution) at low spp: every (pixel, sample) is traced
    exactly once; rays/path sits just below the reference's 2.41 (the GPU skips zombie bounces) and
    about 1 % of the samples die as NaN (the reference counts 1.56 % because it also counts
    zero-throughput paths that turn NaN later; both contribute 0 after de_nan)."""
    nx = ny = 1200
    cam = T.cornell_camera(nx, ny)
    st = cornell.render_device(cam, T.make_params(nx, ny, 4, 15, seed=1, bundle_cull=False))
    assert st["paths"] == nx * ny * 4 and st["culled_paths"] == 0
    assert 1.9 < st["rays"] / st["paths"] < 2.41
    assert 0.005 < st["nan_samples"] / st["paths"] < 0.022
    assert st["kernel_launches"] == 2


@pytest.mark.parametrize("mode", ["parity", "fast"])
@pytest.mark.parametrize("kernel", ["mega", "wavefront"])
def test_pixel_bundle_test_is_exact(T, cornell, mode, kernel):
    """Pixels none of whose rays (any lens point, any jitter) can reach the scene's bounds are
    finished without tracing (tpt_stats.culled_paths). That is exact, not an approximation: the
    image is bit-identical to the one with every path traced, each untraced path is exactly one
    world->hit query that would have returned false (rays differ by culled_paths), and at fov 90
    the shipped camera looks past the box in ~64 % of the frame."""
    m = T.MODE_PARITY if mode == "parity" else T.MODE_FAST
    k = T.KERNEL_WAVEFRONT if kernel == "wavefront" else T.KERNEL_MEGA
    nx, ny, ns = 300, 300, 8
    cam = T.cornell_camera(nx, ny)
    a = cornell.render(cam, T.make_params(nx, ny, ns, 15, mode=m, seed=77, kernel=k, slices=2), want_slices=True)
    b = cornell.render(cam, T.make_params(nx, ny, ns, 15, mode=m, seed=77, kernel=k, slices=2, bundle_cull=False), want_slices=True)
    assert np.array_equal(a.sum_rgb, b.sum_rgb) and np.array_equal(a.rgb8, b.rgb8) and np.array_equal(a.rgb8_slices, b.rgb8_slices)
    sa, sb = a.stats, b.stats
    assert sa["paths"] == sb["paths"] == nx * ny * ns and sb["culled_paths"] == 0
    assert 0.55 < sa["culled_paths"] / sa["paths"] < 0.66
    assert sa["rays"] + sa["culled_paths"] == sb["rays"] and sa["nan_samples"] == sb["nan_samples"]
    # every culled pixel is black in the traced render; no pixel that received light was culled
    culled_px = sa["culled_paths"] // ns
    assert (b.sum_rgb[-1].sum(axis=-1) == 0).sum() >= culled_px
    # a frame-filling camera, the sky background and an interior camera cull nothing
    inside = cornell.render(T.cornell_camera(64, 64, fov=61.93), T.make_params(64, 64, 4, 15, mode=m, kernel=k))
    assert inside.stats["culled_paths"] == 0


def test_fast_and_parity_converge_to_the_same_image(T, cornell):
    nx = ny = 64
    cam = T.cornell_camera(nx, ny, fov=61.93)
    a = cornell.render(cam, T.make_params(nx, ny, 512, 15, mode=T.MODE_PARITY, seed=11))
    b = cornell.render(cam, T.make_params(nx, ny, 512, 15, mode=T.MODE_FAST, seed=12))
    ia, ib = a.sum_rgb[0] / 512, b.sum_rgb[0] / 512
    assert abs(ia.mean() - ib.mean()) < 0.01 * ia.mean()


def test_error_behaviour(T, cornell):
    cam = T.cornell_camera(32, 32)
    with pytest.raises(T.TptError) as e:  # main.cpp:113,127 would SIGFPE (count % 0)
        cornell.render(cam, T.make_params(32, 32, 4, 15, slices=8))
    assert e.value.code == -1
    with pytest.raises(T.TptError):
        cornell.render(cam, T.make_params(32, 32, 4, 15, part_index=2, part_count=2))
    with pytest.raises(T.TptError):
        cornell.render(cam, T.make_params(0, 32, 4, 15))
    with pytest.raises(T.TptError) as e:
        cornell.render(cam, T.make_params(32, 32, 4, 15, kernel=7))
    assert e.value.code == -4
    hs = common.host_scene(T, "cornell_box")
    hs.desc.contents.nodes[0].end_or_prim = 10 ** 6
    with pytest.raises(T.TptError):
        T.Scene(hs)
    hs2 = common.host_scene(T, "cornell_box", lights=[])
    with pytest.raises(T.TptError):
        T.Scene(hs2).render(cam, T.make_params(32, 32, 4, 15))


def test_depth_zero_and_one_sample(T, cornell):
    """edge sizes: max_depth 0 (only directly visible emission), 1x1 image, ns=1."""
    cam = T.cornell_camera(64, 64, fov=61.93)
    r = cornell.render(cam, T.make_params(64, 64, 4, 0, seed=1))
    assert r.stats["rays"] == r.stats["paths"]
    lit = r.sum_rgb[0].sum(axis=-1) > 0
    assert 0 < lit.mean() < 0.2  # only the lamp is visible
    one = cornell.render(T.cornell_camera(1, 1), T.make_params(1, 1, 1, 15, seed=1))
    assert one.stats["paths"] == 1


def test_config5_frame_size(T, cornell):
    """BASELINE configs[4]: 4096x4096 (68.7 G paths at 4096 spp). One sample per pixel here: the
    bin decode, accumulator planes and resolve kernel at 16.7 M pixels, and a 3-way partition of it."""
    nx = ny = 4096
    cam = T.cornell_camera(nx, ny)
    st = cornell.render_device(cam, T.make_params(nx, ny, 1, 15, seed=1, kernel=T.KERNEL_WAVEFRONT))
    assert st["paths"] == nx * ny
    paths = sum(cornell.render_device(cam, T.make_params(nx, ny, 1, 15, seed=1, kernel=T.KERNEL_WAVEFRONT, part_index=i,
                                                         part_count=3))["paths"] for i in range(3))
    assert paths == nx * ny
    r = cornell.fetch(T.make_params(nx, ny, 1, 15), want_sum=False)
    assert r.rgb8.shape == (ny, nx, 3)


def test_pixel_bundle_test_random_cameras(T, gpu):
    """the bundle test against brute force over many cameras: positions inside, beside and behind
    the scene, large and zero apertures, tilted up-vectors, non-square frames, list-rooted and
    BVH-rooted scenes, moving spheres. Images and NaN counts must be identical with the test on and
    off, traced rays must differ by exactly the untraced paths."""
    rng = np.random.default_rng(2024)
    scenes = {name: T.Scene(common.host_scene(T, name)) for name in ("cornell_box", "light_spheres", "random_scene")}
    extent = {"cornell_box": 300.0, "light_spheres": 6.0, "random_scene": 12.0}
    culled_some = 0
    for trial in range(60):
        name = ("cornell_box", "light_spheres", "random_scene")[trial % 3]
        sc, e = scenes[name], extent[name]
        lookfrom = rng.uniform(-3.5 * e, 3.5 * e, 3)
        lookat = rng.uniform(-1.5 * e, 1.5 * e, 3) if trial % 5 else lookfrom + rng.uniform(-1, 1, 3)
        vup = np.array([0.0, 1.0, 0.0]) + (rng.uniform(-0.4, 0.4, 3) if trial % 2 else 0.0)
        nx, ny = [(48, 48), (64, 40), (37, 53)][trial % 3]
        fov = float(rng.uniform(10.0, 120.0))
        aperture = float([0.0, 0.1, 0.5 * e, 2.5 * e][trial % 4])
        focus = float(rng.uniform(0.2 * e, 4.0 * e))
        t1 = 1.0 if name == "random_scene" and trial % 2 else 0.0
        cam = T.make_camera(tuple(lookfrom), tuple(lookat), tuple(vup), fov, nx / ny, aperture, focus, 0.0, t1)
        for mode in (T.MODE_PARITY, T.MODE_FAST):
            a = sc.render(cam, T.make_params(nx, ny, 4, 8, mode=mode, seed=trial, kernel=T.KERNEL_WAVEFRONT))
            b = sc.render(cam, T.make_params(nx, ny, 4, 8, mode=mode, seed=trial, kernel=T.KERNEL_WAVEFRONT, bundle_cull=False))
            assert np.array_equal(a.sum_rgb, b.sum_rgb), (trial, name, mode)
            assert a.stats["rays"] + a.stats["culled_paths"] == b.stats["rays"], (trial, name, mode)
            assert a.stats["nan_samples"] == b.stats["nan_samples"]
        culled_some += a.stats["culled_paths"] > 0
    assert culled_some >= 8  # fires in most Cornell trials; the other two scenes sit on a radius-1000 ground sphere


def test_fp32_peak_probe_is_plausible(T, gpu):
    """the roofline denominator bench.py measures: independent FMA chains must land between half of
    and just above the data-sheet product of this device (148 SMs x 128 lanes x 2 x 1.965 GHz)"""
    r = T.fp32_peak(0)
    assert 30.0 < r["tflops"] < 80.0, r
    assert 20.0 < r["ms"] < 120.0, r


def test_camera_redraws_with_sky_background(T, gpu):
    """fov 90 looks past the Cornell box: the fast wavefront kernel ends those camera paths inside its
    generate step (sky radiance added there, next sample drawn in place) while the megakernel sends
    every camera ray through extend. Same estimator sample by sample: equal path counts, the sky
    pixels equal to the last bits, the rest within fast mode's contraction noise; a frame-filling
    camera (nothing to redraw) likewise."""
    sc = T.Scene(common.host_scene(T, "cornell_box", background=T.BG_SKY))
    nx, ny, ns = 160, 120, 16
    for fov in (90.0, 40.0):
        cam = T.cornell_camera(nx, ny, fov=fov)
        a = sc.render(cam, T.make_params(nx, ny, ns, 15, mode=T.MODE_FAST, seed=9, kernel=T.KERNEL_MEGA))
        b = sc.render(cam, T.make_params(nx, ny, ns, 15, mode=T.MODE_FAST, seed=9, kernel=T.KERNEL_WAVEFRONT))
        assert a.stats["paths"] == b.stats["paths"] == nx * ny * ns
        assert abs(a.stats["rays"] - b.stats["rays"]) <= 2e-3 * a.stats["rays"]
        rel = common.rel_err(b.sum_rgb, a.sum_rgb, 1e-3 * ns)
        assert (rel > 1e-4).any(axis=-1).mean() < 0.02
        if fov == 90.0:  # the frame's corners see only sky: every sample of those pixels was a redraw
            corner_a, corner_b = a.sum_rgb[0, :8, :8], b.sum_rgb[0, :8, :8]
            assert corner_a.min() > 0.0 and np.allclose(corner_a, corner_b, rtol=2e-6, atol=0)


def test_fast_mode_sees_a_moving_sphere_outside_its_t0_box(T, gpu):
    """a list-rooted scene whose moving sphere leaves its t = 0 position. FAST mode prunes
    LIST nodes / culls pixels by node boxes, PARITY replays the reference (which never box-tests a list):
    both must see the sphere along its whole path -- same hits on time-stamped rays, same picture."""
    hs = T.HostScene("moving_list_test")
    sc = T.Scene(hs)
    rng = np.random.default_rng(3)
    n = 20000
    rays = np.zeros((n, 7), np.float32)
    rays[:, 0:3] = (0, 2, 14) + rng.normal(size=(n, 3)) * 0.05
    tt = rng.uniform(0, 1, n).astype(np.float32)
    target = np.stack([-6 + 12 * tt, 3 * tt, np.zeros(n)], axis=1) + rng.normal(size=(n, 3)) * 0.6  # around the sphere at its own time
    rays[:, 3:6] = target - rays[:, 0:3]
    rays[:, 6] = tt
    par = sc.intersect(rays, mode=T.MODE_PARITY)
    fast = sc.intersect(rays, mode=T.MODE_FAST)
    on_sphere = (par["hit"] == 1) & (par["prim"] == 0)
    assert on_sphere.sum() > n // 4 and (on_sphere & (tt > 0.6)).sum() > n // 20  # hits far from the t = 0 box
    assert np.array_equal(par["hit"], fast["hit"]) and np.array_equal(par["prim"][par["hit"] == 1], fast["prim"][par["hit"] == 1])
    cam = T.make_camera((0, 2, 14), (0, 1, 0), (0, 1, 0), 50.0, 1.5, 0.0, 10.0, 0.0, 1.0)
    lights = None
    for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
        a = sc.render(cam, T.make_params(96, 64, 64, 8, mode=T.MODE_PARITY, seed=5, kernel=kernel))
        b = sc.render(cam, T.make_params(96, 64, 64, 8, mode=T.MODE_FAST, seed=5, kernel=kernel))
        c = sc.render(cam, T.make_params(96, 64, 64, 8, mode=T.MODE_FAST, seed=5, kernel=kernel, bundle_cull=False))
        assert np.array_equal(b.sum_rgb, c.sum_rgb)  # the bounds tests drop nothing
        rel = common.rel_err(b.sum_rgb, a.sum_rgb, 1e-3 * 64)
        assert (rel > 1e-4).any(axis=-1).mean() < 0.02
        assert a.stats["rays"] == pytest.approx(b.stats["rays"], rel=2e-3)
