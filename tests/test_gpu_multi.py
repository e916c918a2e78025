"""In-process multi-GPU path (tpt_render_multi): static tile split + work stealing + NVLink gather.
The cross-device test needs >= 2 GPUs and is skipped (never vacuously passed) on a one-GPU box; it
runs under `gpurun --gpus 2` (tools/r02_multi.sh keeps its output under profiles/)."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _multi_vs_single(T, n, mode, kernel):
    m = T.MODE_PARITY if mode == "parity" else T.MODE_FAST
    k = T.KERNEL_MEGA if kernel == "mega" else T.KERNEL_WAVEFRONT
    nx, ny, ns = 300, 200, 16
    cam = T.cornell_camera(nx, ny)
    hs = common.host_scene(T, "cornell_box")
    single = T.Scene(hs, device=0).render(cam, T.make_params(nx, ny, ns, 15, mode=m, seed=3, kernel=k, slices=2, subs=1),
                                          want_slices=True)
    scenes = [T.Scene(hs, device=g) for g in range(n)]
    multi = T.render_multi(scenes, cam, T.make_params(nx, ny, ns, 15, mode=m, seed=3, kernel=k, slices=2, subs=1),
                           want_slices=True)
    assert np.array_equal(multi.sum_rgb, single.sum_rgb)
    assert np.array_equal(multi.rgb8, single.rgb8)
    assert np.array_equal(multi.rgb8_slices, single.rgb8_slices)
    assert multi.stats["paths"] == nx * ny * ns == single.stats["paths"]
    assert multi.stats["rays"] == single.stats["rays"]
    n_static = (8 * n * 7 // 8) // n * n
    assert multi.stats["kernel_launches"] == n + (8 * n - n_static) + n  # per GPU: ONE launch for its static share, the stolen batches, one resolve
    st = multi.stats
    assert st["multi_gpus"] == n and st["multi_batches_total"] == 8 * n
    assert sum(st["multi_batches"]) == 8 * n and len(st["multi_batches"]) == n
    assert sum(st["multi_stolen"]) == 8 * n - (8 * n * 7 // 8) // n * n  # what the static shares leave over
    assert all(b >= (8 * n * 7 // 8) // n for b in st["multi_batches"])  # nobody does less than its static share
    return st


@pytest.mark.parametrize("mode,kernel", [("parity", "mega"), ("fast", "wavefront"), ("parity", "wavefront")])
def test_multi_equals_single_across_gpus(T, gpu, mode, kernel):
    """the REAL multi-GPU path: peer-access pool, NVLink gather kernels, stealing between devices.
    Needs >= 2 GPUs (`gpurun --gpus 2`, tools/gpu_round.sh); on a one-GPU box it is SKIPPED, not passed."""
    if gpu < 2:
        pytest.skip("one GPU visible: the cross-device gather / steal path cannot run here (see test_multi_single_device_path)")
    st = _multi_vs_single(T, min(gpu, 8), mode, kernel)
    assert st["multi_gather_ms"] > 0


def test_multi_single_device_path(T, gpu):
    """n = 1: batching, per-batch work counters, stealing from its own leftovers, resolve of the owned-batch
    mask -- everything of tpt_render_multi except the cross-device gather (covered above with >= 2 GPUs)."""
    _multi_vs_single(T, 1, "fast", "wavefront")


def test_multi_rejects_bad_arguments(T, gpu):
    hs = common.host_scene(T, "cornell_box")
    a, b = T.Scene(hs, device=0), T.Scene(hs, device=0)
    cam = T.cornell_camera(32, 32)
    with pytest.raises(T.TptError):  # two scenes on one device
        T.render_multi([a, b], cam, T.make_params(32, 32, 2, 5))
    with pytest.raises(T.TptError):  # the call partitions the frame itself
        T.render_multi([a], cam, T.make_params(32, 32, 2, 5, part_index=1, part_count=2))
