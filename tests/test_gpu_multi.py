"""In-process multi-GPU path (tpt_render_multi): static tile split + work stealing + NVLink gather.
Runs with however many GPUs are visible (1 on the default test box: the batching, per-batch work
counters, resolve and gather-free path are still exercised; 2+ under `gpurun --gpus N`)."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,kernel", [("parity", "mega"), ("fast", "wavefront")])
def test_multi_equals_single(T, gpu, mode, kernel):
    m = T.MODE_PARITY if mode == "parity" else T.MODE_FAST
    k = T.KERNEL_MEGA if kernel == "mega" else T.KERNEL_WAVEFRONT
    nx, ny, ns = 300, 200, 16
    cam = T.cornell_camera(nx, ny)
    hs = common.host_scene(T, "cornell_box")
    single = T.Scene(hs, device=0).render(cam, T.make_params(nx, ny, ns, 15, mode=m, seed=3, kernel=k, slices=2, subs=1),
                                          want_slices=True)
    n = min(gpu, 8)
    scenes = [T.Scene(hs, device=g) for g in range(n)]
    multi = T.render_multi(scenes, cam, T.make_params(nx, ny, ns, 15, mode=m, seed=3, kernel=k, slices=2, subs=1),
                           want_slices=True)
    assert np.array_equal(multi.sum_rgb, single.sum_rgb)
    assert np.array_equal(multi.rgb8, single.rgb8)
    assert np.array_equal(multi.rgb8_slices, single.rgb8_slices)
    assert multi.stats["paths"] == nx * ny * ns == single.stats["paths"]
    assert multi.stats["rays"] == single.stats["rays"]
    assert multi.stats["kernel_launches"] == 8 * n + n  # 8 batches per GPU + one resolve each


def test_multi_rejects_bad_arguments(T, gpu):
    hs = common.host_scene(T, "cornell_box")
    a, b = T.Scene(hs, device=0), T.Scene(hs, device=0)
    cam = T.cornell_camera(32, 32)
    with pytest.raises(T.TptError):  # two scenes on one device
        T.render_multi([a, b], cam, T.make_params(32, 32, 2, 5))
    with pytest.raises(T.TptError):  # the call partitions the frame itself
        T.render_multi([a], cam, T.make_params(32, 32, 2, 5, part_index=1, part_count=2))
