#!/usr/bin/env python
"""developer tool: is FAST mode unbiased on the random scene programs? Converged frames (48x48, 4096 spp) in FAST and PARITY mode:
difference of the image means against the Monte-Carlo noise of such a frame (two PARITY frames with different seeds)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, tpt_b200 as T
from test_scene_programs import PROGRAM_CAM, PROGRAM_LIGHTS
nx = ny = 48
ns, depth = 4096, 12
cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
for name in [f"program:{s}" for s in (1, 2, 3, 5, 8, 13, 21, 34)] + [f"programm:{s}" for s in (2, 4)] + [f"programL:{s}" for s in (1, 2)] + ["programLm:3"]:
    sc = T.Scene(T.HostScene(name, lights=PROGRAM_LIGHTS))
    def frame(mode, seed):
        r = sc.render(cam, T.make_params(nx, ny, ns, depth, mode=mode, seed=seed, kernel=T.KERNEL_WAVEFRONT))
        img = np.nan_to_num(r.sum_rgb[0] / ns)
        return np.minimum(img, 10.0)  # clamp fireflies for the comparison
    a, a2, b = frame(T.MODE_PARITY, 11), frame(T.MODE_PARITY, 12), frame(T.MODE_FAST, 11)
    noise = abs(a.mean() - a2.mean()) / a.mean()
    rm_noise = np.sqrt(((a - a2) ** 2).mean()) / a.mean()
    print(f"{name}: mean parity {a.mean():.5f} fast {b.mean():.5f}  (fast - parity) / parity = {(b.mean() - a.mean()) / a.mean():+.2e}  [parity vs parity, other seed: {noise:.2e}]  rmse/mean fast-parity {np.sqrt(((a - b) ** 2).mean()) / a.mean():.3f} [parity-parity {rm_noise:.3f}]", flush=True)
