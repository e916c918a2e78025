#!/usr/bin/env python
"""developer tool: FAST vs PARITY renders of the random scene programs under the same Philox stream (what the budgets of
tests/test_gpu_scene_programs.py::test_fast_radiance_tracks_parity were read from)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, tpt_b200 as T
from test_scene_programs import PROGRAM_CAM, PROGRAM_LIGHTS
nx = ny = 48
ns, depth = 16, 12
cam = common.product_camera(T, PROGRAM_CAM, nx, ny)
for fam, seeds in (("program", range(1, 31)), ("programm", range(1, 11)), ("programL", range(1, 9)), ("programLm", range(3, 13, 3))):
    worst4 = worst3 = 0.0
    shifts = []
    for seed in seeds:
        hs = T.HostScene(f"{fam}:{seed}", lights=PROGRAM_LIGHTS)
        try:
            sc = T.Scene(hs)
        except T.TptError:
            continue
        for kernel in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
            a = sc.render(cam, T.make_params(nx, ny, ns, depth, mode=T.MODE_PARITY, seed=5 + seed, kernel=kernel)).sum_rgb
            b = sc.render(cam, T.make_params(nx, ny, ns, depth, mode=T.MODE_FAST, seed=5 + seed, kernel=kernel)).sum_rgb
            rel = common.rel_err(b, a, 1e-3 * ns)
            s4 = float((rel > 1e-4).any(axis=-1).mean()); s3 = float((rel > 1e-3).any(axis=-1).mean())
            worst4 = max(worst4, s4); worst3 = max(worst3, s3)
            shifts.append(abs(float(b.mean()) - float(a.mean())) / max(float(a.mean()), 1e-9))
    print(f"{fam}: worst share of pixels beyond 1e-4 {worst4:.4f}, beyond 1e-3 {worst3:.4f}, worst mean shift {max(shifts):.2e}, median {np.median(shifts):.2e}", flush=True)
