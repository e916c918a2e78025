#!/usr/bin/env python
"""developer tool: replay one trial of tools/gpu_diff_fuzz.py (scene seed trial) and compare kernels / per-sample values"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, oracle_port as P, test_abi_fuzz as F, tpt_b200 as T
scene, seed, want = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
P.lib().tpto_set_rounded_trig(1)
hs = common.host_scene(T, scene, perlin=common.perlin_struct(T, common.golden("textures")), background=T.BG_SKY if scene.startswith("random") else T.BG_BLACK)
src = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
rng = np.random.default_rng(seed)
c = common.RENDER_CASES.get({"cornell_box": "cornell_A", "sphere_cornell_box": "sphere_cornell"}.get(scene, scene))
nx = ny = 48
cam = common.product_camera(T, c["cam"], nx, ny) if c else T.book_camera(nx, ny, fov=20.0, t0=0.0, t1=1.0)
for trial in range(want + 1):
    d, keep = F._clone(T, src)
    muts = F.perturb_geometry(rng, keep, int(rng.integers(1, 6)))
    if trial < want:
        continue
    class H: desc = C.pointer(d)
    sc = T.Scene(H.desc)
    print("mutations", muts)
    p = T.make_params(nx, ny, 4, 12, mode=T.MODE_PARITY, seed=77 + trial, kernel=T.KERNEL_MEGA)
    ref, samples, st = P.render(T, H, cam, p, threads=8, per_sample=True)
    for kname, kern in (("mega", T.KERNEL_MEGA), ("wave", T.KERNEL_WAVEFRONT)):
        p.kernel = kern
        res = sc.render(cam, p)
        rel = common.rel_err(res.sum_rgb, ref, 1e-3 * 4)
        bad = np.argwhere((rel > 1e-4).any(axis=-1))
        print(kname, "bad pixels", bad.tolist(), "rays gpu", res.stats["rays"], "ref", st["rays"])
        for (s, y, x) in bad.tolist():
            print("  pixel", y, x, "gpu", res.sum_rgb[s, y, x], "ref", ref[s, y, x], "samples", samples[y, x].tolist())
            for env in ("TPT_PARITY_SKIP_CULL", "TPT_PARITY_SKIP"):
                os.environ[env] = "0"
                sc2 = T.Scene(H.desc)
                r2 = sc2.render(cam, p)
                print("   ", env, "=0 ->", r2.sum_rgb[s, y, x])
                del os.environ[env]
