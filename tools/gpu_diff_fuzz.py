#!/usr/bin/env python
"""developer tool: differential fuzzing of PARITY mode against the plain-C restatement (oracle/tpt_oracle.c).
Valid flattened scenes get their geometry / material / transform floats perturbed (node boxes are left as they
are: both sides consume the same description, consistent or not); then
  * tpt_intersect_batch (parity) vs tpto_hit_batch on camera + interior + secondary rays: records compared bit for bit
  * a small parity render vs tpto_render under the same Philox stream: pixels beyond 1e-4 counted
usage: gpu_diff_fuzz.py scene seed trials"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import oracle_port as P  # noqa: E402
import raygen  # noqa: E402
import test_abi_fuzz as F  # noqa: E402
import tpt_b200 as T  # noqa: E402


class Holder:  # what oracle_port / T.Scene expect of a host scene
    def __init__(self, desc, keep, max_depth):
        self.desc, self._keep, self.max_depth = C.pointer(desc), keep, max_depth


perturb = F.perturb_geometry


def main():
    scene, seed, trials = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    if os.environ.get("TPTO_ROUNDED_TRIG") == "1":  # diagnostic: the restatement rounds sin / cos of a float like the CUDA parity path
        P.lib().tpto_set_rounded_trig(1)
    hs = common.host_scene(T, scene, perlin=common.perlin_struct(T, common.golden("textures")),
                           lights=common.TEXTURED_LIGHTS if scene == "textured_lit" else None,
                           background=T.BG_SKY if scene.startswith("random") else T.BG_BLACK)
    src = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
    rng = np.random.default_rng(seed)
    media = scene in ("cornell_box_smoke", "oneweek_final")
    c = common.RENDER_CASES.get({"cornell_box": "cornell_A", "sphere_cornell_box": "sphere_cornell", "cornell_box_smoke": "cornell_smoke"}.get(scene, scene))
    nx = ny = 48
    cam = common.product_camera(T, c["cam"], nx, ny) if c else T.book_camera(nx, ny, fov=20.0, t0=0.0, t1=1.0)
    tot = dict(trials=0, hit_rays=0, hit_diff=0, pixels=0, pix_bad=0, worst=0.0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", f"diff_fuzz_{scene}_{seed}.log"), "w")
    for trial in range(trials):
        d, keep = F._clone(T, src)
        # every third trial keeps the description as the flattener made it (boxes consistent: the parity walk's cull is on)
        muts = perturb(rng, keep, int(rng.integers(1, 6))) if trial % 3 else []
        h = Holder(d, keep, hs.max_depth)
        try:
            sc = T.Scene(h.desc)
        except Exception as e:  # refused descriptions are test_abi_fuzz's business
            continue
        tot["trials"] += 1
        if not media:
            rays = raygen.primary_batch(scene if scene in raygen.SCENE_INFO else "cornell_box", 3000, 1500, seed=seed * 1000 + trial)
            for gen in range(2):
                got = sc.intersect(rays, mode=T.MODE_PARITY)
                exp = P.hit_batch(T, h, rays)
                both = (got["hit"] == 1) & (exp["hit"] == 1)
                differ = (got["hit"] != exp["hit"]) | (both & ((got["prim"] != exp["prim"]) | ~common.same_float(got["t"], exp["t"])
                                                               | ~common.same_float(got["p"], exp["p"]).all(axis=1)
                                                               | ~common.same_float(got["n"], exp["n"]).all(axis=1)))
                nd = int(differ.sum())
                tot["hit_rays"] += len(rays)
                tot["hit_diff"] += nd
                if nd:
                    log.write(f"HITS trial {trial} gen {gen}: {nd} records differ; {muts!r}\n")
                    log.flush()
                rays = raygen.secondary_rays(got, np.random.default_rng(trial))
                if len(rays) == 0:
                    break
        if os.environ.get("FUZZ_CAMERA", "1") == "1":  # a different thin-lens camera per trial; shutters also OUTSIDE the moving spheres' [0, 1]
            eye, lookat = raygen.SCENE_INFO.get(scene, raygen.SCENE_INFO["cornell_box"])[1:3]
            eye = np.array(eye, np.float64) + rng.normal(0, 0.05 * np.linalg.norm(np.array(eye) - np.array(lookat)), 3)
            t0 = float(rng.uniform(-0.5, 1.0))
            cam = T.make_camera(tuple(eye), tuple(lookat), (0, 1, 0), float(rng.uniform(15, 90)), 1.0, float(rng.choice([0.0, 0.1, 0.3])),
                                float(np.linalg.norm(eye - np.array(lookat))), t0, t0 + float(rng.uniform(0, 1.2)))
        p = T.make_params(nx, ny, 4, int(rng.integers(1, 20)), mode=T.MODE_PARITY, seed=77 + trial, kernel=T.KERNEL_WAVEFRONT if trial % 2 else T.KERNEL_MEGA)
        ref, _, _ = P.render(T, h, cam, p, threads=8)
        res = sc.render(cam, p)
        rel = common.rel_err(res.sum_rgb, ref, 1e-3 * 4)
        bad = int((rel > 1e-4).any(axis=-1).sum())
        tot["pixels"] += nx * ny
        tot["pix_bad"] += bad
        tot["worst"] = max(tot["worst"], float(rel.max()))
        if bad:
            log.write(f"RENDER trial {trial}: {bad} pixels beyond 1e-4 (worst {float(rel.max()):.3g}); {muts!r}\n")
            log.flush()
        sc.close()
    log.write(f"done {tot}\n")
    print(scene, seed, tot)


if __name__ == "__main__":
    main()
