#!/usr/bin/env python
"""developer tool: OBSERVED fast-mode deviations from the reference, per scene (not budgets).

gate 1: fast-mode tpt_intersect_batch against the golden hit records the reference produced
        (tests/golden/hits_*.npz): hit/miss and closest-object mismatches on the non-adversarial
        rays, each mismatching ray saved with both records so that it can be classified by hand.
gate 2: fast-mode render under the injected stream against the golden radiance: outlier pixels
        beyond 1e-4 relative, mean shift.
Writes gpurun_out/r02_fast_mismatch.json and gpurun_out/r02_fast_mismatch_rays.npz."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import raygen  # noqa: E402
import tpt_b200 as T  # noqa: E402

N_PRIMARY = {"cornell_box": 3000, "sphere_cornell_box": 1200, "random_scene": 3000, "random_scene_list": 800,
             "two_perlin_spheres": 600, "light_spheres": 600, "earth": 600, "textured_lit": 800}


def main():
    out = {"gate1": {}, "gate2": {}}
    keep = {}
    for scene in common.HIT_SCENES:
        g = common.golden("hits_" + scene)
        rays, exp = g["rays"], g["hits"]
        n_adv = len(raygen.adversarial_rays(*raygen.SCENE_INFO[scene][:2]))
        ok = np.isfinite(rays).all(axis=1)  # secondary rays spawned from a NaN hit record of the adversarial block
        n_nonfinite = int((~ok).sum())
        ok[N_PRIMARY[scene]:N_PRIMARY[scene] + n_adv] = False
        sc = T.Scene(common.host_scene(T, scene))
        got = sc.intersect(rays, mode=T.MODE_FAST)
        par = sc.intersect(rays, mode=T.MODE_PARITY)
        hit_mis = (got["hit"] != exp["hit"])
        both = (got["hit"] == 1) & (exp["hit"] == 1)
        prim_mis = both & (got["prim"] != exp["prim"])
        bad = (hit_mis | prim_mis)
        same = both & ~prim_mis & np.isfinite(exp["t"]) & np.isfinite(got["t"])
        rel_t = np.abs(got["t"][same].astype(np.float64) - exp["t"][same]) / np.maximum(np.abs(exp["t"][same]), 1e-6)
        dt_tie = np.abs(got["t"][prim_mis].astype(np.float64) - exp["t"][prim_mis]) / np.maximum(np.abs(exp["t"][prim_mis]), 1e-6)
        out["gate1"][scene] = {
            "rays": int(ok.sum()), "adversarial_rays": int(n_adv), "non_finite_rays": n_nonfinite,
            "hit_mismatch": int((hit_mis & ok).sum()), "prim_mismatch": int((prim_mis & ok).sum()),
            "hit_mismatch_adversarial": int((hit_mis & ~ok).sum()), "prim_mismatch_adversarial": int((prim_mis & ~ok).sum()),
            "parity_hit_mismatch_all": int((par["hit"] != exp["hit"]).sum()),
            "parity_prim_mismatch_all": int(((par["hit"] == 1) & (exp["hit"] == 1) & (par["prim"] != exp["prim"])).sum()),
            "t_rel_err_max_same_prim": float(rel_t.max(initial=0)), "t_rel_err_p99": float(np.percentile(rel_t, 99)) if len(rel_t) else 0.0,
            "prim_mismatch_t_rel_diff": [float(x) for x in dt_tie[:40]],
        }
        idx = np.nonzero(bad)[0]
        keep[scene + "_idx"] = idx
        keep[scene + "_rays"] = rays[idx]
        keep[scene + "_got"] = got[idx]
        keep[scene + "_exp"] = exp[idx]
        keep[scene + "_adversarial"] = ~ok[idx]
        print(scene, json.dumps(out["gate1"][scene]), flush=True)
    for case in ["cornell_A", "cornell_B", "light_spheres", "textured_lit", "sphere_cornell", "random_scene"]:
        c = common.RENDER_CASES[case]
        g = common.golden("render_" + case)
        perlin = common.perlin_struct(T, g)
        hs = common.host_scene(T, c["scene"], perlin=perlin, lights=c.get("lights"))
        sc = T.Scene(hs)
        cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
        for kname, kernel in (("mega", T.KERNEL_MEGA), ("wavefront", T.KERNEL_WAVEFRONT)):
            p = T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], mode=T.MODE_FAST, slices=c.get("slices", 1), seed=c["seed"], kernel=kernel)
            res = sc.render(cam, p)
            rel = common.rel_err(res.sum_rgb, g["sum_rgb"], 1e-3 * c["ns"])
            badpix = (rel > 1e-4).any(axis=-1)
            out["gate2"][f"{case}/{kname}"] = {
                "pixels": int(badpix.size), "outliers_1e-4": int(badpix.sum()), "outlier_share": float(badpix.mean()),
                "outliers_1e-3": int((rel > 1e-3).any(axis=-1).sum()), "outliers_1e-2": int((rel > 1e-2).any(axis=-1).sum()),
                "mean_fast": float(res.sum_rgb.mean()), "mean_ref": float(g["sum_rgb"].mean()),
                "mean_shift_rel": float(abs(res.sum_rgb.mean() - g["sum_rgb"].mean()) / max(g["sum_rgb"].mean(), 1e-9)),
            }
            print(case, kname, json.dumps(out["gate2"][f"{case}/{kname}"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_fast_mismatch.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "r02_fast_mismatch_rays.npz"), **keep)


if __name__ == "__main__":
    main()
