#!/usr/bin/env python
"""developer tool: FAST-mode hit records (normal, point, material) against the restatement on the random scene programs"""
import os, sys, collections
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, oracle_port as P, tpt_b200 as T
from test_scene_programs import program_rays
tot = bad_n = bad_p = bad_m = 0
by_kind = collections.Counter()
for fam, seeds in (("program", range(1, 41)), ("programL", range(1, 9))):
    for seed in seeds:
        hs = T.HostScene(f"{fam}:{seed}")
        d = hs.desc.contents
        try:
            sc = T.Scene(hs)
        except T.TptError:
            continue
        rays = program_rays(seed, 1500)
        rays = rays[np.isfinite(rays).all(axis=1)]
        f = sc.intersect(rays, mode=T.MODE_FAST)
        e = P.hit_batch(T, hs, rays)
        same = (f["hit"] == 1) & (e["hit"] == 1) & (f["prim"] == e["prim"]) & np.isfinite(e["n"]).all(axis=1) & np.isfinite(f["n"]).all(axis=1)
        dn = np.abs(f["n"][same] - e["n"][same]).max(axis=1)
        scale = np.maximum(np.linalg.norm(e["p"][same], axis=1), 1.0)
        dp = np.abs(f["p"][same] - e["p"][same]).max(axis=1) / scale
        tot += int(same.sum()); bad_n += int((dn > 1e-3).sum()); bad_p += int((dp > 1e-4).sum()); bad_m += int((f["mat"][same] != e["mat"][same]).sum())
        for prim in e["prim"][same][dn > 1e-3]:
            pr = d.prims[int(prim)]
            by_kind[(pr.kind, pr.flags & 1, d.chains[pr.chain].n_ops)] += 1
        if (dn > 1e-3).any():
            i = np.nonzero(dn > 1e-3)[0][0]
            print(f"{fam}:{seed}: {int((dn > 1e-3).sum())} normals differ, e.g. prim {int(e['prim'][same][i])} fast {f['n'][same][i]} ref {e['n'][same][i]}", flush=True)
print("records compared", tot, "normals beyond 1e-3:", bad_n, "points beyond 1e-4 rel:", bad_p, "material ids differing:", bad_m)
print("normal differences by (prim kind, flipped, chain length):", dict(by_kind))
