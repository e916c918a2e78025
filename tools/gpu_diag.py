import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T
import common, raygen
from test_gpu_hits import compare_hits
for scene in common.HIT_SCENES:
    g = common.golden("hits_" + scene)
    rays, exp = g["rays"], g["hits"]
    sc = T.Scene(common.host_scene(T, scene))
    got = sc.intersect(rays, mode=T.MODE_PARITY)
    st = compare_hits(got, exp)
    print(scene, "PARITY", st)
    same = (got['hit']==1)&(exp['hit']==1)
    bad = same & ~common.same_float(got['t'], exp['t'])
    for i in np.nonzero(bad)[0][:5]:
        print("   ray", rays[i], "got", got[i], "exp", exp[i])
    badn = same & ~(common.same_float(got['n'], exp['n']).all(axis=1))
    for i in np.nonzero(badn)[0][:5]:
        print("   N ray", rays[i], "got", got[i], "exp", exp[i])
    ok = np.isfinite(rays).all(axis=1) & (np.abs(rays[:, 3:6]).max(axis=1) > 1e-20) & (np.abs(rays[:, 3:6]).max(axis=1) < 1e20)
    gf = sc.intersect(rays[ok], mode=T.MODE_FAST); e = exp[ok]; r = rays[ok]
    both = (gf['hit']==1)&(e['hit']==1)&(gf['prim']==e['prim'])
    dp = np.linalg.norm(gf['p'][both].astype(np.float64)-e['p'][both], axis=1)
    ext = raygen.SCENE_INFO[scene][0]
    dist = np.abs(e['t'][both].astype(np.float64))*np.linalg.norm(r[both][:,3:6].astype(np.float64),axis=1)
    print(scene, "FAST  max |dp|", dp.max(), "extent", ext, "max dp/dist", (dp/np.maximum(dist,1e-3*ext)).max(), "p99 dp/ext", np.percentile(dp,99)/ext)
    w = np.argsort(-dp)[:3]
    for i in w:
        print("    worst: dp", dp[i], "dist", dist[i], "prim", e['prim'][both][i], "t", e['t'][both][i], gf['t'][both][i])
