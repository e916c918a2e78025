import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T
import common, raygen
scene="sphere_cornell_box"
g = common.golden("hits_" + scene)
rays, exp = g["rays"], g["hits"]
sc = T.Scene(common.host_scene(T, scene))
gf = sc.intersect(rays, mode=T.MODE_FAST)
gp = sc.intersect(rays, mode=T.MODE_PARITY)
bad = np.nonzero((gf['hit']==1)&(exp['hit']==1)&(gf['prim']!=exp['prim']))[0]
print(len(bad), "mismatches; index range", bad.min(), bad.max(), "n rays", len(rays))
for i in bad[:12]:
    print(i, "ray", rays[i], "\n    fast", gf[i]['prim'], gf[i]['t'], "ref", exp[i]['prim'], exp[i]['t'], "parity", gp[i]['prim'], gp[i]['t'])
