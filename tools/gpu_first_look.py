# developer tool (run on a GPU box via gpurun): first end-to-end look on the GPU
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T
import oracle_ref as O

print("devices", T.device_count())
print("philox", [hex(x) for x in T.philox([0,0,0,0],[0,0])])
hs = T.HostScene("cornell_box")
sc = T.Scene(hs)
rs = O.RefScene("cornell_box")
rng = np.random.default_rng(1)
n = 200000
rays = np.zeros((n,7),np.float32)
rays[:,0:3] = rng.uniform(-290,290,(n,3))
rays[:,3:6] = rng.normal(size=(n,3))
for mode in (T.MODE_PARITY, T.MODE_FAST):
    g = sc.intersect(rays, mode=mode)
    r = rs.hit_batch(rays)
    print("mode",mode,"hit mismatch", int((g['hit']!=r['hit']).sum()), "prim mismatch", int((g['prim']!=r['prim']).sum()),
          "mat mismatch", int((g['mat']!=r['mat']).sum()))
    both = (g['hit']==1)&(r['hit']==1)
    for f in ('t','p','n','u','v'):
        d = np.abs(g[f][both].astype(np.float64)-r[f][both]).max()
        print("   max abs diff",f,d, "bit-exact" if np.array_equal(g[f][both], r[f][both]) else "")
nx=ny=64; ns=16; depth=15
cam = T.cornell_camera(nx,ny)
for mode in (T.MODE_PARITY, T.MODE_FAST):
    p = T.make_params(nx,ny,ns,depth,mode=mode,seed=1234)
    res = sc.render(cam,p)
    ref,_,st = rs.render(O.CORNELL_CAM,nx,ny,ns,depth,seed=1234)
    g = res.sum_rgb[0]; r = ref[0]
    denom = np.maximum(np.abs(r),1e-3)
    rel = np.abs(g-r)/denom
    print("mode",mode,"render stats",res.stats['paths'],res.stats['rays'],res.stats['nan_samples'],res.stats['render_ms'],"ms; ref rays",st['rays'])
    print("   pixels rel>1e-4:", int((rel.max(axis=2)>1e-4).sum()),"of",nx*ny,"max rel",rel.max(),"mean g",g.mean(),"mean r",r.mean(), "bit-exact", np.array_equal(g,r))
# throughput look
nx=ny=1200
cam = T.cornell_camera(nx,ny)
for ns in (64,):
    for mode in (T.MODE_FAST,T.MODE_PARITY):
        p = T.make_params(nx,ny,ns,15,mode=mode,seed=1)
        st = sc.render_device(cam,p)
        st = sc.render_device(cam,p)
        print("mode",mode,"ns",ns,"ms",st['render_ms'],"Mpaths/s",st['paths']/st['render_ms']/1e3,"Mrays/s",st['rays']/st['render_ms']/1e3,"blocks",st['blocks'], "rays/path", st['rays']/st['paths'])
