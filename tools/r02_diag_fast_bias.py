#!/usr/bin/env python
"""developer tool: where does FAST mode lose energy on a scene program? (fast - parity) / parity of converged means under variations"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common, tpt_b200 as T, test_abi_fuzz as F
from test_scene_programs import PROGRAM_CAM, PROGRAM_LIGHTS
name = sys.argv[1] if len(sys.argv) > 1 else "program:2"
nx = ny = 48
ns = 2048
cam = common.product_camera(T, PROGRAM_CAM, nx, ny)

def mean(sc, mode, depth=12, kernel=T.KERNEL_WAVEFRONT, cull=True, seed=11):
    r = sc.render(cam, T.make_params(nx, ny, ns, depth, mode=mode, seed=seed, kernel=kernel, bundle_cull=cull))
    return float(np.minimum(np.nan_to_num(r.sum_rgb[0] / ns), 10.0).mean())

def report(tag, sc, **kw):
    a, b = mean(sc, T.MODE_PARITY, **kw), mean(sc, T.MODE_FAST, **kw)
    print(f"{tag:46s} parity {a:.5f} fast {b:.5f} rel {(b - a) / max(a, 1e-12):+.2e}", flush=True)

hs = T.HostScene(name, lights=PROGRAM_LIGHTS)
sc = T.Scene(hs)
report("baseline (wavefront, cull on)", sc)
report("mega kernel", sc, kernel=T.KERNEL_MEGA)
report("bundle cull off", sc, cull=False)
for depth in (0, 1, 2, 3, 5):
    report(f"depth limit {depth}", sc, depth=depth)
report("lights: lamp only", T.Scene(T.HostScene(name, lights=PROGRAM_LIGHTS[:1])))
report("lights: sphere shape only", T.Scene(T.HostScene(name, lights=PROGRAM_LIGHTS[1:])))
for env in ("TPT_NO_LEAN", "TPT_SMALL_OPEN_BLOCKS"):
    os.environ[env] = "1" if env == "TPT_NO_LEAN" else "0"
    report(f"{env}={os.environ[env]}", T.Scene(hs))
    del os.environ[env]
src = hs.desc.contents
def variant(tag, edit):
    d, keep = F._clone(T, src)
    edit(d, keep)
    class H: desc = C.pointer(d)
    try:
        report(tag, T.Scene(H.desc))
    except Exception as e:
        print(tag, "refused:", e)
def mats_to(kind_from, to_lambert_tex):
    def edit(d, keep):
        for i in range(d.n_materials):
            m = keep["materials"][i]
            if m.kind == kind_from:
                m.kind = 0
                m.texture = to_lambert_tex
    return edit
def tex_constant(d, keep):
    for i in range(d.n_textures):
        t = keep["textures"][i]
        if t.kind == 1:
            t.kind = 0
            t.color[0] = t.color[1] = t.color[2] = 0.5
variant("dielectric -> lambertian(tex 1)", mats_to(2, 1))
variant("metal -> lambertian(tex 1)", mats_to(1, 1))
variant("checker -> constant 0.5", tex_constant)
def all_lambert(d, keep):
    mats_to(2, 1)(d, keep); mats_to(1, 1)(d, keep); tex_constant(d, keep)
variant("all of the three", all_lambert)
for k in range(1, src.n_prims):
    def drop(d, keep, k=k):
        keep["prims"][k].p[0] += 5000.0 if keep["prims"][k].kind == 0 else 0.0  # move a sphere far away (node boxes unchanged: fast and parity both follow the description...)
    if src.prims[k].kind == 0 and src.n_prims <= 12:
        pass
print("prims:", [(i, src.prims[i].kind, src.materials[src.prims[i].material].kind, src.prims[i].chain, src.prims[i].flags) for i in range(src.n_prims)])
print("materials:", [(i, src.materials[i].kind, src.materials[i].texture) for i in range(src.n_materials)])
print("textures:", [(i, src.textures[i].kind) for i in range(src.n_textures)])
