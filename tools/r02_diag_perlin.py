# developer tool: Perlin-textured scene at a larger size: CUDA parity vs the live injected reference vs the plain-C restatement
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T, common, oracle_port as P, oracle_ref as O
nx = ny = 500
ns = 2
cam_args = dict(common.BOOK_CAM, vfov=40.0)
rs = O.RefScene("light_spheres")
rv, px, py, pz = rs.perlin_tables()
perlin = common.perlin_struct(T, dict(ranvec=rv, perm_x=px, perm_y=py, perm_z=pz))
ref, _, st = rs.render(cam_args, nx, ny, ns, 15, seed=7)
hs = T.HostScene("light_spheres", perlin=perlin)
cam = common.product_camera(T, cam_args, nx, ny)
p = T.make_params(nx, ny, ns, 15, mode=T.MODE_PARITY, seed=7, kernel=T.KERNEL_WAVEFRONT)
gpu = T.Scene(hs).render(cam, p).sum_rgb
port, _, _ = P.render(T, hs, cam, p, threads=16)
def bad(a, b):
    rel = common.rel_err(a, b, 1e-3 * ns)
    return int((rel > 1e-4).any(axis=-1).sum()), float(rel.max())
print("pixels", nx * ny, "mean", float(ref.mean()))
print("gpu  vs reference:", bad(gpu, ref))
print("port vs reference:", bad(port, ref))
print("gpu  vs port     :", bad(gpu, port))
