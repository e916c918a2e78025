# developer tool: where does parity mode differ from the plain-C restatement on random_scene + sky at full resolution?
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T, common, oracle_port as P
nx = ny = 1600
hs = T.HostScene("random_scene", perlin=common.perlin_struct(T, common.golden("textures")), background=T.BG_SKY)
cam = common.product_camera(T, common.BOOK_CAM, nx, ny)
p = T.make_params(nx, ny, 1, 15, mode=T.MODE_PARITY, seed=99, kernel=T.KERNEL_WAVEFRONT)
ref, samples, st = P.render(T, hs, cam, p, threads=16)
a = T.Scene(hs).render(cam, p).sum_rgb
os.environ["TPT_PARITY_SKIP"] = "0"
b = T.Scene(hs).render(cam, p).sum_rgb
print("skip == frame replay:", np.array_equal(a, b), int((a != b).any(axis=-1).sum()))
for name, img in (("skip", a), ("frame", b)):
    rel = common.rel_err(img, ref, 1e-3)
    bad = np.argwhere((rel > 1e-4).any(axis=-1))
    print(name, "outliers", len(bad))
    for s_, j, i in bad[:8]:
        print("  pixel", i, j, "gpu", img[s_, j, i], "port", ref[s_, j, i])
