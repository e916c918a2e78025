# developer tool: tiny renders of every kernel variant for compute-sanitizer (memcheck / racecheck / initcheck)
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T, common, raygen
oneweek_cam = lambda nx, ny: T.make_camera((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, nx / ny, 0.0, 10.0, 0.0, 1.0)
for scene, camf in (("cornell_box", T.cornell_camera), ("random_scene", T.book_camera), ("textured_lit", T.book_camera),
                    ("cornell_box_smoke", T.cornell_camera), ("oneweek_final", oneweek_cam)):
    hs = common.host_scene(T, scene, perlin=common.perlin_struct(T, common.golden("textures")),
                           lights=common.TEXTURED_LIGHTS if scene == "textured_lit" else None)
    sc = T.Scene(hs)
    cam = camf(40, 24)
    for mode in (T.MODE_PARITY, T.MODE_FAST):
        for kern in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
            for cull in (True, False):  # kernel builds with and without the pixel-bundle bounds test
                r = sc.render(cam, T.make_params(40, 24, 4, 8, mode=mode, seed=1, kernel=kern, slices=2, bundle_cull=cull), want_slices=True)
                print(scene, mode, kern, cull, r.stats["paths"], r.stats["culled_paths"], float(r.sum_rgb.mean()))
        if scene not in ("cornell_box_smoke", "oneweek_final"):  # tpt_intersect_batch refuses media scenes
            sc.intersect(raygen.primary_batch(scene, 200, 200, seed=1), mode=mode)
    sc.close()
scs = [T.Scene(common.host_scene(T, "cornell_box"), device=0)]
r = T.render_multi(scs, T.cornell_camera(40, 24), T.make_params(40, 24, 4, 8, kernel=T.KERNEL_WAVEFRONT))
print("multi", r.stats["paths"])
