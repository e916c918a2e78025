# developer tool: throughput of the BVH-scene paths (random_scene, oneweek_final), fast mode, both kernels
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tpt_b200 as T
import common
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
perlin = common.perlin_struct(T, common.golden("textures"))
for scene in ("random_scene", "oneweek_final", "cornell_box_smoke"):
    bg = T.BG_SKY if scene == "random_scene" else T.BG_BLACK
    sc = T.Scene(common.host_scene(T, scene, perlin=perlin, background=bg))
    if scene == "random_scene":
        cam = T.book_camera(1600, 1600, fov=20.0, t0=0.0, t1=1.0)
    elif scene == "oneweek_final":
        cam = T.make_camera((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)
    else:
        cam = T.cornell_camera(1600, 1600, fov=40.0)
    for kern in (0, 1):
        p = T.make_params(1600, 1600, spp, 15, mode=T.MODE_FAST, seed=1, kernel=kern)
        st = sc.render_device(cam, p); st = sc.render_device(cam, p)
        print(f"{scene} kernel {kern}: {st['render_ms']:.2f} ms  {st['paths']/st['render_ms']/1e3:.0f} Mpaths/s  {st['rays']/st['render_ms']/1e3:.0f} Mrays/s rays/path {st['rays']/st['paths']:.3f}", flush=True)
# the README's other two published cases and BASELINE configs[2] at their full sizes (sky background, fov 20, depth 50)
for label, scene, spp in (("README.md:31 no-BVH 1600x1600x100spp (290 s published)", "random_scene_list", 100),
                          ("README.md:34 BVH 1600x1600x300spp (80.1 s published)", "random_scene", 300),
                          ("configs[2]a two_perlin_spheres 1600x1600x256spp", "two_perlin_spheres", 256),
                          ("configs[2]b earth 1600x1600x256spp", "earth", 256)):
    sc = T.Scene(common.host_scene(T, scene, perlin=perlin, background=T.BG_SKY))
    cam = T.book_camera(1600, 1600, fov=20.0, t0=0.0, t1=1.0 if scene.startswith("random") else 0.0)
    p = T.make_params(1600, 1600, spp, 50, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT)
    st = sc.render_device(cam, p); st = sc.render_device(cam, p)
    print(f"{label}: {st['render_ms'] / 1e3:.3f} s  {st['paths']/st['render_ms']/1e3:.0f} Mpaths/s  {st['rays']/st['render_ms']/1e3:.0f} Mrays/s rays/path {st['rays']/st['paths']:.3f}", flush=True)
