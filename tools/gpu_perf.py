# developer tool: perf probe (run on a GPU box via gpurun)
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tpt_b200 as T
import common
kernels = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["0"])]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256
hs = common.host_scene(T, "cornell_box")
sc = T.Scene(hs)
c = common.RENDER_CASES["cornell_A"]; gold = common.golden("render_cornell_A")
cam = common.product_camera(T, c["cam"], c["nx"], c["ny"])
for kern in kernels:
    for mode in (T.MODE_PARITY, T.MODE_FAST):
        try:
            res = sc.render(cam, T.make_params(c["nx"], c["ny"], c["ns"], c["depth"], mode=mode, seed=c["seed"], kernel=kern))
        except T.TptError as e:
            print("kernel", kern, "mode", mode, "ERR", e); continue
        rel = common.rel_err(res.sum_rgb, gold["sum_rgb"], 1e-3 * c["ns"])
        print("kernel", kern, "mode", mode, "golden outliers", int((rel > 1e-4).any(axis=-1).sum()), "of", c["nx"]*c["ny"], "mean", res.sum_rgb.mean(), gold["sum_rgb"].mean())
    for variant, fov, depth in (("A", 90.0, 15), ("B", 61.93, 50)):
        cam2 = T.cornell_camera(1200, 1200, fov=fov)
        for mode in (T.MODE_FAST, T.MODE_PARITY):
            p = T.make_params(1200, 1200, spp if mode == T.MODE_FAST else max(16, spp // 8), depth, mode=mode, seed=1, kernel=kern)
            try:
                st = sc.render_device(cam2, p); st = sc.render_device(cam2, p)
            except T.TptError as e:
                print("kernel", kern, "ERR", e); continue
            print(f"kernel {kern} variant {variant} mode {mode}: {st['render_ms']:.2f} ms  {st['paths']/st['render_ms']/1e3:.0f} Mpaths/s  {st['rays']/st['render_ms']/1e3:.0f} Mrays/s  rays/path {st['rays']/st['paths']:.3f} blocks {st['blocks']}")
for scene, camf in (("random_scene", T.book_camera), ("random_scene_list", T.book_camera), ("textured_lit", T.book_camera),
                    ("cornell_box_smoke", T.cornell_camera), ("oneweek_final", lambda nx, ny, fov: T.make_camera((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0))):
    s2 = T.Scene(common.host_scene(T, scene, perlin=common.perlin_struct(T, common.golden("textures")), lights=common.TEXTURED_LIGHTS if scene == "textured_lit" else None))
    cam3 = camf(800, 800, fov=20.0 if "random" in scene else 50.0) if scene != "oneweek_final" else camf(800, 800, 40.0)
    for mode, kern in ((T.MODE_FAST, 0), (T.MODE_FAST, 1), (T.MODE_PARITY, 0), (T.MODE_PARITY, 1)):
        p = T.make_params(800, 800, 16, 15, mode=mode, seed=1, kernel=kern)
        st = s2.render_device(cam3, p); st = s2.render_device(cam3, p)
        print(f"{scene} mode {mode} kernel {kern}: {st['render_ms']:.2f} ms  {st['paths']/st['render_ms']/1e3:.0f} Mpaths/s  {st['rays']/st['render_ms']/1e3:.0f} Mrays/s rays/path {st['rays']/st['paths']:.3f}")
