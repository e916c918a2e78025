#!/bin/bash
# developer tool: BVH / textured-scene throughput of every tuning build under gpurun_variants/ (fast mode, wavefront kernel)
mkdir -p gpurun_out
cat > /tmp/bt.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
perlin = common.perlin_struct(T, common.golden("textures"))
out = []
for scene, spp in (("random_scene", 64), ("random_scene_list", 64), ("oneweek_final", 32), ("cornell_box_smoke", 32), ("two_perlin_spheres", 128), ("earth", 128)):
    img = common.earth_jpg_decoded() if scene in ("earth", "oneweek_final") else None
    sc = T.Scene(T.HostScene(scene, image=img, perlin=perlin))
    if scene == "oneweek_final":
        cam = T.make_camera((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)
    elif scene == "cornell_box_smoke":
        cam = T.cornell_camera(1600, 1600, fov=40.0)
    else:
        cam = T.book_camera(1600, 1600, fov=20.0)
    best = 0
    for i in range(3):
        st = sc.render_device(cam, T.make_params(1600, 1600, spp, 15, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT))
        best = max(best, st["paths"] / st["render_ms"] / 1e3)
    out.append(f"{scene} {best:.0f}")
import hashlib
for scene in ("random_scene", "oneweek_final"):
    img = common.earth_jpg_decoded() if scene == "oneweek_final" else None
    sc = T.Scene(T.HostScene(scene, image=img, perlin=perlin, background=T.BG_SKY if scene == "random_scene" else T.BG_BLACK))
    cam = T.book_camera(256, 256, fov=20.0, t0=0.0, t1=1.0) if scene == "random_scene" else T.make_camera((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)
    res = sc.render(cam, T.make_params(256, 256, 16, 15, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT))
    out.append(f"{scene}_check {float(res.sum_rgb.mean()):.6f} {hashlib.sha1(res.sum_rgb.tobytes()).hexdigest()[:10]}")
print("  ".join(out))
PY
for d in gpurun_variants/v*; do
  echo "== $(cat $d/flags.txt)"
  TPT_LIBTPT=$d/libtpt.so python /tmp/bt.py 2>&1 | tail -1
done
