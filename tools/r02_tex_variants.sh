#!/bin/bash
# developer tool: textured small scenes, background / depth / CTA-shape A/B
cat > /tmp/tt.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
perlin = common.perlin_struct(T, common.golden("textures"))
out = []
for scene in ("two_perlin_spheres", "earth", "light_spheres"):
    for bg, depth in ((T.BG_SKY, 50), (T.BG_BLACK, 15)):
        img = common.earth_jpg_decoded() if scene == "earth" else None
        sc = T.Scene(T.HostScene(scene, image=img, perlin=perlin, background=bg))
        cam = T.book_camera(1600, 1600, fov=20.0)
        best = 0
        for i in range(3):
            st = sc.render_device(cam, T.make_params(1600, 1600, 128, depth, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT))
            best = max(best, st["paths"] / st["render_ms"] / 1e3)
        out.append(f"{scene}/{'sky' if bg else 'black'}/d{depth} {best:.0f} ({st['rays']/st['paths']:.2f} r/p)")
print("  ".join(out))
PY
for d in gpurun_variants/v*; do
  echo "== $(cat $d/flags.txt)"
  TPT_LIBTPT=$d/libtpt.so python /tmp/tt.py 2>&1 | tail -1
done
