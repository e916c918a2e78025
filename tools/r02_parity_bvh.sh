#!/bin/bash
# developer tool: parity-mode throughput of the BVH scenes (skip walk with / without the "behind the best hit" cull)
cat > /tmp/pb.py <<'PY'
import sys, os, hashlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
perlin = common.perlin_struct(T, common.golden("textures"))
out = []
for scene in ("random_scene", "random_scene_list"):
    sc = T.Scene(common.host_scene(T, scene, perlin=perlin, background=T.BG_SKY))
    cam = T.book_camera(1600, 1600, fov=20.0, t0=0.0, t1=1.0)
    best = 0
    for i in range(2):
        st = sc.render_device(cam, T.make_params(1600, 1600, 16, 15, mode=T.MODE_PARITY, seed=1, kernel=T.KERNEL_WAVEFRONT))
        best = max(best, st["paths"] / st["render_ms"] / 1e3)
    res = sc.render(T.book_camera(400, 400, fov=20.0, t0=0.0, t1=1.0), T.make_params(400, 400, 16, 15, mode=T.MODE_PARITY, seed=1, kernel=T.KERNEL_WAVEFRONT))
    out.append(f"{scene} {best:.0f} sha {hashlib.sha1(res.sum_rgb.tobytes()).hexdigest()[:12]}")
print("  ".join(out))
PY
for cfg in "1 1" "0 1" "0 0"; do set -- $cfg; echo "== TPT_PARITY_SKIP_ORDERED=$1 TPT_PARITY_SKIP_CULL=$2"; TPT_PARITY_SKIP_ORDERED=$1 TPT_PARITY_SKIP_CULL=$2 python /tmp/pb.py 2>&1 | tail -1; done
