#!/bin/bash
# developer tool: ncu --set full captures of the BASELINE configs 2 and 3 scenes (BVH traversal, Perlin tables, image texture)
cat > /tmp/run_scene.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
scene, kern = sys.argv[1], int(sys.argv[2])
perlin = common.perlin_struct(T, common.golden("textures"))
sc = T.Scene(common.host_scene(T, scene, perlin=perlin, background=T.BG_SKY))
cam = T.book_camera(1600, 1600, fov=20.0, t0=0.0, t1=1.0 if scene.startswith("random") else 0.0)
for i in range(2):
    st = sc.render_device(cam, T.make_params(1600, 1600, 32, 15, mode=T.MODE_FAST, seed=1, kernel=kern))
print(scene, kern, st["render_ms"], "ms", st["paths"]/st["render_ms"]/1e3, "Mpaths/s", st["rays"]/st["render_ms"]/1e3, "Mrays/s")
PY
for sc in ${SCENES:-random_scene two_perlin_spheres earth}; do
  ncu --set full --clock-control none -k regex:render_ -s 1 -c 1 -f -o gpurun_out/scene_${sc}_k${KERN:-0} python /tmp/run_scene.py $sc ${KERN:-0} > gpurun_out/scene_$sc.log 2>&1
  tail -1 gpurun_out/scene_$sc.log
  [ -n "$QUICK" ] || python /tmp/run_scene.py $sc 0 | tail -1
  [ -n "$QUICK" ] || python /tmp/run_scene.py $sc 1 | tail -1
done
