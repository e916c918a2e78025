#!/bin/bash
cat > /tmp/run_variant.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
kern = int(sys.argv[1]); spp = int(sys.argv[2]); fov=float(sys.argv[3]); depth=int(sys.argv[4])
sc = T.Scene(common.host_scene(T, "cornell_box"))
cam = T.cornell_camera(1200, 1200, fov=fov)
for i in range(2):
    st = sc.render_device(cam, T.make_params(1200, 1200, spp, depth, mode=T.MODE_FAST, seed=1, kernel=kern))
print(kern, st["render_ms"], st["paths"]/st["render_ms"]/1e3)
PY
for k in ${KERNELS:-0 1}; do
  ncu --set full --clock-control none --import-source on -k regex:render_ -s 1 -c 1 -f -o gpurun_out/full_k${k} python /tmp/run_variant.py $k ${1:-32} 90 15 > gpurun_out/full_k${k}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
