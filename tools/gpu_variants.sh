#!/bin/bash
# developer tool: measure every tuning build under gpurun_variants/
cat > /tmp/run_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
sc = T.Scene(common.host_scene(T, "cornell_box"))
out=[]
for fov, depth in ((90.0,15),(61.93,50)):
    cam = T.cornell_camera(1200, 1200, fov=fov)
    best=0
    for i in range(3):
        st = sc.render_device(cam, T.make_params(1200, 1200, 256, depth, mode=T.MODE_FAST, seed=1, kernel=int(os.environ.get("KERN","1"))))
        best=max(best, st["paths"]/st["render_ms"]/1e3)
    out.append(f"{best:.0f}")
print(" ".join(out), "Mpaths/s (A, B) blocks", st["blocks"])
PY
for d in gpurun_variants/v*; do
  echo -n "$(cat $d/flags.txt): "
  TPT_LIBTPT=$d/libtpt.so python /tmp/run_one.py 2>&1 | tail -1
done
