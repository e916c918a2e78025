#!/bin/bash
# developer tool: measure every tuning build under gpurun_variants/*/libtpt.so, then the in-tree
# library with and without the folded walls (TPT_SMALL_OPEN_BLOCKS)
cat > /tmp/run_one.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
sc = T.Scene(common.host_scene(T, "cornell_box"))
out=[]
for fov, depth in ((90.0,15),(61.93,50)):
    cam = T.cornell_camera(1200, 1200, fov=fov)
    best=0
    for i in range(4):
        st = sc.render_device(cam, T.make_params(1200, 1200, 256, depth, mode=T.MODE_FAST, seed=1, kernel=int(os.environ.get("KERN","1")), bundle_cull=False))
        best=max(best, st["paths"]/st["render_ms"]/1e3)
    out.append(f"{best:.0f}")
cam = T.cornell_camera(1200, 1200, fov=90.0)
st = sc.render_device(cam, T.make_params(1200, 1200, 64, 15, mode=T.MODE_PARITY, seed=1, kernel=1, bundle_cull=False))
st = sc.render_device(cam, T.make_params(1200, 1200, 64, 15, mode=T.MODE_PARITY, seed=1, kernel=1, bundle_cull=False))
out.append(f"{st['paths']/st['render_ms']/1e3:.0f}")
print(" ".join(out), "Mpaths/s (A fast, B fast, A parity) blocks", st["blocks"])
PY
for d in gpurun_variants/*/; do
  [ -f $d/libtpt.so ] || continue
  echo -n "$(basename $d) [$(cat $d/flags.txt 2>/dev/null)]: "
  TPT_LIBTPT=$d/libtpt.so python /tmp/run_one.py 2>&1 | tail -1
done
echo -n "in-tree, walls as five rects: "; TPT_SMALL_OPEN_BLOCKS=0 python /tmp/run_one.py 2>&1 | tail -1
echo -n "in-tree, full kernels (TPT_NO_LEAN=1): "; TPT_NO_LEAN=1 python /tmp/run_one.py 2>&1 | tail -1
echo -n "in-tree (default): "; python /tmp/run_one.py 2>&1 | tail -1
echo -n "in-tree megakernel: "; KERN=0 python /tmp/run_one.py 2>&1 | tail -1
