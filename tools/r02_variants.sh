#!/bin/bash
# developer tool: time every tuning build under gpurun_variants/ (Cornell A/B, parity and fast, wavefront kernel)
mkdir -p gpurun_out
cat > /tmp/vt.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import tpt_b200 as T, common
sc = T.Scene(common.host_scene(T, "cornell_box"))
out = []
for mode, spp in ((T.MODE_PARITY, 256), (T.MODE_FAST, 1024)):
    for fov, depth in ((90.0, 15), (61.93, 50)):
        cam = T.cornell_camera(1200, 1200, fov=fov)
        best = 0
        for i in range(3):
            st = sc.render_device(cam, T.make_params(1200, 1200, spp, depth, mode=mode, seed=1, kernel=T.KERNEL_WAVEFRONT, bundle_cull=False))
            best = max(best, st["paths"] / st["render_ms"] / 1e3)
        out.append(f"{'parity' if mode == T.MODE_PARITY else 'fast'}{'A' if depth == 15 else 'B'} {best:.0f}")
cam = T.cornell_camera(1200, 1200, fov=90.0)
for mode, spp, name in ((T.MODE_FAST, 1024, "fastA+cull"), (T.MODE_PARITY, 256, "parityA+cull")):
    best = 0
    for i in range(3):
        st = sc.render_device(cam, T.make_params(1200, 1200, spp, 15, mode=mode, seed=1, kernel=T.KERNEL_WAVEFRONT, bundle_cull=True))
        best = max(best, st["paths"] / st["render_ms"] / 1e3)
    out.append(f"{name} {best:.0f}")
print("  ".join(out))
PY
for d in gpurun_variants/v*; do
  echo "== $(cat $d/flags.txt)"
  TPT_LIBTPT=$d/libtpt.so python /tmp/vt.py 2>&1 | tail -1
done
