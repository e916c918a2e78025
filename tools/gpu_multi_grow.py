# developer tool: tpt_render_multi with growing accumulator sizes (pool growth after peer access was granted)
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import tpt_b200 as T, common
n = T.device_count()
hs = common.host_scene(T, "cornell_box")
cam = T.cornell_camera(1200, 1200)
scenes = [T.Scene(hs, device=g) for g in range(n)]
for subs in (0, 4, 64, 128, 250):
    p = T.make_params(1200, 1200, 2048, 15, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT, subs=subs)
    try:
        r = T.render_multi(scenes, cam, p)
        print("subs", subs, "ok", r.stats["render_ms"], flush=True)
    except Exception as e:
        print("subs", subs, "FAILED", e, flush=True)
scenes2 = [T.Scene(hs, device=g) for g in range(n)]
try:
    r = T.render_multi(scenes2, cam, T.make_params(1200, 1200, 2048, 15, subs=250, kernel=T.KERNEL_WAVEFRONT))
    print("second scene set ok", flush=True)
except Exception as e:
    print("second scene set FAILED", e, flush=True)
