import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import time, numpy as np
import tpt_b200 as T, common
n = T.device_count()
hs = common.host_scene(T, "cornell_box")
cam = T.cornell_camera(1200, 1200)
for ng in sorted(set([1, 2, n])):
    if ng > n: continue
    scenes = [T.Scene(hs, device=g) for g in range(ng)]
    p = T.make_params(1200, 1200, 2048, 15, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT)
    for it in range(3):
        t0 = time.perf_counter()
        r = T.render_multi(scenes, cam, p)
        dt = time.perf_counter() - t0
    st = r.stats
    print(f"render_multi {ng} GPU: wall {dt*1e3:.1f} ms  slowest-GPU render {st['render_ms']:.1f} ms gather {st['resolve_ms']:.1f} ms  {st['paths']/dt/1e6:.0f} Mpaths/s e2e  launches {st['kernel_launches']} batches/gpu {list(st.get('reserved', []))[:4] if 'reserved' in st else ''}")
