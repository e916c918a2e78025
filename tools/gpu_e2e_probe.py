import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ctypes as C
import numpy as np
import tpt_b200 as T, common
import torch
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda"); torch.cuda.synchronize()
hs = T.HostScene("cornell_box")
cam = T.cornell_camera(1200, 1200)
p = T.make_params(1200, 1200, 2048, 15, mode=T.MODE_FAST, seed=1, kernel=T.KERNEL_WAVEFRONT)
sc = T.Scene(hs)
for i in range(2): st = sc.render_device(cam, p)
print("resident render_ms", st["render_ms"], "wall", st["wall_ms"])
for i in range(5):
    flush.zero_(); torch.cuda.synchronize()
    t0 = time.perf_counter(); s = T.Scene(hs); t1 = time.perf_counter()
    r = s.render(cam, p); t2 = time.perf_counter()
    st = s.stats()
    s.close(); t3 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} ms  render call {1e3*(t2-t1):.1f} ms (kernel {st['render_ms']:.1f}, resolve {st['resolve_ms']:.2f}, d2h {st['d2h_ms']:.1f}, wall {st['wall_ms']:.1f})  close {1e3*(t3-t2):.1f} ms  blocks {st['blocks']}")
