#!/usr/bin/env python
"""developer tool: mutated scene descriptions through the WHOLE of tpt_scene_create + tiny renders + intersect on a GPU.
tests/test_abi_fuzz.py covers the host-only validation; this covers what runs after it (blob, SAH BVH, flat tree, kernels).
usage: gpu_fuzz.py scene seed trials   (one process per scene: a crash or a sticky CUDA error ends it, the last mutation
is in gpurun_out/fuzz_<scene>_<seed>.log)"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
import test_abi_fuzz as F  # noqa: E402
import tpt_b200 as T  # noqa: E402

scene, seed, trials = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
hs = common.host_scene(T, scene, perlin=common.perlin_struct(T, common.golden("textures")),
                       lights=common.TEXTURED_LIGHTS if scene == "textured_lit" else None)
lib = T.lib()
src = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
rng = np.random.default_rng(seed)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", f"fuzz_{scene}_{seed}.log"), "w")
cam = T.cornell_camera(16, 16) if "cornell" in scene else T.book_camera(16, 16, fov=20.0, t0=0.0, t1=1.0)
counts = {}
for trial in range(trials):
    d, keep = F._clone(T, src)
    muts = []
    for _ in range(int(rng.integers(1, 4))):
        aname = list(keep)[int(rng.integers(len(keep)))]
        buf = keep[aname]
        words = np.frombuffer(buf, dtype=np.uint32)
        rec = C.sizeof(buf._type_) // 4
        n_words = (len(buf) - 1) * rec
        i = int(rng.integers(n_words))
        old = int(words[i])
        mode = rng.integers(0, 4)
        if mode == 0:
            words[i] = int(rng.choice(F.INTERESTING)) & 0xFFFFFFFF
        elif mode == 1:
            words[i] = (old ^ (1 << int(rng.integers(32)))) & 0xFFFFFFFF
        elif mode == 2:
            words[i] = np.float32(rng.choice([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e38, -1e38, 1e-45])).view(np.uint32)
        else:
            words[i] = int(rng.integers(-4, 600)) & 0xFFFFFFFF
        muts.append((aname, i // rec, i % rec, hex(old), hex(int(words[i]))))
    log.seek(0)
    log.truncate()
    log.write(f"trial {trial} {muts!r}\n")
    log.flush()
    os.fsync(log.fileno())
    handle = C.c_void_p()
    rc = lib.tpt_scene_create(C.byref(d), 0, C.byref(handle))
    counts[rc] = counts.get(rc, 0) + 1
    if rc != 0:
        continue
    try:
        for mode in (T.MODE_PARITY, T.MODE_FAST):
            for kern in (T.KERNEL_MEGA, T.KERNEL_WAVEFRONT):
                p = T.make_params(16, 16, 2, 4, mode=mode, seed=1, kernel=kern)
                img = T.Image()
                sums = np.zeros((16, 16, 3), np.float32)
                img.sum_rgb = sums.ctypes.data_as(C.POINTER(C.c_float))
                r = lib.tpt_render(handle, C.byref(cam), C.byref(p), C.byref(img))
                counts[("render", r)] = counts.get(("render", r), 0) + 1
                if r not in (0, -4):
                    print("RENDER ERROR", r, lib.tpt_last_error(), muts, flush=True)
                    sys.exit(3)
    finally:
        lib.tpt_scene_destroy(handle)
log.seek(0)
log.truncate()
log.write("completed\n")
print(scene, seed, "completed", counts)
