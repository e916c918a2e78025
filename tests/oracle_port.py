"""ctypes access to the plain-C restatement oracle/tpt_oracle.c (test infrastructure only).
It consumes the same flattened tpt_scene_desc the CUDA library consumes."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libtptoracle.so")
_lib = None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB)
        L.tpto_hit_batch.restype = None
        L.tpto_render.restype = C.c_int
        L.tpto_texture_value.restype = None
        L.tpto_quantise.restype = None
        L.tpto_philox4x32_10.restype = None
        _lib = L
    return _lib


def hit_batch(T, host_scene, rays, tmin=0.001, tmax=3.4028234663852886e38):
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 7)
    out = np.zeros(len(rays), T.HIT_DTYPE)
    lib().tpto_hit_batch(host_scene.desc, C.c_void_p(rays.ctypes.data), C.c_size_t(len(rays)), C.c_float(tmin),
                         C.c_float(tmax), C.c_void_p(out.ctypes.data))
    return out


def render(T, host_scene, cam, params, threads=4, per_sample=False):
    sl = max(1, params.slices)
    out = np.zeros((sl, params.ny, params.nx, 3), np.float32)
    samples = np.zeros((params.ny, params.nx, params.ns, 3), np.float32) if per_sample else None
    stats = (C.c_uint64 * 2)()
    rc = lib().tpto_render(host_scene.desc, C.byref(cam), C.byref(params), C.c_int(threads), C.c_void_p(out.ctypes.data),
                           C.c_void_p(samples.ctypes.data) if per_sample else None, stats)
    if rc != 0:
        raise RuntimeError(f"tpto_render failed ({rc})")
    return out, samples, dict(rays=int(stats[0]), draws=int(stats[1]))


def texture_value(T, host_scene, tex, uvp):
    uvp = np.ascontiguousarray(uvp, np.float32).reshape(-1, 5)
    out = np.zeros((len(uvp), 3), np.float32)
    lib().tpto_texture_value(host_scene.desc, C.c_int(tex), C.c_void_p(uvp.ctypes.data), C.c_int(len(uvp)),
                             C.c_void_p(out.ctypes.data))
    return out
