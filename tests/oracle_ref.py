"""ctypes access to the CHECKERS under oracle/ (test infrastructure only).

* ``oracle/_ref/libtptref.so``      the unmodified reference sources + oracle/ref_harness.cc
* ``oracle/_ref/libtptref_det.so``  same, with drand_r re-bound to the injected Philox stream
Both are built by ``oracle/Makefile`` from /root/reference (in the build container); on the GPU
box the prebuilt files travel with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


class RefLight(C.Structure):
    _fields_ = [("kind", C.c_int), ("p", C.c_float * 5)]


class RefRenderArgs(C.Structure):
    _fields_ = [("lookfrom", C.c_float * 3), ("lookat", C.c_float * 3), ("vup", C.c_float * 3),
                ("vfov", C.c_float), ("aspect", C.c_float), ("aperture", C.c_float), ("focus_dist", C.c_float),
                ("t0", C.c_float), ("t1", C.c_float),
                ("nx", C.c_int32), ("ny", C.c_int32), ("ns", C.c_int32), ("max_depth", C.c_int32),
                ("slices", C.c_int32), ("n_lights", C.c_int32), ("lights", RefLight * 8),
                ("deterministic", C.c_int32), ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32),
                ("threads", C.c_int32), ("count_rays", C.c_int32),
                ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32)]


class RefStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("draws", C.c_uint64), ("seconds", C.c_double),
                ("threads", C.c_int32)]


class RefCamera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left", C.c_float * 3), ("vertical", C.c_float * 3),
                ("horizontal", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


REF_HIT_DTYPE = np.dtype([("hit", np.int32), ("prim", np.int32), ("mat", np.int32), ("t", np.float32),
                          ("u", np.float32), ("v", np.float32), ("p", np.float32, 3), ("n", np.float32, 3)])

REFERENCE_LIGHTS = [(0, (-100.0, 100.0, -150.0, -50.0, 298.0)), (1, (120.0, -50.0, 40.0, 120.0, 0.0))]

_libs = {}


def available(det: bool = True) -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libtptref_det.so" if det else "libtptref.so"))


def ref_lib(det: bool = True) -> C.CDLL:
    name = "libtptref_det.so" if det else "libtptref.so"
    if name not in _libs:
        L = C.CDLL(os.path.join(REF_DIR, name))
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_scene_num_leaves.argtypes = [C.c_void_p]
        L.ref_scene_num_materials.argtypes = [C.c_void_p]
        L.ref_hit_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        L.ref_hit_batch.restype = None
        L.ref_camera_make.argtypes = [C.POINTER(C.c_float)] * 3 + [C.c_float] * 6 + [C.POINTER(RefCamera)]
        L.ref_camera_make.restype = None
        L.ref_render.argtypes = [C.c_void_p, C.POINTER(RefRenderArgs), C.c_void_p, C.c_void_p, C.POINTER(RefStats)]
        L.ref_get_perlin.argtypes = [C.c_void_p] * 4
        L.ref_get_perlin.restype = None
        L.ref_set_perlin.argtypes = [C.c_void_p] * 4
        L.ref_set_perlin.restype = None
        L.ref_perlin_turb.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.ref_perlin_turb.restype = None
        L.ref_image_value.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_image_value.restype = None
        L.ref_checker_value.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_checker_value.restype = None
        L.ref_load_image.restype = C.POINTER(C.c_uint8)
        L.ref_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        if det:
            L.ref_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
            L.ref_philox4x32_10.restype = None
        _libs[name] = L
    return _libs[name]


class RefScene:
    def __init__(self, name: str, det: bool = True, image: np.ndarray | None = None):
        self.lib = ref_lib(det)
        self.det = det
        img_p, w, h = None, 0, 0
        if image is not None:
            image = np.ascontiguousarray(image, dtype=np.uint8)
            h, w = image.shape[:2]
            img_p = image.ctypes.data
        self._img = image
        self.h = self.lib.ref_scene_create(name.encode(), img_p, w, h)
        if not self.h:
            raise RuntimeError(f"reference scene {name} could not be built")
        self.n_leaves = self.lib.ref_scene_num_leaves(self.h)
        self.n_materials = self.lib.ref_scene_num_materials(self.h)

    def perlin_tables(self):
        """(ranvec[256,3] f32, perm_x, perm_y, perm_z int32[256]) -- the live static arrays."""
        rv = np.zeros((256, 3), np.float32)
        px, py, pz = (np.zeros(256, np.int32) for _ in range(3))
        self.lib.ref_get_perlin(rv.ctypes.data, px.ctypes.data, py.ctypes.data, pz.ctypes.data)
        return rv, px, py, pz

    def hit_batch(self, rays: np.ndarray, tmin=0.001, tmax=3.4028234663852886e38) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 7)
        out = np.zeros(rays.shape[0], dtype=REF_HIT_DTYPE)
        self.lib.ref_hit_batch(self.h, rays.ctypes.data, rays.shape[0], tmin, tmax, out.ctypes.data)
        return out

    def render(self, cam_args: dict, nx, ny, ns, max_depth, slices=1, lights=REFERENCE_LIGHTS, deterministic=True,
               seed=0x5EED, threads=0, count_rays=True, window=None, per_sample=False):
        a = RefRenderArgs()
        a.lookfrom[:] = cam_args["lookfrom"]
        a.lookat[:] = cam_args["lookat"]
        a.vup[:] = cam_args.get("vup", (0, 1, 0))
        a.vfov = cam_args["vfov"]
        a.aspect = cam_args.get("aspect", float(nx) / float(ny))
        a.aperture = cam_args["aperture"]
        a.focus_dist = cam_args["focus_dist"]
        a.t0, a.t1 = cam_args.get("t0", 0.0), cam_args.get("t1", 0.0)
        a.nx, a.ny, a.ns, a.max_depth, a.slices = nx, ny, ns, max_depth, slices
        a.n_lights = len(lights)
        for i, (kind, p) in enumerate(lights):
            a.lights[i].kind = kind
            for k in range(5):
                a.lights[i].p[k] = p[k] if k < len(p) else 0.0
        a.deterministic = 1 if deterministic else 0
        a.seed_lo, a.seed_hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        a.threads = threads
        a.count_rays = 1 if count_rays else 0
        x0, y0, x1, y1 = window if window else (0, 0, nx, ny)
        a.x0, a.y0, a.x1, a.y1 = x0, y0, x1, y1
        out = np.zeros((slices, ny, nx, 3), np.float32)
        samples = np.zeros((ny, nx, ns, 3), np.float32) if per_sample else None
        st = RefStats()
        rc = self.lib.ref_render(self.h, C.byref(a), out.ctypes.data,
                                 samples.ctypes.data if per_sample else None, C.byref(st))
        if rc != 0:
            raise RuntimeError(f"ref_render failed ({rc})")
        stats = dict(paths=st.paths, rays=st.rays, draws=st.draws, seconds=st.seconds, threads=st.threads)
        return out, samples, stats


def ref_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, t0, t1, det=True) -> RefCamera:
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    out = RefCamera()
    ref_lib(det).ref_camera_make(f3(lookfrom), f3(lookat), f3(vup), vfov, aspect, aperture, focus_dist, t0, t1,
                                 C.byref(out))
    return out


CORNELL_CAM = dict(lookfrom=(0, 0, 800), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=90.0, aperture=0.1, focus_dist=10.0)
BOOK_CAM = dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=20.0, aperture=0.1,
                focus_dist=float(np.sqrt(np.float32(182.0))))
