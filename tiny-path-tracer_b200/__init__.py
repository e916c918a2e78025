"""tiny-path-tracer_b200 -- Python plumbing over the C-ABI of include/tpt.h.

The product is two native libraries built in-tree:

* ``lib/libtpt.so``       hand-written sm_100a CUDA kernels + the extern "C" layer (csrc/)
* ``lib/libtpt_host.so``  the C++ scene-description front end + flattener (host/)

This module only binds them with ctypes for the tests and ``bench.py``; it contains no
rendering logic and no CPU fallback: if ``libtpt.so`` is missing, or no B200 is visible, the
calls raise instead of computing anything on the host.

The package directory name contains a hyphen; import it with
``importlib.import_module("tiny-path-tracer_b200")`` (see ``tpt_b200.py`` at the repo root).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
LIB_DIR = os.path.join(_HERE, "lib")
LIBTPT = os.environ.get("TPT_LIBTPT", os.path.join(LIB_DIR, "libtpt.so"))  # override: tuning builds only
LIBHOST = os.path.join(LIB_DIR, "libtpt_host.so")

TPT_API_VERSION = 2
TPT_MAX_GPUS = 16
TPT_TILE = 16
MODE_PARITY, MODE_FAST = 0, 1
KERNEL_MEGA, KERNEL_WAVEFRONT = 0, 1
BG_BLACK, BG_SKY = 0, 1
LIGHT_XZ_RECT, LIGHT_SPHERE, LIGHT_OTHER = 0, 1, 2
STATUS = {0: "TPT_OK", -1: "TPT_ERR_INVALID", -2: "TPT_ERR_CUDA", -3: "TPT_ERR_NO_DEVICE",
          -4: "TPT_ERR_UNSUPPORTED", -5: "TPT_ERR_NOMEM"}

# every symbol include/tpt.h declares (checked by tests/test_abi.py)
C_ABI_SYMBOLS = [
    "tpt_api_version", "tpt_device_count", "tpt_device_warm", "tpt_last_error", "tpt_scene_create", "tpt_scene_destroy",
    "tpt_intersect_batch", "tpt_render", "tpt_render_device", "tpt_render_fetch", "tpt_get_stats",
    "tpt_device_buffers", "tpt_render_multi",
    "tpt_debug_philox", "tpt_debug_texture", "tpt_debug_small_scene", "tpt_debug_fp32_peak",
]


class TptError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


# ----------------------------------------------------------------------------- ctypes mirrors
class Node(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("kind", C.c_int32), ("bmax", C.c_float * 3), ("end_or_prim", C.c_int32)]


class XformOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_float), ("b", C.c_float), ("c", C.c_float)]


class Chain(C.Structure):
    _fields_ = [("first_op", C.c_int32), ("n_ops", C.c_int32)]


class Prim(C.Structure):
    _fields_ = [("kind", C.c_int32), ("material", C.c_int32), ("chain", C.c_int32), ("flags", C.c_int32),
                ("p", C.c_float * 12)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture", C.c_int32), ("albedo", C.c_float * 3), ("fuzz", C.c_float),
                ("ref_idx", C.c_float), ("pad", C.c_int32)]


class Texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("color", C.c_float * 3), ("odd", C.c_int32), ("even", C.c_int32),
                ("scale", C.c_float), ("image", C.c_int32)]


class ImageDesc(C.Structure):
    _fields_ = [("rgb", C.POINTER(C.c_uint8)), ("width", C.c_int32), ("height", C.c_int32)]


class PerlinTables(C.Structure):
    _fields_ = [("ranvec", (C.c_float * 3) * 256), ("perm_x", C.c_int32 * 256), ("perm_y", C.c_int32 * 256),
                ("perm_z", C.c_int32 * 256)]


class Light(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p", C.c_float * 5), ("pad", C.c_int32 * 2)]


class SceneDesc(C.Structure):
    _fields_ = [("api_version", C.c_int32), ("n_nodes", C.c_int32), ("n_prims", C.c_int32), ("n_chains", C.c_int32),
                ("n_xform_ops", C.c_int32), ("n_materials", C.c_int32), ("n_textures", C.c_int32),
                ("n_images", C.c_int32), ("n_lights", C.c_int32),
                ("nodes", C.POINTER(Node)), ("prims", C.POINTER(Prim)), ("chains", C.POINTER(Chain)),
                ("xform_ops", C.POINTER(XformOp)), ("materials", C.POINTER(Material)),
                ("textures", C.POINTER(Texture)), ("images", C.POINTER(ImageDesc)),
                ("perlin", C.POINTER(PerlinTables)), ("lights", C.POINTER(Light)),
                ("background", C.c_int32), ("n_root_nodes", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("lower_left_corner", C.c_float * 3), ("vertical", C.c_float * 3),
                ("horizontal", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("lens_radius", C.c_float), ("time0", C.c_float), ("time1", C.c_float)]


class RenderParams(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("ns", C.c_int32), ("max_depth", C.c_int32),
                ("slices", C.c_int32), ("mode", C.c_int32), ("kernel", C.c_int32),
                ("seed_lo", C.c_uint32), ("seed_hi", C.c_uint32), ("t_min", C.c_float),
                ("part_index", C.c_int32), ("part_count", C.c_int32), ("device", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class Image(C.Structure):
    _fields_ = [("sum_rgb", C.POINTER(C.c_float)), ("rgb8", C.POINTER(C.c_uint8)),
                ("rgb8_slices", C.POINTER(C.c_uint8))]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("nan_samples", C.c_uint64),
                ("render_ms", C.c_double), ("resolve_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("wall_ms", C.c_double), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_int32), ("sm_count", C.c_int32),
                ("blocks", C.c_int32), ("threads_per_block", C.c_int32), ("reserved", C.c_int32 * 4),
                ("culled_paths", C.c_uint64),
                ("multi_gpus", C.c_int32), ("multi_batches_total", C.c_int32), ("multi_batches", C.c_int32 * 16),
                ("multi_stolen", C.c_int32 * 16), ("multi_busy_ms", C.c_double * 16), ("multi_gather_ms", C.c_double)]


RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("time", np.float32)])
HIT_DTYPE = np.dtype([("hit", np.int32), ("prim", np.int32), ("mat", np.int32), ("t", np.float32),
                      ("u", np.float32), ("v", np.float32), ("p", np.float32, 3), ("n", np.float32, 3)])
assert RAY_DTYPE.itemsize == 28 and HIT_DTYPE.itemsize == 48

# ----------------------------------------------------------------------------------- loading
_lib = None
_host = None


def build(verbose: bool = False) -> None:
    """Compile both native libraries in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    for sub in ("csrc", "host"):
        r = subprocess.run(["make", "-C", os.path.join(_HERE, sub)], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout[-4000:], r.stderr[-4000:])
        if r.returncode != 0:
            raise RuntimeError(f"building {sub} failed")


def lib() -> C.CDLL:
    """The CUDA core. Fails loudly when the extension is missing -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBTPT):
            raise RuntimeError(f"{LIBTPT} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is the only implementation; there is no CPU fallback)")
        L = C.CDLL(LIBTPT)
        L.tpt_last_error.restype = C.c_char_p
        L.tpt_scene_create.argtypes = [C.POINTER(SceneDesc), C.c_int, C.POINTER(C.c_void_p)]
        L.tpt_scene_destroy.argtypes = [C.c_void_p]
        L.tpt_scene_destroy.restype = None
        L.tpt_intersect_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_int,
                                          C.c_void_p]
        L.tpt_render.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderParams), C.POINTER(Image)]
        L.tpt_render_device.argtypes = [C.c_void_p, C.POINTER(Camera), C.POINTER(RenderParams)]
        L.tpt_render_fetch.argtypes = [C.c_void_p, C.POINTER(Image)]
        L.tpt_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.tpt_render_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Camera), C.POINTER(RenderParams),
                                       C.POINTER(Image)]
        L.tpt_device_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.tpt_debug_philox.argtypes = [C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.tpt_debug_texture.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        if "TPT_LIBTPT" not in os.environ or hasattr(L, "tpt_debug_fp32_peak"):  # tuning builds of older sources lack the probes
            L.tpt_debug_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
            L.tpt_debug_small_scene.argtypes = [C.POINTER(SceneDesc), C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def host() -> C.CDLL:
    global _host
    if _host is None:
        if not os.path.exists(LIBHOST):
            raise RuntimeError(f"{LIBHOST} is missing: run __graft_entry__.build()")
        H = C.CDLL(LIBHOST)
        H.tpt_host_last_error.restype = C.c_char_p
        H.tpt_host_build_scene.restype = C.c_void_p
        H.tpt_host_build_scene.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int]
        H.tpt_host_scene_desc.restype = C.POINTER(SceneDesc)
        H.tpt_host_scene_desc.argtypes = [C.c_void_p]
        H.tpt_host_scene_max_depth.argtypes = [C.c_void_p]
        H.tpt_host_scene_free.argtypes = [C.c_void_p]
        H.tpt_host_scene_free.restype = None
        H.tpt_host_make_camera.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                           C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                           C.POINTER(Camera)]
        H.tpt_host_make_camera.restype = None
        _host = H
    return _host


def _check(rc: int) -> None:
    if rc != 0:
        raise TptError(rc, lib().tpt_last_error().decode())


def device_count() -> int:
    return lib().tpt_device_count()


# ------------------------------------------------------------------------------ host front end
# the reference's hard-coded light-sampling list, main.cpp:99-106
REFERENCE_LIGHTS = [(LIGHT_XZ_RECT, (-100.0, 100.0, -150.0, -50.0, 298.0)), (LIGHT_SPHERE, (120.0, -50.0, 40.0, 120.0, 0.0))]


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def make_camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, t0=0.0, t1=0.0) -> Camera:
    """camera_with_blur ctor on the host (C++ front end, src/camera.cc:2-21 semantics)."""
    cam = Camera()
    host().tpt_host_make_camera(_f3(lookfrom), _f3(lookat), _f3(vup), vfov, aspect, aperture, focus_dist, t0, t1,
                                C.byref(cam))
    return cam


class HostScene:
    """A scene built with the C++ front-end classes (host/) and flattened to a tpt_scene_desc."""

    def __init__(self, name: str, image: np.ndarray | None = None, perlin: PerlinTables | None = None,
                 lights=None, background: int = BG_BLACK):
        img_p, w, h = None, 0, 0
        if image is not None:
            image = np.ascontiguousarray(image, dtype=np.uint8)
            h, w = image.shape[0], image.shape[1]
            img_p = image.ctypes.data
        self._keep = (image, perlin)
        larr, n = None, 0
        if lights is not None:
            n = len(lights)
            larr = (Light * max(n, 1))()
            for i, (kind, p) in enumerate(lights):
                larr[i].kind = kind
                for k, x in enumerate(p):
                    larr[i].p[k] = x
        self._h = host().tpt_host_build_scene(name.encode(), img_p, w, h,
                                              C.byref(perlin) if perlin is not None else None,
                                              larr, n, background)
        if not self._h:
            raise RuntimeError("host scene build failed: " + host().tpt_host_last_error().decode())
        self.name = name
        self.desc = host().tpt_host_scene_desc(self._h)
        self.max_depth = host().tpt_host_scene_max_depth(self._h)

    def close(self):
        if self._h:
            host().tpt_host_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------- GPU scene
@dataclass
class RenderResult:
    sum_rgb: np.ndarray | None      # [slices, ny, nx, 3] float32, row 0 = bottom row
    rgb8: np.ndarray | None         # [ny, nx, 3] uint8
    rgb8_slices: np.ndarray | None  # [slices, ny, nx, 3] uint8
    stats: dict


def make_params(nx, ny, ns, max_depth, mode=MODE_FAST, slices=1, seed=0x5EED, t_min=0.001, part_index=0,
                part_count=1, device=0, kernel=KERNEL_MEGA, subs=0, bundle_cull=True) -> RenderParams:
    p = RenderParams()
    p.nx, p.ny, p.ns, p.max_depth = nx, ny, ns, max_depth
    p.slices, p.mode, p.kernel = slices, mode, kernel
    p.seed_lo, p.seed_hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    p.t_min = t_min
    p.part_index, p.part_count, p.device = part_index, part_count, device
    p.reserved[1] = subs
    p.reserved[2] = 0 if bundle_cull else 1  # 1: trace every path, also those whose pixel cannot see the scene
    return p


class Scene:
    """Device-resident scene (tpt_scene*)."""

    def __init__(self, desc, device: int = 0):
        self._s = C.c_void_p()
        d = desc.desc if isinstance(desc, HostScene) else desc
        self._keep = desc
        _check(lib().tpt_scene_create(d, device, C.byref(self._s)))

    def close(self):
        if self._s:
            lib().tpt_scene_destroy(self._s)
            self._s = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def intersect(self, rays: np.ndarray, tmin: float = 0.001, tmax: float = 3.4028234663852886e38,
                  mode: int = MODE_PARITY) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 7)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        _check(lib().tpt_intersect_batch(self._s, rays.ctypes.data, rays.shape[0], tmin, tmax, mode,
                                         out.ctypes.data))
        return out

    def stats(self) -> dict:
        st = Stats()
        _check(lib().tpt_get_stats(self._s, C.byref(st)))
        arrays = ("reserved", "multi_batches", "multi_stolen", "multi_busy_ms")
        d = {k: getattr(st, k) for k, _ in Stats._fields_ if k not in arrays}
        d["reserved"] = [int(x) for x in st.reserved]
        n = int(st.multi_gpus)
        d["multi_batches"] = [int(x) for x in st.multi_batches][:n]
        d["multi_stolen"] = [int(x) for x in st.multi_stolen][:n]
        d["multi_busy_ms"] = [float(x) for x in st.multi_busy_ms][:n]
        return d

    def device_buffers(self):
        """(sum_ptr, sum_bytes, rgb8_ptr, rgb8_bytes): device addresses of the last render's products."""
        sp, rp = C.c_void_p(), C.c_void_p()
        sb, rb = C.c_size_t(), C.c_size_t()
        _check(lib().tpt_device_buffers(self._s, C.byref(sp), C.byref(sb), C.byref(rp), C.byref(rb)))
        return sp.value, sb.value, rp.value, rb.value

    def render_device(self, cam: Camera, params: RenderParams) -> dict:
        _check(lib().tpt_render_device(self._s, C.byref(cam), C.byref(params)))
        return self.stats()

    def _buffers(self, params, want_sum, want_rgb8, want_slices):
        sl = max(1, params.slices)
        img = Image()
        s = np.zeros((sl, params.ny, params.nx, 3), np.float32) if want_sum else None
        r = np.zeros((params.ny, params.nx, 3), np.uint8) if want_rgb8 else None
        rs = np.zeros((sl, params.ny, params.nx, 3), np.uint8) if want_slices else None
        if s is not None:
            img.sum_rgb = s.ctypes.data_as(C.POINTER(C.c_float))
        if r is not None:
            img.rgb8 = r.ctypes.data_as(C.POINTER(C.c_uint8))
        if rs is not None:
            img.rgb8_slices = rs.ctypes.data_as(C.POINTER(C.c_uint8))
        return img, s, r, rs

    def fetch(self, params: RenderParams, want_sum=True, want_rgb8=True, want_slices=False) -> RenderResult:
        img, s, r, rs = self._buffers(params, want_sum, want_rgb8, want_slices)
        _check(lib().tpt_render_fetch(self._s, C.byref(img)))
        return RenderResult(s, r, rs, self.stats())

    def render(self, cam: Camera, params: RenderParams, want_sum=True, want_rgb8=True,
               want_slices=False) -> RenderResult:
        """The reference-facing call: host buffers in, host buffers out."""
        img, s, r, rs = self._buffers(params, want_sum, want_rgb8, want_slices)
        _check(lib().tpt_render(self._s, C.byref(cam), C.byref(params), C.byref(img)))
        return RenderResult(s, r, rs, self.stats())

    def texture_value(self, texture: int, uvp: np.ndarray, mode: int = MODE_PARITY) -> np.ndarray:
        uvp = np.ascontiguousarray(uvp, dtype=np.float32).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3), np.float32)
        _check(lib().tpt_debug_texture(self._s, texture, uvp.ctypes.data, uvp.shape[0], mode, out.ctypes.data))
        return out


def render_multi(scenes, cam: Camera, params: RenderParams, want_sum=True, want_rgb8=True,
                 want_slices=False) -> RenderResult:
    """tpt_render_multi: one Scene per GPU, static tile split + work stealing + NVLink gather."""
    img, s, r, rs = scenes[0]._buffers(params, want_sum, want_rgb8, want_slices)
    arr = (C.c_void_p * len(scenes))(*[sc._s for sc in scenes])
    _check(lib().tpt_render_multi(arr, len(scenes), C.byref(cam), C.byref(params), C.byref(img)))
    return RenderResult(s, r, rs, scenes[0].stats())


def small_scene_summary(desc) -> dict:
    """tpt_debug_small_scene: the small-scene table tpt_scene_create would build (host only)."""
    d = desc.desc if isinstance(desc, HostScene) else desc
    out = (C.c_int32 * 64)()
    _check(lib().tpt_debug_small_scene(d, out))
    blocks = [{"faces": [int(out[8 + 8 * b + f]) for f in range(6)], "chain": int(out[8 + 8 * b + 6]),
               "open": bool(out[8 + 8 * b + 7])} for b in range(int(out[4]))]
    return {"enabled": bool(out[0]), "groups": int(out[1]), "rects": int(out[2]), "spheres": int(out[3]),
            "blocks": blocks}


def fp32_peak(device: int = 0) -> dict:
    """tpt_debug_fp32_peak: measured FP32 FMA throughput of the device (TFLOP/s)."""
    tf, ms = C.c_double(), C.c_double()
    _check(lib().tpt_debug_fp32_peak(device, C.byref(tf), C.byref(ms)))
    return {"tflops": tf.value, "ms": ms.value}


def philox(ctr, key, device: int = 0):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    _check(lib().tpt_debug_philox(device, c, k, o))
    return [int(x) for x in o]


# ------------------------------------------------------------------------- reference configs
def cornell_camera(nx, ny, fov=90.0, aperture=0.1, t0=0.0, t1=0.0) -> Camera:
    """main.cpp:87-91: lookfrom (0,0,800), lookat 0, focus 10."""
    return make_camera((0, 0, 800), (0, 0, 0), (0, 1, 0), fov, float(nx) / float(ny), aperture, 10.0, t0, t1)


def book_camera(nx, ny, fov=20.0, aperture=0.1, t0=0.0, t1=0.0) -> Camera:
    """the commented alternative main.cpp:82-84: lookfrom (13,2,3), focus = |lookfrom|."""
    d = float(np.float32(np.sqrt(np.float32(13 * 13 + 2 * 2 + 3 * 3))))
    return make_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), fov, float(nx) / float(ny), aperture, d, t0, t1)
