"""Shared helpers for the parity tests."""
import ctypes as C
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FLT_MAX = 3.4028234663852886e38

# the gate-2 cases of tests/golden/make_golden.py (kept in sync by test_golden.py)
BOOK_FOCUS = float(np.sqrt(np.float32(182.0)))
CORNELL_CAM = dict(lookfrom=(0, 0, 800), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=90.0, aperture=0.1, focus_dist=10.0)
BOOK_CAM = dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=20.0, aperture=0.1, focus_dist=BOOK_FOCUS)
TEXTURED_LIGHTS = [(0, (-2.0, 2.0, -2.0, 2.0, 7.0)), (1, (-3.0, 6.0, 4.0, 2.0, 0.0))]
RENDER_CASES = {
    "cornell_A": dict(scene="cornell_box", cam=CORNELL_CAM, nx=48, ny=48, ns=8, depth=15, seed=2024),
    "cornell_B": dict(scene="cornell_box", cam=dict(CORNELL_CAM, vfov=61.93), nx=32, ny=32, ns=8, depth=50, seed=77),
    "cornell_slices": dict(scene="cornell_box", cam=CORNELL_CAM, nx=24, ny=24, ns=12, depth=15, seed=5, slices=4),
    "sphere_cornell": dict(scene="sphere_cornell_box", cam=CORNELL_CAM, nx=32, ny=32, ns=8, depth=15, seed=9),
    "random_scene": dict(scene="random_scene", cam=dict(BOOK_CAM, t0=0.0, t1=1.0), nx=48, ny=32, ns=4, depth=15, seed=3),
    "light_spheres": dict(scene="light_spheres", cam=dict(BOOK_CAM, vfov=40.0), nx=40, ny=40, ns=8, depth=15, seed=4),
    "textured_lit": dict(scene="textured_lit", cam=dict(BOOK_CAM, vfov=50.0), nx=40, ny=40, ns=8, depth=15, seed=6,
                         lights=TEXTURED_LIGHTS),
    "cornell_smoke": dict(scene="cornell_box_smoke", cam=dict(CORNELL_CAM, vfov=61.93), nx=40, ny=40, ns=8, depth=15, seed=12),
    "oneweek_final": dict(scene="oneweek_final", cam=dict(lookfrom=(478, 278, -600), lookat=(278, 278, 0), vup=(0, 1, 0), vfov=40.0,
                          aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0), nx=36, ny=36, ns=4, depth=10, seed=21),
}
HIT_SCENES = ["cornell_box", "sphere_cornell_box", "random_scene", "random_scene_list", "two_perlin_spheres",
              "light_spheres", "earth", "textured_lit"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def earth_small():
    return golden("earth_small")["rgb"]


def perlin_struct(T, g):
    """tpt_perlin_tables from a golden file / reference table dump."""
    pt = T.PerlinTables()
    rv = np.ascontiguousarray(g["ranvec"], np.float32)
    C.memmove(pt.ranvec, rv.ctypes.data, rv.nbytes)
    for name, dst in (("perm_x", pt.perm_x), ("perm_y", pt.perm_y), ("perm_z", pt.perm_z)):
        a = np.ascontiguousarray(g[name], np.int32)
        C.memmove(dst, a.ctypes.data, a.nbytes)
    return pt


def earth_jpg_decoded():
    """tests/golden/earthmap.jpg as the reference's stb_image decodes it (what oneweek_final() loads)"""
    return golden("earth_jpg_decoded")["rgb"]


def host_scene(T, scene, perlin=None, lights=None, background=0):
    img = earth_small() if scene in ("earth", "textured_lit") else None
    if scene == "oneweek_final":
        img = earth_jpg_decoded()
    return T.HostScene(scene, image=img, perlin=perlin, lights=lights, background=background)


def product_camera(T, cam, nx, ny):
    return T.make_camera(cam["lookfrom"], cam["lookat"], cam.get("vup", (0, 1, 0)), cam["vfov"],
                         cam.get("aspect", float(nx) / float(ny)), cam["aperture"], cam["focus_dist"],
                         cam.get("t0", 0.0), cam.get("t1", 0.0))


def rel_err(got, ref, floor):
    """|got-ref| / max(|ref|, floor), elementwise"""
    got = got.astype(np.float64)
    ref = ref.astype(np.float64)
    return np.abs(got - ref) / np.maximum(np.abs(ref), floor)


def same_float(a, b):
    """bitwise-equal floats, with NaN == NaN"""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
