"""Generates the committed golden fixtures under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref, i.e. /root/reference compiled unmodified + oracle/ref_harness.cc).

Run in the build container (where /root/reference exists):   python tests/golden/make_golden.py
The fixtures are what the GPU tests compare against on the GPU box, where /root/reference does
not exist; the CPU tests re-derive them from oracle/_ref to pin the harness (and the C port).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_ref as O  # noqa: E402
import raygen  # noqa: E402

EARTHMAP = "/root/reference/resources/earthmap.jpg"


def small_earth():
    """earthmap.jpg decoded by the reference's own stb_image wrapper (src/utils.cc:236-240),
    decimated 4x so the fixture stays small (the texel addressing code is size independent)."""
    L = O.ref_lib(True)
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    p = L.ref_load_image(EARTHMAP.encode(), C.byref(w), C.byref(h), C.byref(ch))
    assert p and ch.value == 3
    full = np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    return full, np.ascontiguousarray(full[::4, ::4])


def oneweek_fixture(earth):
    """oneweek_final() (src/utils.cc:359-414) does stbi_load("earthmap.jpg") from the CWD: write the
    decimated picture as tests/golden/earthmap.jpg (a derived 256x128 file), let the reference
    decode THAT, and keep the decoded bytes for the product side."""
    from PIL import Image
    jpg = os.path.join(HERE, "earthmap.jpg")
    Image.fromarray(earth).save(jpg, quality=90, subsampling=0)  # 4:4:4 like the reference's earthmap.jpg
    L = O.ref_lib(True)
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    p = L.ref_load_image(jpg.encode(), C.byref(w), C.byref(h), C.byref(ch))
    assert p and ch.value == 3
    return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()


def main():
    full, earth = small_earth()
    np.savez_compressed(os.path.join(HERE, "earth_small.npz"), rgb=earth, full_shape=np.array(full.shape),
                        full_checksum=np.array([int(full.astype(np.uint64).sum())]))
    earth_jpg = oneweek_fixture(earth)
    np.savez_compressed(os.path.join(HERE, "earth_jpg_decoded.npz"), rgb=earth_jpg)
    # ---- gate 1: hit records ------------------------------------------------------------
    for scene, (n_cam, n_int) in {"cornell_box": (1500, 1500), "sphere_cornell_box": (600, 600),
                                  "random_scene": (1500, 1500), "random_scene_list": (400, 400),
                                  "two_perlin_spheres": (300, 300), "light_spheres": (300, 300),
                                  "earth": (300, 300), "textured_lit": (400, 400)}.items():
        img = earth if scene in ("earth", "textured_lit") else None
        rs = O.RefScene(scene, image=img)
        rays = raygen.primary_batch(scene, n_cam, n_int, seed=7)
        h1 = rs.hit_batch(rays)
        sec = raygen.secondary_rays(h1, np.random.default_rng(11))[:1500]
        rays = np.concatenate([rays, sec], axis=0)
        hits = rs.hit_batch(rays)
        np.savez_compressed(os.path.join(HERE, f"hits_{scene}.npz"), rays=rays, hits=hits,
                            n_leaves=np.array([rs.n_leaves]))
        print(scene, "rays", len(rays), "hit fraction", hits["hit"].mean())
    # ---- gate 2: radiance under the injected Philox stream ---------------------------------
    cases = {
        "cornell_A": dict(scene="cornell_box", cam=O.CORNELL_CAM, nx=48, ny=48, ns=8, depth=15, seed=2024),
        "cornell_B": dict(scene="cornell_box", cam=dict(O.CORNELL_CAM, vfov=61.93), nx=32, ny=32, ns=8, depth=50, seed=77),
        "cornell_slices": dict(scene="cornell_box", cam=O.CORNELL_CAM, nx=24, ny=24, ns=12, depth=15, seed=5, slices=4),
        "sphere_cornell": dict(scene="sphere_cornell_box", cam=O.CORNELL_CAM, nx=32, ny=32, ns=8, depth=15, seed=9),
        "random_scene": dict(scene="random_scene", cam=dict(O.BOOK_CAM, t0=0.0, t1=1.0), nx=48, ny=32, ns=4, depth=15, seed=3),
        "light_spheres": dict(scene="light_spheres", cam=dict(O.BOOK_CAM, vfov=40.0), nx=40, ny=40, ns=8, depth=15, seed=4),
        "textured_lit": dict(scene="textured_lit", cam=dict(O.BOOK_CAM, vfov=50.0), nx=40, ny=40, ns=8, depth=15, seed=6,
                             lights=[(0, (-2.0, 2.0, -2.0, 2.0, 7.0)), (1, (-3.0, 6.0, 4.0, 2.0, 0.0))]),
        "cornell_smoke": dict(scene="cornell_box_smoke", cam=dict(O.CORNELL_CAM, vfov=61.93), nx=40, ny=40, ns=8, depth=15, seed=12),
        "oneweek_final": dict(scene="oneweek_final", cam=dict(lookfrom=(478, 278, -600), lookat=(278, 278, 0), vup=(0, 1, 0), vfov=40.0,
                              aperture=0.0, focus_dist=10.0, t0=0.0, t1=1.0), nx=36, ny=36, ns=4, depth=10, seed=21),
    }
    for name, c in cases.items():
        img = earth if c["scene"] in ("earth", "textured_lit") else None
        cwd = os.getcwd()
        os.chdir(HERE)  # oneweek_final() loads ./earthmap.jpg
        rs = O.RefScene(c["scene"], image=img)
        os.chdir(cwd)
        rv, px, py, pz = rs.perlin_tables()
        out, samples, st = rs.render(c["cam"], c["nx"], c["ny"], c["ns"], c["depth"], slices=c.get("slices", 1),
                                     lights=c.get("lights", O.REFERENCE_LIGHTS), seed=c["seed"], per_sample=True)
        np.savez_compressed(os.path.join(HERE, f"render_{name}.npz"), sum_rgb=out, samples=samples,
                            rays=np.array([st["rays"]]), draws=np.array([st["draws"]]),
                            ranvec=rv, perm_x=px, perm_y=py, perm_z=pz)
        print(name, "mean", out[-1].mean(), "rays/path", st["rays"] / st["paths"], "nonzero", (out[-1] > 0).mean())
    # ---- textures ---------------------------------------------------------------------------
    rs = O.RefScene("two_perlin_spheres")
    rv, px, py, pz = rs.perlin_tables()
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-20, 20, (500, 3)), rng.uniform(-1000, 1000, (200, 3)),
                          np.array([[0, 0, 0], [1, 2, 3], [-0.5, 255.5, 256.0], [-1e-3, 1e-3, 0.999]])]).astype(np.float32)
    turb = np.zeros((len(pts), 3), np.float32)
    rs.lib.ref_perlin_turb(pts.ctypes.data, len(pts), C.c_float(2.0), turb.ctypes.data)
    chk = np.zeros((len(pts), 3), np.float32)
    rs.lib.ref_checker_value(pts.ctypes.data, len(pts), chk.ctypes.data)
    uv = np.concatenate([rng.uniform(-0.1, 1.1, (500, 2)), np.array([[0, 0], [1, 1], [0, 1], [1, 0], [0.5, 0.5]])]).astype(np.float32)
    imv = np.zeros((len(uv), 3), np.float32)
    rs.lib.ref_image_value(earth.ctypes.data, earth.shape[1], earth.shape[0], uv.ctypes.data, len(uv), imv.ctypes.data)
    np.savez_compressed(os.path.join(HERE, "textures.npz"), pts=pts, turb_scale2=turb, checker=chk, uv=uv, image=imv,
                        ranvec=rv, perm_x=px, perm_y=py, perm_z=pz)
    # ---- camera -----------------------------------------------------------------------------
    cams = [((0, 0, 800), (0, 0, 0), (0, 1, 0), 90.0, 1.0, 0.1, 10.0, 0.0, 0.0),
            ((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.0, 13.49, 0.0, 1.0),
            ((278, 278, -800), (278, 278, 0), (0, 1, 0), 61.93, 2.0, 0.2, 7.5, 0.25, 0.75)]
    rows = []
    for a in cams:
        c = O.ref_camera(*a)
        rows.append(np.frombuffer(bytes(c), np.float32).copy())
    flat_args = np.array([[*a[0], *a[1], *a[2], *a[3:]] for a in cams], np.float64)
    np.savez_compressed(os.path.join(HERE, "cameras.npz"), args=flat_args, fields=np.array(rows))
    # ---- Philox known answers (Random123 kat_vectors, philox4x32-10) -------------------------
    kat = np.array([[0, 0, 0, 0, 0, 0, 0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8],
                    [0xffffffff] * 6 + [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
                    [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0,
                     0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "philox_kat.npz"), kat=kat)


if __name__ == "__main__":
    main()
