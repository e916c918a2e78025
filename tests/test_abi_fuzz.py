"""The drop-in boundary takes a caller-built description (include/tpt.h, tpt_scene_desc): whatever the caller
hands over, the library must answer with an error code, never with a crash. Valid descriptions produced by
the host flattener are mutated word by word (indices, counts, kinds, ends, NaNs) and pushed through the
host-only part of tpt_scene_create -- validate_desc + the small-scene / block folding (tpt_debug_small_scene);
no device is needed, so this runs in the CPU suite. A segfault here kills the test process: that is the check."""
import ctypes as C
import zlib

import numpy as np
import pytest

import common

ARRAYS = [("nodes", "n_nodes"), ("prims", "n_prims"), ("chains", "n_chains"), ("xform_ops", "n_xform_ops"),
          ("materials", "n_materials"), ("textures", "n_textures"), ("lights", "n_lights")]
COUNTS = ["n_nodes", "n_prims", "n_chains", "n_xform_ops", "n_materials", "n_textures", "n_images", "n_lights",
          "n_root_nodes", "background", "api_version"]


def _clone(T, src):
    """deep copy of a description into buffers this test owns (one spare element behind every array)"""
    d = T.SceneDesc()
    C.memmove(C.byref(d), C.byref(src), C.sizeof(T.SceneDesc))
    keep = {}
    for name, count in ARRAYS:
        n = getattr(src, count)
        ptr = getattr(src, name)
        if n <= 0 or not ptr:
            continue
        elem = type(ptr.contents)
        buf = (elem * (n + 1))()
        C.memmove(buf, ptr, n * C.sizeof(elem))
        keep[name] = buf
        setattr(d, name, C.cast(buf, C.POINTER(elem)))
    return d, keep


def perturb_geometry(rng, keep, n_mut):
    """benign mutations: scale / shift single floats of primitives, materials, transform ops"""
    muts = []
    for _ in range(n_mut):
        aname = str(rng.choice([a for a in ("prims", "prims", "prims", "materials", "xform_ops") if a in keep]))
        buf = keep[aname]
        rec = C.sizeof(buf._type_) // 4
        f = np.frombuffer(buf, dtype=np.float32)
        n_rec = len(buf) - 1
        r = int(rng.integers(n_rec))
        if aname == "prims":
            w = int(rng.integers(4, 13))  # p[0..8] live in words 4..12 of tpt_prim (kind, material, chain, flags first)
            if int(np.frombuffer(buf, dtype=np.int32)[r * rec]) == 5 and w in (5, 6):
                continue  # a medium's boundary range: node indices, not floats
        elif aname == "materials":
            w = int(rng.integers(2, 7))  # albedo, fuzz, ref_idx
        else:
            w = int(rng.integers(1, 4))
        i = r * rec + w
        old = float(f[i])
        if not np.isfinite(old) or abs(old) > 1e6:
            continue
        mode = rng.integers(0, 3)
        if mode == 0:
            f[i] = np.float32(old * rng.uniform(0.5, 1.5))
        elif mode == 1:
            f[i] = np.float32(old + rng.normal(0, 0.05 * max(1.0, abs(old))))
        else:  # snap onto another record's value: coincident planes / centres, the tie cases
            f[i] = f[int(rng.integers(n_rec)) * rec + w]
        muts.append((aname, r, w, old, float(f[i])))
    return muts



INTERESTING = [-1, 0, 1, 2, 3, 7, 8, 31, 32, 33, 48, 49, 255, 256, 0x100, 0x10000, 0x7fffffff, -0x80000000]


@pytest.mark.parametrize("scene", ["cornell_box", "sphere_cornell_box", "random_scene", "cornell_box_smoke", "oneweek_final", "textured_lit"])
def test_mutated_descriptions_are_answered_not_crashed_on(T, scene):
    hs = common.host_scene(T, scene, perlin=common.perlin_struct(T, common.golden("textures")),
                           lights=common.TEXTURED_LIGHTS if scene == "textured_lit" else None)
    lib = T.lib()
    lib.tpt_debug_small_scene.argtypes = [C.POINTER(T.SceneDesc), C.POINTER(C.c_int32)]
    lib.tpt_debug_small_scene.restype = C.c_int
    out = (C.c_int32 * 64)()
    src = hs.desc.contents if hasattr(hs.desc, "contents") else hs.desc
    d0, keep0 = _clone(T, src)
    assert lib.tpt_debug_small_scene(C.byref(d0), out) == 0  # the clone itself is valid
    rng = np.random.default_rng(zlib.crc32(scene.encode()))  # (not hash(): that one changes from process to process)
    refused = accepted = 0
    for trial in range(1500):
        d, keep = _clone(T, src)
        for _ in range(int(rng.integers(1, 4))):
            what = rng.integers(0, 10)
            if what == 0:  # a count / header word
                name = COUNTS[int(rng.integers(len(COUNTS)))]
                cur = getattr(d, name)
                setattr(d, name, int(rng.choice([cur - 1, cur + 1, cur // 2, 0, -1] + INTERESTING)))
                # a count may not run past the buffer this test allocated (the caller's promise, not the library's)
                for aname, cname in ARRAYS:
                    if cname == name and aname in keep:
                        setattr(d, name, min(getattr(d, name), len(keep[aname])))
            elif what == 1 and keep:  # a null array pointer
                aname = list(keep)[int(rng.integers(len(keep)))]
                setattr(d, aname, C.cast(None, type(getattr(d, aname))))
            elif keep:  # one 32-bit word of one record
                aname = list(keep)[int(rng.integers(len(keep)))]
                buf = keep[aname]
                words = np.frombuffer(buf, dtype=np.uint32)
                n_words = (len(buf) - 1) * C.sizeof(buf._type_) // 4
                i = int(rng.integers(n_words))
                mode = rng.integers(0, 4)
                if mode == 0:
                    words[i] = int(rng.choice(INTERESTING)) & 0xFFFFFFFF
                elif mode == 1:
                    words[i] = (int(words[i]) ^ (1 << int(rng.integers(32)))) & 0xFFFFFFFF
                elif mode == 2:
                    words[i] = np.float32(rng.choice([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e38, -1e38, 1e-45])).view(np.uint32)
                else:
                    words[i] = int(rng.integers(-4, 600)) & 0xFFFFFFFF
        rc = lib.tpt_debug_small_scene(C.byref(d), out)
        assert rc in (0, -1, -2, -3, -4, -5, -6), rc
        refused += rc != 0
        accepted += rc == 0
    assert refused > 50 and accepted > 50, (refused, accepted)  # the mutations reach both outcomes
