"""Host C++ front end: scene builders + flattener describe exactly the reference's scenes,
the camera ctor reproduces the reference's fields bit for bit, the ini reader and PPM writers
behave like the reference's I/O."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

SCENES = ["cornell_box", "sphere_cornell_box", "random_scene", "random_scene_list", "two_perlin_spheres",
          "light_spheres"]


class RefLeaf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mat_kind", C.c_int32), ("mat_id", C.c_int32), ("tex_kind", C.c_int32),
                ("p", C.c_float * 12), ("mat", C.c_float * 5)]


def ref_leaves(O, name):
    rs = O.RefScene(name)
    arr = (RefLeaf * 4096)()
    rs.lib.ref_dump_leaves.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    n = rs.lib.ref_dump_leaves(rs.h, arr, 4096)
    return rs, [arr[i] for i in range(n)]


@pytest.mark.parametrize("name", SCENES)
def test_flattened_scene_equals_reference_scene(T, O, name):
    """Same leaves, same order, same parameters (bitwise), same material assignment."""
    hs = T.HostScene(name)
    d = hs.desc.contents
    rs, leaves = ref_leaves(O, name)
    assert d.n_prims == len(leaves) == rs.n_leaves
    assert d.n_materials == rs.n_materials
    for i, rl in enumerate(leaves):
        p = d.prims[i]
        assert p.kind == rl.kind, (name, i)
        assert p.material == rl.mat_id, (name, i)
        n_par = {0: 4, 1: 9, 2: 5, 3: 5, 4: 5}[rl.kind]
        got = np.array(p.p[:n_par], np.float32)
        exp = np.array(rl.p[:n_par], np.float32)
        assert got.tobytes() == exp.tobytes(), (name, i, got, exp)
        m = d.materials[p.material]
        assert m.kind == rl.mat_kind, (name, i)
        if rl.mat_kind == 1:
            assert np.array(list(m.albedo) + [m.fuzz], np.float32).tobytes() == np.array(rl.mat[:4], np.float32).tobytes()
        if rl.mat_kind == 2:
            assert np.float32(m.ref_idx) == np.float32(rl.mat[0])
        if rl.mat_kind in (0, 3):
            t = d.textures[m.texture]
            assert t.kind == rl.tex_kind
            if t.kind == 0:
                assert np.array(t.color[:], np.float32).tobytes() == np.array(rl.mat[:3], np.float32).tobytes()
            if t.kind == 2:
                assert np.float32(t.scale) == np.float32(rl.mat[0])


@pytest.mark.parametrize("name", SCENES)
def test_preorder_structure(T, name):
    hs = T.HostScene(name)
    d = hs.desc.contents
    ends = []
    leaf_prims = []
    skip_until = -1  # inside a duplicated sub-tree (second child of a one-element bvh_node)
    for i in range(d.n_nodes):
        while ends and ends[-1] == i:
            ends.pop()
        n = d.nodes[i]
        k = n.kind & 0xFF
        assert 0 <= (n.kind >> 16) < d.n_chains
        if k == 2:
            assert 0 <= n.end_or_prim < d.n_prims
            if not (n.kind & 0x100) and i >= skip_until:
                leaf_prims.append(n.end_or_prim)
        else:
            assert i < n.end_or_prim <= (ends[-1] if ends else d.n_nodes)
            ends.append(n.end_or_prim)
            if (n.kind & 0x100) and i >= skip_until:
                skip_until = n.end_or_prim
    # every primitive appears exactly once among the non-duplicate leaves, in DFS order
    assert leaf_prims == list(range(d.n_prims))
    assert d.chains[0].n_ops == 0
    assert d.n_lights == 2  # main.cpp:99-106


def test_cornell_box_layout(T):
    hs = T.HostScene("cornell_box")
    d = hs.desc.contents
    assert (d.n_prims, d.n_chains, d.n_xform_ops, d.n_materials) == (19, 3, 4, 6)
    # two boxes: translate(rotate_y(box)) -> chain = [TRANSLATE, ROTATE_Y]
    for c in (1, 2):
        ch = d.chains[c]
        assert ch.n_ops == 2
        assert d.xform_ops[ch.first_op].kind == 0 and d.xform_ops[ch.first_op + 1].kind == 1
    flips = [d.prims[i].flags & 1 for i in range(d.n_prims)]
    assert sum(flips) == 3 + 3 + 3  # right wall, ceiling, lamp + three faces per box


def test_constant_medium_flattening(T):
    """cornell_box_smoke (src/utils.cc:321-357): three constant_medium objects become MEDIUM
    primitives in the root list; their boundaries (two transformed boxes, one sphere) are kept
    behind the root tree and are reachable only through them."""
    import ctypes as C
    hs = T.HostScene("cornell_box_smoke")
    d = hs.desc.contents
    assert d.n_root_nodes == 10 and d.n_nodes == 25
    med = [i for i in range(d.n_prims) if d.prims[i].kind == 5]
    assert len(med) == 3
    dens = [np.float32(d.prims[i].p[0]) for i in med]
    assert dens == [np.float32(0.05), np.float32(0.01), np.float32(0.0001)]
    covered = []
    for i in med:
        first, end = (C.c_int32 * 2).from_buffer_copy(bytes(d.prims[i].p)[4:12])
        assert d.n_root_nodes <= first < end <= d.n_nodes
        covered += list(range(first, end))
        assert d.materials[d.prims[i].material].kind == 5  # isotropic (an absorber at HEAD)
    assert covered == list(range(d.n_root_nodes, d.n_nodes))
    root_leaves = [d.nodes[i].end_or_prim for i in range(d.n_root_nodes) if (d.nodes[i].kind & 0xFF) == 2]
    assert set(med) <= set(root_leaves)
    for i in range(d.n_root_nodes, d.n_nodes):  # boundary primitives never appear in the root tree
        if (d.nodes[i].kind & 0xFF) == 2:
            assert d.nodes[i].end_or_prim not in root_leaves


@pytest.mark.parametrize("args", [
    dict(lookfrom=(0, 0, 800), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=90.0, aspect=1.0, aperture=0.1, focus=10.0, t0=0.0, t1=0.0),
    dict(lookfrom=(13, 2, 3), lookat=(0, 0, 0), vup=(0, 1, 0), vfov=20.0, aspect=1.5, aperture=0.0, focus=13.49, t0=0.0, t1=1.0),
    dict(lookfrom=(278, 278, -800), lookat=(278, 278, 0), vup=(0, 1, 0), vfov=61.93, aspect=2.0, aperture=0.2, focus=7.5, t0=0.25, t1=0.75),
])
def test_camera_fields_bit_exact(T, O, args):
    cam = T.make_camera(args["lookfrom"], args["lookat"], args["vup"], args["vfov"], args["aspect"],
                        args["aperture"], args["focus"], args["t0"], args["t1"])
    ref = O.ref_camera(args["lookfrom"], args["lookat"], args["vup"], args["vfov"], args["aspect"],
                       args["aperture"], args["focus"], args["t0"], args["t1"])
    pairs = [("origin", "origin"), ("lower_left_corner", "lower_left"), ("vertical", "vertical"),
             ("horizontal", "horizontal"), ("u", "u"), ("v", "v"), ("w", "w")]
    for a, b in pairs:
        assert bytes(getattr(cam, a)) == bytes(getattr(ref, b)), a
    for f in ("lens_radius", "time0", "time1"):
        assert np.float32(getattr(cam, f)) == np.float32(getattr(ref, f))


def test_ppm_writers_match_reference_formats(T, tmp_path):
    """main picture: one pixel per line 'r g b \\n' top row first; bonus: one long line."""
    nx, ny = 3, 2
    img = ((np.arange(nx * ny * 3).reshape(ny, nx, 3) * 13) % 256).astype(np.uint8)
    H = T.host()
    H.tpt_host_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    p1, p2 = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
    assert H.tpt_host_write_ppm(p1.encode(), img.ctypes.data, nx, ny, 0) == 0
    assert H.tpt_host_write_ppm(p2.encode(), img.ctypes.data, nx, ny, 1) == 0
    rows_top_first = img[::-1]
    exp1 = "P3\n3 2\n255\n" + "".join(f"{p[0]} {p[1]} {p[2]} \n" for r in rows_top_first for p in r)
    exp2 = "P3\n3 2\n255\n" + "".join(f"{p[0]} {p[1]} {p[2]} " for r in rows_top_first for p in r)
    assert open(p1).read() == exp1
    assert open(p2).read() == exp2


INI_PROBE = r'''
#include "inipp.h"
#include <iostream>
#include <sstream>
int main() {
  std::istringstream in("[DEFAULT]\nwidth = 400\n height=  400 \nsample = 50\nfov = 90.0\n; comment\nbad line\n"
                        "width = 7\n[BLUR]\naperture = 0.1\n[CAM_MOTION]\nstart_time = 0.0\nend_time = x\n");
  inipp::Ini<char> ini; ini.parse(in);
  int nx = 1, ny = 2, ns = 3, depth = 50; float fov = 1, ap = 0.2f, t1 = 1.0f;
  inipp::extract(ini.sections["DEFAULT"]["width"], nx);
  inipp::extract(ini.sections["DEFAULT"]["height"], ny);
  inipp::extract(ini.sections["DEFAULT"]["sample"], ns);
  inipp::extract(ini.sections["DEFAULT"]["recur_depth"], depth);   // missing -> default kept
  inipp::extract(ini.sections["DEFAULT"]["fov"], fov);
  inipp::extract(ini.sections["BLUR"]["aperture"], ap);
  inipp::extract(ini.sections["CAM_MOTION"]["end_time"], t1);      // unparsable -> default kept
  std::cout << nx << " " << ny << " " << ns << " " << depth << " " << fov << " " << ap << " " << t1 << " "
            << ini.errors.size() << "\n";
  ini.generate(std::cout);
}
'''


def test_ini_reader_semantics(T, tmp_path):
    """config.ini goes through inipp itself (third_party/inipp.h, vendored verbatim): the behaviours
    main() relies on -- defaults kept for missing / unparsable values, bad lines collected, generate()
    echoing the sections sorted (main.cpp:42-58)."""
    src = tmp_path / "probe.cc"
    src.write_text(INI_PROBE)
    exe = tmp_path / "probe"
    subprocess.check_call(["g++", "-std=c++14", "-I", os.path.join(T.REPO_ROOT, "tiny-path-tracer_b200", "third_party"),
                           str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).splitlines()
    assert out[0] == "400 400 50 50 90 0.1 1 2"  # two bad lines: 'bad line' and the duplicate width
    echoed = [l for l in out[1:] if l]
    assert echoed == ["[BLUR]", "aperture=0.1", "[CAM_MOTION]", "end_time=x", "start_time=0.0", "[DEFAULT]",
                      "fov=90.0", "height=400", "recur_depth=", "sample=50", "width=400"]  # operator[] inserted recur_depth
    ref = "/root/reference/third_party/inipp.h"
    if os.path.exists(ref):  # verbatim, as third_party/README.md says
        for name in ("inipp.h", "stb_image.h"):
            mine = open(os.path.join(T.REPO_ROOT, "tiny-path-tracer_b200", "third_party", name), "rb").read()
            assert mine == open(os.path.join(os.path.dirname(ref), name), "rb").read()


def test_builtin_jpeg_decoder_matches_stb(T, O):
    """load_image_texture is the reference's wrapper around stbi_load (third_party/stb_image.h, vendored
    verbatim). The bytes must be the ones the reference yields (they are what the GPU samples): checked on
    the derived fixture tests/golden/earthmap.jpg against the decode stored in the golden set, and, where
    the reference tree is present, on the real resources/earthmap.jpg against the reference's build live."""
    import common
    H = T.host()
    H.tpt_host_load_image.restype = C.POINTER(C.c_uint8)
    H.tpt_host_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    H.tpt_host_free.argtypes = [C.c_void_p]

    def mine(path):
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        p = H.tpt_host_load_image(path.encode(), C.byref(w), C.byref(h), C.byref(ch))
        assert p and ch.value == 3
        a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
        H.tpt_host_free(p)
        return a

    assert np.array_equal(mine(os.path.join(common.GOLDEN, "earthmap.jpg")), common.earth_jpg_decoded())
    real = "/root/reference/resources/earthmap.jpg"
    if os.path.exists(real):
        L = O.ref_lib(True)
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        q = L.ref_load_image(real.encode(), C.byref(w), C.byref(h), C.byref(ch))
        ref = np.ctypeslib.as_array(q, shape=(h.value, w.value, 3)).copy()
        assert ref.shape == (512, 1024, 3) and np.array_equal(mine(real), ref)


def _smooth_picture(nx, ny, seed):
    """a picture with gradients, edges and mild noise (bottom-up, like the library's rgb8)"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:ny, 0:nx].astype(np.float32)
    img = np.stack([128 + 100 * np.sin(x / 9.0 + seed), 128 + 100 * np.cos(y / 7.0), 255.0 * x / max(nx - 1, 1)], axis=-1)
    img[ny // 4: ny // 2, nx // 3: nx // 2] = (250, 20, 20)
    img += rng.normal(0, 2.0, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def test_contact_sheet_jpeg_without_imagemagick(T, tmp_path):
    """the `convert a.ppm b.ppm ... +append img.jpg` stage (main.cpp:224-245) done by
    the front end itself. The file must be a baseline JPEG any decoder reads (PIL here, and the
    vendored stb_image), the pictures left to right with the top row first, at quality-92
    fidelity; odd sizes exercise the partial edge blocks."""
    from PIL import Image
    H = T.host()
    nx, ny = 53, 37
    pics = [_smooth_picture(nx, ny, s) for s in (1, 2, 3)]
    arr = (C.c_void_p * 3)(*[p.ctypes.data for p in pics])
    H.tpt_host_write_contact_sheet.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    path = str(tmp_path / "img.jpg")
    assert H.tpt_host_write_contact_sheet(path.encode(), arr, 3, nx, ny, 92) == 0
    want = np.concatenate([p[::-1] for p in pics], axis=1)  # top row first, side by side
    im = Image.open(path)
    assert im.format == "JPEG" and im.size == (3 * nx, ny) and im.mode == "RGB"
    got = np.asarray(im)
    assert _psnr(got, want) > 34.0, _psnr(got, want)
    # each panel is where `+append` puts it
    for i, p in enumerate(pics):
        assert _psnr(got[:, i * nx:(i + 1) * nx], p[::-1]) > 33.0
    # stb_image (the decoder that reads earthmap.jpg) agrees with PIL to rounding
    H.tpt_host_load_image.restype = C.POINTER(C.c_uint8)
    H.tpt_host_load_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    H.tpt_host_free.argtypes = [C.c_void_p]
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    q = H.tpt_host_load_image(path.encode(), C.byref(w), C.byref(h), C.byref(ch))
    assert q and (w.value, h.value, ch.value) == (3 * nx, ny, 3)
    mine = np.ctypeslib.as_array(q, shape=(ny, 3 * nx, 3)).copy()
    H.tpt_host_free(q)
    assert np.abs(mine.astype(int) - got.astype(int)).max() <= 2
    # quality is the IJG scale: lower quality -> smaller file, lower fidelity
    small = str(tmp_path / "small.jpg")
    assert H.tpt_host_write_contact_sheet(small.encode(), arr, 3, nx, ny, 40) == 0
    assert os.path.getsize(small) < os.path.getsize(path)
    assert 24.0 < _psnr(np.asarray(Image.open(small)), want) < _psnr(got, want)
    # a flat picture (every AC coefficient zero) and a 1x1 picture are valid files too
    flat = np.full((8, 8, 3), 77, np.uint8)
    one = str(tmp_path / "flat.jpg")
    assert H.tpt_host_write_contact_sheet(one.encode(), (C.c_void_p * 1)(flat.ctypes.data), 1, 8, 8, 92) == 0
    assert np.abs(np.asarray(Image.open(one)).astype(int) - 77).max() <= 1
    dot = np.array([[[10, 200, 90]]], np.uint8)
    assert H.tpt_host_write_contact_sheet(one.encode(), (C.c_void_p * 1)(dot.ctypes.data), 1, 1, 1, 92) == 0
    assert np.abs(np.asarray(Image.open(one)).astype(int)[0, 0] - dot[0, 0]).max() <= 3


def test_binary_ppm_writer(T, tmp_path):
    from PIL import Image
    H = T.host()
    nx, ny = 5, 4
    img = _smooth_picture(nx, ny, 4)
    H.tpt_host_write_ppm_binary.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    p = str(tmp_path / "a.ppm")
    assert H.tpt_host_write_ppm_binary(p.encode(), img.ctypes.data, nx, ny) == 0
    raw = open(p, "rb").read()
    assert raw.startswith(b"P6\n5 4\n255\n") and len(raw) == 11 + nx * ny * 3
    assert np.array_equal(np.asarray(Image.open(p)), img[::-1])


def test_light_list_derived_from_the_scene(T):
    """emissive primitives the light-sampling list can represent. cornell_box has one
    lamp, flip_normal(xz_rect(-100,100,-150,-50,298)) -- the first of the two shapes main.cpp:99-106
    hard-codes (the second, the r=120 sphere around the glass ball, is not a lamp); light_spheres
    has two lit spheres and a lit xy_rect (not representable: hitable::random is only overridden for
    xz_rect and sphere); random_scene has none."""
    H = T.host()
    H.tpt_host_derive_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    out = (T.Light * 8)()

    def lights(name):
        hs = T.HostScene(name)
        n = H.tpt_host_derive_lights(hs._h, out, 8)
        return [(out[i].kind, [round(float(v), 3) for v in out[i].p]) for i in range(n)]

    assert lights("cornell_box") == [(0, [-100.0, 100.0, -150.0, -50.0, 298.0])]  # TPT_LIGHT_XZ_RECT
    ls = lights("light_spheres")
    assert ls == [(1, [-3.0, 1.0, 2.0, 1.0, 0.0]), (1, [-3.0, 1.0, -2.0, 1.0, 0.0])]  # the two emissive spheres only
    assert lights("random_scene") == []


def test_list_boxes_cover_moving_children(T):
    """A hitable_list's node box is the union of its children's node boxes: a moving_sphere leaf spans its
    own [time0, time1] (hitable_list::bounding_box(0, 0) would only cover its t = 0 position). The FAST
    walk prunes LIST nodes by that box and the scene-bounds tests read the root's."""
    hs = T.HostScene("moving_list_test")
    d = hs.desc.contents
    root = d.nodes[0]
    assert (root.kind & 0xff) == 1 and root.end_or_prim == d.n_nodes
    lo = np.min([list(d.nodes[i].bmin) for i in range(1, d.n_nodes)], axis=0)
    hi = np.max([list(d.nodes[i].bmax) for i in range(1, d.n_nodes)], axis=0)
    assert np.array_equal(np.array(list(root.bmin), np.float32), lo.astype(np.float32))
    assert np.array_equal(np.array(list(root.bmax), np.float32), hi.astype(np.float32))
    assert root.bmax[0] >= 7.0 and root.bmax[1] >= 4.0  # the sphere's t = 1 position (6, 3, 0), radius 1
    # a list whose children live in different transform spaces cannot be united: unbounded, never pruned
    hs2 = T.HostScene("cornell_box")
    d2 = hs2.desc.contents
    for i in range(d2.n_nodes):
        n = d2.nodes[i]
        if (n.kind & 0xff) == 1:  # the `box` lists: six faces in the box's own chain
            kids = [d2.nodes[k] for k in range(i + 1, n.end_or_prim)]
            assert all((k.kind >> 16) == (n.kind >> 16) for k in kids)
            assert n.bmin[0] <= min(k.bmin[0] for k in kids) and n.bmax[0] >= max(k.bmax[0] for k in kids)
