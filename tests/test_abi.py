"""The C-ABI library loads and exports every symbol include/tpt.h declares; struct layouts of
the Python mirrors match the header's documented sizes. No compute calls."""
import ctypes as C
import os
import re

import pytest


def header_functions(root):
    text = open(os.path.join(root, "include", "tpt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tpt_[a-z_0-9]+)\s*\(", text)))


def test_header_lists_expected_entry_points(T):
    fns = header_functions(T.REPO_ROOT)
    assert sorted(T.C_ABI_SYMBOLS) == fns


def test_library_exports_every_declared_symbol(T):
    lib = T.lib()
    for name in header_functions(T.REPO_ROOT):
        assert hasattr(lib, name), f"libtpt.so does not export {name}"
    assert lib.tpt_api_version() == T.TPT_API_VERSION


def test_struct_layouts(T):
    assert C.sizeof(T.Node) == 32
    assert C.sizeof(T.Prim) == 64
    assert C.sizeof(T.Material) == 32
    assert C.sizeof(T.Texture) == 32
    assert C.sizeof(T.Light) == 32
    assert C.sizeof(T.XformOp) == 16
    assert C.sizeof(T.Chain) == 8
    assert C.sizeof(T.Camera) == 24 * 4
    assert C.sizeof(T.PerlinTables) == 256 * 3 * 4 + 3 * 256 * 4
    assert T.RAY_DTYPE.itemsize == 28 and T.HIT_DTYPE.itemsize == 48


def test_no_cpu_fallback(T):
    """Without a device the compute entry points must fail loudly (TPT_ERR_NO_DEVICE)."""
    if T.device_count() > 0:
        pytest.skip("a GPU is visible here")
    hs = T.HostScene("cornell_box")
    with pytest.raises(T.TptError) as e:
        T.Scene(hs)
    assert e.value.code == -3
    with pytest.raises(T.TptError):
        T.philox([0, 0, 0, 0], [0, 0])


def test_product_does_not_reference_oracle(T):
    """The product path must never import, link or execute anything under oracle/."""
    pkg = os.path.join(T.REPO_ROOT, "tiny-path-tracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in text and "oracle_ref" not in text and "libtptref" not in text, (dirpath, f)


def _medium_desc(T, nesting):
    """hand-made description: root list = [one constant_medium]; its boundary = a sphere wrapped in
    `nesting` hitable_lists (kept behind the root tree, as the flattener lays media out)"""
    n_nodes = 2 + nesting + 1
    nodes = (T.Node * n_nodes)()
    prims = (T.Prim * 2)()
    nodes[0].kind, nodes[0].end_or_prim = 1, 2           # root LIST
    nodes[1].kind, nodes[1].end_or_prim = 2, 0           # LEAF -> prim 0 (the medium)
    for k in range(nesting):                             # boundary: LIST(LIST(...(sphere)))
        nodes[2 + k].kind, nodes[2 + k].end_or_prim = 1, n_nodes
    nodes[2 + nesting].kind, nodes[2 + nesting].end_or_prim = 2, 1
    prims[0].kind, prims[0].material = 5, 0              # TPT_PRIM_MEDIUM
    prims[0].p[0] = 0.01
    first_end = (C.c_int32 * 2)(2, n_nodes)
    C.memmove(C.addressof(prims[0].p) + 4, first_end, 8)
    prims[1].kind, prims[1].material = 0, 0              # sphere
    prims[1].p[3] = 1.0
    chains, mats, texs = (T.Chain * 1)(), (T.Material * 1)(), (T.Texture * 1)()
    mats[0].kind = 5                                     # isotropic
    lights = (T.Light * 1)()
    d = T.SceneDesc()
    d.api_version = T.TPT_API_VERSION
    d.n_nodes, d.n_prims, d.n_chains, d.n_materials, d.n_textures, d.n_lights = n_nodes, 2, 1, 1, 1, 1
    d.nodes, d.prims, d.chains, d.materials, d.textures, d.lights = nodes, prims, chains, mats, texs, lights
    d.n_root_nodes = 2
    return d, (nodes, prims, chains, mats, texs, lights)


def test_deep_medium_boundary_is_refused(T):
    """the parity walk of a medium's boundary keeps 8 frames; a boundary nested deeper must be
    refused at upload (TPT_ERR_UNSUPPORTED), not overflow a device-side array. Validation runs before any
    device is touched, so this holds with and without a GPU."""
    lib = T.lib()
    for nesting, refused in ((3, False), (7, False), (8, True), (20, True)):
        d, keep = _medium_desc(T, nesting)
        out = C.c_void_p()
        rc = lib.tpt_scene_create(C.byref(d), 0, C.byref(out))
        if refused:
            assert rc == -4 and b"boundary nests deeper" in lib.tpt_last_error(), (nesting, rc, lib.tpt_last_error())
        else:
            assert rc in (0, -3), (nesting, rc, lib.tpt_last_error())  # fine: created, or no device here
            if rc == 0:
                lib.tpt_scene_destroy(out)
