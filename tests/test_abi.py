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
