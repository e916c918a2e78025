"""Small-scene table of the warp-uniform closest hit (csrc/tpt_api.cu build_small_scene): which
primitives are folded into blocks, and that the block rule itself -- a ray can meet the faces of
an axis-aligned block only where it enters and where it leaves the block; a crossing whose face is
absent is skipped -- gives the same closest hit as testing the faces one rectangle at a time the
way the reference does (src/rect_box.cc:8-24,51-68,75-91 under hitable_list::hit,
src/hitable_list.cc:38-51). Host only: the device code itself is checked against the reference's
hit records in tests/test_gpu_hits.py::test_fast_hits_vs_golden."""
import numpy as np

import common

XY, XZ, YZ = 2, 3, 4  # TPT_PRIM_XY_RECT.. (include/tpt.h)
AXES = {XY: (2, 0, 1), XZ: (1, 0, 2), YZ: (0, 1, 2)}  # plane axis, in-plane axes of p[0..1], p[2..3]


def rect_prims(T, hs):
    d = hs.desc.contents
    out = {}
    for i in range(d.n_prims):
        p = d.prims[i]
        if p.kind in AXES:
            out[i] = (p.kind, p.chain, [float(x) for x in p.p[:5]])
    return out


def test_cornell_walls_become_one_open_block(T):
    hs = common.host_scene(T, "cornell_box")
    s = T.small_scene_summary(hs)
    rects = rect_prims(T, hs)
    assert s["enabled"] and s["groups"] == 3 and s["spheres"] == 1
    assert s["rects"] == 1  # only the lamp stays a loose rect
    assert [b["open"] for b in s["blocks"]] == [True, False, False]
    walls = s["blocks"][0]
    assert walls["chain"] == 0 and walls["faces"][5] == -1 and sum(f >= 0 for f in walls["faces"]) == 5
    lo, hi = [-300.0] * 3, [300.0] * 3
    for slot, f in enumerate(walls["faces"]):
        if f < 0:
            continue
        kind, chain, p = rects[f]
        axis, a, b = AXES[kind]
        assert chain == 0 and slot // 2 == axis and p[4] == (lo, hi)[slot % 2][axis]
        assert p[:4] == [lo[a], hi[a], lo[b], hi[b]]
    used = {f for blk in s["blocks"] for f in blk["faces"] if f >= 0}
    assert len(used) == 5 + 6 + 6  # every face belongs to exactly one block
    for blk in s["blocks"][1:]:  # the two `box` objects: all six faces, their own chains
        assert all(f >= 0 for f in blk["faces"]) and blk["chain"] in (1, 2)


def test_scenes_without_blocks_are_left_alone(T):
    s = T.small_scene_summary(common.host_scene(T, "two_perlin_spheres"))
    assert s["enabled"] and s["blocks"] == [] and s["spheres"] == 2 and s["rects"] == 0
    big = T.small_scene_summary(common.host_scene(T, "random_scene"))
    assert not big["enabled"]  # 485 leaves: SAH BVH path, no constant-bank table


def rects_closest(o, d, faces, tmin):
    """hitable_list::hit over rect::hit, float32"""
    best = np.full(len(o), np.float32(np.inf), np.float32)
    prim = np.full(len(o), -1, np.int64)
    with np.errstate(all="ignore"):
        for pid, (kind, p) in faces.items():
            axis, a, b = AXES[kind]
            t = (np.float32(p[4]) - o[:, axis]) / d[:, axis]
            pa = o[:, a] + t * d[:, a]
            pb = o[:, b] + t * d[:, b]
            ok = (t >= tmin) & (t <= best) & (pa >= p[0]) & (pa <= p[1]) & (pb >= p[2]) & (pb <= p[3])
            best = np.where(ok, t, best)
            prim = np.where(ok, pid, prim)
    return best, prim


def block_closest(o, d, lo, hi, face, tmin):
    """the open-block rule of closest_hit_uniform (csrc/tpt_device.cuh), float32"""
    with np.errstate(all="ignore"):
        inv = np.float32(1.0) / d
        t0 = (lo - o) * inv
        t1 = (hi - o) * inv
    near, far = np.fmin(t0, t1), np.fmax(t0, t1)
    tn, tf = near.max(axis=1), far.min(axis=1)
    best = np.full(len(o), np.float32(np.inf), np.float32)
    prim = np.full(len(o), -1, np.int64)
    face = np.asarray(face)
    for i in range(len(o)):
        if not (tn[i] <= tf[i] and tf[i] >= tmin):
            continue
        settled = False
        if tn[i] >= tmin:
            axis = int(np.argmax(near[i] == tn[i]))
            f = face[2 * axis + (0 if tn[i] == t0[i, axis] else 1)]
            if f >= 0:
                best[i], prim[i], settled = tn[i], f, True
        if not settled:
            axis = int(np.argmax(far[i] == tf[i]))
            f = face[2 * axis + (0 if tf[i] == t0[i, axis] else 1)]
            if f >= 0:
                best[i], prim[i] = tf[i], f
    return best, prim


def test_open_block_rule_equals_rect_by_rect(T):
    hs = common.host_scene(T, "cornell_box")
    walls = T.small_scene_summary(hs)["blocks"][0]
    rects = rect_prims(T, hs)
    faces = {f: (rects[f][0], rects[f][2]) for f in walls["faces"] if f >= 0}
    lo = np.full(3, -300.0, np.float32)
    hi = np.full(3, 300.0, np.float32)
    rng = np.random.default_rng(11)
    n = 6000
    o = np.concatenate([rng.uniform(-299, 299, (n // 2, 3)),               # inside the room
                        rng.uniform(-900, 900, (n // 2, 3))]).astype(np.float32)  # anywhere around it
    o[: n // 8, 2] = 800.0  # the camera's side, looking through the absent front wall
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d[n // 8: n // 4, 1] = 0.0  # rays parallel to floor and ceiling
    tmin = np.float32(0.001)
    bt, bp = rects_closest(o, d, faces, tmin)
    gt, gp = block_closest(o, d, lo, hi, walls["faces"], tmin)
    assert (bp >= 0).sum() > n // 3 and (bp < 0).sum() > n // 10  # both outcomes are exercised
    agree = bp == gp
    # the two forms round t identically ((k - o) / d vs (k - o) * (1 / d) differ by an ulp at most):
    # a disagreement is only legitimate within a few ulps of an edge, where two walls tie
    assert (~agree).sum() <= 2, int((~agree).sum())
    hit = agree & (bp >= 0)
    assert np.allclose(bt[hit], gt[hit], rtol=3e-7, atol=0)


# ---------------------------------------------------------------------------------------------
# hand-made scene descriptions (a hitable_list of loose rects under the identity chain)
# ---------------------------------------------------------------------------------------------
def rect_list_desc(T, rects):
    """rects: (kind, a0, a1, b0, b1, k). Returns (desc, keep-alive)."""
    import ctypes as C
    n = len(rects)
    nodes = (T.Node * (n + 1))()
    prims = (T.Prim * n)()
    nodes[0].kind = 1  # TPT_NODE_LIST, chain 0
    nodes[0].end_or_prim = n + 1
    for i, (kind, a0, a1, b0, b1, k) in enumerate(rects):
        nodes[i + 1].kind = 2  # TPT_NODE_LEAF
        nodes[i + 1].end_or_prim = i
        prims[i].kind, prims[i].material, prims[i].chain, prims[i].flags = kind, 0, 0, 0
        for j, x in enumerate((a0, a1, b0, b1, k)):
            prims[i].p[j] = x
    chains = (T.Chain * 1)()
    mats = (T.Material * 1)()
    texs = (T.Texture * 1)()
    d = T.SceneDesc()
    d.api_version = T.TPT_API_VERSION
    d.n_nodes, d.n_prims, d.n_chains, d.n_materials, d.n_textures = n + 1, n, 1, 1, 1
    d.nodes, d.prims, d.chains, d.materials, d.textures = nodes, prims, chains, mats, texs
    return d, (nodes, prims, chains, mats, texs)


def unit_room(faces="xXyYzZ", lo=-1.0, hi=2.0):
    """faces of the block [lo, hi]^3 as loose rects: x/X = yz_rect at lo/hi, y/Y = xz_rect, z/Z = xy_rect"""
    kinds = {"x": YZ, "y": XZ, "z": XY}
    return [(kinds[f.lower()], lo, hi, lo, hi, lo if f.islower() else hi) for f in faces]


def summary_of(T, rects):
    import ctypes as C
    d, keep = rect_list_desc(T, rects)
    return T.small_scene_summary(C.pointer(d))


def test_six_loose_faces_make_a_closed_block(T):
    s = summary_of(T, unit_room("xXyYzZ"))
    assert s["enabled"] and s["rects"] == 0 and len(s["blocks"]) == 1
    assert s["blocks"][0]["open"] is False and sorted(s["blocks"][0]["faces"]) == [0, 1, 2, 3, 4, 5]
    # slot order is -x +x -y +y -z +z whatever the order the rects come in
    s = summary_of(T, unit_room("ZyxYzX"))
    assert s["blocks"][0]["faces"] == [2, 5, 1, 3, 4, 0]


def test_four_faces_fold_three_do_not(T):
    s = summary_of(T, unit_room("xXyz"))
    assert len(s["blocks"]) == 1 and s["blocks"][0]["open"] and s["rects"] == 0
    assert s["blocks"][0]["faces"] == [0, 1, 2, -1, 3, -1]
    s = summary_of(T, unit_room("xyz"))
    assert s["blocks"] == [] and s["rects"] == 3


def test_only_whole_faces_join_a_block(T):
    room = unit_room("xXyYz")
    lamp = (XZ, -0.5, 0.5, -0.5, 0.5, 1.9)        # inside the room, not on a face plane
    patch = (XY, -1.0, 1.0, -1.0, 2.0, 2.0)       # on the +z plane but not the whole face
    twin = (YZ, -1.0, 2.0, -1.0, 2.0, -1.0)       # a second rect on the -x face
    s = summary_of(T, room + [lamp, patch, twin])
    assert len(s["blocks"]) == 1 and s["blocks"][0]["open"]
    assert s["blocks"][0]["faces"] == [0, 1, 2, 3, 4, -1]
    assert s["rects"] == 3  # lamp, patch and the twin stay rectangle tests


def test_two_separate_rooms(T):
    s = summary_of(T, unit_room("xXyYz") + unit_room("xXyYzZ", lo=10.0, hi=11.0))
    assert len(s["blocks"]) == 2 and s["rects"] == 0
    assert sorted(b["open"] for b in s["blocks"]) == [False, True]
    assert sorted(f for b in s["blocks"] for f in b["faces"] if f >= 0) == list(range(11))
