"""the acceleration structure of B200RT_CAST_BVH, built on the host (csrc/rt_bvh_build.h) - the structural
invariants the traversal (csrc/rt_bvh.cuh) relies on: a spatial tree over the well-shaped triangles and a tree over the
unit normals of all of them.  No GPU: the parity of the cast itself is tests/test_gpu_bvh.py."""
import ctypes as C
import sys

import numpy as np

import scene_util

KAPPA_MAX = 32.0


def build(b200rt, world, which):
    lib = b200rt.load_library()
    sc = world.scene()
    n_nodes, n_idx, depth, leaves = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    args = (C.byref(n_nodes), C.byref(n_idx), C.byref(depth), C.byref(leaves))
    assert lib.b200rt_dev_build_bvh(C.byref(sc), which, None, 0, None, *args) == b200rt.OK
    per = 12 if which == 0 else 8
    nodes = np.zeros((max(n_nodes.value, 1), per), dtype=np.float32)
    perm = np.zeros(max(n_idx.value, 1), dtype=np.uint32)
    assert lib.b200rt_dev_build_bvh(C.byref(sc), which, nodes.ctypes.data_as(C.POINTER(C.c_float)), n_nodes.value,
                                    perm.ctypes.data_as(C.POINTER(C.c_uint32)), *args) == b200rt.OK
    return nodes[: n_nodes.value], perm[: n_idx.value], depth.value, leaves.value


def tri_positions(sc):
    if sc.n_triangles == 0:
        return np.zeros((0, 3, 3), dtype=np.float64)
    t = np.ctypeslib.as_array(C.cast(sc.triangles, C.POINTER(C.c_float)), shape=(sc.n_triangles, 25))
    return np.stack([t[:, 0:3], t[:, 8:11], t[:, 16:19]], axis=1).astype(np.float64)        # [tri][vertex][xyz]


def shape_of(pos):
    """kappa = 1 / sin(theta_min / 2) and the diameter of every triangle"""
    kappa = np.full(len(pos), np.inf)
    diam = np.zeros(len(pos))
    for i in range(len(pos)):
        s = 1.0
        for k in range(3):
            a, b, c = pos[i, k], pos[i, (k + 1) % 3], pos[i, (k + 2) % 3]
            x, y = b - a, c - a
            lx, ly = np.linalg.norm(x), np.linalg.norm(y)
            diam[i] = max(diam[i], lx, ly)
            if not (lx > 0 and ly > 0):
                s = 0.0
                break
            s = min(s, np.sqrt(0.5 * (1 - np.clip(np.dot(x, y) / (lx * ly), -1, 1))))
        if s > 0:
            kappa[i] = 1.0 / s
    return kappa, diam


def walk(nodes, perm, which, on_leaf):
    """depth-first over the tree; returns per-node (lo, hi, extra) and checks that children lie inside their parents"""
    sys.setrecursionlimit(20000)
    u = nodes.view(np.uint32)
    stats = {"leaves": 0, "tris": []}

    def rec(k):
        if which == 0:
            lo, rho, hi, w1, w2, axis = nodes[k, 0:3], nodes[k, 3], nodes[k, 4:7], int(u[k, 8]), int(u[k, 9]), int(u[k, 10])
            assert axis < 3 and rho >= 0
        else:
            lo, hi, w1, w2, rho = nodes[k, 0:3], nodes[k, 4:7], int(u[k, 3]), int(u[k, 7]), 0.0
        if w2 & 0x80000000:
            cnt = w2 & 0x7fffffff
            assert 1 <= cnt <= 4
            stats["leaves"] += 1
            tris = perm[w1:w1 + cnt].tolist()
            stats["tris"] += tris
            on_leaf(tris, lo, hi, rho)
            return lo, hi, rho
        for child in (w1, w2):
            clo, chi, crho = rec(child)
            assert (clo >= lo).all() and (chi <= hi).all() and crho <= rho
        return lo, hi, rho

    if len(nodes):
        rec(0)
    return stats


def check_trees(b200rt, world):
    sc = world.scene()
    n = sc.n_triangles
    pos = tri_positions(sc)
    kappa, diam = shape_of(pos)
    nrm = np.cross(pos[:, 1] - pos[:, 0], pos[:, 2] - pos[:, 1]) if n else np.zeros((0, 3))
    ln = np.linalg.norm(nrm, axis=1, keepdims=True) if n else np.zeros((0, 1))
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = nrm / ln
    regular = np.isfinite(pos).all(axis=(1, 2)) & (ln[:, 0] > 0) & (kappa <= KAPPA_MAX) if n else np.zeros(0, dtype=bool)

    s_nodes, s_perm, s_depth, s_leaves = build(b200rt, world, 0)
    assert sorted(s_perm.tolist()) == np.where(regular)[0].tolist()            # the well-shaped triangles, each once
    assert s_depth <= 62

    def spatial_leaf(tris, lo, hi, rho):
        for i in tris:
            assert (pos[i] >= lo).all() and (pos[i] <= hi).all()              # the exact box of the vertices
            assert rho >= 1e-4 * kappa[i] * diam[i] * (1 - 1e-5)                # rho_geom = 1e-4 kappa E'

    st = walk(s_nodes, s_perm, 0, spatial_leaf)
    assert sorted(st["tris"]) == np.where(regular)[0].tolist() and st["leaves"] == s_leaves

    n_nodes, n_perm, n_depth, n_leaves = build(b200rt, world, 1)
    flagged = (n_perm & 0x80000000) != 0                                        # not in the spatial tree: tested by every ray
    n_perm = n_perm & 0x7fffffff
    assert sorted(n_perm.tolist()) == list(range(n)) and n_depth <= 62        # every triangle, each once
    assert sorted(n_perm[flagged].tolist()) == np.where(~regular)[0].tolist()

    def normal_leaf(tris, lo, hi, rho):
        for i in tris:
            if regular[i]:
                assert (nrm[i] >= lo - 1e-6).all() and (nrm[i] <= hi + 1e-6).all()
            else:
                assert (lo == -1).all() and (hi == 1).all()                      # reached by every ray

    st = walk(n_nodes, n_perm, 1, normal_leaf)
    assert sorted(st["tris"]) == list(range(n)) and st["leaves"] == n_leaves
    return s_nodes, n_nodes, regular


def test_bvh_fixture_scene(b200rt, fixture_world):
    s_nodes, n_nodes, regular = check_trees(b200rt, fixture_world)
    assert regular.all() and len(s_nodes) >= 31 and len(n_nodes) >= 31


def test_bvh_heightfield_and_degenerates(b200rt, tmp_path):
    world, ntri = scene_util.fixture_plus_mesh(b200rt, tmp_path, n=65)           # 8192 + 64 triangles
    o = b200rt.ObjectProxy(world, 0)
    o.push_flat_triangle([[0, 0, 1], [0, 0, 1], [0, 0, 1]])                      # zero area: NaN normal in the reference
    o.push_flat_triangle([[0, 0, 0], [1, 0, 0], [2, 0, 0]])                      # collinear
    o.push_flat_triangle([[0, 0, 0], [1, 0, 0], [1, 1e-7, 0]])                   # a needle: kappa beyond the cap
    s_nodes, n_nodes, regular = check_trees(b200rt, world)
    assert (~regular).sum() == 3
    # the band |n.dir| < g around a great circle crosses few leaves of the normal tree: count them for some directions
    u = n_nodes.view(np.uint32)
    leaf = (u[:, 7] & 0x80000000) != 0
    rng = np.random.default_rng(0)
    crossed = []
    for _ in range(16):
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        lo = np.minimum(n_nodes[leaf, 0:3] * d, n_nodes[leaf, 4:7] * d).sum(axis=1)
        hi = np.maximum(n_nodes[leaf, 0:3] * d, n_nodes[leaf, 4:7] * d).sum(axis=1)
        crossed.append(int(((lo <= 5e-6) & (hi >= -5e-6)).sum()))
    assert np.mean(crossed) < 0.1 * leaf.sum(), (np.mean(crossed), leaf.sum())


def test_bvh_empty_and_single(b200rt):
    w = b200rt.World()
    o = w.push_object(b200rt.color_material())
    check_trees(b200rt, w)
    o.push_flat_triangle([[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    s_nodes, n_nodes, regular = check_trees(b200rt, w)
    assert len(s_nodes) == 1 and len(n_nodes) == 1
