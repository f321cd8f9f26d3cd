"""GPU parity, early split clipping (TwoPassLbvh compiled with USE_PRIM_SPLITTING, TwoPassLbvh.cpp:23-28 →
Utility::doEarlySplitClipping, Utility.cpp:456-538): the references the device emits (boxes + triangle ids, in the reference's
FIFO order) and every buffer of the build over them, byte for byte against the oracle — which is itself pinned to the reference's
own function (tests/test_oracle_split.py)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh, random_tris
from b2bvh import capi, types as T
from test_gpu_lbvh import assert_same_struct, h32

pytestmark = pytest.mark.gpu
KA = json.load(open(os.path.join(GOLDEN, "split_known_answers.json")))


def check_split(ctx, oracle, tris, sa, **kw):
    tree = ctx.build(capi.TWO_PASS_LBVH, tris, split_sa_max=sa, **kw)
    g = ctx.fetch(tree)
    o = oracle.build_lbvh(tris, split_sa_max=sa)
    m = o["refs"].size
    assert tree.n_prims == m and tree.n_triangles == tris.size and tree.n_internal == m - 1
    assert_same_struct(g["boxes"], o["boxes"], "reference boxes")
    assert np.array_equal(g["prim_idx"], o["prim_idx"]), "reference triangle ids"
    assert_same_struct(g["scene"], o["scene"], "scene box")
    assert np.array_equal(g["keys"], o["keys"]) and np.array_equal(g["vals"], np.arange(m, dtype=np.uint32))
    assert np.array_equal(g["skeys"], o["skeys"]) and np.array_equal(g["svals"], o["svals"]), "sorted pairs"
    assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes")
    assert np.array_equal(g["parents"], o["parents"]), "parent indices"
    assert g["n_wide"] == o["wide_count"]
    assert_same_struct(g["wide"], o["wide"], "bvh4 nodes")
    assert_same_struct(g["wide_leaves"], o["wide_leaves"], "bvh4 leaves")
    assert np.float32(ctx.tree_cost(tree)) == np.float32(o["cost"])
    return tree, g, o


@pytest.mark.parametrize("key", [k for k in KA if not k.startswith("_")])
def test_split_golden_cases(ctx, oracle, key):
    kind, n, seed, sa = key.rsplit("_", 3)
    tris = load_mesh(kind) if n == "None" else random_tris(int(n), int(seed), kind)
    tree, g, o = check_split(ctx, oracle, tris, float(sa))
    ka = KA[key]
    refs = np.zeros(tree.n_prims, dtype=T.PRIM_REF)
    refs["primIdx"] = g["prim_idx"]; refs["mn"] = g["boxes"]["mn"]; refs["mx"] = g["boxes"]["mx"]
    assert tree.n_prims == ka["n_refs"] and h32(oracle, refs) == ka["refs_fnv"]
    assert h32(oracle, g["nodes"]) == ka["nodes_fnv"] and h32(oracle, g["wide"]) == ka["wide_fnv"]


@pytest.mark.parametrize("kind,n,seed,sa", [("uniform", 100_003, 51, 6.0), ("uniform", 511, 52, 2.0), ("uniform", 513, 53, 0.9), ("uniform", 20_000, 54, 1e9),
                                            ("clustered", 20_000, 55, 2e-4), ("anisotropic", 30_000, 56, 0.004)])
def test_split_synthetic(ctx, oracle, kind, n, seed, sa):
    tree, g, o = check_split(ctx, oracle, random_tris(n, seed, kind), sa)
    if sa >= 1e9:  # nothing to split: one generation, references == triangles
        assert tree.n_split_levels == 1 and tree.n_prims == n


def test_split_two_kernel_variant_and_repeat(ctx, oracle):
    tris = random_tris(30_000, 57)
    _, g1, _ = check_split(ctx, oracle, tris, 3.0, karras_two_kernel=True)
    _, g2, _ = check_split(ctx, oracle, tris, 3.0)
    assert g1["nodes"].tobytes() == g2["nodes"].tobytes()
    # a build without splitting on the same context afterwards is untouched by the split buffers
    tree = ctx.build(capi.TWO_PASS_LBVH, tris)
    g = ctx.fetch(tree)
    o = oracle.build_lbvh(tris)
    assert tree.n_prims == tris.size and g["prim_idx"] is None and g["nodes"].tobytes() == o["nodes"].tobytes()


def test_split_full_size_properties(ctx):
    """1 M triangles, ~3 M references: every reference fits saMax, fragments keep their triangle's id and lie inside its box,
    leaves name triangles, root box = scene box; sizes the oracle would take minutes for."""
    n = 1_000_000
    d = ctx.synth_uniform(n, 0xB20010)
    half = float(np.float32(1000.0 * n ** (-1.0 / 3.0)))
    sa = 6.0 * half * half  # roughly the median primitive-box area
    tree = ctx.build(capi.TWO_PASS_LBVH, d, n=n, tris_on_device=True, split_sa_max=sa)
    m = tree.n_prims
    assert m > n and tree.n_triangles == n
    boxes = ctx.download(tree.d_triangleAabb, T.AABB, m)
    prim = ctx.download(tree.d_primRefIdx, np.uint32, m)
    e = boxes["mx"] - boxes["mn"]
    area = np.float32(2) * ((e[:, 0] * e[:, 1] + e[:, 0] * e[:, 2]) + e[:, 1] * e[:, 2])
    assert (area <= np.float32(sa)).all()
    assert np.array_equal(np.unique(prim), np.arange(n, dtype=np.uint32))
    tb_tree = ctx.build(capi.TWO_PASS_LBVH, d, n=n, tris_on_device=True, collapse=False)
    tb = ctx.download(tb_tree.d_triangleAabb, T.AABB, n)
    assert (boxes["mn"] >= tb["mn"][prim]).all() and (boxes["mx"] <= tb["mx"][prim]).all()
    vol = np.zeros(n); np.add.at(vol, prim, np.prod(e.astype(np.float64), axis=1))
    assert np.allclose(vol, np.prod((tb["mx"] - tb["mn"]).astype(np.float64), axis=1), rtol=1e-4, atol=1e-9)
    tree = ctx.build(capi.TWO_PASS_LBVH, d, n=n, tris_on_device=True, split_sa_max=sa)
    nodes = ctx.download(tree.d_bvhNodes, T.BVH2_NODE, 2 * m - 1)
    svals = ctx.download(tree.d_sortedMortonCodeValues, np.uint32, m)
    assert np.array_equal(nodes["left"][m - 1:], prim[svals])
    scene = ctx.download(tree.d_sceneExtents, T.AABB, 1)
    assert np.array_equal(nodes["mn"][0], scene["mn"][0]) and np.array_equal(nodes["mx"][0], scene["mx"][0])
    wl = ctx.download(tree.d_wideLeafNodes, T.PRIM_NODE, m)
    assert np.array_equal(wl["primIdx"], prim[svals])
    ctx.free(d)


def test_split_traversal(ctx, oracle):
    """Rays through a tree over split references hit the same triangles at the same distance as through the unsplit tree."""
    tris = random_tris(20_000, 58)
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 400.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.6))
    d_rays, _ = ctx.generate_rays(cam, 256, 256)
    rays = ctx.download(d_rays, T.RAY, 256 * 256)
    base = ctx.build(capi.TWO_PASS_LBVH, tris, collapse=False)
    hits0, _, _ = ctx.traverse(base, d_rays, 256 * 256, tr)
    tree = ctx.build(capi.TWO_PASS_LBVH, tris, split_sa_max=2.0)
    g = ctx.fetch(tree)
    o_hits, cnt = oracle.traverse(rays, g["nodes"], None, tris, tr, 0, tree.n_prims)
    for kernel in (capi.TRAVERSE_WHILE, capi.TRAVERSE_IFIF, capi.TRAVERSE_WIDE4):
        hits, _, _ = ctx.traverse(tree, d_rays, 256 * 256, tr, kernel)
        assert np.array_equal(hits["t"].view(np.uint32), o_hits["t"].view(np.uint32))
        assert np.array_equal(hits["t"].view(np.uint32), hits0["t"].view(np.uint32))
        assert (hits["primIdx"] != hits0["primIdx"]).mean() < 1e-4
    assert cnt > 1000
    ctx.free(d_rays)


def test_split_argument_errors(ctx):
    tris = random_tris(1000, 59)
    for algo in (capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC):
        with pytest.raises(capi.B2bvhError, match="TWO_PASS"):
            ctx.build(algo, tris, split_sa_max=1.0)
    bad = tris.copy()
    bad["v"][3, 0, 0] = np.inf
    with pytest.raises(capi.B2bvhError, match="infinite or NaN"):
        ctx.build(capi.TWO_PASS_LBVH, bad, split_sa_max=1.0)
    # the context is still usable
    tree = ctx.build(capi.TWO_PASS_LBVH, tris, split_sa_max=1.0)
    assert tree.n_prims > tris.size
