"""GPU parity, 60-bit Morton variant (north star "30/60-bit Morton coding"; b2bvh_build_opts.morton_bits = 60):
codes, the two-digit LSD sort, the hierarchy over 64-bit keys and the collapse, byte for byte against the oracle's definition
(plain 20-bit-per-axis interleave, stable sort by the 64-bit code, the Karras / Apetrei rules one word wider)."""
import numpy as np
import pytest

from conftest import load_mesh, random_tris
from b2bvh import capi, types as T
from test_gpu_lbvh import assert_same_struct, same_cost

pytestmark = pytest.mark.gpu


def check60(ctx, oracle, tris, algo, **kw):
    n = tris.size
    tree = ctx.build(algo, tris, morton_bits=60, **kw)
    g = ctx.fetch(tree)
    assert tree.morton_bits == 60
    if algo in (capi.PLOCPP, capi.HPLOC):
        o = oracle.build_ploc(tris, hierarchical=(algo == capi.HPLOC), morton_bits=60)
    else:
        o = oracle.build_lbvh(tris, single_pass=(algo == capi.SINGLE_PASS_LBVH), morton_bits=60)
    assert np.array_equal(g["keys64"], o["keys"]), "60-bit codes"
    assert np.array_equal(g["keys"], (o["keys"] >> np.uint64(30)).astype(np.uint32)), "upper 30 bits"
    assert np.array_equal(g["skeys64"], o["skeys"]) and np.array_equal(g["svals"], o["svals"]), "sorted (code, index) pairs"
    assert g["root"] == o["root"]
    assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes")
    if algo == capi.TWO_PASS_LBVH:
        _, parents = oracle.lbvh_karras(o["refs"], o["skeys"], o["svals"])
        assert np.array_equal(g["parents"], parents)
    if algo in (capi.PLOCPP, capi.HPLOC):
        assert_same_struct(g["leaves"], o["leaves"], "ploc leaves")
    assert g["n_wide"] == o["wide_count"]
    assert_same_struct(g["wide"], o["wide"], "bvh4 nodes")
    assert_same_struct(g["wide_leaves"], o["wide_leaves"], "bvh4 leaves")
    assert same_cost(ctx.tree_cost(tree), o["cost"])
    return tree, g, o


ALGOS = [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC]
SYNTH = [("uniform", 2, 81), ("uniform", 3, 82), ("uniform", 1000, 83), ("uniform", 100_003, 84), ("clustered", 20_000, 85), ("flat", 5000, 86),
         ("duplicate", 700, 87), ("anisotropic", 30_000, 88)]


@pytest.mark.parametrize("algo", ALGOS, ids=["twopass", "singlepass", "ploc", "hploc"])
@pytest.mark.parametrize("kind,n,seed", SYNTH, ids=[f"{k}-{n}" for k, n, _ in SYNTH])
def test_morton60_synthetic(ctx, oracle, algo, kind, n, seed):
    check60(ctx, oracle, random_tris(n, seed, kind), algo)


@pytest.mark.parametrize("algo", ALGOS[:2], ids=["twopass", "singlepass"])
@pytest.mark.parametrize("kind,n,seed", [("uniform", 70_001, 94), ("clustered", 50_000, 95), ("duplicate", 9000, 96)])
def test_morton60_second_merge_level_forced(ctx, oracle, algo, kind, n, seed):
    """The group kernel (second merge level, automatic only from 2^23 primitives) instantiated for 64-bit keys."""
    check60(ctx, oracle, random_tris(n, seed, kind), algo, lbvh_second_level=1)
    check60(ctx, oracle, random_tris(n, seed, kind), algo, lbvh_second_level=2)


@pytest.mark.parametrize("mesh", ["cornellbox", "bunny"])
def test_morton60_meshes(ctx, oracle, mesh):
    tris = load_mesh(mesh)
    if tris is None:
        pytest.skip(f"{mesh} not staged")
    for algo in ALGOS[:2]:
        check60(ctx, oracle, tris, algo)


def test_morton60_resolves_what_30_bits_cannot(ctx):
    """2 M uniform triangles: the sorted 60-bit codes are non-decreasing, (code, index) strictly increasing, the values a permutation, and
    far fewer neighbours share a code than share its upper 30 bits (thousands of 30-bit collisions at this size).  (numpy points: the
    bench's synth_uniform_v1 stream repeats triangles — the reference's lcg/randf yields 24-bit sequences that depend on 24 seed bits.)"""
    n = 2_000_000
    rng = np.random.default_rng(93)
    c = rng.uniform(-1000, 1000, size=(n, 1, 3))
    tris = T.triangles_from_array((c + rng.uniform(-1, 1, size=(n, 3, 3))).astype(np.float32))
    d = ctx.upload(tris)
    tree = ctx.build(capi.SINGLE_PASS_LBVH, d, n=n, tris_on_device=True, morton_bits=60)
    sk = ctx.download(tree.d_sortedMortonCodeKeys64, np.uint64, n)
    sv = ctx.download(tree.d_sortedMortonCodeValues, np.uint32, n)
    k = ctx.download(tree.d_mortonCodeKeys64, np.uint64, n)
    assert (sk[1:] >= sk[:-1]).all() and np.array_equal(sk, k[sv])
    same = sk[1:] == sk[:-1]
    assert (sv[1:][same] > sv[:-1][same]).all()                      # stable
    assert np.array_equal(np.sort(sv), np.arange(n, dtype=np.uint32))  # a permutation
    dup60 = int(same.sum())
    hi = sk >> np.uint64(30)
    dup30 = int((hi[1:] == hi[:-1]).sum())
    assert dup60 * 100 < dup30 and dup30 > 0
    nodes = ctx.download(tree.d_bvhNodes, T.BVH2_NODE, 2 * n - 1)
    scene = ctx.download(tree.d_sceneExtents, T.AABB, 1)
    assert np.array_equal(nodes["mn"][tree.root], scene["mn"][0]) and np.array_equal(nodes["mx"][tree.root], scene["mx"][0])
    ctx.free(d)


def test_morton60_argument_errors(ctx):
    tris = random_tris(1000, 89)
    with pytest.raises(capi.B2bvhError, match="morton_bits"):
        ctx.build(capi.TWO_PASS_LBVH, tris, morton_bits=64)
    with pytest.raises(capi.B2bvhError, match="karras_two_kernel"):
        ctx.build(capi.TWO_PASS_LBVH, tris, morton_bits=60, karras_two_kernel=True)
