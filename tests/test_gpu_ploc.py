"""GPU parity, PLOC++ and H-PLOC builders: Bvh2 nodes (canonical numbering of the oracle), leaf PrimRefs, Bvh4 and SAH cost
are compared byte for byte with the CPU oracle; the cost is also checked against the reference's README."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh, random_tris
from b2bvh import capi
from test_gpu_lbvh import assert_same_struct, h32, same_cost

pytestmark = pytest.mark.gpu
KA = json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def check_ploc(ctx, oracle, tris, algo, **opts):
    n = tris.size
    tree = ctx.build(algo, tris, **opts)
    g = ctx.fetch(tree)
    o = oracle.build_ploc(tris, hierarchical=(algo == capi.HPLOC))
    assert np.array_equal(g["skeys"], o["skeys"]) and np.array_equal(g["svals"], o["svals"])
    assert_same_struct(g["leaves"], o["leaves"], "leaf PrimRefs")
    assert g["root"] == 0
    assert oracle.check_bvh2(g["nodes"], g["leaves"], 0, n), "gpu tree is not a valid Bvh2"
    assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes")
    assert tree.n_iterations == (o["stats"]["merge_calls"] if algo == capi.HPLOC else o["stats"]["iterations"])
    assert g["n_wide"] == o["wide_count"]
    assert_same_struct(g["wide"], o["wide"], "bvh4 nodes")
    assert_same_struct(g["wide_leaves"], o["wide_leaves"], "bvh4 leaves")
    assert same_cost(ctx.tree_cost(tree), o["cost"])
    return tree, g, o


SYNTH = [("uniform", 2, 1), ("uniform", 3, 2), ("uniform", 17, 3), ("uniform", 33, 4), ("uniform", 1024, 5), ("uniform", 1025, 6), ("uniform", 5000, 7),
         ("uniform", 100_003, 8), ("clustered", 30_000, 9), ("flat", 4000, 10), ("duplicate", 900, 11), ("anisotropic", 20_000, 12)]


@pytest.mark.parametrize("algo", [capi.PLOCPP, capi.HPLOC], ids=["ploc", "hploc"])
@pytest.mark.parametrize("kind,n,seed", SYNTH, ids=[f"{k}-{n}" for k, n, _ in SYNTH])
def test_synthetic(ctx, oracle, algo, kind, n, seed):
    check_ploc(ctx, oracle, random_tris(n, seed, kind), algo)


# PLOC++ merge kernel: windows decide 480 clusters, the tail kernel takes over at 1024; with few CTAs a chunk spans many windows
PLOC_EDGES = [("uniform", 1026, 21, 0), ("uniform", 1504, 22, 0), ("uniform", 1505, 23, 0), ("uniform", 2400, 24, 2), ("uniform", 20_011, 25, 1),
              ("uniform", 20_011, 25, 3), ("clustered", 50_000, 26, 7), ("uniform", 300_007, 27, 0), ("anisotropic", 100_000, 28, 16)]


@pytest.mark.parametrize("kind,n,seed,ctas", PLOC_EDGES, ids=[f"{k}-{n}-ctas{c}" for k, n, _, c in PLOC_EDGES])
def test_ploc_merge_chunks(ctx, oracle, kind, n, seed, ctas):
    check_ploc(ctx, oracle, random_tris(n, seed, kind), capi.PLOCPP, merge_max_ctas=ctas)


@pytest.mark.parametrize("algo,key", [(capi.PLOCPP, "ploc"), (capi.HPLOC, "hploc")], ids=["ploc", "hploc"])
@pytest.mark.parametrize("mesh", ["cornellbox", "bunny", "sponza"])
def test_reference_meshes(ctx, oracle, algo, key, mesh):
    tris = load_mesh(mesh)
    if tris is None:
        pytest.skip(f"{mesh} not staged")
    tree, g, o = check_ploc(ctx, oracle, tris, algo)
    ka = KA[mesh]
    assert h32(oracle, g["nodes"]) == ka[f"{key}_nodes_fnv"] and h32(oracle, g["wide"]) == ka[f"{key}_wide_fnv"]
    assert np.float32(ctx.tree_cost(tree)) == np.float32(ka[f"{key}_cost"])
    if "readme" in ka:
        assert ctx.tree_cost(tree) == pytest.approx(ka["readme"][key], rel=1e-4)


@pytest.mark.parametrize("algo", [capi.PLOCPP, capi.HPLOC], ids=["ploc", "hploc"])
def test_large_synthetic_properties(ctx, oracle, algo):
    """1M primitives: valid tree, every box is the union of its children, repeatable bit for bit."""
    n = 1_000_000
    tris = oracle.synth_uniform(n, 0x00B20010)
    tree = ctx.build(algo, tris)
    g = ctx.fetch(tree)
    assert oracle.check_bvh2(g["nodes"], g["leaves"], 0, n) and oracle.check_bvh4(g["wide"], g["wide_leaves"], 0, n)
    nodes, leaves = g["nodes"], g["leaves"]
    allmn = np.concatenate([nodes["mn"], leaves["mn"]]); allmx = np.concatenate([nodes["mx"], leaves["mx"]])
    l, r = nodes["left"].astype(np.int64), nodes["right"].astype(np.int64)
    assert np.array_equal(nodes["mn"], np.minimum(allmn[l], allmn[r])) and np.array_equal(nodes["mx"], np.maximum(allmx[l], allmx[r]))
    assert ctx.tree_cost(tree) > 1.0
    g2 = ctx.fetch(ctx.build(algo, tris))
    assert g2["nodes"].tobytes() == nodes.tobytes() and g2["wide"].tobytes() == g["wide"].tobytes()


HPLOC_TILE = [("uniform", 129, 51), ("uniform", 257, 52), ("uniform", 1025, 53), ("uniform", 5000, 54), ("uniform", 100_003, 55), ("clustered", 30_000, 56),
              ("flat", 4000, 57), ("duplicate", 900, 58), ("anisotropic", 20_000, 59)]


@pytest.mark.parametrize("kind,n,seed", HPLOC_TILE, ids=[f"{k}-{n}" for k, n, _ in HPLOC_TILE])
def test_hploc_tile_phase_forced(ctx, oracle, kind, n, seed):
    """hploc_tile_kernel (automatic only from 2^20 primitives) forced on small inputs: same bytes and the same number of merge calls as the oracle."""
    check_ploc(ctx, oracle, random_tris(n, seed, kind), capi.HPLOC, lbvh_second_level=1)


@pytest.mark.parametrize("mesh", ["bunny", "sponza"])
def test_hploc_tile_phase_forced_on_meshes(ctx, oracle, mesh):
    tris = load_mesh(mesh)
    if tris is None:
        pytest.skip(f"{mesh} not staged")
    check_ploc(ctx, oracle, tris, capi.HPLOC, lbvh_second_level=1)
    check_ploc(ctx, oracle, tris, capi.HPLOC, lbvh_second_level=2)
