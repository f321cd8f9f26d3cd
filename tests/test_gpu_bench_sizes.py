"""GPU parity at the sizes the benchmark quotes: the 10 M synth_uniform_v1 workload of bench.py byte for byte
against the oracle (sorted pairs, Bvh2 nodes of both numberings, parents, Bvh4 nodes and leaves, cost) and against the frozen hashes;
buddha for all four builders; 100 M single-pass LBVH against the oracle when the box has the host memory for it."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh
from b2bvh import capi
from test_gpu_lbvh import assert_same_struct, check_lbvh, h32
from test_gpu_ploc import check_ploc

pytestmark = pytest.mark.gpu
KA = json.load(open(os.path.join(GOLDEN, "large_known_answers.json")))
SEED = 0x00B20010


def test_bench_workload_10m_singlepass(ctx, oracle):
    ka = KA["synth_uniform_v1_10M"]
    tris = oracle.synth_uniform(10_000_000, SEED)
    tree, g, o = check_lbvh(ctx, oracle, tris, capi.SINGLE_PASS_LBVH)
    assert oracle.fnv1a(g["skeys"], g["svals"]) == ka["sorted_kv_fnv"]
    assert h32(oracle, g["nodes"]) == ka["apetrei_nodes_fnv"] and g["root"] == ka["apetrei_root"]
    assert g["n_wide"] == ka["lbvh_wide_count"] and h32(oracle, g["wide"]) == ka["lbvh_wide_fnv"] and h32(oracle, g["wide_leaves"]) == ka["lbvh_wide_leaves_fnv"]
    assert np.float32(ctx.tree_cost(tree)) == np.float32(ka["lbvh_cost"])
    # the way bench.py runs it: triangles generated on the device, graph replay — same bytes
    d = ctx.synth_uniform(10_000_000, SEED)
    try:
        for _ in range(2):
            t2 = ctx.build(capi.SINGLE_PASS_LBVH, d, n=10_000_000, tris_on_device=True, use_graph=True)
        for nm, dt, cnt, ptr in (("nodes", g["nodes"].dtype, g["nodes"].size, t2.d_bvhNodes), ("wide", g["wide"].dtype, t2.n_wide, t2.d_wideBvhNodes),
                                 ("wide_leaves", g["wide_leaves"].dtype, g["wide_leaves"].size, t2.d_wideLeafNodes)):
            assert ctx.download(ptr, dt, cnt).tobytes() == g[nm].tobytes(), nm
    finally:
        ctx.free(d)


def test_bench_workload_10m_twopass(ctx, oracle):
    ka = KA["synth_uniform_v1_10M"]
    tris = oracle.synth_uniform(10_000_000, SEED)
    g = ctx.fetch(ctx.build(capi.TWO_PASS_LBVH, tris))
    assert oracle.fnv1a(g["skeys"], g["svals"]) == ka["sorted_kv_fnv"]
    assert h32(oracle, g["nodes"]) == ka["karras_nodes_fnv"] and h32(oracle, g["parents"]) == ka["karras_parents_fnv"]
    assert g["n_wide"] == ka["lbvh_wide_count"] and h32(oracle, g["wide"]) == ka["lbvh_wide_fnv"]
    nodes, parents = oracle.lbvh_karras(oracle.primrefs(tris)[0], g["skeys"], g["svals"])
    assert_same_struct(g["nodes"], nodes, "karras nodes at 10 M")
    assert np.array_equal(g["parents"], parents)


@pytest.mark.parametrize("algo,key", [(capi.TWO_PASS_LBVH, "lbvh"), (capi.SINGLE_PASS_LBVH, "lbvh"), (capi.PLOCPP, "ploc"), (capi.HPLOC, "hploc")],
                         ids=["twopass", "singlepass", "ploc", "hploc"])
def test_buddha(ctx, oracle, algo, key):
    tris = load_mesh("buddha")
    if tris is None:
        pytest.skip("buddha not staged")
    ka = KA["buddha"]
    if key == "lbvh":
        tree, g, o = check_lbvh(ctx, oracle, tris, algo)
        assert h32(oracle, g["nodes"]) == ka["karras_nodes_fnv" if algo == capi.TWO_PASS_LBVH else "apetrei_nodes_fnv"]
    else:
        tree, g, o = check_ploc(ctx, oracle, tris, algo)
        assert h32(oracle, g["nodes"]) == ka[f"{key}_nodes_fnv"]
    assert g["n_wide"] == ka[f"{key}_wide_count"] == ka["survey_probe"][f"{key}_wide_count"] and h32(oracle, g["wide"]) == ka[f"{key}_wide_fnv"]
    assert np.float32(ctx.tree_cost(tree)) == np.float32(ka[f"{key}_cost"])
    assert ctx.tree_cost(tree) == pytest.approx(ka["survey_probe"][f"{key}_cost"], rel=1e-6)


def test_100m_singlepass_against_oracle(ctx, oracle):
    """BASELINE configs[4]'s size on ONE GPU, memcmp against the oracle.  Needs ~60 GB of host memory and a few minutes of CPU."""
    import psutil
    if psutil.virtual_memory().available < 100 * (1 << 30):
        pytest.skip("less than 100 GB of host memory available")
    n = 100_000_000
    d = ctx.synth_uniform(n, 0x00B20100)
    try:
        tree = ctx.build(capi.SINGLE_PASS_LBVH, d, n=n, tris_on_device=True)
        tris = oracle.synth_uniform(n, 0x00B20100)
        o = oracle.build_lbvh(tris, single_pass=True)
        del tris
        assert tree.root == o["root"] and tree.n_wide == o["wide_count"]
        for nm, ref, ptr in (("sorted keys", o["skeys"], tree.d_sortedMortonCodeKeys), ("sorted values", o["svals"], tree.d_sortedMortonCodeValues),
                             ("bvh2 nodes", o["nodes"], tree.d_bvhNodes), ("bvh4 nodes", o["wide"], tree.d_wideBvhNodes),
                             ("bvh4 leaves", o["wide_leaves"], tree.d_wideLeafNodes)):
            got = ctx.download(ptr, ref.dtype, ref.size)
            assert_same_struct(got, ref, nm + " at 100 M")
            del got
        assert np.float32(ctx.tree_cost(tree)) == np.float32(o["cost"])
    finally:
        ctx.free(d)
