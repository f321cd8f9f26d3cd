"""GPU parity, seeded fuzz: random sizes (dense around the tile / window / tail boundaries of the kernels), input kinds, builder
options and launch modes; every buffer of every builder byte for byte against the oracle."""
import numpy as np
import pytest

from conftest import random_tris
from b2bvh import capi
from test_gpu_lbvh import assert_same_struct

pytestmark = pytest.mark.gpu

BOUNDARIES = [256, 480, 512, 960, 1024, 2048, 4096, 7680, 8192]
KINDS = ["uniform", "clustered", "flat", "anisotropic", "duplicate"]


def cases(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(count):
        if rng.random() < 0.6:
            n = int(rng.choice(BOUNDARIES)) * int(rng.integers(1, 5)) + int(rng.integers(-3, 4))
        else:
            n = int(rng.integers(2, 40_000))
        n = max(2, n)
        kind = KINDS[int(rng.integers(0, len(KINDS)))]
        if kind == "duplicate":
            n = min(n, 3000)  # all keys equal: the PLOC oracle is quadratic-ish in iterations there
        out.append((n, kind, int(rng.integers(0, 1 << 30)), int(rng.integers(0, 5)), bool(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("n,kind,seed,ctas,graph", cases(20261017, 24), ids=lambda v: str(v))
def test_fuzz_all_builders(ctx, oracle, n, kind, seed, ctas, graph):
    tris = random_tris(n, seed, kind)
    d = ctx.upload(tris)  # device triangles: the graph path needs a stable device (or pinned) pointer
    for algo in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH):
        single = algo == capi.SINGLE_PASS_LBVH
        o = oracle.build_lbvh(tris, single_pass=single)
        g = ctx.fetch(ctx.build(algo, d, n=n, tris_on_device=True, lbvh_second_level=1 + (seed & 1), use_graph=graph))
        assert np.array_equal(g["skeys"], o["skeys"]) and np.array_equal(g["svals"], o["svals"])
        assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes")
        assert_same_struct(g["wide"], o["wide"], "bvh4 nodes")
        assert_same_struct(g["wide_leaves"], o["wide_leaves"], "bvh4 leaves")
        assert g["root"] == o.get("root", 0)
    for algo in (capi.PLOCPP, capi.HPLOC):
        o = oracle.build_ploc(tris, hierarchical=(algo == capi.HPLOC))
        g = ctx.fetch(ctx.build(algo, d, n=n, tris_on_device=True, merge_max_ctas=ctas, use_graph=graph))
        assert_same_struct(g["leaves"], o["leaves"], "leaf PrimRefs")
        assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes")
        assert_same_struct(g["wide"], o["wide"], "bvh4 nodes")
        assert_same_struct(g["wide_leaves"], o["wide_leaves"], "bvh4 leaves")
    ctx.free(d)


@pytest.mark.parametrize("n,kind,seed,ctas,graph", cases(20261018, 16), ids=lambda v: str(v))
def test_fuzz_widened_paths(ctx, oracle, n, kind, seed, ctas, graph):
    """The same fuzz over the paths either side: 60-bit codes through all four builders (graph and plain launches, forced second merge
    level), early split with a threshold that cuts roughly every second box, and the input cut into batched items of random sizes."""
    tris = random_tris(n, seed, kind)
    d = ctx.upload(tris)
    for algo in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH):
        o = oracle.build_lbvh(tris, single_pass=(algo == capi.SINGLE_PASS_LBVH), morton_bits=60)
        g = ctx.fetch(ctx.build(algo, d, n=n, tris_on_device=True, lbvh_second_level=1 + (seed & 1), use_graph=graph, morton_bits=60))
        assert np.array_equal(g["skeys64"], o["skeys"]) and np.array_equal(g["svals"], o["svals"])
        assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes (60-bit)")
        assert_same_struct(g["wide"], o["wide"], "bvh4 nodes (60-bit)")
        assert g["root"] == o.get("root", 0)
    for algo in (capi.PLOCPP, capi.HPLOC):
        o = oracle.build_ploc(tris, hierarchical=(algo == capi.HPLOC), morton_bits=60)
        g = ctx.fetch(ctx.build(algo, d, n=n, tris_on_device=True, merge_max_ctas=ctas, use_graph=graph, morton_bits=60))
        assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes (60-bit)")
        assert_same_struct(g["wide"], o["wide"], "bvh4 nodes (60-bit)")
    # early split: threshold = median box area (a box of area 0 is never cut; duplicate / flat inputs give small reference counts)
    _, boxes, _ = oracle.primrefs(tris)
    e = boxes["mx"] - boxes["mn"]
    area = np.float32(2) * ((e[:, 0] * e[:, 1] + e[:, 0] * e[:, 2]) + e[:, 1] * e[:, 2])
    sa = float(np.median(area))
    if sa > 0 and np.isfinite(sa):
        o = oracle.build_lbvh(tris, split_sa_max=sa)
        t = ctx.build(capi.TWO_PASS_LBVH, d, n=n, tris_on_device=True, split_sa_max=sa)
        g = ctx.fetch(t)
        assert t.n_prims == o["refs"].size and np.array_equal(g["prim_idx"], o["prim_idx"])
        assert_same_struct(g["boxes"], o["boxes"], "reference boxes")
        assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes (split)")
        assert_same_struct(g["wide"], o["wide"], "bvh4 nodes (split)")
    # batched items of random sizes covering the whole input
    rng = np.random.default_rng(seed)
    counts = []
    left = n
    while left:
        c = int(min(left, rng.integers(1, 33)))
        counts.append(c); left -= c
    counts = np.array(counts, dtype=np.uint32)
    gb, ob = ctx.fetch_batch(ctx.build_batched(d, counts, tris_on_device=True)), oracle.build_batched(tris, counts)
    assert_same_struct(gb["nodes"], ob["nodes"], "batched nodes")
    assert_same_struct(gb["leaves"], ob["leaves"], "batched leaves")
    assert np.array_equal(gb["roots"], ob["roots"]) and gb["scenes"].tobytes() == ob["scenes"].tobytes()
    ctx.free(d)
