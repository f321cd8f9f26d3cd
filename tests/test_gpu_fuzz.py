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
