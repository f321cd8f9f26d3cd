"""GPU: what the boundary does with inputs the reference never sees in its own scenes — argument errors (status + message, nothing launched) and
degenerate geometry (zero-area triangles, a scene on a line, every triangle the same one, a scene that is one point, huge and denormal
coordinates) — every buffer of all four builders byte for byte against the oracle, which restates the reference's arithmetic for exactly these
cases (division by a zero extent, saturating conversions: DESIGN.md section 4)."""
import numpy as np
import pytest

from conftest import random_tris
from b2bvh import capi, types as T
from test_gpu_lbvh import check_lbvh
from test_gpu_ploc import check_ploc
from test_gpu_morton60 import check60

pytestmark = pytest.mark.gpu


def degenerate(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "points":        # zero-area triangles: boxes without extent
        v = np.repeat(rng.uniform(-50, 50, size=(n, 1, 3)), 3, axis=1)
    elif kind == "line":        # the whole scene on one axis: two flat scene extents
        v = np.zeros((n, 3, 3))
        v[:, :, 0] = rng.uniform(-100, 100, size=(n, 1)) + rng.uniform(-0.5, 0.5, size=(n, 3))
        v[:, :, 1] = 2.0
        v[:, :, 2] = -7.0
    elif kind == "coincident":  # every triangle the same one, no outlier: all centroids equal, every key equal
        v = np.tile(rng.uniform(-1, 1, size=(1, 3, 3)), (n, 1, 1))
    elif kind == "one_point":   # the scene is a single point: every extent is zero
        v = np.tile(rng.uniform(-1, 1, size=(1, 1, 3)), (n, 3, 1))
    elif kind == "huge":        # areas near the top of the float range
        v = rng.uniform(-1, 1, size=(n, 1, 3)) * 1e15 + rng.uniform(-1, 1, size=(n, 3, 3)) * 1e13
    elif kind == "denormal":    # box areas are denormal numbers (no flush to zero on either side)
        v = rng.uniform(-1, 1, size=(n, 1, 3)) * 1e-19 + rng.uniform(-1, 1, size=(n, 3, 3)) * 1e-21
    elif kind == "mixed_scale": # one far outlier squeezes everything else into a few Morton cells
        v = rng.uniform(-1, 1, size=(n, 1, 3)) + rng.uniform(-0.01, 0.01, size=(n, 3, 3))
        v[n // 2] += 1e7
    else:
        raise ValueError(kind)
    return T.triangles_from_array(v.astype(np.float32))


KINDS = ["points", "line", "coincident", "one_point", "huge", "denormal", "mixed_scale"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH], ids=["twopass", "singlepass"])
def test_degenerate_geometry_lbvh(ctx, oracle, algo, kind):
    for n, seed in ((2, 401), (97, 402), (3001, 403)):
        check_lbvh(ctx, oracle, degenerate(kind, n, seed), algo)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("algo", [capi.PLOCPP, capi.HPLOC], ids=["ploc", "hploc"])
def test_degenerate_geometry_ploc(ctx, oracle, algo, kind):
    for n, seed in ((2, 411), (97, 412), (3001, 413)):
        check_ploc(ctx, oracle, degenerate(kind, n, seed), algo)


@pytest.mark.parametrize("kind", KINDS)
def test_degenerate_geometry_morton60_split_and_graph(ctx, oracle, kind):
    """The same inputs through the 60-bit codes (all four builders), through a replayed graph, and — where triangles have area — through early split clipping."""
    tris = degenerate(kind, 2500, 431)
    for algo in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC):
        check60(ctx, oracle, tris, algo)
    check_lbvh(ctx, oracle, tris, capi.SINGLE_PASS_LBVH, use_graph=True)
    check_lbvh(ctx, oracle, tris, capi.SINGLE_PASS_LBVH, use_graph=True)
    check_ploc(ctx, oracle, tris, capi.HPLOC, lbvh_second_level=1)  # H-PLOC's tile phase


def test_argument_errors_are_statuses_with_messages(ctx):
    tris = random_tris(64, 421)
    with pytest.raises(capi.B2bvhError, match="at least 2 primitives"):
        ctx.build(capi.SINGLE_PASS_LBVH, tris[:1])
    with pytest.raises(capi.B2bvhError, match="at least 2 primitives"):
        ctx.build(capi.PLOCPP, tris[:0])
    with pytest.raises(capi.B2bvhError):
        ctx.build(7, tris)                      # no such builder
    with pytest.raises(capi.B2bvhError, match="morton_bits"):
        ctx.build(capi.SINGLE_PASS_LBVH, tris, morton_bits=48)
    with pytest.raises(capi.B2bvhError):
        ctx.sort_pairs(np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint32))
    # the context is still usable after every refusal
    t = ctx.build(capi.SINGLE_PASS_LBVH, tris)
    assert t.n_prims == 64 and t.n_internal == 63


@pytest.mark.parametrize("kind", ["coincident", "points", "line", "mixed_scale", "one_point"])
@pytest.mark.parametrize("algo", [capi.SINGLE_PASS_LBVH, capi.PLOCPP], ids=["singlepass", "ploc"])
def test_degenerate_geometry_traversal(ctx, oracle, algo, kind):
    """Primary rays against the same degenerate scenes: HitInfo of the while-while kernel bit for bit against the oracle (zero-area triangles are
    never hit; among coincident triangles the first one met in traversal order wins on both sides), the other kernels agree on t."""
    from test_gpu_traverse import PI, compare_hits
    tris = degenerate(kind, 1500, 441)
    n = tris.size
    tr = T.make_transform([0.0, 2.5, -3.0], [3.0, 3.0, 3.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 2.5, 5.8, 0.0], oracle.qt_rotation([0.0, 0.0, 1.0, -1.57]), np.float32(45.0) * PI / np.float32(180.0))
    size = 96
    tree = ctx.build(algo, tris)
    g = ctx.fetch(tree)
    d_rays, _ = ctx.generate_rays(cam, size, size)
    rays = ctx.download(d_rays, T.RAY, size * size)
    o_hits, o_cnt = oracle.traverse(rays, g["nodes"], g["leaves"], tris, tr, tree.root, n)
    hits, _, _ = ctx.traverse(tree, d_rays, size * size, tr, capi.TRAVERSE_WHILE)
    compare_hits(hits, o_hits)
    assert int((hits["primIdx"] != 0xFFFFFFFF).sum()) == o_cnt
    if kind == "coincident":
        assert o_cnt > 0
    if kind in ("points", "one_point"):
        assert o_cnt == 0
    for k in (capi.TRAVERSE_SPECULATIVE_WHILE, capi.TRAVERSE_IFIF, capi.TRAVERSE_RESTART_TRAIL, capi.TRAVERSE_WIDE4):
        try:
            h2, _, _ = ctx.traverse(tree, d_rays, size * size, tr, k)
        except capi.B2bvhError as e:
            # PLOC++ over coincident triangles builds a tree deeper than the 64-bit restart trail (or a stack) holds: refused loudly, never a wrong image
            assert "deeper than" in str(e), e
            continue
        assert np.array_equal(h2["t"].view(np.uint32), hits["t"].view(np.uint32)), k
    ctx.free(d_rays)
