"""GPU parity, ray generation + while-while / speculative-while traversal: HitInfo per ray against the oracle's CPU
traversal on the SAME ray buffer (rays are generated once and shared).  primIdx exact; t, u, v
bit-exact (the kernels evaluate the reference's expressions in the reference's order, no FMA)."""
import numpy as np
import pytest

from conftest import load_mesh, random_tris
from b2bvh import capi, types as T

pytestmark = pytest.mark.gpu
PI = np.float32(3.14159265358979323846)

PRESETS = {  # Common.h:26-77
    "cornellbox": dict(t=[0.0, 0.0, -5.0], s=[1.0, 1.0, 1.0], q=None, eye=[0.0, 2.5, 5.8, 0.0], cq=[0.0, 0.0, 1.0, -1.57]),
    "bunny": dict(t=[0.0, 0.0, -3.0], s=[3.0, 3.0, 3.0], q=None, eye=[0.0, 2.5, 5.8, 0.0], cq=[0.0, 0.0, 1.0, -1.57]),
    "sponza": dict(t=[0.0, 0.0, -3.0], s=[1.0, 1.0, 1.0], q=[1.0, 0.0, 0.0, 1.57], eye=[-20.0, 18.5, 10.8, 0.0], cq=[0.0, 1.0, 0.0, -1.57]),
}


def scene(oracle, name):
    p = PRESETS[name]
    q = [0.0, 0.0, 0.0, 1.0] if p["q"] is None else oracle.qt_rotation(p["q"])
    tr = T.make_transform(p["t"], p["s"], q)
    cam = T.make_camera(p["eye"], oracle.qt_rotation(p["cq"]), np.float32(45.0) * PI / np.float32(180.0))
    return tr, cam


def compare_hits(g, o, exact=True):
    assert np.array_equal(g["primIdx"], o["primIdx"]), f"{int((g['primIdx'] != o['primIdx']).sum())} rays hit a different primitive"
    hit = o["primIdx"] != 0xFFFFFFFF
    for f in ("t", "uv"):
        a, b = g[f][hit], o[f][hit]
        if exact:
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f
        else:
            assert np.allclose(a, b, rtol=1e-6, atol=0)


def test_generate_rays_matches_oracle(ctx, oracle):
    for name in PRESETS:
        _, cam = scene(oracle, name)
        d_rays, _ = ctx.generate_rays(cam, 128, 128)
        g = ctx.download(d_rays, T.RAY, 128 * 128)
        ctx.free(d_rays)
        o = oracle.generate_rays(cam, 128, 128)
        for f in ("origin", "direction", "tMin", "tMax"):
            assert np.array_equal(g[f].view(np.uint32), o[f].view(np.uint32)), (name, f)


@pytest.mark.parametrize("mesh,size", [("cornellbox", 512), ("bunny", 256), ("sponza", 128)])
@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC], ids=["twopass", "singlepass", "ploc", "hploc"])
def test_traversal_matches_oracle(ctx, oracle, mesh, size, algo):
    tris = load_mesh(mesh)
    if tris is None:
        pytest.skip(f"{mesh} not staged")
    n = tris.size
    tr, cam = scene(oracle, mesh)
    tree = ctx.build(algo, tris, collapse=False)
    g = ctx.fetch(tree)
    d_rays, _ = ctx.generate_rays(cam, size, size)
    rays = ctx.download(d_rays, T.RAY, size * size)
    o_hits, o_cnt = oracle.traverse(rays, g["nodes"], g["leaves"], tris, tr, tree.root, n)
    hits, rgba, ms = ctx.traverse(tree, d_rays, size * size, tr, capi.TRAVERSE_WHILE, want_rgba=True)
    compare_hits(hits, o_hits)
    assert int((hits["primIdx"] != 0xFFFFFFFF).sum()) == o_cnt and o_cnt > 0
    assert np.array_equal(rgba[:, 3] == 255, hits["primIdx"] != 0xFFFFFFFF)
    # speculative variant: same closest hits (leaf tests are only re-ordered)
    hits2, _, _ = ctx.traverse(tree, d_rays, size * size, tr, capi.TRAVERSE_SPECULATIVE_WHILE)
    assert np.array_equal(hits2["t"].view(np.uint32), hits["t"].view(np.uint32))
    assert (hits2["primIdx"] != hits["primIdx"]).mean() < 1e-4
    ctx.free(d_rays)


@pytest.mark.parametrize("mesh,size", [("cornellbox", 256), ("bunny", 192), ("sponza", 128)])
@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC], ids=["twopass", "singlepass", "ploc", "hploc"])
def test_ifif_restart_trail_and_bvh4_traversal(ctx, oracle, mesh, size, algo):
    """The reference's if-if and restart-trail kernels and the Bvh4 walk: HitInfo and the per-ray triangle-test counter
    (rayCounter) bit-exact against the oracle's restatement of each; all of them find the hits of the while-while kernel."""
    tris = load_mesh(mesh)
    if tris is None:
        pytest.skip(f"{mesh} not staged")
    n = tris.size
    tr, cam = scene(oracle, mesh)
    tree = ctx.build(algo, tris, collapse=True)
    g = ctx.fetch(tree)
    d_rays, _ = ctx.generate_rays(cam, size, size)
    rays = ctx.download(d_rays, T.RAY, size * size)
    ref_hits, _, _ = ctx.traverse(tree, d_rays, size * size, tr, capi.TRAVERSE_WHILE)
    for kind, kernel in ((0, capi.TRAVERSE_IFIF), (1, capi.TRAVERSE_RESTART_TRAIL)):
        o_hits, o_cnt, o_counter = oracle.traverse_kind(kind, rays, g["nodes"], g["leaves"], tris, tr, tree.root, n)
        hits, rgba, _, counter = ctx.traverse(tree, d_rays, size * size, tr, kernel, want_rgba=True, want_counter=True)
        compare_hits(hits, o_hits)
        assert np.array_equal(counter, o_counter), f"kernel {kernel}: ray counters differ on {int((counter != o_counter).sum())} rays"
        assert np.array_equal(rgba[:, 3] == 255, hits["primIdx"] != 0xFFFFFFFF)
        assert np.array_equal(hits["t"].view(np.uint32), ref_hits["t"].view(np.uint32))
        assert capi.heat_map(counter).tobytes() == oracle.heat_map(o_counter).tobytes()
    o_hits, o_cnt, o_counter = oracle.traverse_wide4(rays, g["wide"], g["nodes"], g["leaves"], tris, tr, n)
    hits, _, _, counter = ctx.traverse(tree, d_rays, size * size, tr, capi.TRAVERSE_WIDE4, want_counter=True)
    compare_hits(hits, o_hits)
    assert np.array_equal(counter, o_counter)
    assert o_cnt > 0 and (hits["t"].view(np.uint32) != ref_hits["t"].view(np.uint32)).mean() < 1e-4
    ctx.free(d_rays)


def test_traversal_argument_errors(ctx):
    tris = random_tris(2000, 9)
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 400.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.6))
    d_rays, _ = ctx.generate_rays(cam, 16, 16)
    tree = ctx.build(capi.TWO_PASS_LBVH, tris, collapse=False)
    with pytest.raises(capi.B2bvhError):
        ctx.traverse(tree, d_rays, 256, tr, capi.TRAVERSE_WIDE4)  # no 4-wide nodes were built
    with pytest.raises(capi.B2bvhError):
        ctx.traverse(tree, d_rays, 256, tr, capi.TRAVERSE_WHILE, want_counter=True)  # the while-while kernels keep no counter
    with pytest.raises(capi.B2bvhError):
        ctx.traverse(tree, d_rays, 256, tr, 7)
    ctx.free(d_rays)


def test_traversal_synthetic_scene(ctx, oracle):
    tris = random_tris(20_000, 5)
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 400.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.6))
    tree = ctx.build(capi.TWO_PASS_LBVH, tris, collapse=False)
    g = ctx.fetch(tree)
    d_rays, _ = ctx.generate_rays(cam, 256, 256)
    rays = ctx.download(d_rays, T.RAY, 256 * 256)
    o_hits, cnt = oracle.traverse(rays, g["nodes"], None, tris, tr, 0, tris.size)
    hits, _, _ = ctx.traverse(tree, d_rays, 256 * 256, tr)
    compare_hits(hits, o_hits)
    assert cnt > 1000
    ctx.free(d_rays)


def test_top_level_tree_matches_oracle(ctx, oracle):
    rng = np.random.default_rng(3)
    for g in (1, 2, 3, 8, 64):
        lo = rng.uniform(-100, 100, size=(g, 3)).astype(np.float32)
        boxes = np.zeros(g, dtype=T.AABB)
        boxes["mn"] = lo
        boxes["mx"] = lo + rng.uniform(1, 50, size=(g, 3)).astype(np.float32)
        d_in = ctx.upload(boxes)
        d_out = ctx.alloc((2 * g - 1) * 32)
        capi.check(ctx.lib.b2bvh_top_level(ctx.h, d_in, g, d_out), "b2bvh_top_level")
        got = ctx.download(d_out, T.BVH2_NODE, 2 * g - 1)
        assert got.tobytes() == oracle.top_level(boxes).tobytes(), g
        ctx.free(d_in); ctx.free(d_out)
