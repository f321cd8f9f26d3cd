"""GPU parity, LBVH builders: every buffer a build leaves on the device is compared byte for byte with the CPU oracle
(integer stages bit-exact; AABBs are min/max only, so also bit-exact; SAH cost identical as float32).  All calls go
through the C ABI (b2bvh.capi -> libb2bvh.so)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_mesh, random_tris
from b2bvh import capi, types as T

pytestmark = pytest.mark.gpu
KA = json.load(open(os.path.join(GOLDEN, "known_answers.json")))


def h32(orc, a):
    return orc.fnv1a(np.ascontiguousarray(a).view(np.uint32).reshape(-1))


def same_cost(a, b):
    """float32 equality; a scene whose root box has no area (a line, a point) costs 0/0 = NaN on both sides"""
    a, b = np.float32(a), np.float32(b)
    return bool(a == b or (np.isnan(a) and np.isnan(b)))


def assert_same_struct(a, b, what):
    assert a.dtype == b.dtype and a.shape == b.shape, what
    if a.tobytes() != b.tobytes():
        bad = np.nonzero(np.frombuffer(a.tobytes(), np.uint8).reshape(a.size, -1) != np.frombuffer(b.tobytes(), np.uint8).reshape(b.size, -1))[0]
        raise AssertionError(f"{what}: first differing element {bad[0]} of {a.size}: gpu={a[bad[0]]} oracle={b[bad[0]]}")


def check_lbvh(ctx, oracle, tris, algo, **kw):
    n = tris.size
    single = algo == capi.SINGLE_PASS_LBVH
    tree = ctx.build(algo, tris, **kw)
    g = ctx.fetch(tree)
    o = oracle.build_lbvh(tris, single_pass=single)
    assert_same_struct(g["boxes"], o["boxes"], "primitive boxes")
    assert_same_struct(g["scene"], o["scene"], "scene box")
    assert np.array_equal(g["keys"], o["keys"]), "morton keys"
    assert np.array_equal(g["vals"], np.arange(n, dtype=np.uint32))
    assert np.array_equal(g["skeys"], o["skeys"]) and np.array_equal(g["svals"], o["svals"]), "sorted pairs"
    assert g["root"] == o["root"]
    assert_same_struct(g["nodes"], o["nodes"], "bvh2 nodes")
    if not single:
        _, parents = oracle.lbvh_karras(o["refs"], o["skeys"], o["svals"])
        assert np.array_equal(g["parents"], parents), "parent indices"
    assert g["n_wide"] == o["wide_count"]
    assert_same_struct(g["wide"], o["wide"], "bvh4 nodes")
    assert_same_struct(g["wide_leaves"], o["wide_leaves"], "bvh4 leaves")
    cost = ctx.tree_cost(tree)
    assert same_cost(cost, o["cost"]), (cost, o["cost"])
    return tree, g, o


SYNTH = [("uniform", 2, 1), ("uniform", 3, 2), ("uniform", 33, 3), ("uniform", 1000, 4), ("uniform", 8192, 5), ("uniform", 8193, 6),
         ("uniform", 100_003, 7), ("clustered", 20_000, 8), ("flat", 5000, 9), ("duplicate", 700, 10), ("anisotropic", 30_000, 11)]


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH], ids=["twopass", "singlepass"])
@pytest.mark.parametrize("kind,n,seed", SYNTH, ids=[f"{k}-{n}" for k, n, _ in SYNTH])
def test_synthetic(ctx, oracle, algo, kind, n, seed):
    check_lbvh(ctx, oracle, random_tris(n, seed, kind), algo)


SECOND_LEVEL = [("uniform", 300, 31), ("uniform", 8193, 32), ("uniform", 100_003, 33), ("clustered", 20_000, 34), ("duplicate", 9000, 35),
                ("anisotropic", 30_000, 36)]


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH], ids=["twopass", "singlepass"])
@pytest.mark.parametrize("kind,n,seed", SECOND_LEVEL, ids=[f"{k}-{n}" for k, n, _ in SECOND_LEVEL])
def test_second_merge_level_forced(ctx, oracle, algo, kind, n, seed):
    """lbvh_group_kernel (automatic only from 2^20 primitives) forced on small inputs: same bytes as the oracle."""
    check_lbvh(ctx, oracle, random_tris(n, seed, kind), algo, lbvh_second_level=1)


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH], ids=["twopass", "singlepass"])
def test_second_merge_level_on_sponza(ctx, oracle, algo):
    tris = load_mesh("sponza")
    if tris is None:
        pytest.skip("sponza not staged")
    check_lbvh(ctx, oracle, tris, algo, lbvh_second_level=1)


@pytest.mark.parametrize("kind,n,seed", [("uniform", 5000, 21), ("clustered", 40_000, 22)])
def test_twopass_two_kernel_variant(ctx, oracle, kind, n, seed):
    check_lbvh(ctx, oracle, random_tris(n, seed, kind), capi.TWO_PASS_LBVH, karras_two_kernel=True)


def test_exchange_words_are_left_clean_between_builds(oracle):
    """The climb's exchange words (SLOT_MEET) are filled once and every build leaves them as it found them: sizes going up and down, both numberings,
    the all-global variant's neighbours (two-kernel Karras, second merge level) in between, graph replays — every tree against the oracle."""
    own = capi.Context(0)
    try:
        seq = [(30_000, capi.SINGLE_PASS_LBVH, {}), (700, capi.TWO_PASS_LBVH, {}), (30_000, capi.TWO_PASS_LBVH, {}), (2, capi.SINGLE_PASS_LBVH, {}),
               (90_001, capi.SINGLE_PASS_LBVH, {"lbvh_second_level": 1}), (5000, capi.TWO_PASS_LBVH, {"karras_two_kernel": True}),
               (90_001, capi.TWO_PASS_LBVH, {"lbvh_second_level": 2}), (257, capi.SINGLE_PASS_LBVH, {"use_graph": True}), (257, capi.SINGLE_PASS_LBVH, {"use_graph": True}),
               (120_000, capi.SINGLE_PASS_LBVH, {}), (30_000, capi.SINGLE_PASS_LBVH, {})]
        for i, (n, algo, kw) in enumerate(seq):
            check_lbvh(own, oracle, random_tris(n, 300 + i, "clustered" if i % 3 == 1 else "uniform"), algo, **kw)
    finally:
        own.close()


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH], ids=["twopass", "singlepass"])
@pytest.mark.parametrize("mesh", ["cornellbox", "bunny", "sponza"])
def test_reference_meshes(ctx, oracle, algo, mesh):
    tris = load_mesh(mesh)
    if tris is None:
        pytest.skip(f"{mesh} not staged")
    tree, g, o = check_lbvh(ctx, oracle, tris, algo)
    ka = KA[mesh]
    assert oracle.fnv1a(g["skeys"], g["svals"]) == ka["sorted_kv_fnv"]
    assert h32(oracle, g["wide"]) == ka["lbvh_wide_fnv"] and g["n_wide"] == ka["lbvh_wide_count"]
    assert np.float32(ctx.tree_cost(tree)) == np.float32(ka["lbvh_cost"])
    if "readme" in ka:
        assert ctx.tree_cost(tree) == pytest.approx(ka["readme"]["lbvh"], rel=1e-5)


def test_device_resident_input_and_repeatability(ctx, oracle):
    tris = random_tris(50_000, 31)
    d = ctx.upload(tris)
    try:
        t1 = ctx.build(capi.SINGLE_PASS_LBVH, d, n=tris.size, tris_on_device=True)
        a = ctx.fetch(t1)
        t2 = ctx.build(capi.SINGLE_PASS_LBVH, tris)
        b = ctx.fetch(t2)
        for k in ("nodes", "wide", "wide_leaves", "skeys", "svals"):
            assert a[k].tobytes() == b[k].tobytes(), k
    finally:
        ctx.free(d)


def test_large_synthetic_properties(ctx, oracle):
    """Full-size properties (2M primitives of synth_uniform_v1): sortedness + stability, permutation, root box, tree validity, and
    the Karras and Apetrei numberings describe the same ordered tree."""
    n = 2_000_000
    tris = oracle.synth_uniform(n, 0x00B20010)
    tree = ctx.build(capi.TWO_PASS_LBVH, tris)
    g = ctx.fetch(tree)
    sk, sv = g["skeys"], g["svals"]
    assert np.all(sk[1:] >= sk[:-1])
    eq = sk[1:] == sk[:-1]
    assert np.all(sv[1:][eq] > sv[:-1][eq]), "stable: equal keys keep index order"
    assert np.array_equal(np.sort(sv), np.arange(n, dtype=np.uint32))
    assert np.array_equal(g["keys"][sv], sk)
    assert oracle.check_root_aabb(g["nodes"], 0, n) and oracle.check_bvh2(g["nodes"], None, 0, n)
    assert oracle.check_bvh4(g["wide"], g["wide_leaves"], 0, n)
    o_nodes, _ = oracle.lbvh_karras(oracle.primrefs(tris)[0], sk, sv)
    assert g["nodes"].tobytes() == o_nodes.tobytes()
    t2 = ctx.build(capi.SINGLE_PASS_LBVH, tris)
    g2 = ctx.fetch(t2)
    assert g2["wide"].tobytes() == g["wide"].tobytes() and g2["wide_leaves"].tobytes() == g["wide_leaves"].tobytes()
    assert np.float32(ctx.tree_cost(t2)) == np.float32(ctx.tree_cost(tree))


@pytest.mark.parametrize("n", [1, 5, 3071, 3072, 8191, 8192, 8193, 70_001, 1_000_000])
@pytest.mark.parametrize("bits", [(0, 32), (0, 30), (4, 20), (0, 8)])
def test_sort_pairs_is_stable_sort(ctx, n, bits):
    """Oro::RadixSort contract (Orochi Test/RadixSort/main.cpp:130,239): equal to std::stable_sort on the selected bits."""
    rng = np.random.default_rng(n * 31 + bits[1])
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if n > 100:
        keys[rng.integers(0, n, size=n // 3)] = keys[0]  # heavy duplicates
    vals = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    ko, vo = ctx.sort_pairs(keys, vals, *bits)
    mask = np.uint32(((1 << (bits[1] - bits[0])) - 1) << bits[0])
    order = np.argsort(keys & mask, kind="stable")
    assert np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])


def test_device_synth_generator_matches_oracle(ctx, oracle):
    from b2bvh import types as T
    n_total, first, count = 10_000_000, 9_999_000, 1000
    d = ctx.synth_uniform(n_total, 0x00B20010, first=first, count=count)
    try:
        g = ctx.download(d, T.TRIANGLE, count)
    finally:
        ctx.free(d)
    assert g.tobytes() == oracle.synth_uniform(n_total, 0x00B20010, first=first, count=count).tobytes()


def test_clustered_generator_and_build(ctx, oracle):
    """synth_clustered_v1 (many primitives per Morton cell): device stream == oracle stream bit for bit, and the builders agree with the
    oracle on it (thousands of equal keys, ordered by index)."""
    n = 60_000
    d = ctx.synth_uniform(n, 0x00B20010, clustered=True)
    dev = ctx.download(d, T.TRIANGLE, n)
    tris = oracle.synth_clustered(n, 0x00B20010)
    assert dev.tobytes() == tris.tobytes()
    part = ctx.synth_uniform(n, 0x00B20010, first=1234, count=777, clustered=True)
    assert ctx.download(part, T.TRIANGLE, 777).tobytes() == tris[1234:1234 + 777].tobytes()
    tree, g, o = check_lbvh(ctx, oracle, tris, capi.SINGLE_PASS_LBVH)
    assert int((g["skeys"][1:] == g["skeys"][:-1]).sum()) > 100
    check_lbvh(ctx, oracle, tris, capi.TWO_PASS_LBVH)
    ctx.free(d); ctx.free(part)


def test_profiler_reports_every_launch(ctx):
    tris = random_tris(10_000, 77)
    ctx.profile(True)
    tree = ctx.build(capi.SINGLE_PASS_LBVH, tris)
    ctx.sync()
    entries = ctx.profile_entries()
    ctx.profile(False)
    assert len(entries) == tree.n_launches
    names = [e[0] for e in entries]
    assert names[:3] == ["primref_extents", "morton30", "radix_count"] and names.count("radix_scatter") == 4
    assert all(ms >= 0 for _, ms in entries)


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC], ids=["twopass", "singlepass", "ploc", "hploc"])
def test_graph_replay_builds_the_same_tree(ctx, oracle, algo):
    """opts.use_graph: the launch sequence is captured once and replayed for the same (algorithm, size, options, triangle
    pointer); every buffer equals the plain build, a different size or pointer re-captures, stage times are still reported."""
    from b2bvh import types as T
    tris = random_tris(30_011, 41)
    d = ctx.upload(tris)
    plain = ctx.fetch(ctx.build(algo, d, n=tris.size, tris_on_device=True))
    for rep in range(3):  # capture + two replays
        tree = ctx.build(algo, d, n=tris.size, tris_on_device=True, use_graph=True)
        g = ctx.fetch(tree)
        for k in ("skeys", "svals", "nodes", "wide", "wide_leaves"):
            assert g[k].tobytes() == plain[k].tobytes(), (rep, k)
        assert g["root"] == plain["root"] and g["n_wide"] == plain["n_wide"]
        assert tree.build_ms > 0 and tree.stage_ms[capi.T_SORT] > 0 and tree.n_launches > 5
    tris2 = random_tris(12_345, 42)  # another size and pointer on the same context: the cached graph must not be replayed
    d2 = ctx.upload(tris2)
    a = ctx.fetch(ctx.build(algo, d2, n=tris2.size, tris_on_device=True, use_graph=True))
    b = ctx.fetch(ctx.build(algo, d2, n=tris2.size, tris_on_device=True))
    assert a["nodes"].tobytes() == b["nodes"].tobytes() and a["wide"].tobytes() == b["wide"].tobytes()
    h = ctx.fetch(ctx.build(algo, tris, use_graph=True))  # host triangles: the upload is part of the graph
    assert h["nodes"].tobytes() == plain["nodes"].tobytes()
    ctx.free(d); ctx.free(d2)


def test_baseline_config_10m_properties(ctx):
    """BASELINE configs[3] at full size — 10 M synth_uniform_v1 triangles, single-pass LBVH + Bvh4 collapse (the bench workload) —
    through size-independent properties, vectorised: sorted + stable + a permutation; every Bvh2 box is the union of its children;
    every node has exactly one parent and the root none; the root box is the scene box; Bvh4: child slots are packed, every wide node
    but the root is the child of exactly one wide node, every leaf slot appears once, parents point back."""
    n = 10_000_000
    d = ctx.synth_uniform(n, 0x00B20010)
    tree = ctx.build(capi.SINGLE_PASS_LBVH, d, n=n, tris_on_device=True)
    nInt = n - 1
    sk = ctx.download(tree.d_sortedMortonCodeKeys, np.uint32, n)
    sv = ctx.download(tree.d_sortedMortonCodeValues, np.uint32, n)
    keys = ctx.download(tree.d_mortonCodeKeys, np.uint32, n)
    assert np.all(sk[1:] >= sk[:-1])
    eq = sk[1:] == sk[:-1]
    assert np.all(sv[1:][eq] > sv[:-1][eq])
    seen = np.zeros(n, dtype=np.uint8); seen[sv] = 1
    assert seen.all() and np.array_equal(keys[sv], sk)
    del keys, seen, eq
    nodes = ctx.download(tree.d_bvhNodes, T.BVH2_NODE, 2 * n - 1)
    boxes = ctx.download(tree.d_triangleAabb, T.AABB, n)
    assert np.array_equal(nodes["left"][nInt:], sv) and (nodes["right"][nInt:] == 0xFFFFFFFF).all()
    assert np.array_equal(nodes["mn"][nInt:], boxes["mn"][sv]) and np.array_equal(nodes["mx"][nInt:], boxes["mx"][sv])
    l, r = nodes["left"][:nInt].astype(np.int64), nodes["right"][:nInt].astype(np.int64)
    assert l.max() < 2 * n - 1 and r.max() < 2 * n - 1
    assert np.array_equal(nodes["mn"][:nInt], np.minimum(nodes["mn"][l], nodes["mn"][r]))
    assert np.array_equal(nodes["mx"][:nInt], np.maximum(nodes["mx"][l], nodes["mx"][r]))
    refs = np.bincount(np.concatenate([l, r]), minlength=2 * n - 1)
    assert refs[tree.root] == 0 and (np.delete(refs, tree.root) == 1).all()
    scene = ctx.download(tree.d_sceneExtents, T.AABB, 1)
    assert np.array_equal(nodes["mn"][tree.root], scene["mn"][0]) and np.array_equal(nodes["mx"][tree.root], scene["mx"][0])
    del refs, boxes
    nw = tree.n_wide
    wide = ctx.download(tree.d_wideBvhNodes, T.BVH4_NODE, nw)
    wl = ctx.download(tree.d_wideLeafNodes, T.PRIM_NODE, n)
    ch = wide["child"]
    valid = ch != 0xFFFFFFFF
    cnt = valid.sum(axis=1)
    assert np.array_equal(cnt, wide["childCount"]) and cnt.min() >= 2
    assert (valid[:, :-1] >= valid[:, 1:]).all()                      # packed: no hole before a used slot
    internal = valid & (ch < nInt)
    leaf = valid & (ch >= nInt)
    ic = ch[internal]
    assert ic.max() < nw and np.array_equal(np.sort(ic), np.arange(1, nw, dtype=np.uint32))   # BFS numbering: every non-root node once
    owner = np.repeat(np.arange(nw, dtype=np.uint32), 4).reshape(nw, 4)
    assert np.array_equal(wide["parent"][ic], owner[internal]) and wide["parent"][0] == 0xFFFFFFFF
    lc = ch[leaf] - nInt
    assert np.array_equal(np.sort(lc), np.arange(n, dtype=np.uint32))
    assert np.array_equal(wl["primIdx"][lc], sv[lc]) and np.array_equal(wl["parent"][lc], owner[leaf])
    # internal child boxes of a wide node are the Bvh2 boxes... of SOME Bvh2 node: check containment in the root box and non-emptiness
    ib = wide["aabb"][internal]
    assert (ib[:, :3] <= ib[:, 3:]).all() and (ib[:, :3] >= scene["mn"][0]).all() and (ib[:, 3:] <= scene["mx"][0]).all()
    ctx.free(d)


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC], ids=["twopass", "singlepass", "ploc", "hploc"])
def test_deferred_synchronisation_and_root_box(ctx, oracle, algo):
    """defer_sync: b2bvh_build returns after enqueueing; b2bvh_build_finish delivers root / wide-node count / times; the root box is left
    on the device by the build itself.  Same bytes as the synchronous call."""
    tris = random_tris(50_000, 121)
    ref_tree = ctx.build(algo, tris)
    want = ctx.fetch(ref_tree)
    d_box = ctx.alloc(24)
    tree = ctx.build(algo, tris, defer_sync=True, d_root_box_out=d_box)
    assert tree.n_prims == tris.size and tree.d_bvhNodes and tree.n_wide == 0 and tree.build_ms == 0.0   # not known yet
    ctx.build_finish(tree)
    assert tree.root == ref_tree.root and tree.n_wide == ref_tree.n_wide and tree.build_ms > 0
    got = ctx.fetch(tree)
    assert got["nodes"].tobytes() == want["nodes"].tobytes() and got["wide"].tobytes() == want["wide"].tobytes()
    box = ctx.download(d_box, np.float32, 6)
    root = got["nodes"][tree.root]
    assert np.array_equal(box, np.concatenate([root["mn"], root["mx"]]))
    with pytest.raises(capi.B2bvhError, match="no build is waiting"):
        ctx.build_finish(tree)
    ctx.free(d_box)


@pytest.mark.parametrize("kind,n,seed", [("uniform", 12_000, 161), ("clustered", 8000, 162), ("duplicate", 1500, 163)])
def test_hierarchy_stage_per_range_reproduces_the_one_gpu_tree(ctx, oracle, kind, n, seed):
    """b2bvh_lbvh_from_sorted64 — the device's tile / group / climb kernels over 64-bit keys — used the way a rank of the globally sorted
    build would use it (oracle.lbvh_by_ranges_with_unchanged_builder: range + ghost leaves, keys = code << 32 | global position): after
    the host-side stitching of the few left-over clusters the nodes equal the one-GPU build byte for byte, for both numberings."""
    tris = random_tris(n, seed, kind)
    refs, boxes, _ = oracle.primrefs(tris)
    d_boxes = ctx.upload(boxes)
    rng = np.random.default_rng(seed)
    cuts = sorted(rng.choice(np.arange(1, n), size=3, replace=False).tolist())
    ranges = list(zip([0] + cuts, cuts + [n]))
    for karras in (True, False):
        want = ctx.fetch(ctx.build(capi.TWO_PASS_LBVH if karras else capi.SINGLE_PASS_LBVH, tris, collapse=False))
        nodes, root, leftovers = oracle.lbvh_by_ranges_with_unchanged_builder(
            tris, refs, want["skeys"], want["svals"], ranges, karras, builder=lambda k64, v: ctx.lbvh_from_sorted64(k64, v, d_boxes, karras))
        assert nodes.tobytes() == want["nodes"].tobytes() and root == want["root"], (karras, leftovers)
    ctx.free(d_boxes)


@pytest.mark.parametrize("kind,n,seed", [("uniform", 12_000, 171), ("clustered", 8000, 172), ("duplicate", 1500, 173), ("uniform", 40, 174)])
def test_range_trees_extracted_on_the_device(ctx, oracle, kind, n, seed):
    """Both device building blocks of the globally sorted build together (b2bvh_lbvh_from_sorted64 + b2bvh_range_extract), rank by rank
    on one GPU: the ghost-free nodes come back with global indices, the spine artefacts are dropped, the left-over clusters (a handful per
    rank) carry position range, node and box; after stitching them the node array equals the one-GPU build byte for byte."""
    tris = random_tris(n, seed, kind)
    refs, boxes, _ = oracle.primrefs(tris)
    d_boxes = ctx.upload(boxes)
    rng = np.random.default_rng(seed)
    cuts = sorted(rng.choice(np.arange(2, n - 1), size=3, replace=False).tolist())
    ranges = [(a, b) for a, b in zip([0] + cuts, cuts + [n]) if b - a >= 1]
    nint = n - 1
    for karras in (True, False):
        want = ctx.fetch(ctx.build(capi.TWO_PASS_LBVH if karras else capi.SINGLE_PASS_LBVH, tris, collapse=False))
        sk, sv = want["skeys"], want["svals"]
        nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE)
        nodes["left"][:] = 0xFFFFFFFF; nodes["right"][:] = 0xFFFFFFFF
        clusters, per_rank = [], []
        for a, b in ranges:
            a2, b2 = max(a - 1, 0), min(b + 1, n)
            m = b2 - a2
            k64 = (sk[a2:b2].astype(np.uint64) << np.uint64(32)) | np.arange(a2, b2, dtype=np.uint64)
            out, cl = ctx.range_tree(k64, sv[a2:b2], d_boxes, karras, a > 0, b < n, a2, n)
            inner = out[:m - 1]
            ok = inner["left"] != 0xFFFFFFFF
            nodes[a2:a2 + m - 1][ok] = inner[ok]                       # internal node i of the range is global node a2 + i
            nodes[nint + a:nint + b] = out[m - 1 + (a - a2):m - 1 + (a - a2) + (b - a)]   # the rank's own leaves (ghosts dropped)
            assert (cl["lo"][1:] == cl["hi"][:-1]).all() and cl["lo"][0] == a and cl["hi"][-1] == b   # the left-overs tile the range
            for c in cl:
                assert np.array_equal(nodes[c["node"]]["mn"], c["mn"]) and np.array_equal(nodes[c["node"]]["mx"], c["mx"])
            clusters += [(int(c["lo"]), int(c["hi"]), int(c["node"])) for c in cl]
            per_rank.append(cl.size)
        root = oracle.stitch_leftovers(nodes, clusters, sk, karras)
        assert nodes.tobytes() == want["nodes"].tobytes() and root == want["root"], (karras, per_rank)
        assert max(per_rank) <= 128
    ctx.free(d_boxes)

