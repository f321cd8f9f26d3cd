"""Batched builder, CPU side: the oracle's plain Morton code against the reference's own computeMortonCode
(oracle/_ref/libref_emul.so, when built), and structural properties of orc_batched_build."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_mesh, random_tris

KA = json.load(open(os.path.join(GOLDEN, "batched_known_answers.json")))
HAVE_REF_KERNEL = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_emul_mt.so")) and os.path.isdir("/root/reference/src")


def h32(orc, a):
    return orc.fnv1a(np.ascontiguousarray(a).view(np.uint32).reshape(-1))


@pytest.mark.parametrize("key", [k for k in KA if not k.startswith("_")])
def test_batched_known_answers(oracle, key):
    kind, pc, items, seed = key.rsplit("_", 3)
    pc, items = int(pc), int(items)
    b = oracle.build_batched(random_tris(items * pc, int(seed), kind), np.full(items, pc, dtype=np.uint32))
    ka = KA[key]
    assert h32(oracle, b["nodes"]) == ka["nodes_fnv"] and h32(oracle, b["leaves"]) == ka["leaves_fnv"]
    assert h32(oracle, b["roots"]) == ka["roots_fnv"] and h32(oracle, b["scenes"]) == ka["scenes_fnv"]


@pytest.mark.skipif(not HAVE_REF_KERNEL, reason="oracle/_ref/libref_emul_mt.so with the batched kernel needs /root/reference")
@pytest.mark.parametrize("kind,pc,items,seed", [("uniform", 32, 64, 111), ("uniform", 5, 64, 112), ("clustered", 31, 30, 113), ("duplicate", 16, 30, 114), ("flat", 32, 30, 115)])
def test_batched_matches_reference_kernel(oracle, kind, pc, items, seed):
    """The reference's own BatchedBuildKernelLbvh (BatchedBuildKernel.h:218-312), run block by block by the fiber emulator with one token
    removed (`__shared__` in a parameter list) and ExtentCacheSize defined: on the input as given its topology, roots and scene boxes
    equal the restatement's (its boxes do not: the kernel joins sorted codes to unsorted leaves, :241-242); on the same items pre-sorted by
    code — where that defect is invisible — nodes, leaves, roots and scene boxes are byte-identical."""
    import ref
    tris = random_tris(items * pc, seed, kind)
    c = np.full(items, pc, dtype=np.uint32)
    b = oracle.build_batched(tris, c)
    n3, _, r3, s3 = ref.batched_build_mt(tris, items, pc)
    assert np.array_equal(n3["left"], b["nodes"]["left"]) and np.array_equal(n3["right"], b["nodes"]["right"]) and np.array_equal(r3, b["roots"])
    assert s3.tobytes() == b["scenes"].tobytes()
    perm = (b["leaves"]["primIdx"].reshape(items, pc) + (np.arange(items) * pc)[:, None]).reshape(-1)
    ts = np.ascontiguousarray(tris[perm])
    b2 = oracle.build_batched(ts, c)
    assert np.array_equal(b2["leaves"]["primIdx"].reshape(items, pc), np.tile(np.arange(pc, dtype=np.uint32), (items, 1)))
    n_, l_, r_, s_ = ref.batched_build_mt(ts, items, pc)
    assert n_.tobytes() == b2["nodes"].tobytes() and l_.tobytes() == b2["leaves"].tobytes()
    assert np.array_equal(r_, b2["roots"]) and s_.tobytes() == b2["scenes"].tobytes()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_emul.so")), reason="oracle/_ref not built (needs /root/reference)")
def test_plain_morton_matches_reference_function(oracle):
    import ref
    rng = np.random.default_rng(71)
    pts = rng.uniform(-0.05, 1.05, size=(5000, 3)).astype(np.float32)
    pts[:6] = [[0, 0, 0], [1, 1, 1], [0.999999, 0.5, 0.25], [np.nan, 0.5, 1.0], [-1, 2, 0.5], [np.inf, -np.inf, 0.1]]
    for p in pts:
        assert oracle.morton_plain(p) == ref.morton_plain(p), p
    assert oracle.morton_plain([1, 1, 1]) == 0x3FFFFFFF and oracle.morton_plain([0, 0, 0]) == 0


def test_plain_morton_known_answers(oracle):
    # 10 bits per axis, x highest: code = interleave(x,y,z) with x<<2 | y<<1 | z  (BatchedBuildKernel.h:98-110)
    assert oracle.morton_plain([1.0 / 1024, 0, 0]) == 4 and oracle.morton_plain([0, 1.0 / 1024, 0]) == 2 and oracle.morton_plain([0, 0, 1.0 / 1024]) == 1
    assert oracle.morton_plain([0.5, 0, 0]) == 4 << 27 and oracle.morton_plain([0, 0, 0.5]) == 1 << 27


def test_batched_items_equal_single_builds(oracle):
    """One item of the batch == the single-pass (Apetrei) builder's procedure on that item with the plain code."""
    rng = np.random.default_rng(72)
    counts = rng.integers(1, 33, size=60).astype(np.uint32)
    tris = random_tris(int(counts.sum()), 72)
    b = oracle.build_batched(tris, counts)
    for it in (0, 17, 59):
        n = int(counts[it]); lo = int(b["leaf_off"][it]); no = int(b["node_off"][it])
        item = np.ascontiguousarray(tris[lo:lo + n])
        _, boxes, scene = oracle.primrefs(item)
        assert b["scenes"][it].tobytes() == scene[0].tobytes()
        leaves = b["leaves"][lo:lo + n]
        assert sorted(leaves["primIdx"].tolist()) == list(range(n))
        assert np.array_equal(leaves["mn"], boxes["mn"][leaves["primIdx"]])
        if n > 1:
            root = b["nodes"][no + int(b["roots"][it])]
            assert np.array_equal(root["mn"], scene["mn"][0]) and np.array_equal(root["mx"], scene["mx"][0])
            kids = np.concatenate([b["nodes"][no:no + n - 1]["left"], b["nodes"][no:no + n - 1]["right"]])
            assert sorted(kids.tolist() + [int(b["roots"][it])]) == list(range(2 * n - 1))
