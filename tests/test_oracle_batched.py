"""Batched builder, CPU side: the oracle's plain Morton code against the reference's own computeMortonCode
(oracle/_ref/libref_emul.so, when built), and structural properties of orc_batched_build."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_mesh, random_tris


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
