"""Executable specification of the globally sorted multi-GPU build (the next step of SURVEY §8f 4): on the CPU, the
procedure G ranks would follow — exchange by key interval + local stable sorts, then range-wise cluster merging with the left-over
clusters finished after a gather — yields exactly the one-GPU sorted order and the one-GPU tree (both numberings, both key widths)."""
import numpy as np
import pytest

from conftest import random_tris


@pytest.mark.parametrize("kind,n,seed", [("uniform", 4000, 131), ("clustered", 3000, 132), ("duplicate", 500, 133), ("anisotropic", 2500, 134)])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_exchange_and_local_sorts_equal_the_global_stable_sort(oracle, kind, n, seed, world):
    tris = random_tris(n, seed, kind)
    refs, _, scene = oracle.primrefs(tris)
    keys, vals = oracle.morton_codes(refs, scene)
    sk, sv = oracle.sort_kv(keys, vals)
    rng = np.random.default_rng(seed)
    # splitters from a sample, as a sample sort would pick them (any non-decreasing choice must work, including repeated ones)
    sample = np.sort(rng.choice(keys, size=min(n, 64 * world), replace=False))
    splitters = sample[(np.arange(1, world) * sample.size) // world]
    gk, gv, counts = oracle.global_sort_by_exchange(keys, world, splitters)
    assert np.array_equal(gk, sk) and np.array_equal(gv, sv) and sum(counts) == n
    gk, gv, _ = oracle.global_sort_by_exchange(keys, world, np.full(world - 1, splitters[0]))  # degenerate: everything on two ranks
    assert np.array_equal(gk, sk) and np.array_equal(gv, sv)


@pytest.mark.parametrize("kind,n,seed", [("uniform", 2500, 141), ("clustered", 2000, 142), ("duplicate", 400, 143), ("flat", 1200, 144)])
@pytest.mark.parametrize("bits", [30, 60])
def test_rangewise_merging_builds_the_one_gpu_tree(oracle, kind, n, seed, bits):
    tris = random_tris(n, seed, kind)
    rng = np.random.default_rng(seed)
    cuts = sorted(rng.choice(np.arange(1, n), size=int(rng.integers(1, 8)), replace=False).tolist())
    ranges = list(zip([0] + cuts, cuts + [n]))
    for karras in (True, False):
        o = oracle.build_lbvh(tris, single_pass=not karras, morton_bits=bits)
        nodes, root, leftovers = oracle.lbvh_by_ranges(o["refs"], o["skeys"], o["svals"], ranges, karras)
        assert nodes.tobytes() == o["nodes"].tobytes() and root == o["root"]
        # what crosses the wire in the gather: a few dozen 48-byte records per rank, independent of n
        assert max(leftovers) <= 2 * (96 if bits == 60 else 64)


@pytest.mark.parametrize("kind,n,seed", [("uniform", 2000, 151), ("clustered", 1500, 152), ("duplicate", 300, 153), ("anisotropic", 1800, 154)])
def test_rangewise_build_with_an_unchanged_single_range_builder(oracle, kind, n, seed):
    """The variant that needs no new hierarchy kernel: every rank runs the ordinary 64-bit-key builder over its range + one ghost leaf
    per inner edge, with keys widened to (code << 32 | global position); ghost-free nodes are the one-GPU tree's nodes."""
    tris = random_tris(n, seed, kind)
    rng = np.random.default_rng(seed)
    cuts = sorted(rng.choice(np.arange(1, n), size=int(rng.integers(1, 8)), replace=False).tolist())
    ranges = list(zip([0] + cuts, cuts + [n]))
    for karras in (True, False):
        o = oracle.build_lbvh(tris, single_pass=not karras)
        nodes, root, leftovers = oracle.lbvh_by_ranges_with_unchanged_builder(tris, o["refs"], o["skeys"], o["svals"], ranges, karras)
        assert nodes.tobytes() == o["nodes"].tobytes() and root == o["root"] and max(leftovers) <= 128

