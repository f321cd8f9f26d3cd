"""GPU parity, batched builder (BatchedBvhBuilder::build, BatchedBuilder.cpp:16-77 + BatchedBuildKernelLbvh,
BatchedBuildKernel.h:218-312): every item's nodes, leaves, root and scene box byte for byte against the oracle's restatement
(oracle orc_batched_build: plain Morton code pinned to the reference's computeMortonCode, Apetrei build pinned to the emulated
SinglePassLbvh kernel, stable sort)."""
import numpy as np
import pytest

from conftest import load_mesh, random_tris
from b2bvh import capi, types as T
from test_gpu_lbvh import assert_same_struct

pytestmark = pytest.mark.gpu


def check_batch(ctx, oracle, tris, counts):
    b = ctx.build_batched(tris, counts)
    g = ctx.fetch_batch(b)
    o = oracle.build_batched(tris, counts)
    assert b.n_items == len(counts) and b.n_prims_total == tris.size and b.n_nodes_total == tris.size - len(counts)
    assert np.array_equal(g["leaf_off"], o["leaf_off"]) and np.array_equal(g["node_off"], o["node_off"])
    assert_same_struct(g["scenes"], o["scenes"], "item scene boxes")
    assert_same_struct(g["leaves"], o["leaves"], "item leaves (sorted PrimRefs)")
    assert np.array_equal(g["roots"], o["roots"]), "item roots"
    assert_same_struct(g["nodes"], o["nodes"], "item nodes")
    return b, g, o


def test_batched_golden_cases(ctx, oracle):
    """The committed known answers (each checked against the emulated reference kernel when they were generated)."""
    import json
    import os
    from conftest import GOLDEN
    from test_gpu_lbvh import h32
    ka = json.load(open(os.path.join(GOLDEN, "batched_known_answers.json")))
    for key, want in ka.items():
        if key.startswith("_"):
            continue
        kind, pc, items, seed = key.rsplit("_", 3)
        pc, items = int(pc), int(items)
        _, g, _ = check_batch(ctx, oracle, random_tris(items * pc, int(seed), kind), np.full(items, pc, dtype=np.uint32))
        assert h32(oracle, g["nodes"]) == want["nodes_fnv"] and h32(oracle, g["leaves"]) == want["leaves_fnv"] and h32(oracle, g["roots"]) == want["roots_fnv"]


def test_batched_cornell_box_copies(ctx, oracle):
    """main.cpp:38-52: the cornell box (32 triangles) as every item of the batch."""
    box = load_mesh("cornellbox")
    assert box.size == 32
    tris = np.ascontiguousarray(np.tile(box, 64))
    _, g, _ = check_batch(ctx, oracle, tris, np.full(64, 32, dtype=np.uint32))
    assert (g["roots"] == g["roots"][0]).all() and g["nodes"][:31].tobytes() == g["nodes"][31:62].tobytes()


@pytest.mark.parametrize("kind,seed", [("uniform", 61), ("clustered", 62), ("flat", 63), ("duplicate", 64), ("anisotropic", 65)])
def test_batched_ragged_items(ctx, oracle, kind, seed):
    """Ragged item sizes 1..32 (the reference's offsets only hold for equal sizes), every input kind incl. equal codes and zero extents."""
    rng = np.random.default_rng(seed)
    counts = rng.integers(1, 33, size=700).astype(np.uint32)
    counts[:8] = [1, 2, 3, 31, 32, 32, 1, 2]
    tris = random_tris(int(counts.sum()), seed, kind)
    check_batch(ctx, oracle, tris, counts)


def test_batched_items_are_valid_trees(ctx):
    """200 K items (6.4 M triangles): each root box is the item's scene box, every node box is the union of its children,
    every leaf slot of an item holds a distinct primitive of the item — size-independent properties at a size the oracle skips."""
    n_items = 200_000
    counts = np.full(n_items, 32, dtype=np.uint32)
    d = ctx.synth_uniform(n_items * 32, 0xB20010)
    b = ctx.build_batched(d, counts, tris_on_device=True)
    g = ctx.fetch_batch(b)
    nodes = g["nodes"].reshape(n_items, 31)
    leaves = g["leaves"].reshape(n_items, 32)
    rows = np.arange(n_items)
    root = nodes[rows, g["roots"]]
    assert np.array_equal(root["mn"], g["scenes"]["mn"]) and np.array_equal(root["mx"], g["scenes"]["mx"])
    assert np.array_equal(np.sort(leaves["primIdx"], axis=1), np.tile(np.arange(32, dtype=np.uint32), (n_items, 1)))
    # child boxes: index < 31 -> internal node, else leaf (index - 31)
    allmn = np.concatenate([nodes["mn"], leaves["mn"]], axis=1)  # (items, 63, 3): local index space
    allmx = np.concatenate([nodes["mx"], leaves["mx"]], axis=1)
    l, r = nodes["left"].astype(np.int64), nodes["right"].astype(np.int64)
    assert l.max() < 63 and r.max() < 63
    lmn = np.take_along_axis(allmn, l[:, :, None], axis=1); rmn = np.take_along_axis(allmn, r[:, :, None], axis=1)
    lmx = np.take_along_axis(allmx, l[:, :, None], axis=1); rmx = np.take_along_axis(allmx, r[:, :, None], axis=1)
    assert np.array_equal(nodes["mn"], np.minimum(lmn, rmn)) and np.array_equal(nodes["mx"], np.maximum(lmx, rmx))
    # every local index except the root is the child of exactly one node
    seen = np.zeros((n_items, 63), dtype=np.int32)
    np.add.at(seen, (rows[:, None], l), 1); np.add.at(seen, (rows[:, None], r), 1)
    seen[rows, g["roots"]] += 1
    assert (seen == 1).all()
    assert b.build_ms > 0
    ctx.free(d)


def test_batched_argument_errors(ctx):
    tris = random_tris(40, 66)
    with pytest.raises(capi.B2bvhError, match="1..32"):
        ctx.build_batched(tris, np.array([33, 7], dtype=np.uint32))
    with pytest.raises(capi.B2bvhError, match="1..32"):
        ctx.build_batched(tris, np.array([40, 0], dtype=np.uint32)[::-1].copy())
