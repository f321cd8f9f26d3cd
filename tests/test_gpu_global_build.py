"""GPU, the globally sorted multi-GPU build on the device (b2bvh_global_*, b2bvh/sharded.py GlobalBuildDevice): W ranks — here W threads, one
context each on the one GPU of the box, with a thread-backed stand-in for torch.distributed (the NCCL path proper runs in bench.py --gpus N
and tools/global_build_check.py) — produce, node for node, the tree ONE context builds over all triangles."""
import threading

import numpy as np
import pytest

from conftest import random_tris
from b2bvh import capi

pytestmark = pytest.mark.gpu


class ThreadDist:
    """all_reduce / all_gather_into_tensor / all_to_all_single among W threads of one process (tensors on one device)."""

    class ReduceOp:
        MAX = "max"

    class _Done:
        def wait(self):
            return None

    def __init__(self, world):
        import torch
        self.torch, self.world = torch, world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.local = threading.local()

    def _exchange(self, value):
        self.torch.cuda.synchronize()
        self.slots[self.local.rank] = value
        self.barrier.wait()
        got = list(self.slots)
        self.barrier.wait()
        return got

    def all_reduce(self, t, op=None):
        got = self._exchange(t.clone())
        t.copy_(self.torch.stack(got).max(dim=0).values)

    def all_gather_into_tensor(self, out, t):
        got = self._exchange(t.clone())
        out.view(-1).copy_(self.torch.cat([g.reshape(-1) for g in got]))

    def all_to_all_single(self, out, inp, output_split_sizes=None, input_split_sizes=None, async_op=False):
        offs = np.concatenate([[0], np.cumsum(input_split_sizes)]).astype(int)
        got = self._exchange((inp.clone(), offs))
        pieces = [g[0][g[1][self.local.rank]:g[1][self.local.rank + 1]] for g in got]
        if out.numel():
            out.copy_(self.torch.cat(pieces))
        self.torch.cuda.synchronize()
        return self._Done()


def run_global(tris, world, karras, cuts=None, collapse=False):
    import torch
    from b2bvh.sharded import GlobalBuildDevice, assemble_global_tree
    n = tris.size
    cuts = cuts or [(n * r) // world for r in range(world + 1)]
    dist = ThreadDist(world)
    parts, errs = [None] * world, []

    def work(r):
        try:
            dist.local.rank = r
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                ctx = capi.Context(0, stream=stream.cuda_stream)
                gb = GlobalBuildDevice(ctx, dist if world > 1 else None, r, world)
                res = gb.build(np.ascontiguousarray(tris[cuts[r]:cuts[r + 1]]), cuts[r], n, karras=karras)
                stream.synchronize()
                parts[r] = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in res.items()}
                if collapse:
                    w4 = gb.collapse_replicated(res)
                    stream.synchronize()
                    parts[r]["bvh4"] = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in w4.items()}
                ctx.close()
        except Exception as e:  # a dead thread would leave the others in the barrier
            errs.append(e)
            dist.barrier.abort()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return assemble_global_tree(parts, n), parts


CASES = [("uniform", 20_000, 2, None), ("uniform", 50_001, 3, None), ("clustered", 30_000, 4, None), ("duplicate", 3000, 3, None),
         ("uniform", 9000, 3, [0, 100, 100, 9000]), ("anisotropic", 40_000, 8, None), ("uniform", 300_000, 2, None)]


@pytest.mark.parametrize("karras", [False, True], ids=["apetrei", "karras"])
@pytest.mark.parametrize("kind,n,world,cuts", CASES, ids=[f"{k}-{n}-w{w}" + ("-ragged" if c else "") for k, n, w, c in CASES])
def test_ranks_produce_the_one_gpu_tree(ctx, kind, n, world, cuts, karras):
    tris = random_tris(n, 900 + n % 97, kind)
    (nodes, root, written), parts = run_global(tris, world, karras, cuts)
    one = ctx.fetch(ctx.build(capi.TWO_PASS_LBVH if karras else capi.SINGLE_PASS_LBVH, tris, collapse=False))
    ref = one["nodes"].view(np.int32).reshape(-1, 8)
    assert written.all(), f"{(~written).sum()} nodes were produced by no rank"
    assert root == one["root"]
    bad = np.nonzero((nodes != ref).any(axis=1))[0]
    assert bad.size == 0, f"first differing node {bad[0]}: ranks {nodes[bad[0]]} one GPU {ref[bad[0]]}"
    assert sum(p["last"] - p["first"] for p in parts) == n


def test_world_size_one_is_the_plain_build(ctx):
    tris = random_tris(12_345, 7)
    (nodes, root, written), _ = run_global(tris, 1, False)
    one = ctx.fetch(ctx.build(capi.SINGLE_PASS_LBVH, tris, collapse=False))
    assert written.all() and root == one["root"] and nodes.tobytes() == one["nodes"].tobytes()


@pytest.mark.parametrize("karras", [False, True], ids=["apetrei", "karras"])
@pytest.mark.parametrize("kind,n,world,cuts", [("uniform", 50_001, 3, None), ("clustered", 30_000, 4, None), ("uniform", 9000, 3, [0, 100, 100, 9000]), ("uniform", 12_345, 1, None)],
                         ids=["uniform-w3", "clustered-w4", "ragged-w3", "w1"])
def test_replicated_collapse_is_the_one_gpu_bvh4(ctx, kind, n, world, cuts, karras):
    """collapse_replicated: every rank assembles the whole node array from the all-gathered pieces and collapses it — Bvh4 nodes, leaves and count
    equal the one-GPU build's byte for byte, on every rank."""
    tris = random_tris(n, 1200 + n % 89, kind)
    (_, root, _), parts = run_global(tris, world, karras, cuts, collapse=True)
    one = ctx.fetch(ctx.build(capi.TWO_PASS_LBVH if karras else capi.SINGLE_PASS_LBVH, tris))
    for r, p in enumerate(parts):
        b = p["bvh4"]
        assert b["n_wide"] == one["n_wide"], r
        assert b["full_nodes"].tobytes() == one["nodes"].tobytes(), r
        assert b["wide"].tobytes() == one["wide"].tobytes(), r
        assert b["wide_leaves"].tobytes() == one["wide_leaves"].tobytes(), r
