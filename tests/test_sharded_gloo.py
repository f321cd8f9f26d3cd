"""The N>1 host logic on CPU: two processes, gloo backend, per-rank engine replaced by an oracle-backed test double.
Checks the shard ranges, the {-min,max} all-reduce(MAX) -> global box, the root all-gather and the top-level tree against
the sequential restatement of the sharded procedure (oracle.build_sharded)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_tris


class OracleEngine:
    """Test double of b2bvh.sharded.GpuEngine: same interface, CPU oracle inside (tests only)."""

    def __init__(self, orc):
        self.orc = orc

    def tensor(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    def shard_extents(self, tris):
        _, _, scene = self.orc.primrefs(tris)
        return torch.from_numpy(np.concatenate([-scene["mn"][0], scene["mx"][0]]).astype(np.float32))

    def build(self, tris, scene6):
        from b2bvh import types as T
        sc = np.zeros(1, dtype=T.AABB)
        sc["mn"] = scene6[:3]; sc["mx"] = scene6[3:]
        b = self.orc.build_lbvh(tris, single_pass=True, scene_override=sc)
        b["tris"] = tris
        root = b["nodes"][b["root"]]
        return np.concatenate([root["mn"], root["mx"]]), b

    def trace(self, tree, rays, n_rays, transform, prim_offset, kernel=0):
        from b2bvh.sharded import pack_hits
        hits, _ = self.orc.traverse(rays, tree["nodes"], None, tree["tris"], transform, tree["root"], tree["tris"].size)
        key = pack_hits(torch, torch.from_numpy(hits["primIdx"].view(np.int32).copy()), torch.from_numpy(hits["t"].copy()), prim_offset)
        return key, torch.from_numpy(hits["uv"].copy())

    def top_level(self, roots):
        from b2bvh import types as T
        r = roots.numpy().reshape(-1, 6)
        boxes = np.zeros(r.shape[0], dtype=T.AABB)
        boxes["mn"] = r[:, :3]; boxes["mx"] = r[:, 3:]
        return self.orc.top_level(boxes)


def _scene(orc):
    from b2bvh import types as T
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 400.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.6))
    return tr, orc.generate_rays(cam, 48, 48)


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from b2bvh.sharded import ShardedBuild, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris = random_tris(n, 99)
    a, b = shard_range(n, rank, world)
    res = ShardedBuild(OracleEngine(orc), dist, rank, world).build(np.ascontiguousarray(tris[a:b]))
    np.save(os.path.join(out_dir, f"top{rank}.npy"), res["top"].view(np.uint8))
    np.save(os.path.join(out_dir, f"scene{rank}.npy"), res["scene"])
    np.save(os.path.join(out_dir, f"nodes{rank}.npy"), res["tree"]["nodes"].view(np.uint8))
    # primary rays through the sharded tree: replicated rays, one all-reduce(MIN) of packed hits + one all-reduce(SUM) of the barycentrics
    tr, rays = _scene(orc)
    sb = ShardedBuild(OracleEngine(orc), dist, rank, world)
    t, prim, uv = sb.trace(res, rays, rays.size, tr, a)
    np.save(os.path.join(out_dir, f"trace_t{rank}.npy"), t.numpy()); np.save(os.path.join(out_dir, f"trace_p{rank}.npy"), prim.numpy())
    np.save(os.path.join(out_dir, f"trace_uv{rank}.npy"), uv.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_build_two_ranks(oracle, tmp_path, world):
    n = 5001
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    tris = random_tris(n, 99)
    scene, shards, top = oracle.build_sharded(tris, world)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"scene{r}.npy"), np.concatenate([scene["mn"][0], scene["mx"][0]]))
        assert np.load(tmp_path / f"top{r}.npy").tobytes() == top.tobytes()          # every rank holds the same top-level tree
        assert np.load(tmp_path / f"nodes{r}.npy").tobytes() == shards[r]["nodes"].tobytes()  # shard tree in the GLOBAL frame
    # sharded primary rays: every rank ends with the sequential restatement's hits, and those are the hits of ONE tree over all triangles
    tr, rays = _scene(oracle)
    t, prim, uv = oracle.trace_sharded(tris, world, rays, tr)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"trace_t{r}.npy").view(np.uint32), t.view(np.uint32))
        assert np.array_equal(np.load(tmp_path / f"trace_p{r}.npy"), prim) and np.array_equal(np.load(tmp_path / f"trace_uv{r}.npy").view(np.uint32), uv.view(np.uint32))
    whole = oracle.build_lbvh(tris, single_pass=True)
    hits, cnt = oracle.traverse(rays, whole["nodes"], None, tris, tr, whole["root"], n)
    assert cnt > 20 and np.array_equal(hits["primIdx"] != 0xFFFFFFFF, prim >= 0)
    assert np.array_equal(hits["t"][prim >= 0].view(np.uint32), t[prim >= 0].view(np.uint32))
    # shards partition the input
    from b2bvh.sharded import shard_range
    cover = [shard_range(n, r, world) for r in range(world)]
    assert cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))


# ---------------------------------------------------------------- the globally sorted build (GlobalBuild): G ranks -> the ONE-GPU tree
class OracleGlobalEngine:
    """Test double of the device engine of b2bvh.sharded.GlobalBuild: same interface, CPU oracle inside (tests only)."""

    def __init__(self, orc):
        self.orc = orc

    def boxes_and_scene(self, tris):
        _, boxes, scene = self.orc.primrefs(tris)
        b = np.concatenate([boxes["mn"], boxes["mx"]], axis=1).astype(np.float32)
        return torch.from_numpy(b), torch.from_numpy(np.concatenate([-scene["mn"][0], scene["mx"][0]]).astype(np.float32))

    def _aabbs(self, boxes):
        from b2bvh import types as T
        a = np.zeros(boxes.shape[0], dtype=T.AABB)
        bn = boxes.numpy()
        a["mn"] = bn[:, :3]; a["mx"] = bn[:, 3:]
        return a

    def morton(self, boxes, scene):
        from b2bvh import types as T
        sc = np.zeros(1, dtype=T.AABB)
        sc["mn"] = scene[:3].numpy(); sc["mx"] = scene[3:].numpy()
        keys, _ = self.orc.morton_codes(self._aabbs(boxes), sc)
        return torch.from_numpy(keys.astype(np.int64))

    def sort(self, codes):
        k = codes.numpy().astype(np.uint32)
        sk, sv = self.orc.sort_kv(k, np.arange(k.size, dtype=np.uint32))
        return torch.from_numpy(sk.astype(np.int64)), torch.from_numpy(sv.astype(np.int64))

    def range_tree(self, k64, vals, boxes, karras, ghost_left, ghost_right, first_pos, n_global):
        from b2bvh import types as T
        ab = self._aabbs(boxes)
        refs = np.zeros(ab.size, dtype=T.PRIM_REF)
        refs["primIdx"] = np.arange(ab.size); refs["mn"] = ab["mn"]; refs["mx"] = ab["mx"]
        fake = np.zeros(ab.size, dtype=T.TRIANGLE)  # triangles whose box is the received box (the Apetrei restatement boxes triangles itself)
        fake["v"][:, 0] = ab["mn"]; fake["v"][:, 1] = ab["mx"]; fake["v"][:, 2] = ab["mn"]
        out, cl = self.orc.range_tree_reference(fake, refs, k64.numpy().astype(np.uint64), vals.numpy().astype(np.uint32), karras, ghost_left, ghost_right,
                                                first_pos, n_global)
        return torch.from_numpy(out.view(np.uint8).reshape(-1, 32).copy()), torch.from_numpy(cl.view(np.int32).reshape(-1, 12).copy())


def _global_worker(rank, world, port, n, karras, out_dir, kind=None):
    sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    from b2bvh.sharded import GlobalBuild, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tris = random_tris(n, 97, kind or ("clustered" if karras else "uniform"))
    a, b = shard_range(n, rank, world)
    res = GlobalBuild(OracleGlobalEngine(orc), dist, rank, world, sample=32).build(np.ascontiguousarray(tris[a:b]), a, n, karras=karras)
    np.save(os.path.join(out_dir, f"g_nodes{rank}.npy"), res["nodes"].numpy() if res["nodes"] is not None else np.zeros((0, 8), np.int32))
    np.save(os.path.join(out_dir, f"g_leaves{rank}.npy"), res["leaves"].numpy() if res["leaves"] is not None else np.zeros((0, 8), np.int32))
    np.save(os.path.join(out_dir, f"g_meta{rank}.npy"), np.array([res.get("node_first", 0), res["first"], res["last"], res["root"]], dtype=np.int64))
    top = res["top"]
    np.save(os.path.join(out_dir, f"g_top{rank}.npy"),
            np.array([[k, v[0], v[1]] + v[2].view(np.uint32).tolist() for k, v in sorted(top.items())], dtype=np.int64).reshape(-1, 9))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,karras,kind,n", [(2, False, None, 3001), (3, True, None, 3001), (3, False, None, 3001), (3, False, "duplicate", 301), (4, True, "duplicate", 302),
                                                 (4, False, "flat", 7)])
def test_global_build_yields_the_one_gpu_tree(oracle, tmp_path, world, karras, kind, n):
    """G ranks, gloo, oracle-backed engine: exchange by code interval, local stable sorts, range trees with ghosts, gathered left-overs —
    the assembled node array equals the single build over all triangles byte for byte."""
    """(duplicate: all codes but one are equal, so whole ranks receive nothing; flat / 7: fewer primitives than sample slots)"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_global_worker, args=(world, port, n, karras, str(tmp_path), kind), nprocs=world, join=True)
    tris = random_tris(n, 97, kind or ("clustered" if karras else "uniform"))
    want = oracle.build_lbvh(tris, single_pass=not karras)
    nodes = np.zeros((2 * n - 1, 8), dtype=np.int32)
    nint = n - 1
    covered = 0
    tops = [np.load(tmp_path / f"g_top{r}.npy") for r in range(world)]
    for r in range(world):
        assert np.array_equal(tops[r], tops[0])                       # every rank finished the same top of the tree
        nf, a, b, root = np.load(tmp_path / f"g_meta{r}.npy").tolist()
        assert root == want["root"]
        loc = np.load(tmp_path / f"g_nodes{r}.npy")
        ok = loc[:, 0] != -1                                          # artefacts carry INVALID children
        idx = nf + np.nonzero(ok)[0]
        nodes[idx] = loc[ok]
        nodes[nint + a:nint + b] = np.load(tmp_path / f"g_leaves{r}.npy")
        covered += b - a
    assert covered == n
    for row in tops[0]:
        nodes[row[0], 0] = row[1]; nodes[row[0], 1] = row[2]; nodes[row[0], 2:8] = row[3:9].astype(np.uint32).view(np.int32)
    assert nodes.tobytes() == want["nodes"].tobytes()
