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
        root = b["nodes"][b["root"]]
        return np.concatenate([root["mn"], root["mx"]]), b

    def top_level(self, roots):
        from b2bvh import types as T
        r = roots.numpy().reshape(-1, 6)
        boxes = np.zeros(r.shape[0], dtype=T.AABB)
        boxes["mn"] = r[:, :3]; boxes["mx"] = r[:, 3:]
        return self.orc.top_level(boxes)


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
    # shards partition the input
    from b2bvh.sharded import shard_range
    cover = [shard_range(n, r, world) for r in range(world)]
    assert cover[0][0] == 0 and cover[-1][1] == n and all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
