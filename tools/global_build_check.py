"""The globally sorted multi-GPU build over NCCL: every rank checks its slice of the distributed node array (and the gathered top of the
tree) against ONE tree over all triangles that it builds on its own GPU.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/global_build_check.py [--prims 2000000] [--karras]"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi, types as T  # noqa: E402
from b2bvh.sharded import GlobalBuild, GpuGlobalEngine, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prims", dest="n", type=int, default=2_000_000)
    ap.add_argument("--karras", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local, stream=stream.cuda_stream)
    n = a.n
    first, last = shard_range(n, rank, world)
    half = float(np.float32(1000.0 * n ** (-1.0 / 3.0)))
    d_shard = ctx.synth_uniform(n, 0x00B20010, first=first, count=last - first, half=half)
    gb = GlobalBuild(GpuGlobalEngine(ctx), dist if world > 1 else None, rank, world)
    res = gb.build((d_shard, last - first), first, n, karras=a.karras)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    res = gb.build((d_shard, last - first), first, n, karras=a.karras)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1)
    # the ONE tree, on this GPU
    d_all = ctx.synth_uniform(n, 0x00B20010, half=half)
    whole = ctx.build(capi.TWO_PASS_LBVH if a.karras else capi.SINGLE_PASS_LBVH, d_all, n=n, tris_on_device=True, collapse=False)
    want = ctx.download(whole.d_bvhNodes, T.BVH2_NODE, 2 * n - 1).view(np.int32).reshape(-1, 8)
    ok = res["root"] == whole.root
    mine = res["nodes"].cpu().numpy()
    valid = mine[:, 0] != -1
    idx = res["node_first"] + np.nonzero(valid)[0]
    ok = ok and np.array_equal(mine[valid], want[idx])
    ok = ok and np.array_equal(res["leaves"].cpu().numpy(), want[n - 1 + res["first"]:n - 1 + res["last"]])
    for k, (l, r, box) in res["top"].items():
        ok = ok and want[k, 0] == l and want[k, 1] == r and np.array_equal(want[k, 2:8].view(np.float32), box)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}: positions [{res['first']}, {res['last']}) of {n}, {int(valid.sum())} ghost-free nodes + {len(res['top'])} top nodes, slice identical to the one-GPU tree: {ok}; "
          f"{ms:.2f} ms per global build (first version: host-side finishing of the top, no overlap)")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
