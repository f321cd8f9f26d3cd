"""The globally sorted multi-GPU build over NCCL (device path: b2bvh_global_*, GlobalBuildDevice): every rank checks its slice of the
distributed node array and the nodes above the ranks against ONE tree over all triangles that it builds on its own GPU, then times
the build (CUDA events, max over ranks) and reports the bytes each rank puts on the wire.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/global_build_check.py [--prims 2000000] [--karras] [--steps 5]"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi, types as T  # noqa: E402
from b2bvh.sharded import GlobalBuildDevice, check_against_one_tree, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prims", dest="n", type=int, default=2_000_000)
    ap.add_argument("--karras", action="store_true")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x00B20010)
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
    d_shard = ctx.synth_uniform(n, a.seed, first=first, count=last - first)
    gb = GlobalBuildDevice(ctx, dist if world > 1 else None, rank, world)
    res = gb.build((d_shard, last - first), first, n, karras=a.karras)
    d_all = ctx.synth_uniform(n, a.seed)
    ok, nodes_mine, ntop, _ = check_against_one_tree(ctx, res, d_all, n, a.karras)
    ctx.free(d_all)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    for _ in range(2):
        gb.build((d_shard, last - first), first, n, karras=a.karras)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        res = gb.build((d_shard, last - first), first, n, karras=a.karras)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps, float(res["wire_bytes_sent"])], device="cuda", dtype=torch.float64)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = float(tmax[0])
    print(f"rank {rank}: positions [{res['first']}, {res['last']}) of {n}, {nodes_mine} ghost-free nodes + {ntop} nodes above the ranks, identical to the one-GPU tree: {ok}", flush=True)
    if rank == 0:
        wire = float(t[1])
        print(json.dumps({"tool": "global_build_check", "n_gpus": world, "total_prims": n, "numbering": "karras" if a.karras else "apetrei",
                          "identical_to_one_gpu_tree_on_all_ranks": bool(int(flag.item()) == 1), "ms_per_build": ms, "Mprims_s": n / ms / 1e3,
                          "wire_bytes_all_ranks": wire, "wire_GBs_aggregate_over_the_whole_build": wire / (ms * 1e-3) / 1e9,
                          "note": "32 B per primitive that changes rank (code, global id, box); the time is the whole build: extents, codes, partition, exchange, sort, hierarchy, top"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
