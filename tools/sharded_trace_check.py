"""Sharded build + sharded primary rays over NCCL, checked on rank 0 against ONE tree over all triangles on one GPU.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/sharded_trace_check.py [--prims 2000000]"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi, types as T  # noqa: E402
from b2bvh.sharded import GpuEngine, ShardedBuild, shard_range  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prims", dest="n", type=int, default=2_000_000)
    ap.add_argument("--size", type=int, default=512)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx = capi.Context(local, stream=stream.cuda_stream)
    first, last = shard_range(a.n, rank, world)
    half = float(np.float32(1000.0 * a.n ** (-1.0 / 3.0)))
    d_shard = ctx.synth_uniform(a.n, 0x00B20010, first=first, count=last - first, half=half)
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 2600.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.75))
    nr = a.size * a.size
    d_rays, _ = ctx.generate_rays(cam, a.size, a.size)
    sb = ShardedBuild(GpuEngine(ctx, capi.SINGLE_PASS_LBVH), dist if world > 1 else None, rank, world)
    built = sb.build((d_shard, last - first))
    t, prim, uv = sb.trace(built, d_rays, nr, tr, first)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        t, prim, uv = sb.trace(built, d_rays, nr, tr, first)
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if rank == 0:
        import time
        eng = sb.engine
        def timed(f, reps=5):
            stream.synchronize(); t0 = time.perf_counter()
            for _ in range(reps):
                r = f()
            stream.synchronize(); return (time.perf_counter() - t0) / reps * 1e3, r
        ta, (key, uvv) = timed(lambda: eng.trace(built["tree"], d_rays, nr, tr, first))
        hitsbuf = torch.empty((nr, 8), dtype=torch.int32, device="cuda")
        from b2bvh.sharded import pack_hits
        tb, _ = timed(lambda: pack_hits(torch, hitsbuf[:, 0], hitsbuf[:, 1].view(torch.float32), first))
        tc, _ = timed(lambda: ctx.lib.b2bvh_traverse(ctx.h, capi.C.byref(built["tree"]), capi.C.c_void_p(int(d_rays)), nr, tr.ctypes.data_as(capi.C.c_void_p), 0,
                                                     capi.C.c_void_p(hitsbuf.data_ptr()), None, capi.C.byref(capi.C.c_float())))
        print(f"  pieces (host ms): engine.trace {ta:.3f} = b2bvh_traverse {tc:.3f} + pack {tb:.3f}; whole ShardedBuild.trace {ms:.3f}")
    if rank == 0:
        d_all = ctx.synth_uniform(a.n, 0x00B20010, half=half)
        whole = ctx.build(capi.SINGLE_PASS_LBVH, d_all, n=a.n, tris_on_device=True, collapse=False)
        hits, _, ms1 = ctx.traverse(whole, d_rays, nr, tr)
        h = hits["primIdx"] != 0xFFFFFFFF
        p = prim.cpu().numpy()
        ok = np.array_equal(p >= 0, h) and np.array_equal(t.cpu().numpy()[h].view(np.uint32), hits["t"][h].view(np.uint32))
        same_prim = float((p[h] == hits["primIdx"][h]).mean()) if h.any() else 1.0
        print(f"sharded trace world={world} n={a.n} rays={nr}: hits={int(h.sum())} closest-hit distances identical to the one-tree trace: {ok}; "
              f"same primitive on {same_prim:.4f} of the hits; {ms:.3f} ms per sharded trace ({nr / ms / 1e3:.0f} Mray/s) vs {ms1:.3f} ms one tree")
        if not ok:
            sys.exit(1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
