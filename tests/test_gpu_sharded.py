"""GPU, sharded build helpers (world size 1 on the box; the collectives themselves are covered by tests/test_sharded_gloo.py):
a shard built with boxes_ready + the {-min,max} vector left on the device is byte-identical to the same shard built with the
global scene box passed from the host, and both equal the oracle's build of the shard in the global frame."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_tris
from b2bvh import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", [capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.HPLOC], ids=["twopass", "singlepass", "hploc"])
def test_shard_with_device_scene_box(ctx, oracle, algo):
    tris = random_tris(30_000, 77)
    shard = np.ascontiguousarray(tris[5_000:17_000])
    # global frame = box of ALL triangles, delivered the way the all-reduce delivers it: {-min, max} on the device
    v = tris["v"].reshape(-1, 3)
    gmin, gmax = v.min(axis=0).astype(np.float32), v.max(axis=0).astype(np.float32)
    box6_host = np.concatenate([-gmin, gmax]).astype(np.float32)
    d_box6 = ctx.upload(box6_host)
    d_local = ctx.alloc(24)
    try:
        host_tree = ctx.fetch(ctx.build(algo, shard, scene_box=np.concatenate([gmin, gmax])))
        capi.check(ctx.lib.b2bvh_shard_extents(ctx.h, shard.ctypes.data_as(C.c_void_p), shard.size, 0, C.c_void_p(d_local)), "b2bvh_shard_extents")
        local = ctx.download(d_local, np.float32, 6)
        sv = shard["v"].reshape(-1, 3)
        assert np.array_equal(local, np.concatenate([-sv.min(axis=0), sv.max(axis=0)]).astype(np.float32))
        dev_tree = ctx.fetch(ctx.build(algo, shard, boxes_ready=True, d_scene_negmin_max=d_box6))
        for k in ("scene", "boxes", "keys", "skeys", "svals", "nodes", "wide", "wide_leaves"):
            assert host_tree[k].tobytes() == dev_tree[k].tobytes(), k
        if algo != capi.HPLOC:  # the oracle builds the same shard in the same global frame
            from b2bvh import types as T
            ov = np.zeros(1, dtype=T.AABB); ov["mn"][0] = gmin; ov["mx"][0] = gmax
            o = oracle.build_lbvh(shard, single_pass=(algo == capi.SINGLE_PASS_LBVH), scene_override=ov)
            assert np.array_equal(dev_tree["skeys"], o["skeys"]) and dev_tree["nodes"].tobytes() == o["nodes"].tobytes()
            assert dev_tree["wide"].tobytes() == o["wide"].tobytes()
        assert np.array_equal(dev_tree["scene"]["mn"].reshape(3), gmin) and np.array_equal(dev_tree["scene"]["mx"].reshape(3), gmax)
    finally:
        ctx.free(d_box6)
        ctx.free(d_local)


def test_sharded_trace_on_one_gpu(oracle):
    """Primary rays through a 2-shard build: each shard's GpuEngine traces its sub-tree on the device and packs (bits of t, global
    primitive) words; their element-wise minimum — what the all-reduce(MIN) of ShardedBuild.trace computes across ranks (covered with gloo
    in tests/test_sharded_gloo.py) — equals the oracle's sequential restatement, and world = 1 through ShardedBuild equals the plain kernel."""
    import torch
    from b2bvh import types as T
    from b2bvh.sharded import GpuEngine, ShardedBuild, shard_range
    n, world = 40_000, 2
    tris = random_tris(n, 78)
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 400.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.6))
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctxs = [capi.Context(0, stream=stream.cuda_stream) for _ in range(world)]
        d_rays, _ = ctxs[0].generate_rays(cam, 128, 128)
        rays = ctxs[0].download(d_rays, T.RAY, 128 * 128)
        v = tris["v"].reshape(-1, 3)
        box6 = torch.from_numpy(np.concatenate([-v.min(axis=0), v.max(axis=0)]).astype(np.float32)).cuda()
        keys, uvs = [], []
        for r in range(world):
            a, b = shard_range(n, r, world)
            shard = np.ascontiguousarray(tris[a:b])
            eng = GpuEngine(ctxs[r], capi.SINGLE_PASS_LBVH)
            eng.shard_extents(shard)
            _, tree = eng.build(shard, box6)  # enqueued only (defer_sync) ...
            eng.finish(tree)                  # ... the root index arrives here
            k, uv = eng.trace(tree, d_rays, 128 * 128, tr, a)
            keys.append(k); uvs.append(uv)
        best = torch.minimum(keys[0], keys[1])
        uv = torch.where((keys[0] == best).unsqueeze(1), uvs[0], uvs[1])
        stream.synchronize()
        t, prim, ouv = oracle.trace_sharded(tris, world, rays, tr)
        hit = prim >= 0
        assert hit.sum() > 500
        bk = best.cpu().numpy()
        assert np.array_equal(bk != ShardedBuild.MISS, hit)
        assert np.array_equal((bk[hit] >> 32).astype(np.uint32), t[hit].view(np.uint32)) and np.array_equal(bk[hit] & 0xFFFFFFFF, prim[hit])
        assert np.array_equal(uv.cpu().numpy()[hit].view(np.uint32), ouv[hit].view(np.uint32))
        # world = 1: ShardedBuild.trace is the plain kernel
        eng = GpuEngine(ctxs[0], capi.SINGLE_PASS_LBVH)
        sb = ShardedBuild(eng)
        built = sb.build(tris)
        t1, p1, uv1 = sb.trace(built, d_rays, 128 * 128, tr, 0)
        hits, _, _ = ctxs[0].traverse(built["tree"], d_rays, 128 * 128, tr)
        h = hits["primIdx"] != 0xFFFFFFFF
        assert np.array_equal(p1.cpu().numpy() >= 0, h) and np.array_equal(p1.cpu().numpy()[h], hits["primIdx"][h].astype(np.int64))
        assert np.array_equal(t1.cpu().numpy()[h].view(np.uint32), hits["t"][h].view(np.uint32))
        assert np.array_equal(uv1.cpu().numpy()[h].view(np.uint32), hits["uv"][h].view(np.uint32))
        ctxs[0].free(d_rays)
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("karras", [False, True], ids=["apetrei", "karras"])
def test_global_build_engine_on_one_gpu(oracle, karras):
    """GlobalBuild with the device engine at world size 1 (the exchange and the gathers are covered with gloo on 2-3 ranks in
    tests/test_sharded_gloo.py and over NCCL by tools/global_build_check.py): codes, sort, hierarchy stage over widened keys and the
    extraction run on the device and give the nodes of the ordinary build, which equal the oracle's."""
    import torch
    from b2bvh import types as T
    from b2bvh.sharded import GlobalBuild, GpuGlobalEngine
    n = 30_000
    tris = random_tris(n, 79)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        c = capi.Context(0, stream=stream.cuda_stream)
        res = GlobalBuild(GpuGlobalEngine(c)).build(tris, 0, n, karras=karras)
        stream.synchronize()
        o = oracle.build_lbvh(tris, single_pass=not karras)
        want = o["nodes"].view(np.int32).reshape(-1, 8)
        assert res["root"] == o["root"] and res["first"] == 0 and res["last"] == n and res["top"] == {}
        assert np.array_equal(res["nodes"].cpu().numpy(), want[:n - 1]) and np.array_equal(res["leaves"].cpu().numpy(), want[n - 1:])
        c.close()


@pytest.mark.parametrize("world", [2, 3])
def test_build_sharded_c_abi(ctx, oracle, world):
    """b2bvh_build_sharded: ONE host thread over `world` contexts (here all on the box's one GPU): every shard's tree and the top-level tree
    equal the oracle's sharded procedure byte for byte."""
    from b2bvh import types as T
    tris = random_tris(60_001, 91)
    n = tris.size
    ctxs = [capi.Context(0) for _ in range(world)]
    try:
        ranges = [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]
        shards = [np.ascontiguousarray(tris[a:b]) for a, b in ranges]
        handles = (C.c_void_p * world)(*[c.h for c in ctxs])
        ptrs = (C.c_void_p * world)(*[s.ctypes.data for s in shards])
        counts = (C.c_uint32 * world)(*[s.size for s in shards])
        trees = (capi.Tree * world)()
        scene = np.zeros(1, dtype=T.AABB)
        top = np.zeros(2 * world - 1, dtype=T.BVH2_NODE)
        capi.check(ctx.lib.b2bvh_build_sharded(handles, world, capi.SINGLE_PASS_LBVH, ptrs, counts, None, trees, scene.ctypes.data_as(C.c_void_p),
                                               top.ctypes.data_as(C.c_void_p)), "b2bvh_build_sharded")
        o_scene, o_shards, o_top = oracle.build_sharded(tris, world, single_pass=True)
        assert scene.tobytes() == o_scene.tobytes()
        assert top.tobytes() == o_top.tobytes()
        for r in range(world):
            g = ctxs[r].fetch(trees[r])
            for k in ("skeys", "svals", "nodes", "wide", "wide_leaves"):
                assert g[k].tobytes() == o_shards[r][k].tobytes(), (r, k)
            assert g["root"] == o_shards[r]["root"]
    finally:
        for c in ctxs:
            c.close()
