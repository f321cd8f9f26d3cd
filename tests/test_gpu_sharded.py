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
