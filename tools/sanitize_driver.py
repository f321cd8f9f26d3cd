#!/usr/bin/env python
"""The workload compute-sanitizer runs (tools/sanitize.sh): one small build per builder, the paths either side of the hot path and every
traversal kernel — no torch, no oracle, just the C ABI — sized so that the cooperative kernels finish under the sanitizer's slowdown.
Argument: a comma-separated subset of lbvh,ploc,hploc,split,batched,m60,sort,trace,sizes (default: all)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi, types as T  # noqa: E402

what = set((sys.argv[1] if len(sys.argv) > 1 else "lbvh,ploc,hploc,split,batched,m60,sort,trace,sizes").split(","))
N = int(os.environ.get("B2_SANITIZE_N", "40000"))
ctx = capi.Context(0)
d = ctx.synth_uniform(N, 0x00B20010)
dc = ctx.synth_uniform(N, 0x00B20010, clustered=True)
ctx.sync()
done = []


def build(algo, tris=d, n=N, **kw):
    t = ctx.build(algo, tris, n=n, tris_on_device=True, **kw)
    assert t.n_prims >= 2
    return t


if "lbvh" in what:
    for al in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH):
        build(al); build(al, tris=dc); build(al, use_graph=True)
    build(capi.TWO_PASS_LBVH, karras_two_kernel=True)
    done.append("lbvh")
if "ploc" in what:
    build(capi.PLOCPP); build(capi.PLOCPP, tris=dc)
    done.append("ploc")
if "hploc" in what:
    build(capi.HPLOC); build(capi.HPLOC, tris=dc)
    build(capi.HPLOC, lbvh_second_level=1); build(capi.HPLOC, tris=dc, lbvh_second_level=1)  # the tile phase (automatic from 2^20 primitives)
    done.append("hploc")
if "sizes" in what:  # sizes around the tile boundaries of the hierarchy, sort and collapse kernels
    for n in (2, 3, 33, 511, 512, 513, 1025, 7681, 15361):
        for al in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC):
            build(al, n=n)
    done.append("sizes")
if "split" in what:
    half = float(np.float32(1000.0 * N ** (-1.0 / 3.0)))
    build(capi.TWO_PASS_LBVH, split_sa_max=6.0 * half * half)
    done.append("split")
if "batched" in what:
    items = 500
    counts = (np.arange(items, dtype=np.uint32) % 32) + 1
    ctx.build_batched(d, counts, n_total=int(counts.sum()), tris_on_device=True)
    done.append("batched")
if "m60" in what:
    for al in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC):
        build(al, morton_bits=60)
    done.append("m60")
if "sort" in what:
    rng = np.random.default_rng(7)
    for n in (1, 5, 2048, 40001, 1 << 20):
        k = rng.integers(0, 1 << 32, n, dtype=np.uint32)
        ks, vs = ctx.sort_pairs(k, np.arange(n, dtype=np.uint32))
        assert (ks[1:] >= ks[:-1]).all()
    done.append("sort")
if "trace" in what:
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 3000.0, 0.0], np.array([0.0, 0.0, 0.0, 1.0], dtype=np.float32), np.float32(0.8))
    side = 128
    d_rays, _ = ctx.generate_rays(cam, side, side)
    for al in (capi.SINGLE_PASS_LBVH, capi.PLOCPP):
        t = build(al)
        for k in (capi.TRAVERSE_WHILE, capi.TRAVERSE_SPECULATIVE_WHILE, capi.TRAVERSE_IFIF, capi.TRAVERSE_RESTART_TRAIL, capi.TRAVERSE_WIDE4):
            ctx.traverse(t, d_rays, side * side, tr, kernel=k)
    ctx.free(d_rays)
    done.append("trace")
if "large" in what:  # the variants chosen from 2^20 primitives on: 15-pair sort tiles, second merge level, 512-thread numbering CTAs, many-window PLOC++ chunks
    NL = int(os.environ.get("B2_SANITIZE_NL", "1200000"))
    dl = ctx.synth_uniform(NL, 0x00B20010)
    for al in (capi.TWO_PASS_LBVH, capi.SINGLE_PASS_LBVH, capi.PLOCPP, capi.HPLOC):
        build(al, tris=dl, n=NL)
    ctx.free(dl)
    done.append("large")
ctx.sync()
ctx.free(d)
ctx.free(dc)
ctx.close()
print("sanitize_driver ok:", ",".join(done), "N =", N)
