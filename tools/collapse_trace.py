#!/usr/bin/env python
"""development: three builds with a -DCOL_TRACE / -DLBVH_TRACE library (B2BVH_LIB) so that the device-side trace of the last one can be read.
usage: tools/collapse_trace.py <N | mesh name> [clustered] [algo]"""
import lzma, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi, types as T
ctx = capi.Context(0)
algo = int(sys.argv[3]) if len(sys.argv) > 3 else 1
if sys.argv[1].isdigit():
    n = int(sys.argv[1])
    d = ctx.synth_uniform(n, 0x00B20010, clustered=len(sys.argv) > 2 and sys.argv[2] == "clustered")
    kw = dict(n=n, tris_on_device=True)
else:
    p = os.path.join(ROOT, "tests", "golden", sys.argv[1] + ".tri.xz")
    d = T.triangles_from_array(np.frombuffer(lzma.open(p).read(), dtype=np.float32).reshape(-1, 9).copy())
    kw = {}
for i in range(3):
    print("BUILD", i, flush=True)
    t = ctx.build(algo, d, **kw)
    ctx.sync()
