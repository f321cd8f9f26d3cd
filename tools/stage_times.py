#!/usr/bin/env python
"""development: best-of stage times (us) of one builder at N primitives, stream launches and replayed graph.  usage: tools/stage_times.py <algo 0-3> <N> [clustered]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi
algo = int(sys.argv[1]); n = int(sys.argv[2])
ctx = capi.Context(0)
d = ctx.synth_uniform(n, 0x00B20010, clustered=len(sys.argv) > 3)
for graph in (False, True):
    runs = []
    for _ in range(14):
        t = ctx.build(algo, d, n=n, tris_on_device=True, use_graph=graph)
        runs.append([float(x) * 1e3 for x in t.stage_ms[:6]] + [float(t.build_ms) * 1e3 if hasattr(t, "build_ms") else 0.0])
    best = [min(r[i] for r in runs[2:]) for i in range(7)]
    print(("graph " if graph else "stream"), "extents %.1f morton %.1f sort %.1f build %.1f collapse %.1f | total %.1f" % (best[0], best[1], best[2], best[3], best[5], best[6]), flush=True)
