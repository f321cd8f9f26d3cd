#!/usr/bin/env python
"""A/B of library variants (development): for every hip-bvh-construction_b200/variants/libb2bvh_<name>.so run one builder at N primitives in a
fresh process (B2BVH_LIB) and print its best stage times.  usage: tools/variant_bench.py <hploc|ploc|singlepass|twopass> [N]"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, "hip-bvh-construction_b200")
from b2bvh import capi
algo = {"twopass": 0, "singlepass": 1, "ploc": 2, "hploc": 3}[sys.argv[1]]
n = int(sys.argv[2])
ctx = capi.Context(0)
kw = dict(n=n, tris_on_device=True)
if len(sys.argv) > 3 and sys.argv[3].startswith("mesh:"):
    import lzma, numpy as np
    from b2bvh import types as T
    d = T.triangles_from_array(np.frombuffer(lzma.open("tests/golden/" + sys.argv[3][5:] + ".tri.xz").read(), dtype=np.float32).reshape(-1, 9).copy())
    dd = ctx.upload_triangles(d) if hasattr(ctx, "upload_triangles") else None
    kw = {}
else:
    d = ctx.synth_uniform(n, 0x00B20010, clustered=len(sys.argv) > 3 and sys.argv[3] == "clustered")
best = None
for _ in range(12):
    t = ctx.build(algo, d, **kw, use_graph=len(sys.argv) > 4 and sys.argv[4] == "graph", lbvh_second_level=int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    st = [float(x) for x in t.stage_ms[:6]]
    if best is None or st[5] < best[5]:
        best = st
print(json.dumps({"build_ms": best[3], "sort_ms": best[2], "collapse_ms": best[5], "n_wide": int(t.n_wide)}))
'''
which = sys.argv[1] if len(sys.argv) > 1 else "hploc"
n = sys.argv[2] if len(sys.argv) > 2 else "10000000"
kind = sys.argv[3] if len(sys.argv) > 3 else "uniform"
graph = sys.argv[4] if len(sys.argv) > 4 else "stream"
second = sys.argv[5] if len(sys.argv) > 5 else "0"
for lib in sorted(glob.glob(os.path.join(ROOT, "hip-bvh-construction_b200", "variants", "libb2bvh_*.so"))):
    r = subprocess.run([sys.executable, "-c", CHILD, which, n, kind, graph, second], env=dict(os.environ, B2BVH_LIB=lib), capture_output=True, text=True, timeout=120, cwd=ROOT)
    print(os.path.basename(lib), r.stdout.strip() or r.stderr[-300:], flush=True)
