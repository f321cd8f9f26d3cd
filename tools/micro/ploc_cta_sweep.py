import os, sys
sys.path.insert(0, "hip-bvh-construction_b200")
import numpy as np
from b2bvh import capi, types as T
ctx = capi.Context(0)
def run(d, n, label):
    for ctas in (148, 296, 444, 592):
        best = 1e9
        for _ in range(6):
            t = ctx.build(capi.PLOCPP, d, n=n, tris_on_device=True, merge_max_ctas=ctas, use_graph=True)
            best = min(best, t.stage_ms[capi.T_BUILD])
        print(f"{label} ctas={ctas}: build stage {best*1e3:.1f} us")
d = ctx.synth_uniform(10_000_000, 0x00B20010)
for nn in (600_000, 1_000_000, 2_000_000, 4_000_000):
    run(d, nn, f"{nn}")
for m in ():
    tris = T.triangles_from_array(np.fromfile(f"oracle/_ref/meshes/{m}.tri", dtype=np.float32).reshape(-1, 9))
    dm = ctx.upload(tris)
    run(dm, tris.size, m)
