"""LBVH second merge level on/off by input size, builds replayed from a CUDA graph (development probe)."""
import sys
sys.path.insert(0, "hip-bvh-construction_b200")
import numpy as np
from b2bvh import capi, types as T
ctx = capi.Context(0)
def run(d, n, label):
    for lvl in (2, 1):
        best = 1e9
        for _ in range(8):
            t = ctx.build(capi.SINGLE_PASS_LBVH, d, n=n, tris_on_device=True, lbvh_second_level=lvl, use_graph=True)
            best = min(best, t.stage_ms[capi.T_BUILD])
        print(f"{label} second_level={'on' if lvl == 1 else 'off'}: build stage {best*1e3:.1f} us")
d = ctx.synth_uniform(10_000_000, 0x00B20010)
for nn in (3_000_000, 5_000_000, 7_000_000, 10_000_000):
    run(d, nn, str(nn))
for m in ():
    tris = T.triangles_from_array(np.fromfile(f"oracle/_ref/meshes/{m}.tri", dtype=np.float32).reshape(-1, 9))
    dm = ctx.upload(tris)
    run(dm, tris.size, m)
