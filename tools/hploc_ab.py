import sys, os
sys.path.insert(0, 'hip-bvh-construction_b200')
from b2bvh import capi
n = int(sys.argv[1])
ctx = capi.Context(0)
d = ctx.synth_uniform(n, 0x00B20010)
best = 1e9
for _ in range(8):
    t = ctx.build(capi.HPLOC, d, n=n, tris_on_device=True)
    best = min(best, float(t.stage_ms[capi.T_BUILD]))
ctx.profile(True); t = ctx.build(capi.HPLOC, d, n=n, tris_on_device=True); ctx.sync()
ent = {k: round(v, 4) for k, v in ctx.profile_entries() if 'hploc' in k}
print(os.environ.get('B2BVH_HPLOC_WALK_ONLY', 'tile'), n, 'build stage best', round(best, 4), ent, 'calls', t.n_iterations)
