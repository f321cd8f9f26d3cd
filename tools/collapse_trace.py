import sys
sys.path.insert(0, "hip-bvh-construction_b200")
from b2bvh import capi
n = int(sys.argv[1]); ctx = capi.Context(0)
d = ctx.synth_uniform(n, 0x00B20010, clustered=len(sys.argv) > 2)
for i in range(3):
    print("BUILD", i, flush=True)
    t = ctx.build(1, d, n=n, tris_on_device=True)
    ctx.sync()
