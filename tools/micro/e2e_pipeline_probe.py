"""Host-side timeline of the two-context end-to-end pipeline (development probe)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
import torch
from b2bvh import capi

n = 10_000_000
lanes = []
for _ in range(2):
    st = torch.cuda.Stream()
    lanes.append((st, capi.Context(0, stream=st.cuda_stream)))
ctx = lanes[0][1]
d_tris = ctx.synth_uniform(n, 0x00B20010)
h_tris = ctx.pinned(n * 64)
capi.check(ctx.lib.b2bvh_d2h(ctx.h, h_tris, d_tris, n * 64))
out_bytes = (2 * n - 1) * 32 + n * 128 + n * 8
h_out = [ctx.pinned(out_bytes), ctx.pinned(out_bytes)]
mode = sys.argv[1] if len(sys.argv) > 1 else "tree"
t00 = time.perf_counter()
for i in range(10):
    st, c = lanes[i & 1]
    t0 = time.perf_counter(); c.sync(); t1 = time.perf_counter()
    t = c.build(capi.SINGLE_PASS_LBVH, h_tris, n=n, tris_on_device=False)
    t2 = time.perf_counter()
    b0, b1, b2 = (2 * n - 1) * 32, t.n_wide * 128, n * 8
    capi.check(c.lib.b2bvh_d2h_async(c.h, h_out[i & 1], t.d_bvhNodes, b0))
    capi.check(c.lib.b2bvh_d2h_async(c.h, h_out[i & 1] + b0, t.d_wideBvhNodes, b1))
    capi.check(c.lib.b2bvh_d2h_async(c.h, h_out[i & 1] + b0 + b1, t.d_wideLeafNodes, b2))
    t3 = time.perf_counter()
    print(f"step {i} lane {i & 1}: start {1e3 * (t0 - t00):7.2f}  sync {1e3 * (t1 - t0):6.2f}  build {1e3 * (t2 - t1):6.2f} (h2d {t.h2d_ms:.2f}, dev {t.build_ms:.2f})  enqueue d2h {1e3 * (t3 - t2):6.2f}")
for st, c in lanes:
    c.sync()
print(f"total {1e3 * (time.perf_counter() - t00):.2f} ms for 10 steps")
