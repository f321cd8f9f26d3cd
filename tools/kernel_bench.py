"""Per-kernel CUDA-event times of one build (development tool; not part of the bench contract).
usage: python tools/kernel_bench.py [--n 10000000] [--algo singlepass|twopass|ploc|hploc|split|batched] [--reps 5]
(split = TwoPassLbvh over early-split references with saMax = 6 h^2; batched = n/32 items of 32 triangles)
B2BVH_LIB=/path/to/variant.so selects another build of the library."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
from b2bvh import capi  # noqa: E402

if os.environ.get("B2BVH_LIB"):
    capi.LIB_PATH = os.environ["B2BVH_LIB"]

ALGOS = {"twopass": capi.TWO_PASS_LBVH, "singlepass": capi.SINGLE_PASS_LBVH, "ploc": capi.PLOCPP, "hploc": capi.HPLOC}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--algo", default="singlepass")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--two-kernel", action="store_true")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--graph", action="store_true", help="replay the build from a CUDA graph (use_graph)")
    ap.add_argument("--morton60", action="store_true", help="60-bit Morton variant (morton_bits=60)")
    ap.add_argument("--mesh", default="", help="bunny | sponza | buddha (staged under oracle/_ref/meshes) instead of the synthetic soup")
    a = ap.parse_args()
    ctx = capi.Context(0)
    if a.mesh:
        import numpy as np
        from b2bvh import types as T
        tris = T.triangles_from_array(np.fromfile(os.path.join(ROOT, "oracle", "_ref", "meshes", a.mesh + ".tri"), dtype=np.float32).reshape(-1, 9))
        a.n = tris.size
        d = ctx.alloc(tris.nbytes)
        ctx.h2d(d, tris)
    else:
        d = ctx.synth_uniform(a.n, 0x00B20010)
    if a.algo == "batched":
        import numpy as np
        counts = np.full(a.n // 32, 32, dtype=np.uint32)
        for _ in range(a.warmup):
            ctx.build_batched(d, counts, tris_on_device=True)
        best = min(ctx.build_batched(d, counts, tris_on_device=True).build_ms for _ in range(a.reps))
        print(f"batched: items={counts.size} x 32  build_ms(best)={best:.4f} -> {counts.size * 32 / best / 1e3:.1f} Mprims/s, {124 * counts.size * 32 / best / 1e6:.0f} GB/s algorithmic")
        return
    kw = {}
    if a.morton60:
        kw["morton_bits"] = 60
        global_build0 = ctx.build
        kw60 = dict(kw)
        ctx.build = lambda *x, **y: global_build0(*x, **y, **kw60)
        kw = {}
    if a.algo == "split":
        a.algo = "twopass"
        kw["split_sa_max"] = 6.0 * (1000.0 * a.n ** (-1.0 / 3.0)) ** 2
        global_build = ctx.build
        ctx.build = lambda *x, **y: global_build(*x, **{k: v for k, v in y.items() if k != "use_graph"}, **kw)
    for _ in range(a.warmup):
        tree = ctx.build(ALGOS[a.algo], d, n=a.n, tris_on_device=True, karras_two_kernel=a.two_kernel)
    agg = {}
    tot = []
    for _ in range(a.reps):
        ctx.profile(True)
        tree = ctx.build(ALGOS[a.algo], d, n=a.n, tris_on_device=True, karras_two_kernel=a.two_kernel)
        ctx.sync()
        for name, ms in ctx.profile_entries():
            agg.setdefault(name, []).append(ms)
        ctx.profile(False)
        tree = ctx.build(ALGOS[a.algo], d, n=a.n, tris_on_device=True, karras_two_kernel=a.two_kernel, use_graph=a.graph)
        tree = ctx.build(ALGOS[a.algo], d, n=a.n, tris_on_device=True, karras_two_kernel=a.two_kernel, use_graph=a.graph)
        tot.append((tree.build_ms, [tree.stage_ms[k] for k in (0, 1, 2, 3, 5)]))
    if kw:
        print(f"split: references={tree.n_prims} generations={tree.n_split_levels} split_ms={tree.split_ms:.4f}")
    print(f"lib={os.path.basename(capi.LIB_PATH)} algo={a.algo} n={a.n} launches={tree.n_launches} iterations={tree.n_iterations} n_wide={tree.n_wide}")
    best = min(tot)
    print(f"  build_ms(best)={best[0]:.4f}  stages ext/morton/sort/build/collapse = " + " / ".join(f"{x:.4f}" for x in best[1]) +
          f"  -> {a.n / best[0] / 1e3:.1f} Mprims/s")
    for name, v in agg.items():
        per = len(v) // a.reps
        print(f"  {name:28s} launches/build={per:3d}  avg={sum(v) / len(v) * 1e3:9.2f} us  total/build={sum(v) / a.reps * 1e3:10.2f} us")


if __name__ == "__main__":
    main()
