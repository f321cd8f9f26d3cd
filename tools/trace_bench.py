"""Primary-ray timing of the traversal kernels on the reference scenes (development tool).
usage: python tools/trace_bench.py [--mesh bunny|sponza] ; B2BVH_LIB=/path/to/variant.so selects another build of the library."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))
sys.path.insert(0, ROOT)
from b2bvh import capi, types as T  # noqa: E402

if os.environ.get("B2BVH_LIB"):
    capi.LIB_PATH = os.environ["B2BVH_LIB"]
import bench  # noqa: E402  (TRACE_PRESETS, qt_rotation)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="bunny,sponza")
    a = ap.parse_args()
    ctx = capi.Context(0)
    for mesh in a.mesh.split(","):
        p = os.path.join(ROOT, "oracle", "_ref", "meshes", mesh + ".tri")
        mt = T.triangles_from_array(np.fromfile(p, dtype=np.float32).reshape(-1, 9))
        dm = ctx.upload(mt)
        pr = bench.TRACE_PRESETS[mesh]
        tr = T.make_transform(pr["t"], pr["s"], [0.0, 0.0, 0.0, 1.0] if pr["q"] is None else bench.qt_rotation(pr["q"]))
        cam = T.make_camera(pr["eye"], bench.qt_rotation(pr["cq"]), np.float32(45.0) * np.float32(np.pi) / np.float32(180.0))
        d_rays, _ = ctx.generate_rays(cam, 512, 512)
        for nm, al in (("TwoPassLbvh", capi.TWO_PASS_LBVH), ("PLOC++", capi.PLOCPP)):
            t = ctx.build(al, dm, n=mt.size, tris_on_device=True)
            out = []
            for knm, kk in (("while", capi.TRAVERSE_WHILE), ("spec", capi.TRAVERSE_SPECULATIVE_WHILE), ("ifif", capi.TRAVERSE_IFIF), ("trail", capi.TRAVERSE_RESTART_TRAIL),
                            ("bvh4", capi.TRAVERSE_WIDE4)):
                ms = min(ctx.traverse(t, d_rays, 512 * 512, tr, kernel=kk)[2] for _ in range(7))
                out.append(f"{knm} {512 * 512 / ms / 1e3:6.0f}")
            print(f"{os.path.basename(capi.LIB_PATH)} {mesh:7s} {nm:12s} Mray/s: " + "  ".join(out))
        ctx.free(d_rays); ctx.free(dm)


if __name__ == "__main__":
    main()
