"""CPU: the oracle's restatements of the reference's traversal kernels agree with each other on the closest hit — if-if
(== Utility::TraversalLbvhCPU order), restart trail (TraversalKernel.h:49-146, incl. a root that is not node 0) and the
Bvh4 walk — and the heat-map colouring follows Utility.cpp:424-454."""
import numpy as np

from conftest import load_mesh, random_tris
from b2bvh import types as T

PI = np.float32(3.14159265358979323846)


def cornell(oracle):
    tr = T.make_transform([0.0, 0.0, -5.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 2.5, 5.8, 0.0], oracle.qt_rotation([0.0, 0.0, 1.0, -1.57]), np.float32(45.0) * PI / np.float32(180.0))
    return load_mesh("cornellbox"), tr, cam


def check(oracle, tris, tr, cam, size):
    n = tris.size
    rays = oracle.generate_rays(cam, size, size)
    for single in (False, True):  # Karras numbering (root 0) and Apetrei numbering (root anywhere)
        o = oracle.build_lbvh(tris, single_pass=single)
        root = o.get("root", 0)
        ref, cnt = oracle.traverse(rays, o["nodes"], None, tris, tr, root, n)
        h0, c0, k0 = oracle.traverse_kind(0, rays, o["nodes"], None, tris, tr, root, n)
        h1, c1, k1 = oracle.traverse_kind(1, rays, o["nodes"], None, tris, tr, root, n)
        h2, c2, k2 = oracle.traverse_wide4(rays, o["wide"], o["nodes"], None, tris, tr, n)
        assert h0.tobytes() == ref.tobytes() and c0 == cnt and cnt > 0
        for h in (h1, h2):
            assert np.array_equal(h["t"].view(np.uint32), ref["t"].view(np.uint32))
            assert (h["primIdx"] != ref["primIdx"]).mean() < 1e-3
        assert c1 == cnt and c2 == cnt
        assert k0.sum() > 0 and k1.sum() > 0 and k2.sum() > 0
        assert k2.sum() <= 2 * k0.sum() + rays.size  # the Bvh4 walk culls leaves by their boxes as the Bvh2 walk does
    return k0


def test_traversal_variants_agree_cornellbox(oracle):
    tris, tr, cam = cornell(oracle)
    k = check(oracle, tris, tr, cam, 96)
    rgba = oracle.heat_map(k)
    mx = k.max()
    i = int(np.argmax(k))
    assert tuple(rgba[i]) == (150, 255, 255, 255)
    j = int(np.argmin(k))
    assert tuple(rgba[j]) == (int(np.float32(k[j]) / np.float32(mx) * 150), int(np.float32(k[j]) / np.float32(mx) * 255), 255, 255)


def test_traversal_variants_agree_synthetic(oracle):
    tris = random_tris(3000, 4)
    tr = T.make_transform([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    cam = T.make_camera([0.0, 0.0, 400.0, 0.0], [0.0, 0.0, 0.0, 1.0], np.float32(0.6))
    check(oracle, tris, tr, cam, 64)
