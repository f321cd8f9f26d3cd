"""Pins the CPU oracle against the UNMODIFIED reference code compiled from /root/reference into oracle/_ref
(kernels run by the sequential CUDA-thread emulator, host utilities compiled as is).  Skipped when oracle/_ref was never
built (no /root/reference); tests/test_oracle_golden.py then pins the oracle through the committed golden values."""
import numpy as np
import pytest

from conftest import load_mesh, random_tris

import ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def inputs():
    out = [("cornellbox", load_mesh("cornellbox")), ("uniform", random_tris(3000, 1)), ("clustered", random_tris(2000, 2, "clustered")),
           ("flat", random_tris(1500, 3, "flat")), ("anisotropic", random_tris(2500, 4, "anisotropic")), ("duplicate", random_tris(200, 5, "duplicate"))]
    b = load_mesh("bunny")
    if b is not None:
        out.append(("bunny", b))
    return out


@pytest.mark.parametrize("name,tris", inputs(), ids=[n for n, _ in inputs()])
def test_pipeline_matches_reference_kernels(oracle, name, tris):
    n = tris.size
    refs, boxes, scene = oracle.primrefs(tris)
    # PrimRefs: doEarlySplitClipping (host, Utility.cpp:456) and InitPrimRefs (device twin, CommonBlocksKernel.h:80)
    assert ref.early_split(tris).tobytes() == refs.tobytes()
    assert ref.init_primrefs(tris).tobytes() == refs.tobytes()
    # Morton codes: CalculateMortonCodesPrimRef / CalculateMortonCodes
    k_ref, v_ref = ref.morton(refs, scene)
    k, v = oracle.morton_codes(refs, scene)
    assert np.array_equal(k, k_ref) and np.array_equal(v, v_ref)
    k2, _ = ref.morton(boxes, scene)
    assert np.array_equal(k, k2)
    sk, sv = oracle.sort_kv(k, v)
    order = np.argsort(k, kind="stable")
    assert np.array_equal(sk, k[order]) and np.array_equal(sv, v[order])
    # Karras two-pass: InitBvhNodesPrimRef + BvhBuild + FitBvhNodes
    nodes_ref, parents_ref = ref.twopass_build(refs, sk, sv)
    nodes, parents = oracle.lbvh_karras(refs, sk, sv)
    assert nodes.tobytes() == nodes_ref.tobytes()
    assert np.array_equal(parents, parents_ref)
    # Apetrei single-pass: InitBvhNodes + BvhBuildAndFit
    sp_ref, root_ref = ref.singlepass_build(tris, sk, sv)
    sp, root = oracle.lbvh_apetrei(tris, sk, sv)
    assert root == root_ref and sp.tobytes() == sp_ref.tobytes()
    # collapse (LBVH layout) + cost functions + validators
    wide_ref, wl_ref, cnt_ref = ref.collapse(nodes_ref, None, 0, n)
    wide, wl, cnt = oracle.collapse4(nodes, None, 0, n)
    assert cnt == cnt_ref and wl.tobytes() == wl_ref.tobytes()
    for f in ("aabb", "child", "parent", "childCount"):  # bytes 120..127 are padding the reference leaves unset
        assert np.array_equal(wide[f], wide_ref[f]), f
    assert oracle.cost_bvh4(wide, wl, boxes, 0, n) == ref.cost_bvh4(wide_ref, wl_ref, boxes, 0, n)
    assert oracle.cost_lbvh(nodes, 0, n) == ref.cost_lbvh(nodes_ref, 0, n)
    assert ref.check_bvh4(wide, wl, 0, n) and ref.check_root_aabb(nodes, 0, n)
    # PLOC layout: SetupClusters and the PLOC-layout collapse run on the oracle's PLOC++ tree
    pn, pl, _ = oracle.ploc(boxes, sv)
    _, leaves_ref, idx_ref = ref.ploc_setup(boxes, sv)
    assert pl.tobytes() == leaves_ref.tobytes() and np.array_equal(idx_ref, np.arange(n) + n - 1)
    w2_ref, wl2_ref, c2_ref = ref.collapse(pn, pl, 0, n)
    w2, wl2, c2 = oracle.collapse4(pn, pl, 0, n)
    assert c2 == c2_ref and wl2.tobytes() == wl2_ref.tobytes()
    for f in ("aabb", "child", "parent", "childCount"):
        assert np.array_equal(w2[f], w2_ref[f]), f


def test_morton_function_matches_reference_on_random_extents(oracle):
    rng = np.random.default_rng(7)
    for _ in range(300):
        ext = (10.0 ** rng.uniform(-3, 4, size=3)).astype(np.float32)
        if rng.random() < 0.1:
            ext[rng.integers(0, 3)] = 0.0
        cfg = oracle.morton_config(ext)
        for _ in range(20):
            p = rng.uniform(-0.1, 1.1, size=3).astype(np.float32)
            assert oracle.morton_code(p, cfg) == ref.morton_code(p, ext), (ext, p)


def test_synth_rng_matches_reference(oracle):
    import ctypes as C
    t = oracle.synth_uniform(1000, 0x00B20010, first=5, count=3, half=2.0)
    for k in range(3):
        s = C.c_uint32(ref.emul().ref_tea16(5 + k, 0x00B20010))
        r = [ref.emul().ref_randf(C.byref(s)) for _ in range(12)]
        c = [np.float32(-1000.0) + np.float32(2000.0) * np.float32(r[j]) for j in range(3)]
        v = [np.float32(c[j % 3]) + (np.float32(r[3 + j]) - np.float32(0.5)) * np.float32(4.0) for j in range(9)]
        assert np.array_equal(t["v"][k].reshape(-1), np.array(v, dtype=np.float32))


def test_traversal_matches_reference_cpu(oracle):
    tris = load_mesh("cornellbox")
    n = tris.size
    b = oracle.build_lbvh(tris)
    from b2bvh import types as T
    cam = T.make_camera([0.0, 2.5, 5.8, 0.0], oracle.qt_rotation([0.0, 0.0, 1.0, -1.57]), 45.0 * np.float32(np.pi) / 180.0)
    tr = T.make_transform([0.0, 0.0, -5.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0, 1.0])
    w = h = 64
    rays = oracle.generate_rays(cam, w, h)
    hits, cnt = oracle.traverse(rays, b["nodes"], None, tris, tr, 0, n)
    img = ref.traverse_cpu(rays, b["nodes"], tris, tr, w, h, n)
    hit_ref = img[:, 3] != 0
    # the reference CPU traversal writes alpha only for hit pixels
    assert int(hit_ref.sum()) == cnt
    assert np.array_equal(hit_ref, hits["primIdx"] != 0xFFFFFFFF)
