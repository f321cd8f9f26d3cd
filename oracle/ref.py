"""ctypes wrapper around oracle/_ref/*.so — the UNMODIFIED reference code compiled from /root/reference in place
(oracle/Makefile).  TEST INFRASTRUCTURE ONLY.  `available()` is False when the libraries were never built
(fresh clone without /root/reference); tests then fall back to the committed golden vectors."""
import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "hip-bvh-construction_b200"))
from b2bvh import types as T  # noqa: E402

_emul = _util = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_emul.so")) and os.path.exists(os.path.join(_HERE, "_ref", "libref_utility.so"))


def emul():
    global _emul
    if _emul is None:
        _emul = C.CDLL(os.path.join(_HERE, "_ref", "libref_emul.so"))
        for f in ("ref_morton_code", "ref_morton_plain", "ref_tea16", "ref_collapse_lbvh", "ref_collapse_ploc", "ref_singlepass_build"):
            getattr(_emul, f).restype = C.c_uint32
        _emul.ref_randf.restype = C.c_float
    return _emul


def util():
    global _util
    if _util is None:
        _util = C.CDLL(os.path.join(_HERE, "_ref", "libref_utility.so"))
        for f in ("ref_cost_bvh4", "ref_cost_lbvh", "ref_cost_sah"):
            getattr(_util, f).restype = C.c_float
        for f in ("ref_early_split", "ref_early_split_sa", "ref_load_obj"):
            getattr(_util, f).restype = C.c_uint32
    return _util


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(n):
    return C.c_uint32(int(n))


def morton_code(p, ext):
    return int(emul().ref_morton_code((C.c_float * 3)(*map(float, p)), (C.c_float * 3)(*map(float, ext))))


def init_primrefs(tris):
    refs = np.zeros(tris.size, dtype=T.PRIM_REF)
    emul().ref_init_primrefs(_p(refs), _p(tris), _u32(tris.size))
    return refs


def morton(boxes, scene):
    n = boxes.size
    keys = np.zeros(n, dtype=np.uint32); vals = np.zeros(n, dtype=np.uint32)
    if boxes.dtype == T.PRIM_REF:
        emul().ref_morton_primref(_p(boxes), _p(scene), _p(keys), _p(vals), _u32(n))
    else:
        emul().ref_morton_aabb(_p(boxes), _p(scene), _p(keys), _p(vals), _u32(n))
    return keys, vals


def twopass_build(refs, skeys, svals):
    n = skeys.size
    nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE); parents = np.zeros(2 * n - 1, dtype=np.uint32); flags = np.zeros(2 * n - 1, dtype=np.uint32)
    emul().ref_twopass_build(_p(refs), _p(skeys), _p(svals), _u32(n), _p(nodes), _p(parents), _p(flags))
    return nodes, parents


def singlepass_build(tris, skeys, svals):
    n = skeys.size
    nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE)
    root = emul().ref_singlepass_build(_p(tris), _p(skeys), _p(svals), _u32(n), _p(nodes))
    return nodes, int(root)


def ploc_setup(boxes, svals):
    n = svals.size
    nodes = np.zeros(n - 1, dtype=T.BVH2_NODE); leaves = np.zeros(n, dtype=T.PRIM_REF); idx = np.zeros(n, dtype=np.int32)
    sv = svals.copy(); bx = boxes.copy()
    emul().ref_ploc_setup(_p(nodes), _p(leaves), _p(sv), _p(bx), _p(idx), _u32(n))
    return nodes, leaves, idx


_emul_mt = None


def ploc_build_mt(nodes, leaves, idx):
    """Runs the reference Ploc / SinglePassPloc kernels (block-level emulation with one cooperative fiber per GPU thread, ref_shim/cuda_emul_mt.h) on
    the SetupClusters output; returns the internal nodes (numbering inside an iteration is timing dependent, as on a GPU)."""
    global _emul_mt
    if _emul_mt is None:
        _emul_mt = C.CDLL(os.path.join(_HERE, "_ref", "libref_emul_mt.so"))
        _emul_mt.ref_ploc_build_mt.restype = C.c_uint32
    n = leaves.size
    nd = nodes.copy(); lv = leaves.copy(); i0 = idx.astype(np.int32).copy(); i1 = np.full(n, -1, dtype=np.int32)
    launches = _emul_mt.ref_ploc_build_mt(_p(nd), _p(lv), _p(i0), _p(i1), _u32(n))
    return nd, int(launches)


def batched_build_mt(tris, n_items, prim_count):
    """Runs the reference's BatchedBuildKernelLbvh (BatchedBuildKernel.h:218-312; one 32-thread block per item, block-level emulation)
    on n_items items of prim_count triangles each.  Returns nodes, leaves, roots, scenes as the kernel writes them."""
    global _emul_mt
    if _emul_mt is None:
        _emul_mt = C.CDLL(os.path.join(_HERE, "_ref", "libref_emul_mt.so"))
        _emul_mt.ref_ploc_build_mt.restype = C.c_uint32
    assert tris.size == n_items * prim_count and 2 <= prim_count <= 32
    nodes = np.zeros(n_items * (prim_count - 1), dtype=T.BVH2_NODE)
    leaves = np.zeros(n_items * prim_count, dtype=T.PRIM_REF)
    roots = np.zeros(n_items, dtype=np.uint32)
    scenes = np.zeros(n_items, dtype=T.AABB)
    _emul_mt.ref_batched_build_mt(_p(tris), _u32(n_items), _u32(prim_count), _p(nodes), _p(leaves), _p(roots), _p(scenes))
    return nodes, leaves, roots, scenes


def hploc_build_mt(boxes, skeys, svals):
    """Runs the reference's HPloc kernel (HplocKernel.h:257-315) wavefront by wavefront under the block emulator, after its SetupClusters
    semantics (leaf g = {svals[g], boxes[svals[g]]}, nodeIdx[g] = g + n-1, parentIdx = INVALID; the kernel itself is shared with Ploc++'s
    setup in the single-thread emulator).  One statement is added to the compiled copy of the header (a barrier where a lock-step wavefront
    is converged anyway; oracle/Makefile, ref_shim/ref_emul_ploc_mt.cpp).  Returns (nodes, leaves, merged) — numbering follows the atomicAdd order."""
    global _emul_mt
    if _emul_mt is None:
        _emul_mt = C.CDLL(os.path.join(_HERE, "_ref", "libref_emul_mt.so"))
        _emul_mt.ref_ploc_build_mt.restype = C.c_uint32
    _emul_mt.ref_hploc_mt.restype = C.c_uint32
    n = svals.size
    assert (n - 1) % 32 != 0, "the reference launches n-1 threads rounded up to 32: leaf n-1 has no thread when (n-1) % 32 == 0"
    nodes = np.zeros(n - 1, dtype=T.BVH2_NODE)
    nodes["left"] = 0xFFFFFFFF; nodes["right"] = 0xFFFFFFFF
    leaves = np.zeros(n, dtype=T.PRIM_REF)
    leaves["primIdx"] = svals
    leaves["mn"] = boxes["mn"][svals]; leaves["mx"] = boxes["mx"][svals]
    idx = (np.arange(n, dtype=np.uint32) + np.uint32(n - 1)).copy()
    parent = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    keys = np.ascontiguousarray(skeys, dtype=np.uint32).copy()
    merged = _emul_mt.ref_hploc_mt(_p(nodes), _p(leaves), _p(keys), _p(idx), _p(parent), _u32(n))
    return nodes, leaves, int(merged)


def collapse(nodes, leaves, root, n):
    """Runs the reference CollapseToWide4Bvh (LBVH variant when leaves is None, PLOC variant otherwise)."""
    wide = np.zeros(2 * n, dtype=T.BVH4_NODE); wl = np.zeros(n, dtype=T.PRIM_NODE)
    nd = nodes.copy()
    if leaves is None:
        cnt = emul().ref_collapse_lbvh(_p(nd), _u32(root), _u32(n), _p(wide), _p(wl))
    else:
        # the PLOC variant reads bvh2Nodes[leafIdx] out of bounds (Ploc++Kernel.h:395-396, value unused): pad the array
        nd = np.concatenate([nd, np.zeros(n, dtype=T.BVH2_NODE)])
        lv = leaves.copy()
        cnt = emul().ref_collapse_ploc(_p(nd), _p(lv), _u32(root), _u32(n), _p(wide), _p(wl))
    return wide[:cnt].copy(), wl, int(cnt)


def cost_bvh4(wide, wl, prim_boxes, root, n):
    pb = prim_boxes.copy()
    return float(util().ref_cost_bvh4(_p(wide), _p(wl), _p(pb), _u32(root), _u32(wide.size), _u32(n - 1)))


def cost_lbvh(nodes, root, n):
    return float(util().ref_cost_lbvh(_p(nodes), _u32(root), _u32(n), _u32(n - 1)))


def cost_sah(nodes):
    return float(util().ref_cost_sah(_p(nodes), _u32(0), _u32(nodes.size)))


def check_bvh4(wide, wl, root, n):
    return bool(util().ref_check_bvh4(_p(wide), _p(wl), _u32(root), _u32(n - 1)))


def check_root_aabb(nodes, root, n):
    return bool(util().ref_check_root_aabb(_p(nodes), _u32(root), _u32(n), _u32(n - 1)))


def early_split(tris):
    out = np.zeros(tris.size, dtype=T.PRIM_REF)
    cnt = util().ref_early_split(_p(tris), _u32(tris.size), _p(out))
    assert cnt == tris.size
    return out


def morton_plain(p):
    """computeMortonCode (plain 10/10/10, CommonBlocksKernel.h:361-372 == BatchedBuildKernel.h:98-110), the reference's own code."""
    q = (C.c_float * 3)(*[float(x) for x in p])
    return int(emul().ref_morton_plain(q))


def early_split_sa(tris, sa_max):
    """Utility::doEarlySplitClipping(prims, refs, saMax), the reference's own code."""
    cnt = util().ref_early_split_sa(_p(tris), _u32(tris.size), C.c_float(float(sa_max)), None, _u32(0))
    out = np.zeros(cnt, dtype=T.PRIM_REF)
    util().ref_early_split_sa(_p(tris), _u32(tris.size), C.c_float(float(sa_max)), _p(out), _u32(cnt))
    return out


def load_obj(path, mtl_dir):
    n = util().ref_load_obj(path.encode(), mtl_dir.encode(), None, _u32(0))
    out = np.zeros(n, dtype=T.TRIANGLE)
    util().ref_load_obj(path.encode(), mtl_dir.encode(), _p(out), _u32(n))
    return out


def traverse_cpu(rays, nodes, tris, transform, width, height, n):
    dst = np.zeros(width * height * 4, dtype=np.uint8)
    tr = transform.copy()
    util().ref_traverse_cpu(_p(rays), _p(nodes), _u32(nodes.size), _p(tris), _u32(tris.size), _p(tr), _p(dst), _u32(width), _u32(height), _u32(n - 1))
    return dst.reshape(-1, 4)
