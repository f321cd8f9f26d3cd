"""ctypes binding of include/b2bvh.h — the same symbols the C++ host classes bind with dlopen.
No compute happens here; every call goes to libb2bvh.so (hand-written CUDA).  Missing library => ImportError-like
RuntimeError at load(), never a silent fallback."""
import ctypes as C
import os

import numpy as np

from . import types as T

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2BVH_LIB: the same override the C++ host layer honours (host/BvhConstruction.h) — development builds of the library, never a fallback
LIB_PATH = os.environ.get("B2BVH_LIB") or os.path.join(os.path.dirname(_HERE), "libb2bvh.so")

ABI_VERSION = 6  # B2BVH_ABI_VERSION of include/b2bvh.h: the struct layouts mirrored below
TWO_PASS_LBVH, SINGLE_PASS_LBVH, PLOCPP, HPLOC = 0, 1, 2, 3
TRAVERSE_WHILE, TRAVERSE_SPECULATIVE_WHILE, TRAVERSE_IFIF, TRAVERSE_RESTART_TRAIL, TRAVERSE_WIDE4 = 0, 1, 2, 3, 4
T_EXTENTS, T_MORTON, T_SORT, T_BUILD, T_TRAVERSAL, T_COLLAPSE, T_RAYGEN, T_COUNT = range(8)
STAGE_NAMES = ["CalculateCentroidExtentsTime", "CalculateMortonCodesTime", "SortingTime", "BvhBuildTime", "TraversalTime",
               "CollapseTime", "RayGenTime"]


class Aabb(C.Structure):
    _fields_ = [("mn", C.c_float * 3), ("mx", C.c_float * 3)]


class BuildOpts(C.Structure):
    _fields_ = [("collapse", C.c_uint32), ("tris_on_device", C.c_uint32), ("use_scene_box", C.c_uint32), ("scene_box", Aabb),
                ("stage_timing", C.c_uint32), ("karras_two_kernel", C.c_uint32), ("boxes_ready", C.c_uint32),
                ("d_scene_negmin_max", C.c_void_p), ("lbvh_second_level", C.c_uint32), ("merge_max_ctas", C.c_uint32), ("use_graph", C.c_uint32), ("split_sa_max", C.c_float), ("morton_bits", C.c_uint32), ("defer_sync", C.c_uint32), ("d_root_box_out", C.c_void_p)]


class Tree(C.Structure):
    _fields_ = [("algo", C.c_uint32), ("n_prims", C.c_uint32), ("n_internal", C.c_uint32), ("root", C.c_uint32), ("n_wide", C.c_uint32),
                ("leaves_separate", C.c_uint32),
                ("d_triangleBuff", C.c_void_p), ("d_triangleAabb", C.c_void_p), ("d_sceneExtents", C.c_void_p),
                ("d_mortonCodeKeys", C.c_void_p), ("d_mortonCodeValues", C.c_void_p), ("d_sortedMortonCodeKeys", C.c_void_p),
                ("d_sortedMortonCodeValues", C.c_void_p), ("d_bvhNodes", C.c_void_p), ("d_parentIdxs", C.c_void_p),
                ("d_leafNodes", C.c_void_p), ("d_wideBvhNodes", C.c_void_p), ("d_wideLeafNodes", C.c_void_p),
                ("stage_ms", C.c_float * T_COUNT), ("build_ms", C.c_float), ("h2d_ms", C.c_float), ("n_iterations", C.c_uint32),
                ("n_launches", C.c_uint32), ("n_triangles", C.c_uint32), ("n_split_levels", C.c_uint32), ("split_ms", C.c_float),
                ("reserved", C.c_uint32), ("d_primRefIdx", C.c_void_p), ("d_mortonCodeKeys64", C.c_void_p), ("d_sortedMortonCodeKeys64", C.c_void_p),
                ("morton_bits", C.c_uint32), ("reserved4", C.c_uint32)]


class Batch(C.Structure):
    _fields_ = [("n_items", C.c_uint32), ("n_prims_total", C.c_uint32), ("n_nodes_total", C.c_uint32), ("reserved", C.c_uint32),
                ("d_triangles", C.c_void_p), ("d_bvhNodes", C.c_void_p), ("d_primRefs", C.c_void_p), ("d_rootNodes", C.c_void_p),
                ("d_sceneExtents", C.c_void_p), ("d_leafOffsets", C.c_void_p), ("d_nodeOffsets", C.c_void_p), ("build_ms", C.c_float),
                ("h2d_ms", C.c_float)]


# every symbol include/b2bvh.h declares (tests check the library exports each of them)
SYMBOLS = ["b2bvh_ctx_create", "b2bvh_ctx_destroy", "b2bvh_device_name", "b2bvh_device_sm_count", "b2bvh_alloc", "b2bvh_free",
           "b2bvh_memset", "b2bvh_h2d", "b2bvh_d2h", "b2bvh_h2d_async", "b2bvh_d2h_async", "b2bvh_d2d", "b2bvh_sync", "b2bvh_host_alloc_pinned", "b2bvh_host_free_pinned",
           "b2bvh_last_error", "b2bvh_build", "b2bvh_build_finish", "b2bvh_build_batched", "b2bvh_scene_extents", "b2bvh_morton_codes", "b2bvh_sort_pairs", "b2bvh_lbvh_from_sorted64", "b2bvh_range_extract", "b2bvh_generate_rays",
           "b2bvh_traverse", "b2bvh_traverse_ex", "b2bvh_heat_map", "b2bvh_shard_extents", "b2bvh_top_level", "b2bvh_cost_bvh4", "b2bvh_cost_lbvh", "b2bvh_tree_cost",
           "b2bvh_abi_version", "b2bvh_synth_uniform", "b2bvh_synth_clustered", "b2bvh_profile_enable", "b2bvh_profile_count", "b2bvh_profile_entry",
           "b2bvh_global_partition", "b2bvh_global_sort", "b2bvh_global_tree", "b2bvh_global_top", "b2bvh_build_sharded", "b2bvh_global_assemble", "b2bvh_collapse_bvh2"]

_lib = None


class B2bvhError(RuntimeError):
    pass


def load():
    """dlopen libb2bvh.so.  Fails loudly when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2bvhError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         f"(nvcc, sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32, sz, fp = C.c_void_p, C.c_uint32, C.c_size_t, C.POINTER(C.c_float)
    sig = {
        "b2bvh_ctx_create": [C.c_int, vp, C.POINTER(vp)], "b2bvh_ctx_destroy": [vp], "b2bvh_device_name": [vp, C.c_char_p, sz],
        "b2bvh_device_sm_count": [vp, C.POINTER(C.c_int)], "b2bvh_alloc": [vp, sz, C.POINTER(vp)], "b2bvh_free": [vp, vp],
        "b2bvh_memset": [vp, vp, C.c_int, sz], "b2bvh_h2d": [vp, vp, vp, sz], "b2bvh_d2h": [vp, vp, vp, sz], "b2bvh_d2d": [vp, vp, vp, sz], "b2bvh_h2d_async": [vp, vp, vp, sz], "b2bvh_d2h_async": [vp, vp, vp, sz], "b2bvh_sync": [vp],
        "b2bvh_host_alloc_pinned": [sz, C.POINTER(vp)], "b2bvh_host_free_pinned": [vp],
        "b2bvh_build": [vp, C.c_int, vp, u32, C.POINTER(BuildOpts), C.POINTER(Tree)],
        "b2bvh_build_batched": [vp, vp, u32, vp, u32, C.POINTER(Batch)], "b2bvh_build_finish": [vp, C.POINTER(Tree)],
        "b2bvh_scene_extents": [vp, vp, u32, vp, vp], "b2bvh_morton_codes": [vp, vp, vp, u32, vp, vp],
        "b2bvh_sort_pairs": [vp, vp, vp, vp, vp, u32, u32, u32],
        "b2bvh_lbvh_from_sorted64": [vp, vp, vp, vp, u32, C.c_int, vp, vp, C.POINTER(u32)],
        "b2bvh_range_extract": [vp, vp, u32, u32, C.c_int, u32, u32, u32, u32, vp, vp, C.POINTER(u32)],
        "b2bvh_global_partition": [vp, vp, vp, u32, u32, vp, u32, vp, vp, vp, vp], "b2bvh_global_sort": [vp, vp, u32, vp, vp, vp],
        "b2bvh_global_tree": [vp, vp, vp, vp, vp, u32, vp, C.c_int, C.c_int, u32, u32, C.c_int, vp, vp, vp],
        "b2bvh_global_top": [vp, vp, vp, u32, u32, C.c_int, vp, vp],
        "b2bvh_build_sharded": [vp, u32, C.c_int, vp, vp, C.POINTER(BuildOpts), vp, vp, vp],
        "b2bvh_global_assemble": [vp, vp, vp, u32, u32, u32, vp, vp, vp, u32, vp, vp], "b2bvh_collapse_bvh2": [vp, vp, vp, vp, u32, vp, vp, C.POINTER(u32)],
        "b2bvh_generate_rays": [vp, vp, u32, u32, vp, fp],
        "b2bvh_traverse": [vp, C.POINTER(Tree), vp, u32, vp, C.c_int, vp, vp, fp],
        "b2bvh_traverse_ex": [vp, C.POINTER(Tree), vp, u32, vp, C.c_int, vp, vp, vp, fp], "b2bvh_heat_map": [vp, u32, vp],
        "b2bvh_shard_extents": [vp, vp, u32, u32, vp], "b2bvh_top_level": [vp, vp, u32, vp],
        "b2bvh_cost_bvh4": [vp, vp, vp, u32, u32, u32], "b2bvh_cost_lbvh": [vp, u32, u32, u32],
        "b2bvh_tree_cost": [vp, C.POINTER(Tree), fp], "b2bvh_abi_version": [], "b2bvh_last_error": [],
        "b2bvh_synth_uniform": [vp, C.c_uint64, u32, u32, C.c_float, vp], "b2bvh_synth_clustered": [vp, C.c_uint64, u32, u32, C.c_float, vp], "b2bvh_profile_enable": [vp, C.c_int],
        "b2bvh_profile_count": [vp, C.POINTER(C.c_int)], "b2bvh_profile_entry": [vp, C.c_int, C.c_char_p, sz, fp],
    }
    for name, args in sig.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = C.c_int
    lib.b2bvh_last_error.restype = C.c_char_p
    lib.b2bvh_cost_bvh4.restype = C.c_float
    lib.b2bvh_cost_lbvh.restype = C.c_float
    lib.b2bvh_abi_version.restype = C.c_uint32
    if lib.b2bvh_abi_version() != ABI_VERSION:
        raise B2bvhError(f"{LIB_PATH} speaks ABI version {lib.b2bvh_abi_version()}, this mirror was written for {ABI_VERSION} "
                         f"(struct layouts differ): rebuild the library")
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        raise B2bvhError(f"{what} failed with status {status}: {load().b2bvh_last_error().decode()}")


def _hp(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Replaces BvhConstruction::Context (src/Context.cpp:7-22): one device, one stream."""

    def __init__(self, device=0, stream=None):
        self.lib = load()
        h = C.c_void_p()
        check(self.lib.b2bvh_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)), "b2bvh_ctx_create")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2bvh_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def name(self):
        buf = C.create_string_buffer(256)
        check(self.lib.b2bvh_device_name(self.h, buf, 256))
        return buf.value.decode()

    def sm_count(self):
        v = C.c_int()
        check(self.lib.b2bvh_device_sm_count(self.h, C.byref(v)))
        return v.value

    # GpuMemory<T> look-alike primitives
    def alloc(self, nbytes):
        p = C.c_void_p()
        check(self.lib.b2bvh_alloc(self.h, nbytes, C.byref(p)), "b2bvh_alloc")
        return p.value

    def free(self, dptr):
        check(self.lib.b2bvh_free(self.h, C.c_void_p(dptr)), "b2bvh_free")

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.alloc(arr.nbytes)
        check(self.lib.b2bvh_h2d(self.h, C.c_void_p(d), _hp(arr), arr.nbytes), "b2bvh_h2d")
        return d

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        check(self.lib.b2bvh_h2d(self.h, C.c_void_p(dptr), _hp(arr), arr.nbytes), "b2bvh_h2d")

    def download(self, dptr, dtype, count):
        out = np.zeros(count, dtype=dtype)
        if count:
            check(self.lib.b2bvh_d2h(self.h, _hp(out), C.c_void_p(dptr), out.nbytes), "b2bvh_d2h")
        return out

    def sync(self):
        check(self.lib.b2bvh_sync(self.h))

    def pinned(self, nbytes):
        p = C.c_void_p()
        check(self.lib.b2bvh_host_alloc_pinned(nbytes, C.byref(p)), "b2bvh_host_alloc_pinned")
        return p.value

    # ---- stages ----
    def build(self, algo, tris, n=None, collapse=True, tris_on_device=False, scene_box=None, karras_two_kernel=False, boxes_ready=False,
              d_scene_negmin_max=None, lbvh_second_level=0, merge_max_ctas=0, use_graph=False, split_sa_max=0.0, morton_bits=0, defer_sync=False, d_root_box_out=None):
        """tris: TRIANGLE[n] numpy array (host) or an int device/pinned-host pointer (then pass n)."""
        opts = BuildOpts()
        opts.collapse = 1 if collapse else 0
        opts.tris_on_device = 1 if tris_on_device else 0
        opts.stage_timing = 1
        opts.karras_two_kernel = 1 if karras_two_kernel else 0
        opts.boxes_ready = 1 if boxes_ready else 0
        opts.lbvh_second_level = int(lbvh_second_level)
        opts.merge_max_ctas = int(merge_max_ctas)
        opts.use_graph = 1 if use_graph else 0
        opts.split_sa_max = float(split_sa_max)
        opts.morton_bits = int(morton_bits)
        opts.defer_sync = 1 if defer_sync else 0
        if d_root_box_out:
            opts.d_root_box_out = int(d_root_box_out)
        if d_scene_negmin_max:
            opts.d_scene_negmin_max = int(d_scene_negmin_max)
        if scene_box is not None:
            opts.use_scene_box = 1
            sb = np.asarray(scene_box, dtype=np.float32).reshape(6)
            for k in range(3):
                opts.scene_box.mn[k] = float(sb[k])
                opts.scene_box.mx[k] = float(sb[3 + k])
        if isinstance(tris, np.ndarray):
            assert tris.dtype == T.TRIANGLE and tris.flags["C_CONTIGUOUS"]
            n = tris.size
            ptr = _hp(tris)
            self._keepalive = tris
        else:
            ptr = C.c_void_p(int(tris))
        tree = Tree()
        check(self.lib.b2bvh_build(self.h, int(algo), ptr, int(n), C.byref(opts), C.byref(tree)), "b2bvh_build")
        return tree

    def build_batched(self, tris, counts, n_total=None, tris_on_device=False):
        """BatchedBvhBuilder::build: tris = TRIANGLE array of all items back to back (or a device pointer), counts = triangles per item."""
        counts = np.ascontiguousarray(counts, dtype=np.uint32)
        if isinstance(tris, np.ndarray):
            assert tris.dtype == T.TRIANGLE and tris.flags["C_CONTIGUOUS"] and tris.size == int(counts.sum())
            ptr = _hp(tris)
        else:
            ptr = C.c_void_p(int(tris))
        b = Batch()
        check(self.lib.b2bvh_build_batched(self.h, ptr, 1 if tris_on_device else 0, _hp(counts), counts.size, C.byref(b)), "b2bvh_build_batched")
        return b

    def fetch_batch(self, b):
        return dict(nodes=self.download(b.d_bvhNodes, T.BVH2_NODE, b.n_nodes_total), leaves=self.download(b.d_primRefs, T.PRIM_REF, b.n_prims_total),
                    roots=self.download(b.d_rootNodes, np.uint32, b.n_items), scenes=self.download(b.d_sceneExtents, T.AABB, b.n_items),
                    leaf_off=self.download(b.d_leafOffsets, np.uint32, b.n_items + 1), node_off=self.download(b.d_nodeOffsets, np.uint32, b.n_items + 1))

    def build_finish(self, tree):
        """Completes a build enqueued with defer_sync=True: synchronises and fills root, n_wide, times."""
        check(self.lib.b2bvh_build_finish(self.h, C.byref(tree)), "b2bvh_build_finish")
        return tree

    def tree_cost(self, tree):
        c = C.c_float()
        check(self.lib.b2bvh_tree_cost(self.h, C.byref(tree), C.byref(c)), "b2bvh_tree_cost")
        return float(c.value)

    def fetch(self, tree):
        """D2H of everything a build produced (GpuMemory::getData on every member)."""
        n = tree.n_prims
        separate = bool(tree.leaves_separate)
        r = dict(n=n, root=tree.root, n_wide=tree.n_wide)
        r["boxes"] = self.download(tree.d_triangleAabb, T.AABB, n)
        r["scene"] = self.download(tree.d_sceneExtents, T.AABB, 1)
        r["keys"] = self.download(tree.d_mortonCodeKeys, np.uint32, n)
        r["vals"] = self.download(tree.d_mortonCodeValues, np.uint32, n)
        r["keys64"] = self.download(tree.d_mortonCodeKeys64, np.uint64, n) if tree.d_mortonCodeKeys64 else None
        r["skeys64"] = self.download(tree.d_sortedMortonCodeKeys64, np.uint64, n) if tree.d_sortedMortonCodeKeys64 else None
        r["skeys"] = self.download(tree.d_sortedMortonCodeKeys, np.uint32, n)
        r["svals"] = self.download(tree.d_sortedMortonCodeValues, np.uint32, n)
        r["nodes"] = self.download(tree.d_bvhNodes, T.BVH2_NODE, n - 1 if separate else 2 * n - 1)
        r["parents"] = self.download(tree.d_parentIdxs, np.uint32, 2 * n - 1) if tree.d_parentIdxs else None
        r["leaves"] = self.download(tree.d_leafNodes, T.PRIM_REF, n) if separate else None
        r["prim_idx"] = self.download(tree.d_primRefIdx, np.uint32, n) if tree.d_primRefIdx else None
        if tree.n_wide:
            r["wide"] = self.download(tree.d_wideBvhNodes, T.BVH4_NODE, tree.n_wide)
            r["wide_leaves"] = self.download(tree.d_wideLeafNodes, T.PRIM_NODE, n)
        return r

    def sort_pairs(self, keys, vals, start_bit=0, end_bit=32):
        keys = np.ascontiguousarray(keys, dtype=np.uint32)
        n = keys.size
        dk = self.upload(keys)
        dv = self.upload(np.ascontiguousarray(vals, dtype=np.uint32)) if vals is not None else None
        ko, vo = self.alloc(4 * n), self.alloc(4 * n)
        try:
            check(self.lib.b2bvh_sort_pairs(self.h, C.c_void_p(dk), C.c_void_p(dv) if dv else None, C.c_void_p(ko), C.c_void_p(vo), n,
                                            start_bit, end_bit), "b2bvh_sort_pairs")
            return self.download(ko, np.uint32, n), self.download(vo, np.uint32, n)
        finally:
            for p in (dk, dv, ko, vo):
                if p:
                    self.free(p)

    def lbvh_from_sorted64(self, keys64, vals, d_boxes, karras):
        """The hierarchy stage over sorted 64-bit keys (host arrays; d_boxes: device AABB array indexed by vals).  Returns nodes, root."""
        keys64 = np.ascontiguousarray(keys64, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        n = keys64.size
        dk, dv = self.upload(keys64), self.upload(vals)
        dn, dp = self.alloc((2 * n - 1) * 32), self.alloc((2 * n - 1) * 4)
        root = C.c_uint32()
        try:
            check(self.lib.b2bvh_lbvh_from_sorted64(self.h, C.c_void_p(dk), C.c_void_p(dv), C.c_void_p(d_boxes), n, 1 if karras else 0, C.c_void_p(dn),
                                                    C.c_void_p(dp), C.byref(root)), "b2bvh_lbvh_from_sorted64")
            return self.download(dn, T.BVH2_NODE, 2 * n - 1), int(root.value)
        finally:
            for p in (dk, dv, dn, dp):
                self.free(p)

    def range_tree(self, keys64, vals, d_boxes, karras, ghost_left, ghost_right, first_pos, n_global):
        """One rank of the globally sorted build on the device: hierarchy stage over the (ghost-extended) range, then extraction.
        Returns (nodes[2m-1] with global child indices, artefacts invalid; left-over clusters as a CLUSTER array)."""
        keys64 = np.ascontiguousarray(keys64, dtype=np.uint64)
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        m = keys64.size
        dk, dv = self.upload(keys64), self.upload(vals)
        dn, dp, do, dc = self.alloc((2 * m - 1) * 32), self.alloc((2 * m - 1) * 4), self.alloc((2 * m - 1) * 32), self.alloc(256 * 48)
        root, cnt = C.c_uint32(), C.c_uint32()
        try:
            check(self.lib.b2bvh_lbvh_from_sorted64(self.h, C.c_void_p(dk), C.c_void_p(dv), C.c_void_p(d_boxes), m, 1 if karras else 0, C.c_void_p(dn),
                                                    C.c_void_p(dp), C.byref(root)), "b2bvh_lbvh_from_sorted64")
            check(self.lib.b2bvh_range_extract(self.h, C.c_void_p(dn), m, root.value, 1 if karras else 0, 1 if ghost_left else 0, 1 if ghost_right else 0,
                                               int(first_pos), int(n_global), C.c_void_p(do), C.c_void_p(dc), C.byref(cnt)), "b2bvh_range_extract")
            return self.download(do, T.BVH2_NODE, 2 * m - 1), self.download(dc, T.CLUSTER, cnt.value)
        finally:
            for p in (dk, dv, dn, dp, do, dc):
                self.free(p)

    def synth_uniform(self, n_total, seed, first=0, count=None, half=None, clustered=False):
        """synth_uniform_v1 (or synth_clustered_v1) triangles [first, first+count) generated on the device; returns the device pointer."""
        count = n_total if count is None else count
        h = np.float32(1000.0 * float(n_total) ** (-1.0 / 3.0)) if half is None else np.float32(half)
        d = self.alloc(count * 64)
        f = self.lib.b2bvh_synth_clustered if clustered else self.lib.b2bvh_synth_uniform
        check(f(self.h, int(first), int(count), int(seed), C.c_float(float(h)), C.c_void_p(d)), "b2bvh_synth")
        return d

    def profile(self, on=True):
        check(self.lib.b2bvh_profile_enable(self.h, 1 if on else 0))

    def profile_entries(self):
        n = C.c_int()
        check(self.lib.b2bvh_profile_count(self.h, C.byref(n)))
        out = []
        buf = C.create_string_buffer(64)
        ms = C.c_float()
        for i in range(n.value):
            check(self.lib.b2bvh_profile_entry(self.h, i, buf, 64, C.byref(ms)))
            out.append((buf.value.decode(), float(ms.value)))
        return out

    def generate_rays(self, cam, width, height):
        d_cam_host = np.ascontiguousarray(cam)
        d_rays = self.alloc(width * height * 32)
        ms = C.c_float()
        check(self.lib.b2bvh_generate_rays(self.h, _hp(d_cam_host), width, height, C.c_void_p(d_rays), C.byref(ms)), "b2bvh_generate_rays")
        return d_rays, float(ms.value)

    def traverse(self, tree, d_rays, n_rays, transform, kernel=TRAVERSE_WHILE, want_rgba=False, want_counter=False):
        """Returns hits, rgba (or None), ms — and the per-ray triangle-test counter as a fourth value when want_counter."""
        d_hits = self.alloc(n_rays * 32)
        d_rgba = self.alloc(n_rays * 4) if want_rgba else None
        d_cnt = self.alloc(n_rays * 4) if want_counter else None
        ms = C.c_float()
        tr = np.ascontiguousarray(transform)
        try:
            check(self.lib.b2bvh_traverse_ex(self.h, C.byref(tree), C.c_void_p(d_rays), n_rays, _hp(tr), int(kernel), C.c_void_p(d_hits),
                                             C.c_void_p(d_rgba) if d_rgba else None, C.c_void_p(d_cnt) if d_cnt else None, C.byref(ms)), "b2bvh_traverse_ex")
            hits = self.download(d_hits, T.HIT, n_rays)
            rgba = self.download(d_rgba, np.uint8, n_rays * 4).reshape(-1, 4) if d_rgba else None
            if want_counter:
                return hits, rgba, float(ms.value), self.download(d_cnt, np.uint32, n_rays)
            return hits, rgba, float(ms.value)
        finally:
            self.free(d_hits)
            if d_rgba:
                self.free(d_rgba)
            if d_cnt:
                self.free(d_cnt)


def heat_map(counter):
    """Utility::generateTraversalHeatMap colouring of a ray counter (host)."""
    counter = np.ascontiguousarray(counter, dtype=np.uint32)
    rgba = np.zeros((counter.size, 4), dtype=np.uint8)
    check(load().b2bvh_heat_map(_hp(counter), counter.size, _hp(rgba)), "b2bvh_heat_map")
    return rgba


def cost_bvh4(wide, wide_leaves, prim_boxes, n):
    lib = load()
    return float(lib.b2bvh_cost_bvh4(_hp(wide), _hp(wide_leaves), _hp(prim_boxes), 0, wide.size, n - 1))


def cost_lbvh(nodes, root, n):
    lib = load()
    return float(lib.b2bvh_cost_lbvh(_hp(nodes), root, n, n - 1))
