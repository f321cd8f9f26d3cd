"""ctypes wrapper around the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
import this from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never from the product package."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "hip-bvh-construction_b200"))
from b2bvh import types as T  # noqa: E402

_lib = None


def build(verbose=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_area.restype = C.c_float
        _lib.orc_cost_bvh4.restype = C.c_float
        _lib.orc_cost_lbvh.restype = C.c_float
        _lib.orc_cost_bvh2_ploc.restype = C.c_float
        _lib.orc_cost_binned_sah_quirk.restype = C.c_float
        _lib.orc_cost_binned_sah_proper.restype = C.c_float
        _lib.orc_ray_zdir.restype = C.c_float
        _lib.orc_ray_zdir.argtypes = [C.c_float]
        for f in ("orc_fnv1a_words", "orc_lbvh_apetrei", "orc_collapse4", "orc_bvh2_depth", "orc_traverse", "orc_traverse_kind", "orc_traverse_wide4", "orc_binned_sah_build",
                  "orc_morton_code_cfg", "orc_early_split", "orc_morton_plain", "orc_lbvh_apetrei64"):
            getattr(_lib, f).restype = C.c_uint32
        _lib.orc_morton60_point.restype = C.c_uint64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u32(n):
    return C.c_uint32(int(n))


class MortonCfg(C.Structure):
    _fields_ = [("axis", C.c_int * 3), ("pre", C.c_int * 2), ("swap", C.c_int), ("sum", C.c_int), ("nb", C.c_int * 3)]


def fnv1a(keys, vals=None):
    keys = np.ascontiguousarray(keys, dtype=np.uint32)
    if vals is not None:
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
    return int(lib().orc_fnv1a_words(_p(keys), _p(vals), C.c_uint64(keys.size)))


def primrefs(tris):
    n = tris.size
    refs = np.zeros(n, dtype=T.PRIM_REF)
    boxes = np.zeros(n, dtype=T.AABB)
    scene = np.zeros(1, dtype=T.AABB)
    lib().orc_primrefs(_p(tris), _u32(n), _p(refs), _p(boxes), _p(scene))
    return refs, boxes, scene


def early_split(tris, sa_max, max_levels=64):
    """Utility::doEarlySplitClipping with a finite saMax: PRIM_REF[m] in the reference's emission order."""
    args = (_p(tris), _u32(tris.size), C.c_float(float(sa_max)))
    cnt = int(lib().orc_early_split(*args, None, _u32(0), _u32(max_levels)))
    if cnt == 0:
        raise ValueError("early split does not terminate within max_levels (or exceeds 2^30 references)")
    out = np.zeros(cnt, dtype=T.PRIM_REF)
    lib().orc_early_split(*args, _p(out), _u32(cnt), _u32(max_levels))
    return out


def split_cost_boxes(refs):
    table = np.zeros(refs.size, dtype=T.AABB)
    lib().orc_split_cost_boxes(_p(refs), _u32(refs.size), _p(table))
    return table


def scene_of(boxes):
    scene = np.zeros(1, dtype=T.AABB)
    scene["mn"] = boxes["mn"].min(axis=0)
    scene["mx"] = boxes["mx"].max(axis=0)
    return scene


def morton_config(extent):
    cfg = MortonCfg()
    e = (C.c_float * 3)(*[float(x) for x in extent])
    lib().orc_morton_config(e, C.byref(cfg))
    return cfg


def morton_code(p, cfg):
    q = (C.c_float * 3)(*[float(x) for x in p])
    return int(lib().orc_morton_code_cfg(q, C.byref(cfg)))


def morton_codes(boxes, scene):
    """boxes: AABB[n] or PRIM_REF[n]."""
    n = boxes.size
    keys = np.zeros(n, dtype=np.uint32)
    vals = np.zeros(n, dtype=np.uint32)
    if boxes.dtype == T.PRIM_REF:
        base, stride = C.c_void_p(boxes.ctypes.data + 4), 28
    else:
        base, stride = C.c_void_p(boxes.ctypes.data), 24
    lib().orc_morton_codes(base, _u32(stride), _p(scene), _u32(n), _p(keys), _p(vals))
    return keys, vals


def morton60_point(p):
    q = (C.c_float * 3)(*[float(x) for x in p])
    return int(lib().orc_morton60_point(q))


def morton60_codes(boxes, scene):
    """60-bit plain Morton codes (20 bits per axis) of AABB[n] or PRIM_REF[n] centroids: uint64 keys, iota values."""
    n = boxes.size
    keys = np.zeros(n, dtype=np.uint64)
    vals = np.zeros(n, dtype=np.uint32)
    base, stride = (C.c_void_p(boxes.ctypes.data + 4), 28) if boxes.dtype == T.PRIM_REF else (C.c_void_p(boxes.ctypes.data), 24)
    lib().orc_morton60_codes(base, _u32(stride), _p(scene), _u32(n), _p(keys), _p(vals))
    return keys, vals


def sort_kv64(keys, vals):
    n = keys.size
    ko = np.zeros(n, dtype=np.uint64)
    vo = np.zeros(n, dtype=np.uint32)
    lib().orc_sort_kv64(_p(keys), _p(vals), _u32(n), _p(ko), _p(vo))
    return ko, vo


def sort_kv(keys, vals):
    n = keys.size
    ko = np.zeros(n, dtype=np.uint32)
    vo = np.zeros(n, dtype=np.uint32)
    lib().orc_sort_kv(_p(keys), _p(vals), _u32(n), _p(ko), _p(vo))
    return ko, vo


def lbvh_karras(refs, skeys, svals):
    """skeys: uint32 (reference, 30-bit codes) or uint64 (60-bit variant)."""
    n = skeys.size
    nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE)
    parents = np.zeros(2 * n - 1, dtype=np.uint32)
    f = lib().orc_lbvh_karras64 if skeys.dtype == np.uint64 else lib().orc_lbvh_karras
    f(_p(refs), _p(skeys), _p(svals), _u32(n), _p(nodes), _p(parents))
    return nodes, parents


def lbvh_apetrei(tris, skeys, svals):
    n = skeys.size
    nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE)
    f = lib().orc_lbvh_apetrei64 if skeys.dtype == np.uint64 else lib().orc_lbvh_apetrei
    root = f(_p(tris), _p(skeys), _p(svals), _u32(n), _p(nodes))
    return nodes, int(root)


def ploc(boxes, svals):
    n = svals.size
    nodes = np.zeros(n - 1, dtype=T.BVH2_NODE)
    leaves = np.zeros(n, dtype=T.PRIM_REF)
    stats = np.zeros(4, dtype=np.uint32)
    lib().orc_ploc(_p(boxes), _p(svals), _u32(n), _p(nodes), _p(leaves), _p(stats))
    return nodes, leaves, {"iterations": int(stats[0]), "global_iterations": int(stats[1]), "sum_clusters": int(stats[2]) | (int(stats[3]) << 32)}


def hploc(boxes, skeys, svals):
    n = svals.size
    nodes = np.zeros(n - 1, dtype=T.BVH2_NODE)
    leaves = np.zeros(n, dtype=T.PRIM_REF)
    stats = np.zeros(2, dtype=np.uint32)
    f = lib().orc_hploc64 if skeys.dtype == np.uint64 else lib().orc_hploc
    f(_p(boxes), _p(skeys), _p(svals), _u32(n), _p(nodes), _p(leaves), _p(stats))
    return nodes, leaves, {"merge_calls": int(stats[0]), "allocated": int(stats[1])}


def collapse4(nodes, leaves, root, n):
    wide = np.zeros(max(n - 1, 1), dtype=T.BVH4_NODE)
    wl = np.zeros(n, dtype=T.PRIM_NODE)
    cnt = lib().orc_collapse4(_p(nodes), _p(leaves), _u32(root), _u32(n), _p(wide), _p(wl))
    return wide[:cnt].copy(), wl, int(cnt)


def cost_bvh4(wide, wl, prim_boxes, root, n):
    return float(lib().orc_cost_bvh4(_p(wide), _p(wl), _p(prim_boxes), _u32(root), _u32(wide.size), _u32(n - 1)))


def cost_lbvh(nodes, root, n):
    return float(lib().orc_cost_lbvh(_p(nodes), _u32(root), _u32(n), _u32(n - 1)))


def cost_bvh2_ploc(nodes, leaves, root, n):
    return float(lib().orc_cost_bvh2_ploc(_p(nodes), _p(leaves), _u32(root), _u32(n)))


def check_bvh2(nodes, leaves, root, n):
    return bool(lib().orc_check_bvh2(_p(nodes), _p(leaves), _u32(root), _u32(n)))


def check_bvh4(wide, wl, root, n):
    return bool(lib().orc_check_bvh4(_p(wide), _p(wl), _u32(root), _u32(n - 1)))


def check_root_aabb(nodes, root, n):
    return bool(lib().orc_check_root_aabb(_p(nodes), _u32(root), _u32(n), _u32(n - 1)))


def bvh2_depth(nodes, root, n):
    return int(lib().orc_bvh2_depth(_p(nodes), _u32(root), _u32(n)))


def qt_rotation(axis_angle):
    a = (C.c_float * 4)(*[float(x) for x in axis_angle])
    o = (C.c_float * 4)()
    lib().orc_qt_rotation(a, o)
    return np.array(list(o), dtype=np.float32)


def ray_zdir(fov):
    return float(lib().orc_ray_zdir(C.c_float(float(fov))))


def generate_rays(cam, width, height):
    rays = np.zeros(width * height, dtype=T.RAY)
    lib().orc_generate_rays(_p(cam), _u32(width), _u32(height), _p(rays))
    return rays


def traverse(rays, nodes, leaves, tris, transform, root, n):
    hits = np.zeros(rays.size, dtype=T.HIT)
    cnt = lib().orc_traverse(_p(rays), _p(nodes), _p(leaves), _p(tris), _p(transform), _u32(root), _u32(n - 1), _u32(rays.size), _p(hits))
    return hits, int(cnt)


def traverse_kind(kind, rays, nodes, leaves, tris, transform, root, n):
    """kind 0: if-if (stack), 1: restart trail.  Returns hits, hit count, leaf tests per ray."""
    hits = np.zeros(rays.size, dtype=T.HIT)
    counter = np.zeros(rays.size, dtype=np.uint32)
    cnt = lib().orc_traverse_kind(_u32(kind), _p(rays), _p(nodes), _p(leaves), _p(tris), _p(transform), _u32(root), _u32(n - 1), _u32(rays.size), _p(hits),
                                  _p(counter))
    return hits, int(cnt), counter


def traverse_wide4(rays, wide, nodes, leaves, tris, transform, n):
    """Bvh4 traversal; leaf boxes / primitive ids come from the Bvh2 leaf records (nodes, or leaves for the separate-leaf layout)."""
    hits = np.zeros(rays.size, dtype=T.HIT)
    counter = np.zeros(rays.size, dtype=np.uint32)
    cnt = lib().orc_traverse_wide4(_p(rays), _p(wide), _p(nodes), _p(leaves), _p(tris), _p(transform), _u32(n - 1), _u32(rays.size), _p(hits), _p(counter))
    return hits, int(cnt), counter


def heat_map(counter):
    counter = np.ascontiguousarray(counter, dtype=np.uint32)
    rgba = np.zeros((counter.size, 4), dtype=np.uint8)
    lib().orc_heat_map(_p(counter), _u32(counter.size), _p(rgba))
    return rgba


def binned_sah(tris):
    n = tris.size
    nodes = np.zeros(3 * n - 1, dtype=T.SAH_NODE)
    cnt = int(lib().orc_binned_sah_build(_p(tris), _u32(n), _p(nodes)))
    return nodes[:cnt], cnt


def cost_binned_sah(nodes):
    quirk = float(lib().orc_cost_binned_sah_quirk(_p(nodes), _u32(0), _u32(nodes.size)))
    proper = float(lib().orc_cost_binned_sah_proper(_p(nodes), _u32(nodes.size)))
    return quirk, proper


def check_sah(nodes, n):
    return bool(lib().orc_check_sah(_p(nodes), _u32(n)))


def synth_half(n):
    """half-size h = 1000 * N^(-1/3), rounded to float once on the host."""
    return np.float32(1000.0 * float(n) ** (-1.0 / 3.0))


def synth_uniform(n, seed, first=0, count=None, half=None):
    count = n if count is None else count
    out = np.zeros(count, dtype=T.TRIANGLE)
    h = synth_half(n) if half is None else np.float32(half)
    lib().orc_synth_uniform(C.c_uint64(first), _u32(count), _u32(seed), C.c_float(float(h)), _p(out))
    return out


def synth_clustered(n, seed, first=0, count=None, half=None):
    count = n if count is None else count
    out = np.zeros(count, dtype=T.TRIANGLE)
    h = synth_half(n) if half is None else np.float32(half)
    lib().orc_synth_clustered(C.c_uint64(first), _u32(count), _u32(seed), C.c_float(float(h)), _p(out))
    return out


def top_level(root_boxes):
    g = root_boxes.size
    nodes = np.zeros(2 * g - 1, dtype=T.BVH2_NODE)
    lib().orc_top_level(_p(root_boxes), _u32(g), _p(nodes))
    return nodes


# ---- full pipelines (launch order of the reference builders) ----
def build_lbvh(tris, single_pass=False, scene_override=None, split_sa_max=None, morton_bits=30):
    """TwoPassLbvh::build (TwoPassLbvh.cpp:17-197) / SinglePassLbvh::build (SinglePassLbvh.cpp:17-188).
    scene_override: AABB[1] global scene box of a sharded build (replaces the local union for Morton coding).
    split_sa_max: TwoPassLbvh compiled with USE_PRIM_SPLITTING (TwoPassLbvh.cpp:23-28): the build runs over the split references."""
    if split_sa_max is not None:
        return _build_lbvh_split(tris, split_sa_max)
    n = tris.size
    refs, boxes, scene = primrefs(tris)
    if scene_override is not None:
        scene = scene_override
    if morton_bits == 60:
        keys, vals = morton60_codes(refs, scene)
        sk, sv = sort_kv64(keys, vals)
    else:
        keys, vals = morton_codes(refs, scene)
        sk, sv = sort_kv(keys, vals)
    if single_pass:
        nodes, root = lbvh_apetrei(tris, sk, sv)
    else:
        nodes, _ = lbvh_karras(refs, sk, sv)
        root = 0
    wide, wl, cnt = collapse4(nodes, None, root, n)
    cost = cost_bvh4(wide, wl, boxes, 0, n)
    return dict(scene=scene, keys=keys, vals=vals, skeys=sk, svals=sv, nodes=nodes, root=root, wide=wide, wide_leaves=wl,
                wide_count=cnt, cost=cost, boxes=boxes, refs=refs)


def _build_lbvh_split(tris, sa_max):
    refs = early_split(tris, sa_max)
    m = refs.size
    boxes = np.zeros(m, dtype=T.AABB)
    boxes["mn"] = refs["mn"]; boxes["mx"] = refs["mx"]
    scene = scene_of(boxes)                      # CalculatePrimRefExtents over the references
    keys, vals = morton_codes(refs, scene)
    sk, sv = sort_kv(keys, vals)
    nodes, parents = lbvh_karras(refs, sk, sv)   # leaf.left = refs[sv[g]].m_primIdx (InitBvhNodesPrimRef)
    wide, wl, cnt = collapse4(nodes, None, 0, m)
    cost = cost_bvh4(wide, wl, split_cost_boxes(refs), 0, m)
    return dict(scene=scene, keys=keys, vals=vals, skeys=sk, svals=sv, nodes=nodes, parents=parents, root=0, wide=wide, wide_leaves=wl,
                wide_count=cnt, cost=cost, boxes=boxes, refs=refs, prim_idx=refs["primIdx"].copy())


def morton_plain(p):
    q = (C.c_float * 3)(*[float(x) for x in p])
    return int(lib().orc_morton_plain(q))


def build_batched(tris, counts):
    """BatchedBvhBuilder::build (BatchedBuilder.cpp:16-77) restated: tris = all items back to back, counts[item] <= 32.
    Returns nodes (sum(n-1)), leaves (sum(n), sorted order per item), roots, scenes, node/leaf offsets."""
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    assert int(counts.sum()) == tris.size and counts.min() >= 1 and counts.max() <= 32
    nodes = np.zeros(max(int(counts.sum()) - counts.size, 1), dtype=T.BVH2_NODE)
    leaves = np.zeros(tris.size, dtype=T.PRIM_REF)
    roots = np.zeros(counts.size, dtype=np.uint32)
    scenes = np.zeros(counts.size, dtype=T.AABB)
    lib().orc_batched_build(_p(tris), _p(counts), _u32(counts.size), _p(nodes), _p(leaves), _p(roots), _p(scenes))
    leaf_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    node_off = np.concatenate([[0], np.cumsum(counts - 1)]).astype(np.uint32)
    return dict(nodes=nodes[:int(node_off[-1])], leaves=leaves, roots=roots, scenes=scenes, node_off=node_off, leaf_off=leaf_off)


def build_ploc(tris, hierarchical=False, scene_override=None, morton_bits=30):
    """PLOCNew::build (PLOC++Bvh.cpp:16-196) / HPLOC::build (Hploc.cpp:16-165).  morton_bits=60: the 60-bit variant (PLOC++ uses the sorted order alone, H-PLOC walks the 64-bit codes)."""
    n = tris.size
    refs, boxes, scene = primrefs(tris)
    if scene_override is not None:
        scene = scene_override
    if morton_bits == 60:
        keys, vals = morton60_codes(boxes, scene)
        sk, sv = sort_kv64(keys, vals)
    else:
        keys, vals = morton_codes(boxes, scene)
        sk, sv = sort_kv(keys, vals)
    if hierarchical:
        nodes, leaves, stats = hploc(boxes, sk, sv)
    else:
        nodes, leaves, stats = ploc(boxes, sv)
    wide, wl, cnt = collapse4(nodes, leaves, 0, n)
    cost = cost_bvh4(wide, wl, boxes, 0, n)
    return dict(scene=scene, keys=keys, vals=vals, skeys=sk, svals=sv, nodes=nodes, leaves=leaves, root=0, wide=wide,
                wide_leaves=wl, wide_count=cnt, cost=cost, boxes=boxes, stats=stats)


def build_sharded(tris, world, single_pass=True):
    """The sharded procedure of b2bvh/sharded.py restated sequentially: per-shard boxes, global box = union, per-shard LBVH in the
    global frame, top-level tree over the shard roots.  Returns (global scene AABB[1], [per-shard build dicts], top-level nodes)."""
    n = tris.size
    ranges = [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]
    scene = np.zeros(1, dtype=T.AABB)
    scene["mn"] = np.min([primrefs(tris[a:b])[2]["mn"][0] for a, b in ranges], axis=0)
    scene["mx"] = np.max([primrefs(tris[a:b])[2]["mx"][0] for a, b in ranges], axis=0)
    shards = [build_lbvh(np.ascontiguousarray(tris[a:b]), single_pass=single_pass, scene_override=scene) for a, b in ranges]
    roots = np.zeros(world, dtype=T.AABB)
    for r, s in enumerate(shards):
        roots[r]["mn"] = s["nodes"][s["root"]]["mn"]; roots[r]["mx"] = s["nodes"][s["root"]]["mx"]
    return scene, shards, top_level(roots)


def trace_sharded(tris, world, rays, transform, single_pass=True):
    """Primary rays through a sharded build, restated sequentially (b2bvh/sharded.py ShardedBuild.trace): every shard's sub-tree is
    traced on its own, the closest hit per ray is the minimum over (bits of t << 32 | global primitive index).
    Returns t float32[n], prim int64[n] (global index, -1 = miss), uv float32[n,2]."""
    n = tris.size
    _, shards, _ = build_sharded(tris, world, single_pass=single_pass)
    miss = np.int64((1 << 63) - 1)
    best = np.full(rays.size, miss, dtype=np.int64)
    uv = np.zeros((rays.size, 2), dtype=np.float32)
    for r, s in enumerate(shards):
        a, b = (n * r) // world, (n * (r + 1)) // world
        hits, _ = traverse(rays, s["nodes"], None, np.ascontiguousarray(tris[a:b]), transform, s["root"], b - a)
        hit = hits["primIdx"] != 0xFFFFFFFF
        key = (hits["t"].view(np.uint32).astype(np.int64) << 32) | (hits["primIdx"].astype(np.int64) + a)
        key = np.where(hit, key, miss)
        better = key < best
        best = np.where(better, key, best)
        uv[better] = hits["uv"][better]
    hit = best != miss
    t = np.where(hit, (best >> 32).astype(np.uint32).view(np.float32), np.float32(0))
    prim = np.where(hit, best & 0xFFFFFFFF, -1)
    return t.astype(np.float32), prim.astype(np.int64), uv



# ---- executable specification of the globally sorted multi-GPU build (next round's GPU work) ----
def _boundary_depth(skeys, g):
    """Common-prefix length of the augmented keys (key, sorted position) of the sorted leaves g and g+1; -1 outside the array.
    32-bit keys: clz of (key xor) in 0..31, else 32 + clz32(position xor); 64-bit keys: 0..63, else 64 + clz32."""
    n = skeys.size
    if g < 0 or g + 1 >= n:
        return -1
    bits = 64 if skeys.dtype == np.uint64 else 32
    kx = int(skeys[g]) ^ int(skeys[g + 1])
    if kx:
        return bits - kx.bit_length()
    return bits + (32 - (g ^ (g + 1)).bit_length())


def lbvh_by_ranges(refs, skeys, svals, ranges, karras):
    """The one-GPU LBVH built the way G ranks would build it: rank r owns the sorted positions ranges[r] = [a, b) of the GLOBALLY sorted
    sequence, knows the two keys across its edges, and merges clusters exactly as lbvh_tile_kernel does — two neighbours form a node when
    the boundary between them is deeper than both boundaries next to it — but only inside its range; the clusters it is left with (their
    parents straddle a rank boundary) are then concatenated in rank order and finished by the same rule (the 'all-gather + group kernel'
    step).  Returns nodes[2n-1] (LBVH layout), root, and the number of left-over clusters per rank.  Must equal lbvh_karras / lbvh_apetrei."""
    n = skeys.size
    nint = n - 1
    nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE)
    nodes["left"][:] = 0xFFFFFFFF; nodes["right"][:] = 0xFFFFFFFF
    for g in range(n):
        r = refs[svals[g]]
        nodes[nint + g] = (r["primIdx"], 0xFFFFFFFF, r["mn"], r["mx"])
    depth = [_boundary_depth(skeys, g) for g in range(-1, n)]  # depth[g + 1] = boundary between g and g+1
    D = lambda g: depth[g + 1]
    root = [None]

    def merge_rounds(clusters, lo_edge, hi_edge):
        """clusters: list of (lo, hi, node id); merges while some interior boundary is a local maximum of the depths; the boundaries
        outside [lo_edge, hi_edge) are not available (rank edges keep their own depth as the sentinel)."""
        while True:
            merged = False
            out = []
            i = 0
            while i < len(clusters):
                if i + 1 < len(clusters):
                    (lo, mid, idl), (_, hi, idr) = clusters[i], clusters[i + 1]
                    d0, dl, dr = D(mid - 1), D(lo - 1), D(hi - 1)
                    if d0 > dl and d0 > dr and lo >= lo_edge and hi <= hi_edge:
                        is_root = lo == 0 and hi == n
                        if karras:
                            nid = 0 if is_root else (hi - 1 if dr > dl else lo)
                        else:
                            nid = mid - 1
                        nodes[nid]["left"] = idl; nodes[nid]["right"] = idr
                        nodes[nid]["mn"] = np.minimum(nodes[idl]["mn"], nodes[idr]["mn"]); nodes[nid]["mx"] = np.maximum(nodes[idl]["mx"], nodes[idr]["mx"])
                        if is_root:
                            root[0] = nid
                        out.append((lo, hi, nid))
                        i += 2
                        merged = True
                        continue
                out.append(clusters[i])
                i += 1
            clusters = out
            if not merged:
                return clusters

    leftovers, counts = [], []
    for a, b in ranges:
        left = merge_rounds([(g, g + 1, nint + g) for g in range(a, b)], a, b)
        counts.append(len(left))
        leftovers += left
    final = merge_rounds(leftovers, 0, n)
    assert len(final) == 1 and final[0][0] == 0 and final[0][1] == n
    return nodes, (root[0] if root[0] is not None else final[0][2]), counts


def global_sort_by_exchange(keys, world, splitters):
    """The sort step of the globally sorted build, restated: rank r holds keys[shard r] with GLOBAL indices; every pair goes to the rank
    whose key interval [splitters[d-1], splitters[d]) contains its key (equal keys never split), the pieces a rank receives are
    concatenated in SOURCE-RANK order (so equal keys arrive in global-index order), and one local stable sort per rank finishes.  The
    concatenation over ranks must equal the stable sort of the whole array.  Returns (sorted keys, sorted global indices, per-rank counts)."""
    n = keys.size
    splitters = np.asarray(splitters, dtype=keys.dtype)
    assert splitters.size == world - 1 and np.all(splitters[1:] >= splitters[:-1])
    recv_k = [[] for _ in range(world)]
    recv_v = [[] for _ in range(world)]
    for r in range(world):
        a, b = (n * r) // world, (n * (r + 1)) // world
        k = keys[a:b]
        dest = np.searchsorted(splitters, k, side="right")
        for d in range(world):
            m = dest == d
            recv_k[d].append(k[m]); recv_v[d].append(np.arange(a, b, dtype=np.uint32)[m])
    out_k, out_v, counts = [], [], []
    for d in range(world):
        k = np.concatenate(recv_k[d]); v = np.concatenate(recv_v[d])
        order = np.argsort(k, kind="stable")
        out_k.append(k[order]); out_v.append(v[order]); counts.append(k.size)
    return np.concatenate(out_k), np.concatenate(out_v), counts


def lbvh_by_ranges_with_unchanged_builder(tris, refs, skeys, svals, ranges, karras, builder=None):
    """The same range-wise build, but every rank runs an UNCHANGED single-range LBVH builder (here the oracle's; on the device the
    64-bit-key instantiation of the tile / group / climb kernels), which makes the next round's work orchestration only:
      * the rank's keys are widened to (key << 32 | GLOBAL sorted position): they are unique, so the builder's own tie-break on LOCAL
        positions never decides, and the common prefix of two neighbours is exactly the augmented-key depth of the one-GPU build;
      * the range is extended by ONE ghost leaf on each inner side (the neighbour's edge key), so every boundary of a ghost-free node —
        including the two at the rank's edges — is seen with its true depth;
      * of the local tree, the nodes that contain no ghost ARE nodes of the one-GPU tree (local index + offset = global index); the
        nodes containing a ghost (the two spines) are artefacts, and the ghost-free children hanging off them are the rank's left-over
        clusters, finished after the gather as in lbvh_by_ranges.
    skeys: uint32 sorted keys (30-bit codes).  builder(keys64, vals) -> (nodes[2m-1], root): the single-range builder under test
    (default: the oracle's; tests pass the device's b2bvh_lbvh_from_sorted64)."""
    n = skeys.size
    nint = n - 1
    nodes = np.zeros(2 * n - 1, dtype=T.BVH2_NODE)
    nodes["left"][:] = 0xFFFFFFFF; nodes["right"][:] = 0xFFFFFFFF
    for g in range(n):
        r = refs[svals[g]]
        nodes[nint + g] = (r["primIdx"], 0xFFFFFFFF, r["mn"], r["mx"])
    leftovers, counts = [], []
    for a, b in ranges:
        a2, b2 = max(a - 1, 0), min(b + 1, n)
        m = b2 - a2
        k64 = (skeys[a2:b2].astype(np.uint64) << np.uint64(32)) | np.arange(a2, b2, dtype=np.uint64)
        v = np.ascontiguousarray(svals[a2:b2])
        if m == 1:
            leftovers.append((a, b, nint + a)); counts.append(1)
            continue
        if builder is not None:
            loc, lroot = builder(k64, v)
        elif karras:
            loc, _ = lbvh_karras(refs, k64, v)
            lroot = 0
        else:
            loc, lroot = lbvh_apetrei(tris, k64, v)
        ghostL, ghostR = a > 0, b < n
        mine = []

        def to_global(idx):
            return (nint + a2 + (idx - (m - 1))) if idx >= m - 1 else a2 + idx

        def walk(idx, lo, hi):
            """returns after classifying the subtree rooted at local node idx covering local leaves [lo, hi)"""
            has_ghost = (ghostL and lo == 0) or (ghostR and hi == m)
            if idx >= m - 1:                      # a leaf
                if not has_ghost:
                    mine.append((a2 + lo, a2 + hi, to_global(idx)))
                return
            l, r = int(loc[idx]["left"]), int(loc[idx]["right"])
            split = _subtree_size(loc, l, m) + lo
            if not has_ghost:
                g = to_global(idx)
                nodes[g]["left"] = to_global(l); nodes[g]["right"] = to_global(r)
                nodes[g]["mn"] = loc[idx]["mn"]; nodes[g]["mx"] = loc[idx]["mx"]
                _emit_subtree(loc, l, m, to_global, nodes); _emit_subtree(loc, r, m, to_global, nodes)
                mine.append((a2 + lo, a2 + hi, g))
                return
            walk(l, lo, split); walk(r, split, hi)

        walk(lroot, 0, m)
        mine.sort()
        counts.append(len(mine))
        leftovers += mine
    # the gather + final rounds, as in lbvh_by_ranges
    depth = [_boundary_depth(skeys, g) for g in range(-1, n)]
    D = lambda g: depth[g + 1]
    root = None
    clusters = leftovers
    while len(clusters) > 1:
        out, i, merged = [], 0, False
        while i < len(clusters):
            if i + 1 < len(clusters):
                (lo, mid, idl), (_, hi, idr) = clusters[i], clusters[i + 1]
                d0, dl, dr = D(mid - 1), D(lo - 1), D(hi - 1)
                if d0 > dl and d0 > dr:
                    is_root = lo == 0 and hi == n
                    nid = (0 if is_root else (hi - 1 if dr > dl else lo)) if karras else mid - 1
                    nodes[nid]["left"] = idl; nodes[nid]["right"] = idr
                    nodes[nid]["mn"] = np.minimum(nodes[idl]["mn"], nodes[idr]["mn"]); nodes[nid]["mx"] = np.maximum(nodes[idl]["mx"], nodes[idr]["mx"])
                    if is_root:
                        root = nid
                    out.append((lo, hi, nid)); i += 2; merged = True
                    continue
            out.append(clusters[i]); i += 1
        assert merged
        clusters = out
    return nodes, (root if root is not None else clusters[0][2]), counts


def _subtree_size(loc, idx, m):
    """number of leaves under local node idx (iterative)"""
    size, stack = 0, [idx]
    while stack:
        i = stack.pop()
        if i >= m - 1:
            size += 1
        else:
            stack.append(int(loc[i]["left"])); stack.append(int(loc[i]["right"]))
    return size


def _emit_subtree(loc, idx, m, to_global, nodes):
    """copies the internal nodes under local node idx into the global array with renumbered children"""
    stack = [idx]
    while stack:
        i = stack.pop()
        if i >= m - 1:
            continue
        l, r = int(loc[i]["left"]), int(loc[i]["right"])
        g = to_global(i)
        nodes[g]["left"] = to_global(l); nodes[g]["right"] = to_global(r)
        nodes[g]["mn"] = loc[i]["mn"]; nodes[g]["mx"] = loc[i]["mx"]
        stack.append(l); stack.append(r)


def stitch_leftovers(nodes, clusters, skeys, karras):
    """The last step of the globally sorted build: the left-over clusters of all ranks, in position order ((lo, hi, node id) with boxes
    already in nodes[node id]), merged by the usual rule — two neighbours form a node when the boundary between them is deeper than both
    boundaries next to it — until one cluster spans everything.  Writes the new nodes into `nodes`; returns the root index."""
    n = skeys.size
    depth = [_boundary_depth(skeys, g) for g in range(-1, n)]
    D = lambda g: depth[g + 1]
    root = None
    clusters = list(clusters)
    while len(clusters) > 1:
        out, i, merged = [], 0, False
        while i < len(clusters):
            if i + 1 < len(clusters):
                (lo, mid, idl), (_, hi, idr) = clusters[i], clusters[i + 1]
                d0, dl, dr = D(mid - 1), D(lo - 1), D(hi - 1)
                if d0 > dl and d0 > dr:
                    is_root = lo == 0 and hi == n
                    nid = (0 if is_root else (hi - 1 if dr > dl else lo)) if karras else mid - 1
                    nodes[nid]["left"] = idl; nodes[nid]["right"] = idr
                    nodes[nid]["mn"] = np.minimum(nodes[idl]["mn"], nodes[idr]["mn"]); nodes[nid]["mx"] = np.maximum(nodes[idl]["mx"], nodes[idr]["mx"])
                    if is_root:
                        root = nid
                    out.append((lo, hi, nid)); i += 2; merged = True
                    continue
            out.append(clusters[i]); i += 1
        assert merged
        clusters = out
    return root if root is not None else clusters[0][2]


def range_tree_reference(tris, refs, k64, vals, karras, ghost_left, ghost_right, first_pos, n_global):
    """CPU twin of capi.Context.range_tree (b2bvh_lbvh_from_sorted64 + b2bvh_range_extract): local tree over the ghost-extended range,
    ghost-free nodes renumbered to global indices, artefacts invalidated, left-over clusters in position order."""
    m = k64.size
    v = np.ascontiguousarray(vals, dtype=np.uint32)
    if karras:
        loc, _ = lbvh_karras(refs, np.ascontiguousarray(k64), v)
        lroot = 0
    else:
        loc, lroot = lbvh_apetrei(tris, np.ascontiguousarray(k64), v)
    out = loc.copy()
    nint = n_global - 1
    art = np.zeros(m - 1, dtype=bool)
    cl = []
    stack = [(lroot, 0, m)]
    while stack:
        idx, lo, hi = stack.pop()
        has_ghost = (ghost_left and lo == 0) or (ghost_right and hi == m)
        leaf = idx >= m - 1
        if not has_ghost:
            cl.append((first_pos + lo, first_pos + hi, (nint + first_pos + idx - (m - 1)) if leaf else first_pos + idx, loc[idx]["mn"].copy(), loc[idx]["mx"].copy()))
            continue
        if leaf:
            continue
        art[idx] = True
        l, r = int(loc[idx]["left"]), int(loc[idx]["right"])
        split = idx + 1 if not karras else ((l - (m - 1)) if l >= m - 1 else l) + 1
        stack.append((r, split, hi)); stack.append((l, lo, split))
    for i in range(m - 1):
        if art[i]:
            out[i]["left"] = 0xFFFFFFFF; out[i]["right"] = 0xFFFFFFFF
        else:
            for f in ("left", "right"):
                c = int(loc[i][f])
                out[i][f] = (nint + first_pos + c - (m - 1)) if c >= m - 1 else first_pos + c
    c = np.zeros(len(cl), dtype=T.CLUSTER)
    for j, x in enumerate(cl):
        c[j] = (x[0], x[1], x[2], x[3], x[4])
    return out, c
