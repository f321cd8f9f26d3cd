"""Primitive-range sharded build across the GPUs of one node (north star; no reference counterpart — the reference is
single-device, Context.cpp:11).

Rank r owns triangles [r*N/G, (r+1)*N/G).  Per build:
  1. local primitive boxes + local scene box                       (b2bvh_shard_extents, on the GPU)
  2. ONE all-reduce(MAX) of 6 floats {-min, max}  -> global scene box, so every shard codes Morton keys in the same frame
  3. local Morton + sort + hierarchy + refit (+ collapse), unchanged single-GPU path with the global box
  4. ONE all-gather of the G sub-tree root boxes (24 B each)
  5. every rank builds the same top-level tree over the G roots     (b2bvh_top_level, on the GPU)
The collectives carry bytes, not bandwidth; they ride on torch.distributed (NCCL over NVLink on the GPUs; gloo in the
CPU tests, where the per-rank engine is a test double).  No other data-path exchange exists: the result is G sub-trees +
a top tree (not node-identical to a single-GPU build of the whole input; parity is per shard and for the top tree).

Primary rays through the sharded tree: the rays are replicated, every rank traces its own sub-tree with the
single-GPU traversal kernel, and the closest hit per ray is ONE all-reduce(MIN) over packed 64-bit words
(bits of t << 32 | global primitive index; t >= 0, so the bit pattern orders like the value; a miss is all ones) followed by
one all-reduce(SUM) in which only the winning rank contributes the barycentrics."""
import numpy as np


def shard_range(n_total, rank, world):
    """[first, last) of rank's primitive range."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


class ShardedBuild:
    """engine: object with
         shard_extents(tris) -> tensor[6] {-min.xyz, max.xyz} on the collective's device
         build(tris, scene_box6) -> (root_box6 as numpy float32[6], tree); an engine with device_scene = True takes the
                                    reduced {-min, max} tensor itself and returns the root box as a tensor on the same device
         top_level(root_boxes tensor[G*6]) -> top-level nodes
         tensor(np_array) -> tensor on the collective's device
       dist: torch.distributed (initialised) or None for world == 1."""

    def __init__(self, engine, dist=None, rank=0, world=1):
        self.engine, self.dist, self.rank, self.world = engine, dist, rank, world

    def build(self, tris):
        import torch
        box6 = self.engine.shard_extents(tris)
        if self.world > 1:
            self.dist.all_reduce(box6, op=self.dist.ReduceOp.MAX)
        if getattr(self.engine, "device_scene", False):
            # the reduced vector stays on the device: no host round trip between the collective and the build
            scene = box6
            mine, tree = self.engine.build(tris, box6)
        else:
            b = box6.detach().cpu().numpy().astype(np.float32)
            scene = np.concatenate([-b[:3], b[3:]]).astype(np.float32)
            root_box, tree = self.engine.build(tris, scene)
            mine = self.engine.tensor(np.asarray(root_box, dtype=np.float32))
        if self.world > 1:
            roots = torch.empty(self.world * 6, dtype=torch.float32, device=mine.device)
            self.dist.all_gather_into_tensor(roots, mine)
        else:
            roots = mine
        top = self.engine.top_level(roots)
        if hasattr(self.engine, "finish"):
            self.engine.finish(tree)  # the build was only enqueued: this is the step's single host synchronisation
        return dict(scene=scene, tree=tree, roots=roots, top=top)

    MISS = (1 << 63) - 1

    def trace(self, built, rays, n_rays, transform, prim_offset, kernel=0):
        """Closest hits of `rays` (replicated on every rank) through the sharded tree.  built: the dict build() returned;
        prim_offset: first global primitive index of this rank's shard.  Returns (t float32[n], prim int64[n] global index or -1,
        uv float32[n,2]) as tensors on the collective's device, identical on every rank."""
        import torch
        key, uv = self.engine.trace(built["tree"], rays, n_rays, transform, prim_offset, kernel)
        best = key.clone()
        if self.world > 1:
            self.dist.all_reduce(best, op=self.dist.ReduceOp.MIN)
        hit = best != self.MISS
        mine = (key == best) & hit
        uv = torch.where(mine.unsqueeze(1), uv, torch.zeros_like(uv))
        if self.world > 1:
            self.dist.all_reduce(uv, op=self.dist.ReduceOp.SUM)
        t = torch.where(hit, (best >> 32).to(torch.int32).view(torch.float32), torch.full_like(best, 0, dtype=torch.float32))
        prim = torch.where(hit, best & 0xFFFFFFFF, torch.full_like(best, -1))
        return t, prim, uv


def pack_hits(torch, prim_u32_as_i32, t_f32, prim_offset):
    """(t bits << 32) | (primIdx + prim_offset), all ones... MISS for primIdx == 0xFFFFFFFF.  t >= 0: bit order == value order."""
    prim = prim_u32_as_i32.to(torch.int64) & 0xFFFFFFFF
    tb = t_f32.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    key = (tb << 32) | (prim + int(prim_offset))
    return torch.where(prim == 0xFFFFFFFF, torch.full_like(key, ShardedBuild.MISS), key)


class GpuEngine:
    """The real engine: one b2bvh Context on this rank's GPU (stream shared with torch so NCCL and the kernels are ordered)."""

    def __init__(self, ctx, algo, collapse=True):
        import torch
        from . import capi, types as T
        self.ctx, self.algo, self.collapse, self.capi, self.T, self.torch = ctx, algo, collapse, capi, T, torch
        self.box6 = torch.zeros(6, dtype=torch.float32, device="cuda")
        self.root6 = torch.zeros(6, dtype=torch.float32, device="cuda")
        self.top_nodes = None
        self.device_scene = True

    def tensor(self, a):
        return self.torch.from_numpy(a).cuda()

    def _ptr_n(self, tris):
        if isinstance(tris, tuple):  # (device pointer, n)
            return tris[0], tris[1], 1
        return tris.ctypes.data, tris.size, 0

    def shard_extents(self, tris):
        p, n, on_dev = self._ptr_n(tris)
        self.capi.check(self.ctx.lib.b2bvh_shard_extents(self.ctx.h, p, n, on_dev, self.box6.data_ptr()), "b2bvh_shard_extents")
        return self.box6

    def build(self, tris, box6):
        """box6: the all-reduced {-min, max} tensor (device).  The primitive boxes computed by shard_extents are reused."""
        p, n, on_dev = self._ptr_n(tris)
        # enqueued only (defer_sync): the root box lands in root6 at the end of the build, stream-ordered, so the caller's all-gather
        # and top-level tree need no host round trip in between; finish() synchronises
        tree = self.ctx.build(self.algo, p, n=n, tris_on_device=bool(on_dev), collapse=self.collapse, boxes_ready=True,
                              d_scene_negmin_max=box6.data_ptr(), defer_sync=True, d_root_box_out=self.root6.data_ptr())
        return self.root6, tree

    def finish(self, tree):
        self.ctx.build_finish(tree)

    def top_level(self, roots):
        g = roots.numel() // 6
        if self.top_nodes is None or self.top_nodes.numel() < (2 * g - 1) * 8:
            self.top_nodes = self.torch.zeros((2 * g - 1) * 8, dtype=self.torch.float32, device="cuda")
        self.capi.check(self.ctx.lib.b2bvh_top_level(self.ctx.h, roots.data_ptr(), g, self.top_nodes.data_ptr()), "b2bvh_top_level")
        return self.top_nodes

    def trace(self, tree, rays, n_rays, transform, prim_offset, kernel=0):
        """rays: device pointer of RAY[n_rays] (b2bvh_generate_rays).  HitInfo lands in a torch buffer on the shared stream."""
        torch = self.torch
        hits = torch.empty((n_rays, 8), dtype=torch.int32, device="cuda")  # HitInfo: primIdx, t, u, v, pad[4] (32 B)
        tr = np.ascontiguousarray(transform)
        ms = self.capi.C.c_float()
        self.capi.check(self.ctx.lib.b2bvh_traverse(self.ctx.h, self.capi.C.byref(tree), self.capi.C.c_void_p(int(rays)), n_rays,
                                                    tr.ctypes.data_as(self.capi.C.c_void_p), int(kernel), self.capi.C.c_void_p(hits.data_ptr()), None,
                                                    self.capi.C.byref(ms)), "b2bvh_traverse")
        key = pack_hits(torch, hits[:, 0], hits[:, 1].view(torch.float32), prim_offset)
        return key, hits[:, 2:4].view(torch.float32).contiguous()


class GlobalBuild:
    """The globally sorted multi-GPU build: G ranks produce the nodes of the ONE-GPU LBVH over all triangles,
    distributed by sorted position.  Per build:
      1. local boxes, ONE all-reduce(MAX) of {-min, max}, Morton codes in the global frame                      (as ShardedBuild)
      2. splitters from an all-gathered sample; every (code, global id, box) goes to the rank whose code interval holds it —
         ONE all-to-all of counts + three all-to-alls of payload; pieces arrive in source-rank order, so equal codes keep index order
      3. local stable sort; global position of the rank's first leaf from an all-gather of counts and edge codes
      4. the ordinary hierarchy stage over [left ghost] + range + [right ghost] with keys (code << 32 | global position), then the
         extraction: ghost-free nodes with global indices + the rank's left-over clusters                        (engine.range_tree)
      5. ONE all-gather of the left-over clusters (<= 256 x 48 bytes per rank); every rank finishes the same top nodes.
    engine: tensor(np) / boxes_and_scene(tris) -> (boxes [n,6] float32, {-min,max} [6]) / morton(boxes, scene6_minmax) -> int64 codes /
            sort(codes int64) -> (sorted codes int64, permutation int64) / range_tree(k64 int64, vals int64, boxes [k,6], karras, ghostL,
            ghostR, first_pos, n_global) -> (nodes uint8 [(2m-1), 32], clusters int32 [c, 12]).  All tensors live on the collective's device."""

    def __init__(self, engine, dist=None, rank=0, world=1, sample=256):
        self.engine, self.dist, self.rank, self.world, self.sample = engine, dist, rank, world, sample

    def _all_gather(self, t):
        import torch
        if self.world == 1:
            return t.unsqueeze(0)
        flat = t.contiguous().reshape(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat)
        return out.reshape((self.world,) + tuple(t.shape))

    def _all_to_all(self, t, send_counts, recv_counts):
        import torch
        if self.world == 1:
            return t
        out = torch.empty((int(sum(recv_counts)),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_to_all_single(out, t.contiguous(), output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts))
        return out

    def build(self, tris, first_global, n_total, karras=False):
        import torch
        eng, W = self.engine, self.world
        # checked on EVERY rank before the first collective: a rank that raised later would leave the others waiting in the next one
        if n_total < 2:
            raise ValueError("GlobalBuild needs at least two primitives in total")
        boxes, box6 = eng.boxes_and_scene(tris)
        if W > 1:
            self.dist.all_reduce(box6, op=self.dist.ReduceOp.MAX)
        scene = torch.cat([-box6[:3], box6[3:]])
        codes = eng.morton(boxes, scene)                                           # int64, < 2^30
        n_local = codes.numel()
        gids = torch.arange(first_global, first_global + n_local, dtype=torch.int64, device=codes.device)
        # 2. splitters from a sample (any non-decreasing choice is correct; a sample balances the ranks)
        S = self.sample
        # integer arithmetic: float32 linspace rounds its end value up to n_local for shards beyond 2^24 primitives (out-of-bounds read)
        pick = (torch.arange(S, dtype=torch.int64, device=codes.device) * max(n_local - 1, 0)) // max(S - 1, 1)
        samp = codes[pick] if n_local else torch.full((S,), (1 << 62), dtype=torch.int64, device=codes.device)
        allsamp = torch.sort(self._all_gather(samp).reshape(-1)).values
        splitters = allsamp[(torch.arange(1, W, device=codes.device) * allsamp.numel()) // W] if W > 1 else allsamp[:0]
        dest = torch.bucketize(codes, splitters, right=True)                        # equal codes never split
        order = torch.sort(dest, stable=True).indices                               # by destination, local order kept
        send_counts = torch.bincount(dest, minlength=W)
        recv_counts = send_counts.clone()
        if W > 1:
            self.dist.all_to_all_single(recv_counts, send_counts)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        r_codes = self._all_to_all(codes[order], sc, rc)
        r_gids = self._all_to_all(gids[order], sc, rc)
        r_boxes = self._all_to_all(boxes[order], sc, rc)
        cnt = r_codes.numel()
        # 3. local stable sort + where the rank sits in the global order
        s_codes, perm = eng.sort(r_codes) if cnt else (r_codes, r_codes)
        edge = torch.tensor([cnt, int(s_codes[0]) if cnt else 0, int(s_codes[-1]) if cnt else 0], dtype=torch.int64, device=codes.device)
        edges = self._all_gather(edge).cpu().tolist()
        a = sum(e[0] for e in edges[:self.rank])
        b = a + cnt
        prev = [e for e in edges[:self.rank] if e[0]]
        nxt = [e for e in edges[self.rank + 1:] if e[0]]
        res = dict(first=a, last=b, n_total=n_total, karras=karras, nodes=None, leaves=None, clusters=None)
        my = torch.zeros((256, 13), dtype=torch.int64, device=codes.device)      # lo, hi, node, box bits x6 (as int64), depthRight, valid, pad
        if cnt:
            ghostL, ghostR = bool(prev), bool(nxt)
            a2 = a - (1 if ghostL else 0)
            parts_k, parts_v = [], []
            if ghostL:
                parts_k.append(torch.tensor([prev[-1][2]], dtype=torch.int64, device=codes.device)); parts_v.append(torch.zeros(1, dtype=torch.int64, device=codes.device))
            parts_k.append(s_codes); parts_v.append(perm)
            if ghostR:
                parts_k.append(torch.tensor([nxt[0][1]], dtype=torch.int64, device=codes.device)); parts_v.append(torch.zeros(1, dtype=torch.int64, device=codes.device))
            kk, vv = torch.cat(parts_k), torch.cat(parts_v)
            m = kk.numel()
            k64 = (kk << 32) | torch.arange(a2, a2 + m, dtype=torch.int64, device=codes.device)
            if m >= 2:
                nodes_u8, cl = eng.range_tree(k64, vv, r_boxes, karras, ghostL, ghostR, a2, n_total)
                nodes_i32 = nodes_u8.view(torch.int32).reshape(2 * m - 1, 8)
                # leaves name the GLOBAL primitive (the local builder wrote the index into the received arrays)
                own = nodes_i32[m - 1 + (a - a2):m - 1 + (a - a2) + cnt]
                own[:, 0] = r_gids[own[:, 0].long()].to(torch.int32)
                res.update(nodes=nodes_i32[:m - 1], node_first=a2, leaves=own)
                c = cl.shape[0]
                cl64 = cl.to(torch.int64) & 0xFFFFFFFF
                hi_local = (cl64[:, 1] - a2)
                kx = torch.where(hi_local < m, k64[(hi_local - 1).clamp(0, m - 1)] ^ k64[hi_local.clamp(0, m - 1)], torch.zeros_like(hi_local))
                depth = torch.tensor([(64 - int(x).bit_length()) if (h < m) else -1 for x, h in zip(kx.cpu().tolist(), hi_local.cpu().tolist())], dtype=torch.int64,
                                     device=codes.device)
                my[:c, 0:3] = cl64[:, 0:3]; my[:c, 3:9] = cl64[:, 4:10]; my[:c, 9] = depth; my[:c, 10] = 1
            # m < 2 cannot happen here: n_total >= 2 was checked up front, so a rank that holds a leaf also has a neighbour (a ghost)
        # 5. gather the left-overs; every rank finishes the same top of the tree
        allc = self._all_gather(my).reshape(-1, 13).cpu().numpy()
        allc = allc[allc[:, 10] == 1]
        res["top"], res["root"] = _finish_top(allc, n_total, karras)
        return res


def _finish_top(clusters, n_total, karras):
    """clusters: rows (lo, hi, node, box bits x6, depth of the boundary right of the cluster, valid).  The usual rule — two neighbours form a
    node when the boundary between them is deeper than both boundaries next to it.  Returns ({global node index: (left, right, box6 float32)}, root)."""
    cur = [(int(r[0]), int(r[1]), int(r[2]), np.array(r[3:9], dtype=np.uint32).view(np.float32).copy(), int(r[9])) for r in clusters]
    top = {}
    root = cur[0][2] if len(cur) == 1 else None
    while len(cur) > 1:
        out, i, merged = [], 0, False
        while i < len(cur):
            if i + 1 < len(cur):
                (lo, mid, idl, bl, d0), (_, hi, idr, br, dr) = cur[i], cur[i + 1]
                dl = cur[i - 1][4] if i > 0 else -1
                # dl: depth of the boundary left of cluster i == the right depth of its left neighbour IN THE CURRENT LIST
                if out and out[-1][1] == lo:
                    dl = out[-1][4]
                if d0 > dl and d0 > dr:
                    is_root = lo == 0 and hi == n_total
                    nid = (0 if is_root else (hi - 1 if dr > dl else lo)) if karras else mid - 1
                    box = np.concatenate([np.minimum(bl[:3], br[:3]), np.maximum(bl[3:], br[3:])]).astype(np.float32)
                    top[nid] = (idl, idr, box)
                    if is_root:
                        root = nid
                    out.append((lo, hi, nid, box, dr)); i += 2; merged = True
                    continue
            out.append(cur[i]); i += 1
        if not merged:
            raise RuntimeError("left-over clusters do not merge: inconsistent boundary depths")
        cur = out
    return top, root


class GpuGlobalEngine:
    """The device engine of GlobalBuild: every step is a call into libb2bvh.so on torch-owned buffers (the context shares torch's stream,
    so the NCCL collectives and the kernels are ordered)."""

    def __init__(self, ctx):
        import torch
        from . import capi
        self.ctx, self.capi, self.torch, self.C = ctx, capi, torch, capi.C

    def _p(self, t):
        return self.C.c_void_p(t.data_ptr())

    def boxes_and_scene(self, tris):
        """tris: (device pointer, n) or a TRIANGLE numpy array."""
        torch = self.torch
        if isinstance(tris, tuple):
            d, n = tris
        else:
            n = tris.size
            self._tris = torch.from_numpy(tris.view(np.uint8).reshape(n, 64)).cuda()
            d = self._tris.data_ptr()
        boxes = torch.empty((n, 6), dtype=torch.float32, device="cuda")
        scene = torch.empty(6, dtype=torch.float32, device="cuda")
        self.capi.check(self.ctx.lib.b2bvh_scene_extents(self.ctx.h, self.C.c_void_p(int(d)), n, self._p(boxes), self._p(scene)), "b2bvh_scene_extents")
        return boxes, torch.cat([-scene[:3], scene[3:]])

    def morton(self, boxes, scene):
        torch = self.torch
        n = boxes.shape[0]
        keys = torch.empty(n, dtype=torch.int32, device="cuda")
        vals = torch.empty(n, dtype=torch.int32, device="cuda")
        sc = scene.contiguous()
        self.capi.check(self.ctx.lib.b2bvh_morton_codes(self.ctx.h, self._p(boxes), self._p(sc), n, self._p(keys), self._p(vals)), "b2bvh_morton_codes")
        return keys.to(torch.int64)

    def sort(self, codes):
        torch = self.torch
        n = codes.numel()
        kin = codes.to(torch.int32).contiguous()
        kout = torch.empty(n, dtype=torch.int32, device="cuda")
        vout = torch.empty(n, dtype=torch.int32, device="cuda")
        self.capi.check(self.ctx.lib.b2bvh_sort_pairs(self.ctx.h, self._p(kin), None, self._p(kout), self._p(vout), n, 0, 32), "b2bvh_sort_pairs")
        return kout.to(torch.int64), vout.to(torch.int64)

    def range_tree(self, k64, vals, boxes, karras, ghost_left, ghost_right, first_pos, n_global):
        torch, C = self.torch, self.C
        m = k64.numel()
        k = k64.contiguous()
        v = vals.to(torch.int32).contiguous()
        bx = boxes.contiguous()
        nodes = torch.empty((2 * m - 1, 32), dtype=torch.uint8, device="cuda")
        parents = torch.empty(2 * m - 1, dtype=torch.int32, device="cuda")
        out = torch.empty((2 * m - 1, 32), dtype=torch.uint8, device="cuda")
        clusters = torch.zeros((256, 12), dtype=torch.int32, device="cuda")
        root, cnt = C.c_uint32(), C.c_uint32()
        self.capi.check(self.ctx.lib.b2bvh_lbvh_from_sorted64(self.ctx.h, self._p(k), self._p(v), self._p(bx), m, 1 if karras else 0, self._p(nodes), self._p(parents),
                                                              C.byref(root)), "b2bvh_lbvh_from_sorted64")
        self.capi.check(self.ctx.lib.b2bvh_range_extract(self.ctx.h, self._p(nodes), m, root.value, 1 if karras else 0, 1 if ghost_left else 0,
                                                         1 if ghost_right else 0, int(first_pos), int(n_global), self._p(out), self._p(clusters), C.byref(cnt)),
                        "b2bvh_range_extract")
        return out, clusters[:cnt.value]


class GlobalBuildDevice:
    """The globally sorted multi-GPU build with every data-path step on the device behind the C ABI (b2bvh_global_*, csrc/global_build.cu):
    G ranks produce the nodes of the ONE-GPU LBVH over all triangles, distributed by sorted position.  What is left to
    this class is the sequence and the collectives between the steps:
      boxes + local scene box  -> all-reduce(MAX) of {-min,max}  -> Morton codes in the global frame
      256 sampled codes        -> all-gather, sorted             -> G-1 splitters                                        (2 KB: control plane)
      b2bvh_global_partition   -> all-gather of the G send counts -> the G x G count matrix goes to the host (the build's ONE synchronisation:
                                  NCCL takes send / receive sizes from the host) -> three all-to-alls: code, global id, 24-byte box = 32 B/primitive;
                                  the local sort starts as soon as the codes have landed, ids and boxes are still on the wire then
      b2bvh_global_sort        -> all-gather of the ranks' edge codes (8 B each)
      b2bvh_global_tree        -> all-gather of the left-over clusters (<= 256 x 48 B per rank) -> b2bvh_global_top on every rank
    Results stay on the device (torch tensors); nothing is read back here.
    dist: torch.distributed (or an object with all_reduce / all_gather_into_tensor / all_to_all_single and ReduceOp), None for world == 1."""

    def __init__(self, ctx, dist=None, rank=0, world=1, sample=256):
        import torch
        from . import capi
        self.ctx, self.dist, self.rank, self.world, self.sample = ctx, dist, rank, world, sample
        self.torch, self.capi, self.C = torch, capi, capi.C
        self.timing = {}

    def _p(self, t):
        return self.C.c_void_p(t.data_ptr())

    def _all_gather(self, t):
        if self.world == 1:
            return t.reshape(1, -1)
        flat = t.contiguous().reshape(-1)
        out = self.torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat)
        return out.reshape(self.world, -1)

    def build(self, tris, first_global, n_total, karras=False):
        """tris: (device pointer, n) or a TRIANGLE numpy array (this rank's primitives, global indices first_global + i)."""
        torch, capi, C, W = self.torch, self.capi, self.C, self.world
        lib, h = self.ctx.lib, self.ctx.h
        if n_total < 2:  # on every rank, before the first collective
            raise ValueError("GlobalBuildDevice needs at least two primitives in total")
        dev = torch.device("cuda", torch.cuda.current_device())
        i32 = dict(dtype=torch.int32, device=dev)
        if isinstance(tris, tuple):
            d, n = tris
        else:
            n = tris.size
            self._tris = torch.from_numpy(tris.view(np.uint8).reshape(n, 64)).to(dev)
            d = self._tris.data_ptr()
        # ---- boxes, global scene box, codes ----
        boxes = torch.empty((max(n, 1), 6), dtype=torch.float32, device=dev)
        scene = torch.empty(6, dtype=torch.float32, device=dev)
        if n:
            capi.check(lib.b2bvh_scene_extents(h, C.c_void_p(int(d)), n, self._p(boxes), self._p(scene)), "b2bvh_scene_extents")
            box6 = torch.cat([-scene[:3], scene[3:]])
        else:
            box6 = torch.full((6,), -3.0e38, dtype=torch.float32, device=dev)
        if W > 1:
            self.dist.all_reduce(box6, op=self.dist.ReduceOp.MAX)
        scene = torch.cat([-box6[:3], box6[3:]]).contiguous()
        codes = torch.empty(max(n, 1), **i32)
        iota = torch.empty(max(n, 1), **i32)
        if n:
            capi.check(lib.b2bvh_morton_codes(h, self._p(boxes), self._p(scene), n, self._p(codes), self._p(iota)), "b2bvh_morton_codes")
        # ---- splitters from a sample (any non-decreasing choice is correct; a sample balances the ranks) ----
        S = self.sample
        if n:
            pick = (torch.arange(S, dtype=torch.int64, device=dev) * (n - 1)) // max(S - 1, 1)
            samp = codes[pick].to(torch.int64) & 0xFFFFFFFF
        else:
            samp = torch.full((S,), 1 << 62, dtype=torch.int64, device=dev)
        allsamp = torch.sort(self._all_gather(samp).reshape(-1)).values
        if W > 1:
            spl = allsamp[(torch.arange(1, W, device=dev) * allsamp.numel()) // W].clamp(max=0xFFFFFFFF)
            splitters = torch.where(spl >= (1 << 31), spl - (1 << 32), spl).to(torch.int32).contiguous()  # u32 bit patterns in an int32 tensor
        else:
            splitters = torch.zeros(1, **i32)
        # ---- destination, stable partition, payload in send order ----
        s_codes, s_gids = torch.empty(max(n, 1), **i32), torch.empty(max(n, 1), **i32)
        s_boxes = torch.empty((max(n, 1), 6), dtype=torch.float32, device=dev)
        s_counts = torch.zeros(W, **i32)
        capi.check(lib.b2bvh_global_partition(h, self._p(codes), self._p(boxes), n, int(first_global), self._p(splitters), W, self._p(s_codes), self._p(s_gids),
                                              self._p(s_boxes), self._p(s_counts)), "b2bvh_global_partition")
        M = self._all_gather(s_counts).cpu().tolist()  # the build's one host synchronisation: M[src][dst]
        sc = M[self.rank]
        rc = [M[src][self.rank] for src in range(W)]
        cnts = [sum(M[src][r] for src in range(W)) for r in range(W)]
        cnt, a = cnts[self.rank], sum(cnts[:self.rank])
        prev = max((r for r in range(self.rank) if cnts[r]), default=-1)
        nxt = min((r for r in range(self.rank + 1, W) if cnts[r]), default=-1)
        # ---- the exchange: the sort starts when the codes are here ----
        r_codes, r_gids = torch.empty(max(cnt, 1), **i32), torch.empty(max(cnt, 1), **i32)
        r_boxes = torch.empty((max(cnt, 1), 6), dtype=torch.float32, device=dev)
        works = []
        if W > 1:
            for out, src in ((r_codes, s_codes), (r_gids, s_gids), (r_boxes, s_boxes)):
                works.append(self.dist.all_to_all_single(out[:cnt], src[:n], output_split_sizes=rc, input_split_sizes=sc, async_op=True))
            works[0].wait()
        else:
            r_codes, r_gids, r_boxes = s_codes, s_gids, s_boxes
        sorted_codes, perm = torch.empty(max(cnt, 1), **i32), torch.empty(max(cnt, 1), **i32)
        edge2 = torch.zeros(2, **i32)
        capi.check(lib.b2bvh_global_sort(h, self._p(r_codes), cnt, self._p(sorted_codes), self._p(perm), self._p(edge2)), "b2bvh_global_sort")
        all_edges = self._all_gather(edge2).reshape(-1).contiguous()
        for wk in works[1:]:
            wk.wait()
        # ---- the rank's part of the tree ----
        gl, gr = (1 if prev >= 0 else 0), (1 if nxt >= 0 else 0)
        m = cnt + gl + gr if cnt else 0
        nodes = torch.empty((max(2 * m - 1, 1), 8), **i32)
        clusters = torch.zeros(256 * 12, **i32)
        ccount = torch.zeros(1, **i32)
        capi.check(lib.b2bvh_global_tree(h, self._p(sorted_codes), self._p(perm), self._p(r_gids), self._p(r_boxes), cnt, self._p(all_edges), prev, nxt, a, int(n_total),
                                         1 if karras else 0, self._p(nodes), self._p(clusters), self._p(ccount)), "b2bvh_global_tree")
        # ---- the nodes above the ranks ----
        all_clusters = self._all_gather(clusters).reshape(-1).contiguous()
        all_counts = self._all_gather(ccount).reshape(-1).contiguous()
        top = torch.zeros(W * 256 * 12, **i32)
        res3 = torch.zeros(3, **i32)
        capi.check(lib.b2bvh_global_top(h, self._p(all_clusters), self._p(all_counts), W, int(n_total), 1 if karras else 0, self._p(top), self._p(res3)), "b2bvh_global_top")
        self._keep = (boxes, codes, iota, s_codes, s_gids, s_boxes, r_codes, r_gids, r_boxes, sorted_codes, perm, all_edges, all_clusters, all_counts)
        a2 = a - gl
        return dict(first=a, last=a + cnt, n_total=n_total, karras=karras, m=m, node_first=a2, nodes=nodes[:max(m - 1, 0)] if cnt else nodes[:0],
                    leaves=nodes[m - 1 + gl:m - 1 + gl + cnt] if cnt else nodes[:0], top=top.reshape(-1, 12), top_result=res3, sorted_codes=sorted_codes[:cnt],
                    wire_bytes_sent=32 * (n - sc[self.rank]), counts=cnts)


def _collapse_replicated(self, res):
    """The 4-wide tree of the globally sorted build, REPLICATED: the breadth-first wide-node numbering is a property of the whole tree, so every
    rank all-gathers the pieces (64 B per primitive over NVLink, padded to the longest piece), assembles the one-GPU node array on its device
    (b2bvh_global_assemble) and runs the ordinary collapse over it (b2bvh_collapse_bvh2).  Every rank ends up with the Bvh4 ONE GPU would build
    over all triangles — byte-identical nodes, leaves and count.  Returns dict(full_nodes, wide, wide_leaves, n_wide)."""
    torch, capi, C, W = self.torch, self.capi, self.C, self.world
    lib, h = self.ctx.lib, self.ctx.h
    n_total, cnts = res["n_total"], res["counts"]
    dev = res["top"].device
    firsts = [sum(cnts[:r]) for r in range(W)]
    prevs = [max((q for q in range(r) if cnts[q]), default=-1) for r in range(W)]
    nexts = [min((q for q in range(r + 1, W) if cnts[q]), default=-1) for r in range(W)]
    ms = [(cnts[r] + (1 if prevs[r] >= 0 else 0) + (1 if nexts[r] >= 0 else 0)) if cnts[r] else 0 for r in range(W)]
    node_counts = [max(m - 1, 0) for m in ms]
    node_firsts = [firsts[r] - (1 if prevs[r] >= 0 else 0) for r in range(W)]
    max_nodes, max_leaves = max(max(node_counts), 1), max(max(cnts), 1)
    pn = torch.zeros((max_nodes, 8), dtype=torch.int32, device=dev)
    pl = torch.zeros((max_leaves, 8), dtype=torch.int32, device=dev)
    pn[:res["nodes"].shape[0]] = res["nodes"]
    pl[:res["leaves"].shape[0]] = res["leaves"]
    if W > 1:
        all_n = torch.empty((W * max_nodes, 8), dtype=torch.int32, device=dev)
        all_l = torch.empty((W * max_leaves, 8), dtype=torch.int32, device=dev)
        self.dist.all_gather_into_tensor(all_n, pn)
        self.dist.all_gather_into_tensor(all_l, pl)
    else:
        all_n, all_l = pn, pl
    layout = np.array([[node_firsts[r], node_counts[r], firsts[r], cnts[r]] for r in range(W)], dtype=np.uint32)
    full = torch.empty((2 * n_total - 1, 8), dtype=torch.int32, device=dev)
    leaf_prim = torch.empty(n_total, dtype=torch.int32, device=dev)
    top = res["top"].contiguous()
    capi.check(lib.b2bvh_global_assemble(h, self._p(all_n), self._p(all_l), W, max_nodes, max_leaves, layout.ctypes.data_as(C.c_void_p), self._p(top), self._p(res["top_result"]),
                                         int(n_total), self._p(full), self._p(leaf_prim)), "b2bvh_global_assemble")
    wide = torch.empty((n_total, 32), dtype=torch.int32, device=dev)       # Bvh4Node: 128 B
    wide_leaves = torch.empty((n_total, 2), dtype=torch.int32, device=dev)  # PrimNode: 8 B
    n_wide = C.c_uint32()
    d_root = self.C.c_void_p(res["top_result"].data_ptr() + 4)  # result3[1] = root node index, on the device
    capi.check(lib.b2bvh_collapse_bvh2(h, self._p(full), self._p(leaf_prim), d_root, int(n_total), self._p(wide), self._p(wide_leaves), C.byref(n_wide)), "b2bvh_collapse_bvh2")
    self._keep2 = (all_n, all_l, pn, pl, top)
    return dict(full_nodes=full, wide=wide[:n_wide.value], wide_leaves=wide_leaves, n_wide=int(n_wide.value), wire_bytes_sent=(W - 1) * 32 * (max_nodes + max_leaves))


GlobalBuildDevice.collapse_replicated = _collapse_replicated


def assemble_global_tree(parts, n_total):
    """Host-side check helper: the full node array (2n-1, LBVH layout) of the one-GPU tree from every rank's GlobalBuildDevice result
    (nodes / leaves / top as numpy int32 arrays).  Returns (nodes int32 [(2n-1), 8], root)."""
    out = np.zeros((2 * n_total - 1, 8), dtype=np.int32)
    written = np.zeros(2 * n_total - 1, dtype=bool)
    root = None
    for p in parts:
        nd = p["nodes"]
        if nd.shape[0]:
            keep = nd[:, 0] != -1  # artefacts are {INVALID, INVALID, empty}
            idx = p["node_first"] + np.nonzero(keep)[0]
            out[idx] = nd[keep]
            written[idx] = True
        lv = p["leaves"]
        out[n_total - 1 + p["first"]:n_total - 1 + p["first"] + lv.shape[0]] = lv
        written[n_total - 1 + p["first"]:n_total - 1 + p["first"] + lv.shape[0]] = True
    p = parts[0]
    ntop, root, status = [int(x) & 0xFFFFFFFF for x in p["top_result"]]
    if status != 0:
        raise RuntimeError(f"b2bvh_global_top reported status {status}")
    for t in p["top"][:ntop]:
        i = int(t[0])
        out[i, 0], out[i, 1] = t[1], t[2]
        out[i, 2:8] = t[4:10]
        written[i] = True
    return out, root, written


def check_against_one_tree(ctx, res, d_all, n, karras, want_hash=None):
    """Check helper (tools/global_build_check.py, bench.py).  res: GlobalBuildDevice.build result of this rank; d_all: ALL triangles on this GPU.  True when this rank's ghost-free nodes, its leaves
    and the nodes above the ranks equal the corresponding nodes of the tree one context builds over all triangles."""
    from . import capi, types as T
    whole = ctx.build(capi.TWO_PASS_LBVH if karras else capi.SINGLE_PASS_LBVH, d_all, n=n, tris_on_device=True, collapse=False)
    want = ctx.download(whole.d_bvhNodes, T.BVH2_NODE, 2 * n - 1).view(np.int32).reshape(-1, 8)
    ntop, root, status = [int(x) & 0xFFFFFFFF for x in res["top_result"].cpu().numpy()]
    ok = status == 0 and root == whole.root
    mine = res["nodes"].cpu().numpy()
    valid = mine[:, 0] != -1
    idx = res["node_first"] + np.nonzero(valid)[0]
    ok = ok and np.array_equal(mine[valid], want[idx])
    ok = ok and np.array_equal(res["leaves"].cpu().numpy(), want[n - 1 + res["first"]:n - 1 + res["last"]])
    top = res["top"].cpu().numpy()[:ntop]
    ok = ok and np.array_equal(top[:, 1:3], want[top[:, 0], 0:2]) and np.array_equal(top[:, 4:10], want[top[:, 0], 2:8])
    # want_hash: callable(node array) -> value, e.g. the FNV-1a the frozen oracle answers use: ties the one-GPU tree itself to the oracle
    return bool(ok), int(valid.sum()), ntop, (want_hash(want) if want_hash else None)


