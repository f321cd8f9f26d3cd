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
