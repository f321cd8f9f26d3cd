/*
 * collapse.cu — stage S5: Bvh2 -> 4-wide Bvh (Bvh4Node[] + PrimNode[]).
 *
 * Replaces CollapseToWide4Bvh (TwoPassLbvhKernel.h:237-337 for the LBVH layout, Ploc++Kernel.h:364-465 for the
 * separate-leaf layout; host setup TwoPassLbvh.cpp:154-183).  The reference runs ONE persistent launch whose threads
 * spin on a task queue and allocate wide nodes with a global atomicAdd: node numbering is timing dependent and the
 * launch needs every thread resident.  Here the collapse is level-synchronous: the tasks of one level of the wide tree
 * are a contiguous index range, every task expands its Bvh2 node (twice: the internal child with the largest area is
 * replaced by its two children; strict '>' so the first of equals wins, areas without FMA), the number of internal
 * children is prefix-summed across the level (CTA scan + decoupled look-back between tiles), and the children receive
 * consecutive indices in (task, slot) order.  That is exactly breadth-first numbering — what a sequential execution of
 * the reference's task loop produces — so the output is deterministic and comparable with memcmp.
 *
 * Per wide node: <= 7 reads of 32-byte Bvh2 nodes, one 128-byte node written with eight 16-byte stores,
 * 8 B per task record, 8 B per PrimNode.
 */
#include "common.cuh"
#include "lookback.cuh"

#define COL_THREADS 256
#define COL_MAX_LEVELS 4096 /* ranges recorded per level; deeper trees are handled by re-basing (see launcher) */
#define COL_BATCH 24

#define COL_FLAG_AGG 0x40000000u
#define COL_FLAG_INC 0x80000000u
#define COL_VAL_MASK 0x3FFFFFFFu

/* scratch layout:
 *   u32 ctrl[8]:  [0] level counter base, [1] nWide so far (== end of the newest level), [2] newest level size
 *   uint2 range[2]          ping-pong {start,end} of the level being processed / produced
 *   u32 ticket[COL_BATCH]   tile tickets, one per launch of a batch
 *   uint2 tasks[n]          {bvh2 node, parent wide index} per wide node
 *   u32 status[n/256 + COL_MAX_LEVELS + 2]  look-back words                               */
struct CollapseCtrl {
  u32 nWide;
  u32 lastLevelSize;
  u32 pad[2];
  uint2 range[2];  /* range[level & 1] = {start,end} wide-index range of `level` */
  u32 ticket[COL_BATCH];
};

size_t b2_collapse_scratch_bytes(u32 n) {
  return 256 + (size_t)n * sizeof(uint2) + ((size_t)n / COL_THREADS + COL_MAX_LEVELS + 2) * sizeof(u32);
}

__global__ void collapse_init_kernel(CollapseCtrl* ctrl, uint2* tasks, const u32* rootIdx) {
  ctrl->nWide = 1; ctrl->lastLevelSize = 1; ctrl->pad[0] = ctrl->pad[1] = 0;
  ctrl->range[0] = make_uint2(0, 1);
  ctrl->range[1] = make_uint2(1, 1);
  for (int k = 0; k < COL_BATCH; k++) ctrl->ticket[k] = 0;
  tasks[0] = make_uint2(*rootIdx, B2_INVALID);
}

template <bool SEPARATE_LEAVES>
__global__ void __launch_bounds__(COL_THREADS) collapse4_level_kernel(const b2bvh_bvh2_node* __restrict__ nodes,
                                                                     const b2bvh_prim_ref* __restrict__ leaves, u32 nInt,
                                                                     b2bvh_bvh4_node* __restrict__ wide, b2bvh_prim_node* __restrict__ wideLeaves,
                                                                     CollapseCtrl* ctrl, uint2* tasks, u32* status, u32 level, u32 launchInBatch) {
  __shared__ u32 sTile, sTileExcl;
  __shared__ u32 sWarp[COL_THREADS / 32];
  /* range[level & 1] was published by the previous launch and is not written during this one */
  const uint2 range = ctrl->range[level & 1u];
  const u32 start = range.x, end = range.y;
  if (start >= end) { /* tree finished: keep every later level empty */
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctrl->range[(level + 1) & 1u] = make_uint2(end, end); ctrl->lastLevelSize = 0; }
    return;
  }
  const u32 nTiles = (end - start + COL_THREADS - 1) / COL_THREADS;
  /* status words of this level: disjoint from every other level's (floor(start/256) + level is strictly increasing
   * by at least the level's tile count) */
  u32* st = status + (start / COL_THREADS) + level;
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;

  while (true) {
    if (tid == 0) sTile = atomicAdd(&ctrl->ticket[launchInBatch], 1u);
    __syncthreads();
    const u32 tile = sTile;
    if (tile >= nTiles) return;
    const u32 g = start + tile * COL_THREADS + tid;
    const bool active = g < end;

    u32 ch[4] = {B2_INVALID, B2_INVALID, B2_INVALID, B2_INVALID};
    Box bx[4];
    u32 cc = 0, parent = B2_INVALID, nInternal = 0;
    if (active) {
      const uint2 task = tasks[g];
      parent = task.y;
      const Node2 n2 = load_node2_ro(nodes + task.x);
      ch[0] = n2.left; ch[1] = n2.right; cc = 2;
      Node2 cn[4];
      if (ch[0] < nInt) { cn[0] = load_node2_ro(nodes + ch[0]); bx[0] = cn[0].box; }
      if (ch[1] < nInt) { cn[1] = load_node2_ro(nodes + ch[1]); bx[1] = cn[1].box; }
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
        float best = 0.0f;
        int pos = -1;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k < (int)cc && ch[k] < nInt) {
            const float a = box_area(bx[k]);
            if (a > best) { best = a; pos = k; }
          }
        if (pos < 0) break;
        /* replace slot `pos` by its left child, append its right child */
        u32 lc = 0, rc = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (k == pos) { lc = cn[k].left; rc = cn[k].right; }
        Node2 ln, rn;
        if (lc < nInt) ln = load_node2_ro(nodes + lc);
        if (rc < nInt) rn = load_node2_ro(nodes + rc);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (k == pos) { ch[k] = lc; cn[k] = ln; bx[k] = ln.box; }
          if (k == (int)cc) { ch[k] = rc; cn[k] = rn; bx[k] = rn.box; }
        }
        cc++;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) nInternal += (k < (int)cc && ch[k] < nInt) ? 1u : 0u;
    }

    /* CTA exclusive scan of nInternal */
    u32 incl = nInternal;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (l == 31) sWarp[w] = incl;
    __syncthreads();
    u32 warpBase = 0, tileTotal = 0;
#pragma unroll
    for (int k = 0; k < COL_THREADS / 32; k++) { const u32 t = sWarp[k]; if (k < (int)w) warpBase += t; tileTotal += t; }
    const u32 localExcl = warpBase + incl - nInternal;

    /* decoupled look-back across the tiles of this level, warp 0 reads 32 predecessors per round trip */
    if (w == 0) {
      if (l == 0) st_relaxed(st + tile, (tile == 0 ? LB_INC : LB_AGG) | tileTotal);
      const u32 excl = warp_lookback_u32(st, tile);
      if (l == 0) {
        if (tile > 0) st_relaxed(st + tile, LB_INC | (excl + tileTotal));
        sTileExcl = excl;
        if (tile == nTiles - 1) {
          /* last tile of the level: publish the next level's range for the next launch */
          const u32 next = end + excl + tileTotal;
          ctrl->range[(level + 1) & 1u] = make_uint2(end, next);
          ctrl->nWide = next;
          ctrl->lastLevelSize = next - end;
        }
      }
    }
    __syncthreads();

    if (active) {
      u32 nextId = end + sTileExcl + localExcl;
      uint4* out = reinterpret_cast<uint4*>(wide + g);
      u32 outChild[4];
      Box outBox[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        outChild[k] = B2_INVALID;
        outBox[k] = box_empty();
        if (k < (int)cc) {
          if (ch[k] < nInt) {
            outChild[k] = nextId;
            outBox[k] = bx[k];
            tasks[nextId] = make_uint2(ch[k], g);
            nextId++;
          } else {
            outChild[k] = ch[k];
            const u32 slot = ch[k] - nInt;
            const u32 prim = SEPARATE_LEAVES ? __ldg(&leaves[slot].m_primIdx) : __ldg(&nodes[ch[k]].m_leftChildIdx);
            reinterpret_cast<uint2*>(wideLeaves)[slot] = make_uint2(prim, g);
          }
        }
      }
      /* 128 bytes: 4 boxes (24 floats), 4 children, parent, childCount, 2 zero pad words */
      const float* f0 = &outBox[0].lx; const float* f1 = &outBox[1].lx; const float* f2 = &outBox[2].lx; const float* f3 = &outBox[3].lx;
      out[0] = make_uint4(__float_as_uint(f0[0]), __float_as_uint(f0[1]), __float_as_uint(f0[2]), __float_as_uint(f0[3]));
      out[1] = make_uint4(__float_as_uint(f0[4]), __float_as_uint(f0[5]), __float_as_uint(f1[0]), __float_as_uint(f1[1]));
      out[2] = make_uint4(__float_as_uint(f1[2]), __float_as_uint(f1[3]), __float_as_uint(f1[4]), __float_as_uint(f1[5]));
      out[3] = make_uint4(__float_as_uint(f2[0]), __float_as_uint(f2[1]), __float_as_uint(f2[2]), __float_as_uint(f2[3]));
      out[4] = make_uint4(__float_as_uint(f2[4]), __float_as_uint(f2[5]), __float_as_uint(f3[0]), __float_as_uint(f3[1]));
      out[5] = make_uint4(__float_as_uint(f3[2]), __float_as_uint(f3[3]), __float_as_uint(f3[4]), __float_as_uint(f3[5]));
      out[6] = make_uint4(outChild[0], outChild[1], outChild[2], outChild[3]);
      out[7] = make_uint4(parent, cc, 0u, 0u);
    }
    __syncthreads(); /* sTile / sWarp are reused by the next tile */
  }
}

int b2_launch_collapse(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const b2bvh_prim_ref* d_leaves, const u32* d_rootIdx, u32 n,
                       b2bvh_bvh4_node* d_wide, b2bvh_prim_node* d_wideLeaves, void* d_scratch, u32* h_nWide) {
  if (n < 2) return b2_fail(B2BVH_ERR_INVALID, "collapse needs at least 2 primitives");
  unsigned char* base = reinterpret_cast<unsigned char*>(d_scratch);
  CollapseCtrl* ctrl = reinterpret_cast<CollapseCtrl*>(base);
  uint2* tasks = reinterpret_cast<uint2*>(base + 256);
  u32* status = reinterpret_cast<u32*>(base + 256 + (size_t)n * sizeof(uint2));
  const size_t statusWords = (size_t)n / COL_THREADS + COL_MAX_LEVELS + 2;
  B2_CUDA(cudaMemsetAsync(status, 0, statusWords * sizeof(u32), ctx->stream));
  B2_KERNEL(ctx, "collapse_init");
  collapse_init_kernel<<<1, 1, 0, ctx->stream>>>(ctrl, tasks, d_rootIdx);
  B2_LAUNCH_CHECK(ctx);
  const u32 grid = (u32)ctx->sm_count * 4u;
  u32 levels = 0;
  CollapseCtrl h;
  while (true) {
    for (u32 k = 0; k < COL_BATCH; k++) {
      B2_KERNEL(ctx, "collapse4_level");
      if (d_leaves)
        collapse4_level_kernel<true><<<grid, COL_THREADS, 0, ctx->stream>>>(d_nodes, d_leaves, n - 1, d_wide, d_wideLeaves, ctrl, tasks, status, levels + k, k);
      else
        collapse4_level_kernel<false><<<grid, COL_THREADS, 0, ctx->stream>>>(d_nodes, d_leaves, n - 1, d_wide, d_wideLeaves, ctrl, tasks, status, levels + k, k);
      B2_LAUNCH_CHECK(ctx);
    }
    levels += COL_BATCH;
    B2_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.lastLevelSize == 0) break;
    if (levels + COL_BATCH > COL_MAX_LEVELS) return b2_fail(B2BVH_ERR_INTERNAL, "collapse: wide tree deeper than %d levels", COL_MAX_LEVELS);
    B2_CUDA(cudaMemsetAsync(ctrl->ticket, 0, sizeof(u32) * COL_BATCH, ctx->stream));
  }
  *h_nWide = h.nWide;
  return 0;
}
