/*
 * collapse.cu — stage S5: Bvh2 -> 4-wide Bvh (Bvh4Node[] + PrimNode[]).
 *
 * Replaces CollapseToWide4Bvh (TwoPassLbvhKernel.h:237-337 for the LBVH layout, Ploc++Kernel.h:364-465 for the
 * separate-leaf layout; host setup TwoPassLbvh.cpp:154-183).  The reference runs ONE persistent launch whose threads
 * spin on a task queue and allocate wide nodes with a global atomicAdd: node numbering is timing dependent and the
 * launch needs every thread resident.  Here the work is split so that the expensive, latency-bound part is fully
 * parallel and only a light numbering pass is level-synchronous:
 *   1. collapse_expand_kernel   for EVERY internal Bvh2 node, the (up to 4) children it would have as a wide node: twice,
 *                               the internal child with the largest area is replaced by its two children (strict '>' so the
 *                               first of equals wins, areas without FMA).  One thread per node, no synchronisation, 16 B out.
 *   2. collapse_number_kernel   one launch per level of the wide tree: the tasks of a level are a contiguous index range;
 *                               the number of internal children is prefix-summed across the level (CTA scan + warp-window
 *                               look-back between ticketed tiles) and the children get consecutive indices in (task, slot)
 *                               order.  That is breadth-first numbering — what a sequential execution of the reference's
 *                               task loop produces — so the output is deterministic and comparable with memcmp.  Per task:
 *                               8 B task + 16 B expansion read, 4 B + 8 B per child written.
 *   3. collapse_emit_kernel     one thread per wide node: child boxes gathered, the 128-byte node written with eight
 *                               16-byte stores, PrimNode records for leaf children.
 */
#include "common.cuh"
#include "lookback.cuh"

#define COL_THREADS 256
#define COL_MAX_LEVELS 4096 /* ranges recorded per level; deeper trees are handled by re-basing (see launcher) */
#define COL_BATCH 24

#define COL_FLAG_AGG 0x40000000u
#define COL_FLAG_INC 0x80000000u
#define COL_VAL_MASK 0x3FFFFFFFu

/* scratch layout: CollapseCtrl (256 B) | uint4 expansion[n] | uint2 tasks[n] | u32 firstChild[n] | u32 status[n/256 + COL_MAX_LEVELS + 2] */
struct CollapseCtrl {
  u32 nWide;
  u32 lastLevelSize;
  u32 pad[2];
  uint2 range[2];  /* range[level & 1] = {start,end} wide-index range of `level` */
  u32 ticket[COL_BATCH];
};

size_t b2_collapse_scratch_bytes(u32 n) {
  return 256 + (size_t)n * (sizeof(uint4) + sizeof(uint2) + sizeof(u32)) + ((size_t)n / COL_THREADS + COL_MAX_LEVELS + 2) * sizeof(u32);
}

template <bool SEPARATE_LEAVES>
__device__ __forceinline__ Box col_child_box(const b2bvh_bvh2_node* __restrict__ nodes, u32 id) { return load_node2_ro(nodes + id).box; }

/* ---- 1. expansion of every internal node (also resets the level control block) ---- */
__global__ void __launch_bounds__(COL_THREADS) collapse_expand_kernel(const b2bvh_bvh2_node* __restrict__ nodes, u32 nInt, uint4* __restrict__ expansion,
                                                                     CollapseCtrl* ctrl, uint2* tasks, const u32* __restrict__ rootIdx) {
  const u32 i = blockIdx.x * COL_THREADS + threadIdx.x;
  if (i == 0) {
    ctrl->nWide = 1; ctrl->lastLevelSize = 1; ctrl->pad[0] = ctrl->pad[1] = 0;
    ctrl->range[0] = make_uint2(0, 1);
    ctrl->range[1] = make_uint2(1, 1);
    for (int k = 0; k < COL_BATCH; k++) ctrl->ticket[k] = 0;
    tasks[0] = make_uint2(*rootIdx, B2_INVALID);
  }
  if (i >= nInt) return;
  const uint2 top = __ldg(reinterpret_cast<const uint2*>(nodes + i));
  u32 ch[4] = {top.x, top.y, B2_INVALID, B2_INVALID};
  /* first expansion: the internal child with the larger area (left wins ties, a leaf never wins) */
  Node2 c0, c1;
  float a0 = 0.0f, a1 = 0.0f;
  if (ch[0] < nInt) { c0 = load_node2_ro(nodes + ch[0]); a0 = box_area(c0.box); }
  if (ch[1] < nInt) { c1 = load_node2_ro(nodes + ch[1]); a1 = box_area(c1.box); }
  float best = 0.0f;
  int pos = -1;
  if (ch[0] < nInt && a0 > best) { best = a0; pos = 0; }
  if (ch[1] < nInt && a1 > best) { best = a1; pos = 1; }
  if (pos >= 0) {
    const Node2 e = pos == 0 ? c0 : c1;
    /* slots now: [pos] = e.left, [2] = e.right, [1 - pos] unchanged */
    float ar[3];
    u32 id3[3];
    id3[pos] = e.left; id3[1 - pos] = ch[1 - pos]; id3[2] = e.right;
    ar[1 - pos] = pos == 0 ? a1 : a0;
    Node2 g0, g1; /* the two new children: their boxes are needed for the second choice, their children for the second expansion */
    ar[pos] = 0.0f; ar[2] = 0.0f;
    if (e.left < nInt) { g0 = load_node2_ro(nodes + e.left); ar[pos] = box_area(g0.box); }
    if (e.right < nInt) { g1 = load_node2_ro(nodes + e.right); ar[2] = box_area(g1.box); }
    ch[0] = id3[0]; ch[1] = id3[1]; ch[2] = id3[2];
    /* second expansion */
    best = 0.0f;
    int pos2 = -1;
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (id3[k] < nInt && ar[k] > best) { best = ar[k]; pos2 = k; }
    if (pos2 >= 0) {
      u32 l2, r2;
      if (pos2 == pos) { l2 = g0.left; r2 = g0.right; }
      else if (pos2 == 2) { l2 = g1.left; r2 = g1.right; }
      else { const Node2 o = pos == 0 ? c1 : c0; l2 = o.left; r2 = o.right; }
#pragma unroll
      for (int k = 0; k < 3; k++) if (k == pos2) ch[k] = l2;
      ch[3] = r2;
    }
  }
  expansion[i] = make_uint4(ch[0], ch[1], ch[2], ch[3]);
}

/* ---- 2. breadth-first numbering, one launch per level ---- */
__global__ void __launch_bounds__(COL_THREADS) collapse_number_kernel(const uint4* __restrict__ expansion, u32 nInt, CollapseCtrl* ctrl, uint2* tasks,
                                                                     u32* __restrict__ firstChild, u32* status, u32 level, u32 launchInBatch) {
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
    uint4 ex = make_uint4(B2_INVALID, B2_INVALID, B2_INVALID, B2_INVALID);
    if (active) ex = __ldg(expansion + tasks[g].x);
    const u32 nInternal = (ex.x < nInt ? 1u : 0u) + (ex.y < nInt ? 1u : 0u) + (ex.z < nInt ? 1u : 0u) + (ex.w < nInt ? 1u : 0u);

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
      firstChild[g] = nextId;
      if (ex.x < nInt) tasks[nextId++] = make_uint2(ex.x, g);
      if (ex.y < nInt) tasks[nextId++] = make_uint2(ex.y, g);
      if (ex.z < nInt) tasks[nextId++] = make_uint2(ex.z, g);
      if (ex.w < nInt) tasks[nextId++] = make_uint2(ex.w, g);
    }
    __syncthreads(); /* sTile / sWarp are reused by the next tile */
  }
}

/* ---- 3. wide nodes + leaf records ---- */
template <bool SEPARATE_LEAVES>
__global__ void __launch_bounds__(COL_THREADS) collapse_emit_kernel(const b2bvh_bvh2_node* __restrict__ nodes, const u32* __restrict__ sortedVals,
                                                                   u32 nInt, const uint4* __restrict__ expansion, const uint2* __restrict__ tasks,
                                                                   const u32* __restrict__ firstChild, const CollapseCtrl* __restrict__ ctrl,
                                                                   b2bvh_bvh4_node* __restrict__ wide, b2bvh_prim_node* __restrict__ wideLeaves) {
  const u32 nWide = ctrl->nWide;
  for (u32 g = blockIdx.x * COL_THREADS + threadIdx.x; g < nWide; g += gridDim.x * COL_THREADS) {
    const uint2 task = __ldg(tasks + g);
    const uint4 ex = __ldg(expansion + task.x);
    const u32 ch[4] = {ex.x, ex.y, ex.z, ex.w};
    u32 nextId = __ldg(firstChild + g);
    u32 outChild[4];
    Box outBox[4];
    u32 cc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      outChild[k] = B2_INVALID;
      outBox[k] = box_empty();
      if (ch[k] != B2_INVALID) {
        cc++;
        if (ch[k] < nInt) {
          outChild[k] = nextId++;
          outBox[k] = load_node2_ro(nodes + ch[k]).box;
        } else {
          outChild[k] = ch[k];
          const u32 slot = ch[k] - nInt;
          /* leaf slot s holds primitive sortedVals[s] in both layouts (Bvh2 leaf m_leftChildIdx / PrimRef m_primIdx): read the
           * dense 4-byte array instead of one 32-byte node or 28-byte PrimRef per leaf */
          reinterpret_cast<uint2*>(wideLeaves)[slot] = make_uint2(__ldg(sortedVals + slot), g);
        }
      }
    }
    /* 128 bytes: 4 boxes (24 floats), 4 children, parent, childCount, 2 zero pad words */
    uint4* out = reinterpret_cast<uint4*>(wide + g);
    const float* f0 = &outBox[0].lx; const float* f1 = &outBox[1].lx; const float* f2 = &outBox[2].lx; const float* f3 = &outBox[3].lx;
    out[0] = make_uint4(__float_as_uint(f0[0]), __float_as_uint(f0[1]), __float_as_uint(f0[2]), __float_as_uint(f0[3]));
    out[1] = make_uint4(__float_as_uint(f0[4]), __float_as_uint(f0[5]), __float_as_uint(f1[0]), __float_as_uint(f1[1]));
    out[2] = make_uint4(__float_as_uint(f1[2]), __float_as_uint(f1[3]), __float_as_uint(f1[4]), __float_as_uint(f1[5]));
    out[3] = make_uint4(__float_as_uint(f2[0]), __float_as_uint(f2[1]), __float_as_uint(f2[2]), __float_as_uint(f2[3]));
    out[4] = make_uint4(__float_as_uint(f2[4]), __float_as_uint(f2[5]), __float_as_uint(f3[0]), __float_as_uint(f3[1]));
    out[5] = make_uint4(__float_as_uint(f3[2]), __float_as_uint(f3[3]), __float_as_uint(f3[4]), __float_as_uint(f3[5]));
    out[6] = make_uint4(outChild[0], outChild[1], outChild[2], outChild[3]);
    out[7] = make_uint4(task.y, cc, 0u, 0u);
  }
}

int b2_launch_collapse(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const b2bvh_prim_ref* d_leaves, const u32* d_sortedVals, const u32* d_rootIdx, u32 n,
                       b2bvh_bvh4_node* d_wide, b2bvh_prim_node* d_wideLeaves, void* d_scratch, u32* h_nWide) {
  if (n < 2) return b2_fail(B2BVH_ERR_INVALID, "collapse needs at least 2 primitives");
  unsigned char* base = reinterpret_cast<unsigned char*>(d_scratch);
  CollapseCtrl* ctrl = reinterpret_cast<CollapseCtrl*>(base);
  uint4* expansion = reinterpret_cast<uint4*>(base + 256);
  uint2* tasks = reinterpret_cast<uint2*>(base + 256 + (size_t)n * sizeof(uint4));
  u32* firstChild = reinterpret_cast<u32*>(base + 256 + (size_t)n * (sizeof(uint4) + sizeof(uint2)));
  u32* status = firstChild + n;
  const size_t statusWords = (size_t)n / COL_THREADS + COL_MAX_LEVELS + 2;
  const u32 nInt = n - 1;
  B2_CUDA(cudaMemsetAsync(status, 0, statusWords * sizeof(u32), ctx->stream));
  B2_KERNEL(ctx, "collapse_expand");
  collapse_expand_kernel<<<(nInt + COL_THREADS - 1) / COL_THREADS, COL_THREADS, 0, ctx->stream>>>(d_nodes, nInt, expansion, ctrl, tasks, d_rootIdx);
  B2_LAUNCH_CHECK(ctx);
  const u32 cap = (u32)ctx->sm_count * 8u;
  u32 levels = 0;
  CollapseCtrl h;
  while (true) {
    for (u32 k = 0; k < COL_BATCH; k++) {
      B2_KERNEL(ctx, "collapse_number");
      collapse_number_kernel<<<cap, COL_THREADS, 0, ctx->stream>>>(expansion, nInt, ctrl, tasks, firstChild, status, levels + k, k);
      B2_LAUNCH_CHECK(ctx);
    }
    levels += COL_BATCH;
    B2_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h.lastLevelSize == 0) break;
    if (levels + COL_BATCH > COL_MAX_LEVELS) return b2_fail(B2BVH_ERR_INTERNAL, "collapse: wide tree deeper than %d levels", COL_MAX_LEVELS);
    B2_CUDA(cudaMemsetAsync(ctrl->ticket, 0, sizeof(u32) * COL_BATCH, ctx->stream));
  }
  u32 grid = (h.nWide + COL_THREADS - 1) / COL_THREADS;
  if (grid > cap * 2) grid = cap * 2;
  B2_KERNEL(ctx, "collapse_emit");
  if (d_leaves)
    collapse_emit_kernel<true><<<grid, COL_THREADS, 0, ctx->stream>>>(d_nodes, d_sortedVals, nInt, expansion, tasks, firstChild, ctrl, d_wide, d_wideLeaves);
  else
    collapse_emit_kernel<false><<<grid, COL_THREADS, 0, ctx->stream>>>(d_nodes, d_sortedVals, nInt, expansion, tasks, firstChild, ctrl, d_wide, d_wideLeaves);
  B2_LAUNCH_CHECK(ctx);
  *h_nWide = h.nWide;
  return 0;
}
