/*
 * batched.cu — the batched builder: thousands of tiny BVHs (<= 32 triangles each, MaxBatchedBlockSize, Common.h:597) in one
 * launch.  Replaces BatchedBvhBuilder::build (BatchedBuilder.cpp:16-77) + BatchedBuildKernelLbvh (BatchedBuildKernel.h:218-312).
 *
 * The reference gives every item a 32-thread block: block reduction of the item's box through shared memory, plain 10/10/10
 * Morton codes, 32 one-bit split passes with a block prefix sum and two barriers each (a stable LSD sort), then the Apetrei
 * build with atomic counters in shared memory and __threadfence().  Here an item is ONE WARP and nothing leaves its registers
 * but the results:
 *   box + scene box   one triangle per lane (2 x 16-byte + 1 x 4-byte load), six redux.sync reductions on the integer image of the floats
 *   sort              rank of (code, lane) against the 32 codes read back from shared memory — the rank IS the stable sorted position
 *   hierarchy         the sorted leaves are an ordered list of clusters, one per lane.  A cluster [lo, hi) joins the node at the
 *                     deeper of its two boundaries (findParent, BatchedBuildKernel.h:136-159: smaller XOR of the augmented
 *                     keys = longer common prefix); two neighbours that pick the boundary between them become node hi-1 of
 *                     the left one — the index the reference's walk gives it.  All such pairs merge in the same round; the list
 *                     is a mask of live lanes and neighbour decisions are bit tests on ballots, so a round costs 7 shuffles
 *                     (the absorbed neighbour's packed range/index word and box); no atomics, no fences.
 * The shuffle pipe bounds this kernel: the first version (butterfly reductions, rank by 32 shuffles, register compaction after every
 * round: ~260 shuffles per item) ran 0.46 ms for 312 500 items with mio_throttle as the top stall.
 * Eight items per CTA; nodes leave as two 16-byte stores, leaf records (28 bytes) through a shared-memory transpose as
 * consecutive words.  Algorithmic traffic per triangle: 64 B read + 28 B leaf + 32 B node written.
 *
 * Repairs of the reference kernel (which is unfinished: ExtentCacheSize is undefined, main.cpp:38-52 keeps it disabled),
 * restated by the CPU oracle (orc_batched_build): leaves are stored in SORTED order (the reference never permutes them,
 * :241-242); item offsets are prefix sums of the item sizes (the reference multiplies by the item's own size, :234-235); an
 * item of one triangle has no internal node and root 0.  Node and leaf indices are local to the item, as in the reference.
 */
#include "common.cuh"

#ifndef BATCH_WARPS
#define BATCH_WARPS 8
#endif
#define BATCH_MAX 32u

__device__ __forceinline__ u32 morton3d_10(u32 x) { /* BatchedBuildKernel.h:89-96 */
  x = (x * 0x00010001u) & 0xFF0000FFu;
  x = (x * 0x00000101u) & 0x0F00F00Fu;
  x = (x * 0x00000011u) & 0xC30C30C3u;
  x = (x * 0x00000005u) & 0x49249249u;
  return x;
}

struct BatchWarpSmem {
  u64 xb[BATCH_MAX + 1];      /* xb[b] = augmented-key XOR across boundary b (between sorted leaves b-1 and b); ~0 outside */
  __align__(16) u32 raw[BATCH_MAX]; /* codes in input order (rank computation) */
  u32 val[BATCH_MAX];         /* sorted position -> lane (= primitive) */
  u32 key[BATCH_MAX];
  u32 leaf[BATCH_MAX * 7];    /* PrimRef records of the item, staged for consecutive-word stores */
};

__global__ void __launch_bounds__(BATCH_WARPS * 32) batched_lbvh_kernel(const b2bvh_triangle* __restrict__ tris, const u32* __restrict__ leafOff,
                                                                        const u32* __restrict__ nodeOff, u32 nItems, b2bvh_bvh2_node* nodes,
                                                                        b2bvh_prim_ref* leaves, u32* roots, b2bvh_aabb* scenes) {
  __shared__ BatchWarpSmem smem[BATCH_WARPS];
  const u32 lane = lane_id(), warp = threadIdx.x >> 5;
  const u32 item = blockIdx.x * BATCH_WARPS + warp;
  if (item >= nItems) return; /* whole warps leave; nothing below synchronises across warps */
  BatchWarpSmem& S = smem[warp];
  const u32 first = __ldg(leafOff + item);
  const u32 n = __ldg(leafOff + item + 1) - first;
  const u32 nOff = __ldg(nodeOff + item);

  /* ---- primitive boxes + the item's scene box (BatchedBuildKernel.h:237-259) ---- */
  Box b = box_empty();
  if (lane < n) {
    const float4* p = reinterpret_cast<const float4*>(tris + first + lane);
    const float4 a = __ldg(p), c = __ldg(p + 1);
    const float v3z = __ldg(reinterpret_cast<const float*>(p + 2));
    b.lx = fminf(fminf(fminf(b.lx, a.x), a.w), c.z); b.hx = fmaxf(fmaxf(fmaxf(b.hx, a.x), a.w), c.z);
    b.ly = fminf(fminf(fminf(b.ly, a.y), c.x), c.w); b.hy = fmaxf(fmaxf(fmaxf(b.hy, a.y), c.x), c.w);
    b.lz = fminf(fminf(fminf(b.lz, a.z), c.y), v3z); b.hz = fmaxf(fmaxf(fmaxf(b.hz, a.z), c.y), v3z);
  }
  /* six warp reductions in the integer image of the floats (redux.sync; unsigned order == float order, -0 < +0 as fminf/fmaxf
   * order them; the boxes hold no NaN: fminf/fmaxf drop NaN vertices): 6 instructions instead of 30 shuffles — the kernel is bound
   * by the shuffle pipe (profiles/r01s: mio_throttle on top with 260 shuffles per item) */
  Box sc;
  sc.lx = ordered_to_float(__reduce_min_sync(B2_FULL, float_to_ordered(b.lx))); sc.ly = ordered_to_float(__reduce_min_sync(B2_FULL, float_to_ordered(b.ly)));
  sc.lz = ordered_to_float(__reduce_min_sync(B2_FULL, float_to_ordered(b.lz))); sc.hx = ordered_to_float(__reduce_max_sync(B2_FULL, float_to_ordered(b.hx)));
  sc.hy = ordered_to_float(__reduce_max_sync(B2_FULL, float_to_ordered(b.hy))); sc.hz = ordered_to_float(__reduce_max_sync(B2_FULL, float_to_ordered(b.hz)));
  if (lane == 0) store_aabb(scenes + item, sc);

  /* ---- plain Morton code of the normalised centroid (:272-282, computeMortonCode :98-110); idle lanes sort last ---- */
  u32 key = B2_INVALID;
  if (lane < n) {
    const float cx = __fmul_rn(__fadd_rn(b.hx, b.lx), 0.5f), cy = __fmul_rn(__fadd_rn(b.hy, b.ly), 0.5f), cz = __fmul_rn(__fadd_rn(b.hz, b.lz), 0.5f);
    const float px = __fdiv_rn(__fsub_rn(cx, sc.lx), __fsub_rn(sc.hx, sc.lx));
    const float py = __fdiv_rn(__fsub_rn(cy, sc.ly), __fsub_rn(sc.hy, sc.ly));
    const float pz = __fdiv_rn(__fsub_rn(cz, sc.lz), __fsub_rn(sc.hz, sc.lz));
    const float x = fminf(fmaxf(__fmul_rn(px, 1024.0f), 0.0f), 1023.0f); /* 0/0 = NaN -> 0 through fmaxf */
    const float y = fminf(fmaxf(__fmul_rn(py, 1024.0f), 0.0f), 1023.0f);
    const float z = fminf(fmaxf(__fmul_rn(pz, 1024.0f), 0.0f), 1023.0f);
    key = morton3d_10((u32)x) * 4u + morton3d_10((u32)y) * 2u + morton3d_10((u32)z);
  }

  /* ---- stable sort (:285-297): position = number of (code, lane) pairs that order before mine ---- */
  S.raw[lane] = key;
  __syncwarp();
  u32 rank = 0;
#pragma unroll
  for (u32 j = 0; j < 32; j += 4) { /* eight 16-byte broadcast reads instead of 32 shuffles */
    const uint4 k4 = *reinterpret_cast<const uint4*>(S.raw + j);
    rank += (k4.x < key || (k4.x == key && j + 0 < lane)) ? 1u : 0u;
    rank += (k4.y < key || (k4.y == key && j + 1 < lane)) ? 1u : 0u;
    rank += (k4.z < key || (k4.z == key && j + 2 < lane)) ? 1u : 0u;
    rank += (k4.w < key || (k4.w == key && j + 3 < lane)) ? 1u : 0u;
  }
  S.key[rank] = key;
  S.val[rank] = lane;
  __syncwarp();
  const u32 sk = S.key[lane], sv = S.val[lane]; /* lanes >= n keep themselves (all-ones keys stay in lane order) */
  Box lb;
  lb.lx = __shfl_sync(B2_FULL, b.lx, sv); lb.ly = __shfl_sync(B2_FULL, b.ly, sv); lb.lz = __shfl_sync(B2_FULL, b.lz, sv);
  lb.hx = __shfl_sync(B2_FULL, b.hx, sv); lb.hy = __shfl_sync(B2_FULL, b.hy, sv); lb.hz = __shfl_sync(B2_FULL, b.hz, sv);

  /* leaf g = {primitive with the g-th smallest code, its box} */
  if (lane < n) {
    u32* w = S.leaf + lane * 7;
    w[0] = sv; w[1] = __float_as_uint(lb.lx); w[2] = __float_as_uint(lb.ly); w[3] = __float_as_uint(lb.lz);
    w[4] = __float_as_uint(lb.hx); w[5] = __float_as_uint(lb.hy); w[6] = __float_as_uint(lb.hz);
  }
  /* boundary b (1 <= b < n) lies between sorted leaves b-1 and b: findHighestDiffBit (:120-126) */
  {
    const u32 skPrev = __shfl_up_sync(B2_FULL, sk, 1);
    u64 x = ~0ull;
    if (lane >= 1 && lane < n) x = ((u64)(skPrev ^ sk) << 32) | (u64)((lane - 1) ^ lane);
    S.xb[lane] = x;
    if (lane == 0) S.xb[BATCH_MAX] = ~0ull;
    if (lane == 0 && n < BATCH_MAX) S.xb[n] = ~0ull;
  }
  __syncwarp();
  {
    u32* dst = reinterpret_cast<u32*>(leaves + first);
    for (u32 q = lane; q < n * 7; q += 32) dst[q] = S.leaf[q];
  }

  /* ---- hierarchy: rounds of pairwise merges over the ordered cluster list.  A cluster stays in the lane of its first leaf; the list
   * is the set of live lanes (a mask), neighbours are the next live lanes, and every per-round decision that involves a neighbour is
   * a bit test on a ballot — the only shuffles of a round fetch the absorbed neighbour's packed range/index word and its box ---- */
  u32 meta = lane | ((lane + 1u) << 6) | (((n - 1u) + lane) << 12); /* [5:0] lo, [11:6] hi, [17:12] node id (leaf g = n-1+g <= 62) */
  Box cb = lb;
  u32 aliveMask = n >= 32u ? B2_FULL : ((1u << n) - 1u);
  const u32 gt = ~lanemask_lt() << 1; /* lanes above mine */
  while (aliveMask & (aliveMask - 1u)) { /* more than one cluster */
    const bool alive = (aliveMask >> lane) & 1u;
    const u32 lo = meta & 63u, hi = (meta >> 6) & 63u;
    /* findParent (:136-159): right boundary when it is deeper than the left one (or there is no left one) */
    const bool goRight = alive && hi != n && (lo == 0 || S.xb[hi] < S.xb[lo]);
    const u32 grMask = __ballot_sync(B2_FULL, goRight);
    const u32 above = aliveMask & gt;
    const u32 r = above ? (u32)__ffs((int)above) - 1u : lane; /* my right neighbour */
    const bool merge = goRight && above && !((grMask >> r) & 1u); /* it picked the boundary between us as well */
    const u32 mergeMask = __ballot_sync(B2_FULL, merge);
    const u32 nbMeta = __shfl_sync(B2_FULL, meta, r);
    Box nb;
    nb.lx = __shfl_sync(B2_FULL, cb.lx, r); nb.ly = __shfl_sync(B2_FULL, cb.ly, r); nb.lz = __shfl_sync(B2_FULL, cb.lz, r);
    nb.hx = __shfl_sync(B2_FULL, cb.hx, r); nb.hy = __shfl_sync(B2_FULL, cb.hy, r); nb.hz = __shfl_sync(B2_FULL, cb.hz, r);
    if (merge) {
      const u32 node = hi - 1u; /* Apetrei: the node's index is its split position */
      cb = box_union(cb, nb);
      store_node2(nodes + nOff + node, meta >> 12, nbMeta >> 12, cb);
      meta = lo | (nbMeta & (63u << 6)) | (node << 12);
    }
    /* the absorbed clusters are the right neighbours of the merging ones: the live lane just above each set bit of mergeMask */
    const u32 below = aliveMask & lanemask_lt();
    const bool absorbed = alive && below && ((mergeMask >> (31u - (u32)__clz((int)below))) & 1u);
    aliveMask &= ~__ballot_sync(B2_FULL, absorbed);
  }
  const u32 id = meta >> 12; /* lane 0 is never absorbed: it holds the root */
  if (lane == 0) roots[item] = (n == 1) ? 0u : id; /* rootNodes[item] (:311) */
}

int b2_launch_batched(b2bvh_ctx* ctx, const b2bvh_triangle* d_tris, const u32* d_leafOff, const u32* d_nodeOff, u32 nItems, b2bvh_bvh2_node* d_nodes,
                      b2bvh_prim_ref* d_leaves, u32* d_roots, b2bvh_aabb* d_scenes) {
  B2_KERNEL(ctx, "batched_lbvh");
  batched_lbvh_kernel<<<(nItems + BATCH_WARPS - 1) / BATCH_WARPS, BATCH_WARPS * 32, 0, ctx->stream>>>(d_tris, d_leafOff, d_nodeOff, nItems, d_nodes, d_leaves,
                                                                                                     d_roots, d_scenes);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
