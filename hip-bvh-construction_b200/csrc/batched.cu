/*
 * batched.cu — the batched builder: thousands of tiny BVHs (<= 32 triangles each, MaxBatchedBlockSize, Common.h:597) in one
 * launch.  Replaces BatchedBvhBuilder::build (BatchedBuilder.cpp:16-77) + BatchedBuildKernelLbvh (BatchedBuildKernel.h:218-312).
 *
 * The reference gives every item a 32-thread block: block reduction of the item's box through shared memory, plain 10/10/10
 * Morton codes, 32 one-bit split passes with a block prefix sum and two barriers each (a stable LSD sort), then the Apetrei
 * build with atomic counters in shared memory and __threadfence().  Here an item is ONE WARP and nothing leaves its registers
 * but the results:
 *   box + scene box   one triangle per lane (2 x 16-byte + 1 x 4-byte load), min/max butterflies
 *   sort              rank of (code, lane) by 32 broadcasts — the rank IS the stable sorted position
 *   hierarchy         the sorted leaves are an ordered list of clusters, one per lane.  A cluster [lo, hi) joins the node at the
 *                     deeper of its two boundaries (findParent, BatchedBuildKernel.h:136-159: smaller XOR of the augmented
 *                     keys = longer common prefix); two neighbours that pick the boundary between them become node hi-1 of
 *                     the left one — the index the reference's walk gives it.  All such pairs merge in the same round, the
 *                     list is compacted with one ballot (survivor lanes through 32 words of shared memory), and the next round
 *                     starts; no atomics, no fences.
 * Eight items per CTA; nodes leave as two 16-byte stores, leaf records (28 bytes) through a shared-memory transpose as
 * consecutive words.  Algorithmic traffic per triangle: 64 B read + 28 B leaf + 32 B node written.
 *
 * Repairs of the reference kernel (which is unfinished: ExtentCacheSize is undefined, main.cpp:38-52 keeps it disabled),
 * restated by the CPU oracle (orc_batched_build): leaves are stored in SORTED order (the reference never permutes them,
 * :241-242); item offsets are prefix sums of the item sizes (the reference multiplies by the item's own size, :234-235); an
 * item of one triangle has no internal node and root 0.  Node and leaf indices are local to the item, as in the reference.
 */
#include "common.cuh"

#define BATCH_WARPS 8
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
  Box sc = b;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sc.lx = fminf(sc.lx, __shfl_xor_sync(B2_FULL, sc.lx, o)); sc.ly = fminf(sc.ly, __shfl_xor_sync(B2_FULL, sc.ly, o));
    sc.lz = fminf(sc.lz, __shfl_xor_sync(B2_FULL, sc.lz, o)); sc.hx = fmaxf(sc.hx, __shfl_xor_sync(B2_FULL, sc.hx, o));
    sc.hy = fmaxf(sc.hy, __shfl_xor_sync(B2_FULL, sc.hy, o)); sc.hz = fmaxf(sc.hz, __shfl_xor_sync(B2_FULL, sc.hz, o));
  }
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
  u32 rank = 0;
#pragma unroll
  for (u32 j = 0; j < 32; j++) {
    const u32 kj = __shfl_sync(B2_FULL, key, j);
    rank += (kj < key || (kj == key && j < lane)) ? 1u : 0u;
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

  /* ---- hierarchy: rounds of pairwise merges over the ordered cluster list (one cluster per lane) ---- */
  u32 lo = lane, hi = lane + 1, id = (n - 1) + lane, count = n;
  Box cb = lb;
  while (count > 1) {
    const bool active = lane < count;
    /* findParent (:136-159): right boundary when it is deeper than the left one (or there is no left one) */
    const bool goRight = active && hi != n && (lo == 0 || S.xb[hi] < S.xb[lo]);
    const bool nbGoRight = __shfl_down_sync(B2_FULL, goRight, 1);
    const bool merge = active && (lane + 1 < count) && goRight && !nbGoRight; /* my right neighbour picked the same boundary */
    const bool absorbed = __shfl_up_sync(B2_FULL, merge, 1) && lane > 0;
    const u32 nbHi = __shfl_down_sync(B2_FULL, hi, 1), nbId = __shfl_down_sync(B2_FULL, id, 1);
    Box nb;
    nb.lx = __shfl_down_sync(B2_FULL, cb.lx, 1); nb.ly = __shfl_down_sync(B2_FULL, cb.ly, 1); nb.lz = __shfl_down_sync(B2_FULL, cb.lz, 1);
    nb.hx = __shfl_down_sync(B2_FULL, cb.hx, 1); nb.hy = __shfl_down_sync(B2_FULL, cb.hy, 1); nb.hz = __shfl_down_sync(B2_FULL, cb.hz, 1);
    if (merge) {
      const u32 node = hi - 1; /* Apetrei: the node's index is its split position */
      cb = box_union(cb, nb);
      store_node2(nodes + nOff + node, id, nbId, cb);
      id = node;
      hi = nbHi;
    }
    const bool alive = active && !absorbed;
    const u32 aliveMask = __ballot_sync(B2_FULL, alive);
    /* lane of the (lane+1)-th surviving cluster, through shared memory (__fns is a software loop: a fifth of the kernel's instructions) */
    __syncwarp();
    if (alive) S.val[__popc(aliveMask & lanemask_lt())] = lane;
    __syncwarp();
    const u32 src = S.val[lane] & 31u; /* stale past the new count: unused */
    lo = __shfl_sync(B2_FULL, lo, src); hi = __shfl_sync(B2_FULL, hi, src); id = __shfl_sync(B2_FULL, id, src);
    cb.lx = __shfl_sync(B2_FULL, cb.lx, src); cb.ly = __shfl_sync(B2_FULL, cb.ly, src); cb.lz = __shfl_sync(B2_FULL, cb.lz, src);
    cb.hx = __shfl_sync(B2_FULL, cb.hx, src); cb.hy = __shfl_sync(B2_FULL, cb.hy, src); cb.hz = __shfl_sync(B2_FULL, cb.hz, src);
    count = __popc(aliveMask);
  }
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
