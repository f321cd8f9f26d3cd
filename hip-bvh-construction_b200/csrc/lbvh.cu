/*
 * lbvh.cu — stage S4: LBVH hierarchy emit + bottom-up AABB refit.
 *
 * Replaces, for TwoPassLbvh (TwoPassLbvh.cpp:95-143):
 *     InitBvhNodesPrimRef (TwoPassLbvhKernel.h:164), determineRange/findSplit/BvhBuild (:42,:102,:196), FitBvhNodes (:217)
 * and for SinglePassLbvh (SinglePassLbvh.cpp:99-131):
 *     InitBvhNodes (SinglePassLbvhKernel.h:27), findParent (:64), BvhBuildAndFit (:88).
 *
 * Both reference builders produce the SAME ordered radix tree over the augmented keys (key << 32 | index);
 * they differ only in how internal nodes are numbered:
 *     Karras  : a node that is a left child takes the LAST leaf index of its range, a right child the FIRST; root 0.
 *     Apetrei : a node takes the position of its split; the root index is reported.
 * lbvh_fused_kernel builds either numbering in ONE bottom-up pass: a thread per leaf climbs, the two children of
 * a node meet through a single atomic exchange that hands the first arriver's range bound to the second, and the
 * second arriver — which then knows the node's full range, hence its parent, hence its own index — writes the
 * finished 32-byte node with two 16-byte stores.  Boxes are min/max only, so the result is bit-exact.
 * lbvh_karras_emit_kernel + lbvh_refit_kernel keep the reference's two-launch structure (top-down range/split search,
 * then refit) for TwoPassLbvh when karras_two_kernel is requested; results are identical.
 *
 * Traffic per primitive: sorted value 4 + sorted key 4 (+ neighbours from L1/L2) + gathered box 24 + leaf node 32 written
 *   + internal node 32 written + sibling node 32 re-read + 4 (exchange word)  ~ 132 B (+8 when the parent array is written).
 */
#include <stdlib.h>

#include "common.cuh"

#define LBVH_THREADS 256

__device__ __forceinline__ u64 aug_key(const u32* __restrict__ keys, u32 i) { return ((u64)__ldg(keys + i) << 32) | i; }

/* Parent choice for the node covering leaves [lo, hi) (findParent, SinglePassLbvhKernel.h:64-86):
 * returns the split position p of the parent and whether this node is its LEFT child.  The parent splits
 * between leaves p and p+1.  Must not be called for the root. */
__device__ __forceinline__ u32 choose_parent(const u32* __restrict__ keys, u32 n, u32 lo, u32 hi, bool& isLeft) {
  if (lo == 0) { isLeft = true; return hi - 1; }
  if (hi == n) { isLeft = false; return lo - 1; }
  const u64 xr = aug_key(keys, hi - 1) ^ aug_key(keys, hi);
  const u64 xl = aug_key(keys, lo - 1) ^ aug_key(keys, lo);
  isLeft = xr < xl;
  return isLeft ? hi - 1 : lo - 1;
}

/* The climb through GLOBAL memory, starting from a finished node `self` covering [lo, hi) with box `box` whose parent has
 * split `p` (this node being its left child iff isLeft).  Returns when the node is the first to arrive at some parent
 * (the sibling's thread takes over) or when the root has been written. */
template <bool KARRAS>
__device__ __forceinline__ void climb_global(const u32* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, u32* parents, u32* meet, u32* rootOut,
                                             u32 self, u32 lo, u32 hi, Box box, u32 p, bool isLeft) {
  const u32 nInt = n - 1;
  while (true) {
    if (!KARRAS) {
      /* Apetrei numbering: hand our index to the parent slot (the sibling cannot derive it) */
      st_relaxed(reinterpret_cast<u32*>(nodes + p) + (isLeft ? 0 : 1), self);
    }
    const u32 other = atom_exch_acq_rel(meet + p, isLeft ? lo : hi);
    if (other == B2_INVALID) return; /* first arriver: the sibling's thread finishes this node */
    /* second arriver: the node with split p now has its full range */
    if (isLeft) hi = other; else lo = other;
    u32 sib;
    if (KARRAS) {
      /* derivable: left child = p (leaf: p + nInt), right child = p + 1 (leaf: p + 1 + nInt) */
      if (isLeft) sib = (p + 2 == hi) ? p + 1 + nInt : p + 1;
      else sib = (p == lo) ? p + nInt : p;
    } else {
      sib = ld_relaxed(reinterpret_cast<const u32*>(nodes + p) + (isLeft ? 1 : 0));
    }
    box = box_union(box, load_node2_cg(nodes + sib).box);
    const u32 left = isLeft ? self : sib, right = isLeft ? sib : self;
    const bool isRoot = (lo == 0 && hi == n);
    const u32 split = p;
    if (!isRoot) p = choose_parent(keys, n, lo, hi, isLeft); /* parent of the finished node: fixes its Karras index */
    const u32 id = KARRAS ? (isRoot ? 0u : (isLeft ? hi - 1 : lo)) : split;
    store_node2(nodes + id, left, right, box);
    if (parents) { parents[left] = id; parents[right] = id; if (isRoot) parents[id] = B2_INVALID; }
    if (isRoot) { if (rootOut) *rootOut = id; return; }
    self = id;
  }
}

template <bool KARRAS>
__global__ void __launch_bounds__(LBVH_THREADS) lbvh_fused_kernel(const u32* __restrict__ keys, const u32* __restrict__ vals,
                                                                  const b2bvh_aabb* __restrict__ triAabb, u32 n, b2bvh_bvh2_node* nodes,
                                                                  u32* parents, u32* meet /* n-1 words, 0xFFFFFFFF */, u32* rootOut) {
  const u32 g = blockIdx.x * LBVH_THREADS + threadIdx.x;
  if (g >= n) return;
  const u32 nInt = n - 1;
  const u32 prim = __ldg(vals + g);
  const Box box = load_aabb(triAabb + prim);
  store_node2(nodes + nInt + g, prim, B2_INVALID, box);
  if (n == 1) { if (rootOut) *rootOut = 0; return; }
  bool isLeft;
  const u32 p = choose_parent(keys, n, g, g + 1, isLeft);
  climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, nInt + g, g, g + 1, box, p, isLeft);
}

/* ---------------------------------------------------------------- CTA-local climb in shared memory, then the global climb
 * A CTA owns BL consecutive leaves [b0, b1).  Every internal node whose range lies inside [b0, b1) is finished by one of the
 * CTA's own threads, and when BOTH children of a node lie inside, the two arrivals can meet in shared memory: the
 * exchange word, the first arriver's box and its index are shared-memory traffic, and no fence or L2 round trip is paid.
 * That covers all but O(log BL) nodes per CTA.  A finished node whose sibling does NOT lie inside the CTA (its parent
 * straddles a CTA boundary) is left "stranded" in its slot; after a CTA barrier the stranded nodes, and the nodes that
 * reached the CTA boundary directly, continue through the global protocol above.  Every finished node is written to global
 * memory exactly once, with two 16-byte stores, as before; results are identical to lbvh_fused_kernel. */
#define LBVH_BL 512
#define LBVH_CONSUMED 0xFFFFFFFEu

/* A finished node that must continue through global memory: 48 bytes, written by lbvh_block_kernel, consumed by lbvh_climb_kernel */
struct LbvhPending {
  u32 self, lo, hi, pSide; /* pSide = parent split | (isLeft << 31) */
  float box[6];
  u32 pad[2];
};
static_assert(sizeof(LbvhPending) == 48, "LbvhPending layout");

struct LbvhBlockSmem {
  u32 pendCount, pendBase;
  u32 key[LBVH_BL + 2];          /* keys of leaves b0-1 .. b1 */
  u32 meet[LBVH_BL];             /* exchange word of split p (between leaves p and p+1), index p - b0 */
  u32 sibId[LBVH_BL][2];         /* [split][side]: index of the child that arrived from that side (0 = left child) */
  float sibBox[LBVH_BL][2][6];
};

__device__ __forceinline__ u32 choose_parent_smem(const u32* sk, u32 kb /* leaf index of sk[0] */, u32 n, u32 lo, u32 hi, bool& isLeft) {
  if (lo == 0) { isLeft = true; return hi - 1; }
  if (hi == n) { isLeft = false; return lo - 1; }
  const u64 a = ((u64)sk[hi - 1 - kb] << 32) | (hi - 1), b = ((u64)sk[hi - kb] << 32) | hi;
  const u64 c = ((u64)sk[lo - 1 - kb] << 32) | (lo - 1), d = ((u64)sk[lo - kb] << 32) | lo;
  isLeft = (a ^ b) < (c ^ d);
  return isLeft ? hi - 1 : lo - 1;
}

template <bool KARRAS>
__global__ void __launch_bounds__(LBVH_BL) lbvh_block_kernel(const u32* __restrict__ keys, const u32* __restrict__ vals,
                                                             const b2bvh_aabb* __restrict__ triAabb, u32 n, b2bvh_bvh2_node* nodes, u32* parents,
                                                             u32* meet, u32* rootOut, u32* pendingCount, LbvhPending* pending, u32 pendingCap) {
  __shared__ LbvhBlockSmem S;
  const u32 t = threadIdx.x;
  const u32 b0 = blockIdx.x * LBVH_BL, b1 = min(n, b0 + LBVH_BL);
  const u32 g = b0 + t;
  const u32 nInt = n - 1;
  const u32 kb = b0 - 1; /* leaf index of S.key[0] (wraps for b0 == 0; slot 0 is then unused) */
  S.meet[t] = B2_INVALID;
  S.sibId[t][0] = B2_INVALID; S.sibId[t][1] = B2_INVALID;
  if (t == 0) S.pendCount = 0;
  if (g < n) S.key[t + 1] = __ldg(keys + g);
  if (t == 0 && b0 > 0) S.key[0] = __ldg(keys + b0 - 1);
  if (t == 0 && b1 < n) S.key[b1 - b0 + 1] = __ldg(keys + b1);
  Box box = box_empty();
  u32 self = B2_INVALID, lo = 0, hi = 0, p = 0;
  bool isLeft = false, goGlobal = false;
  if (g < n) {
    const u32 prim = __ldg(vals + g);
    box = load_aabb(triAabb + prim);
    store_node2(nodes + nInt + g, prim, B2_INVALID, box);
  }
  __syncthreads();
  if (n == 1) { if (g == 0 && rootOut) *rootOut = 0; return; }

  if (g < n) {
    lo = g; hi = g + 1; self = nInt + g;
    p = choose_parent_smem(S.key, kb, n, lo, hi, isLeft);
    while (true) {
      if (p < b0 || p + 1 >= b1) { goGlobal = true; break; } /* the parent straddles the CTA boundary */
      const u32 slot = p - b0;
      const int side = isLeft ? 0 : 1;
      float* sb = S.sibBox[slot][side];
      sb[0] = box.lx; sb[1] = box.ly; sb[2] = box.lz; sb[3] = box.hx; sb[4] = box.hy; sb[5] = box.hz;
      S.sibId[slot][side] = self;
      __threadfence_block();
      const u32 other = atomicExch(&S.meet[slot], isLeft ? lo : hi);
      if (other == B2_INVALID) break; /* first arriver: waits in its slot for the sibling (or for the escalation below) */
      __threadfence_block();
      S.meet[slot] = LBVH_CONSUMED;
      if (isLeft) hi = other; else lo = other;
      const float* ob = S.sibBox[slot][side ^ 1];
      const u32 sib = S.sibId[slot][side ^ 1];
      box = box_union(box, Box{ob[0], ob[1], ob[2], ob[3], ob[4], ob[5]});
      const u32 left = isLeft ? self : sib, right = isLeft ? sib : self;
      const bool isRoot = (lo == 0 && hi == n);
      const u32 split = p;
      if (!isRoot) p = choose_parent_smem(S.key, kb, n, lo, hi, isLeft);
      const u32 id = KARRAS ? (isRoot ? 0u : (isLeft ? hi - 1 : lo)) : split;
      store_node2(nodes + id, left, right, box);
      if (parents) { parents[left] = id; parents[right] = id; if (isRoot) parents[id] = B2_INVALID; }
      if (isRoot) { if (rootOut) *rootOut = id; break; }
      self = id;
    }
  }
  __syncthreads();
  /* ---- hand over to the global climb (separate launch, so this CTA retires as soon as its shared-memory work is done):
   * (1) this thread's own node if it reached the CTA boundary, (2) the stranded node of slot t ---- */
  const u32 m = S.meet[t];
  const bool stranded = (m != B2_INVALID && m != LBVH_CONSUMED);
  u32 myIdx = 0;
  const u32 mine = (goGlobal ? 1u : 0u) + (stranded ? 1u : 0u);
  if (mine) myIdx = atomicAdd(&S.pendCount, mine);
  __syncthreads();
  if (t == 0 && S.pendCount) S.pendBase = atomicAdd(pendingCount, S.pendCount);
  __syncthreads();
  u32 slotOut = S.pendBase + myIdx;
  if (goGlobal) {
    if (slotOut < pendingCap) {
      uint4* q = reinterpret_cast<uint4*>(pending + slotOut);
      q[0] = make_uint4(self, lo, hi, p | (isLeft ? 0x80000000u : 0u));
      q[1] = make_uint4(__float_as_uint(box.lx), __float_as_uint(box.ly), __float_as_uint(box.lz), __float_as_uint(box.hx));
      q[2] = make_uint4(__float_as_uint(box.hy), __float_as_uint(box.hz), 0u, 0u);
    } else {
      climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, self, lo, hi, box, p, isLeft); /* overflow of the hand-over list */
    }
    slotOut++;
  }
  if (stranded) {
    const u32 sp = b0 + t; /* split of the stranded node's parent */
    const int side = (S.sibId[t][0] != B2_INVALID) ? 0 : 1;
    const float* ob = S.sibBox[t][side];
    const u32 slo = side == 0 ? m : sp + 1, shi = side == 0 ? sp + 1 : m;
    if (slotOut < pendingCap) {
      uint4* q = reinterpret_cast<uint4*>(pending + slotOut);
      q[0] = make_uint4(S.sibId[t][side], slo, shi, sp | (side == 0 ? 0x80000000u : 0u));
      q[1] = make_uint4(__float_as_uint(ob[0]), __float_as_uint(ob[1]), __float_as_uint(ob[2]), __float_as_uint(ob[3]));
      q[2] = make_uint4(__float_as_uint(ob[4]), __float_as_uint(ob[5]), 0u, 0u);
    } else {
      climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, S.sibId[t][side], slo, shi, Box{ob[0], ob[1], ob[2], ob[3], ob[4], ob[5]}, sp, side == 0);
    }
  }
}

template <bool KARRAS>
__global__ void __launch_bounds__(LBVH_THREADS) lbvh_climb_kernel(const u32* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, u32* parents, u32* meet,
                                                                  u32* rootOut, const u32* __restrict__ pendingCount,
                                                                  const LbvhPending* __restrict__ pending, u32 pendingCap) {
  const u32 count = min(*pendingCount, pendingCap);
  for (u32 i = blockIdx.x * LBVH_THREADS + threadIdx.x; i < count; i += gridDim.x * LBVH_THREADS) {
    const uint4* q = reinterpret_cast<const uint4*>(pending + i);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const Box box = Box{__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(c.x), __uint_as_float(c.y)};
    climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, a.x, a.y, a.z, box, a.w & 0x7FFFFFFFu, (a.w >> 31) != 0u);
  }
}

/* ---------------------------------------------------------------- two-launch Karras variant */
__device__ __forceinline__ int delta_aug(const u32* __restrict__ keys, u32 n, u32 i, int j) {
  if (j < 0 || j >= (int)n) return -1;
  return __clzll((long long)(aug_key(keys, i) ^ aug_key(keys, (u32)j)));
}

__global__ void __launch_bounds__(LBVH_THREADS) lbvh_karras_emit_kernel(const u32* __restrict__ keys, const u32* __restrict__ vals,
                                                                        const b2bvh_aabb* __restrict__ triAabb, u32 n, b2bvh_bvh2_node* nodes,
                                                                        u32* parents) {
  const u32 g = blockIdx.x * LBVH_THREADS + threadIdx.x;
  if (g >= n) return;
  const u32 nInt = n - 1;
  {
    const u32 prim = __ldg(vals + g);
    store_node2(nodes + nInt + g, prim, B2_INVALID, load_aabb(triAabb + prim));
  }
  if (g == 0) parents[0] = B2_INVALID;
  if (g >= nInt) return;
  /* range of internal node g: one end is g, the other is found by galloping + bisection on the common-prefix length */
  u32 first, last;
  if (g == 0) { first = 0; last = n - 1; }
  else {
    const int dR = delta_aug(keys, n, g, (int)g + 1), dL = delta_aug(keys, n, g, (int)g - 1);
    const int dir = dR > dL ? 1 : -1;
    const int dMin = min(dL, dR);
    int reach = 2;
    while (delta_aug(keys, n, g, (int)g + dir * reach) > dMin) reach <<= 1;
    int len = 0;
    for (int t = reach >> 1; t > 0; t >>= 1)
      if (delta_aug(keys, n, g, (int)g + (len + t) * dir) > dMin) len += t;
    const u32 other = (u32)((int)g + len * dir);
    first = dir > 0 ? g : other; last = dir > 0 ? other : g;
  }
  /* split: last leaf that shares more than the node's common prefix with `first` */
  const int dNode = delta_aug(keys, n, first, (int)last);
  u32 split = first, stride = last - first;
  do {
    stride = (stride + 1) >> 1;
    const u32 mid = split + stride;
    if (mid < last && delta_aug(keys, n, first, (int)mid) > dNode) split = mid;
  } while (stride > 1);
  const u32 left = (split == first) ? split + nInt : split;
  const u32 right = (split + 1 == last) ? split + 1 + nInt : split + 1;
  store_node2(nodes + g, left, right, box_empty());
  parents[left] = g;
  parents[right] = g;
}

__global__ void __launch_bounds__(LBVH_THREADS) lbvh_refit_kernel(b2bvh_bvh2_node* nodes, const u32* __restrict__ parents, u32* flags, u32 n) {
  const u32 g = blockIdx.x * LBVH_THREADS + threadIdx.x;
  if (g >= n) return;
  const u32 nInt = n - 1;
  u32 self = nInt + g;
  Box box = load_node2_ro(nodes + self).box; /* leaf boxes come from the emit launch */
  u32 p = __ldg(parents + self);
  while (p != B2_INVALID) {
    if (atom_add_acq_rel(flags + p, 1u) == 0) return;
    const uint2 ch = __ldcg(reinterpret_cast<const uint2*>(nodes + p));
    const u32 sib = ch.x == self ? ch.y : ch.x;
    box = box_union(box, load_node2_cg(nodes + sib).box);
    store_node2(nodes + p, ch.x, ch.y, box);
    self = p;
    p = __ldg(parents + p);
  }
}

size_t b2_lbvh_scratch_bytes(u32 n) {
  /* the larger of: fused path (meet words + hand-over list) and two-kernel path (2n-1 flags) */
  const size_t fused = (((size_t)n * 4 + 15) & ~(size_t)15) + 16 + ((size_t)n / 8 + 1024) * sizeof(LbvhPending);
  const size_t two = (2 * (size_t)n - 1) * 4;
  return fused > two ? fused : two;
}

int b2_launch_lbvh_fused(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                         b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch, u32* d_root, int karrasNumbering) {
  if (n > 1) B2_CUDA(cudaMemsetAsync(d_scratch, 0xFF, (size_t)(n - 1) * sizeof(u32), ctx->stream));
  B2_KERNEL(ctx, karrasNumbering ? "lbvh_fused_karras" : "lbvh_fused_apetrei");
  static const bool globalOnly = getenv("B2BVH_LBVH_GLOBAL_ONLY") != nullptr; /* development switch: the all-global-memory variant */
  if (globalOnly) {
    const u32 grid = (n + LBVH_THREADS - 1) / LBVH_THREADS;
    if (karrasNumbering)
      lbvh_fused_kernel<true><<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root);
    else
      lbvh_fused_kernel<false><<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root);
  } else {
    /* scratch: meet[n-1] | (16-byte aligned) pendingCount | pending[cap] */
    const size_t off = (((size_t)n * 4 + 15) & ~(size_t)15);
    u32* pendingCount = reinterpret_cast<u32*>(reinterpret_cast<unsigned char*>(d_scratch) + off);
    LbvhPending* pending = reinterpret_cast<LbvhPending*>(reinterpret_cast<unsigned char*>(d_scratch) + off + 16);
    const u32 cap = n / 8 + 1024;
    B2_CUDA(cudaMemsetAsync(pendingCount, 0, 4, ctx->stream));
    const u32 grid = (n + LBVH_BL - 1) / LBVH_BL;
    if (karrasNumbering)
      lbvh_block_kernel<true><<<grid, LBVH_BL, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root, pendingCount, pending, cap);
    else
      lbvh_block_kernel<false><<<grid, LBVH_BL, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root, pendingCount, pending, cap);
    B2_LAUNCH_CHECK(ctx);
    u32 grid2 = (cap + LBVH_THREADS - 1) / LBVH_THREADS;
    const u32 cap2 = (u32)ctx->sm_count * 8u;
    if (grid2 > cap2) grid2 = cap2;
    B2_KERNEL(ctx, "lbvh_climb");
    if (karrasNumbering)
      lbvh_climb_kernel<true><<<grid2, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, n, d_nodes, d_parents, d_scratch, d_root, pendingCount, pending, cap);
    else
      lbvh_climb_kernel<false><<<grid2, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, n, d_nodes, d_parents, d_scratch, d_root, pendingCount, pending, cap);
  }
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2_launch_lbvh_karras_two_kernel(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                                     b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_flags) {
  const u32 grid = (n + LBVH_THREADS - 1) / LBVH_THREADS;
  B2_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)(2 * (size_t)n - 1) * sizeof(u32), ctx->stream));
  B2_KERNEL(ctx, "lbvh_karras_emit");
  lbvh_karras_emit_kernel<<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents);
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "lbvh_refit");
  lbvh_refit_kernel<<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_nodes, d_parents, d_flags, n);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
