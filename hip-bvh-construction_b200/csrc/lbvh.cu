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

template <bool KARRAS>
__global__ void __launch_bounds__(LBVH_THREADS) lbvh_fused_kernel(const u32* __restrict__ keys, const u32* __restrict__ vals,
                                                                  const b2bvh_aabb* __restrict__ triAabb, u32 n, b2bvh_bvh2_node* nodes,
                                                                  u32* parents, u32* meet /* n-1 words, 0xFFFFFFFF */, u32* rootOut) {
  const u32 g = blockIdx.x * LBVH_THREADS + threadIdx.x;
  if (g >= n) return;
  const u32 nInt = n - 1;
  const u32 prim = __ldg(vals + g);
  Box box = load_aabb(triAabb + prim);
  store_node2(nodes + nInt + g, prim, B2_INVALID, box);
  if (n == 1) { if (rootOut) *rootOut = 0; return; }

  u32 lo = g, hi = g + 1;
  u32 self = nInt + g; /* index of the node this thread currently stands on (already written) */
  bool isLeft;
  u32 p = choose_parent(keys, n, lo, hi, isLeft);
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

int b2_launch_lbvh_fused(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                         b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch, u32* d_root, int karrasNumbering) {
  if (n > 1) B2_CUDA(cudaMemsetAsync(d_scratch, 0xFF, (size_t)(n - 1) * sizeof(u32), ctx->stream));
  const u32 grid = (n + LBVH_THREADS - 1) / LBVH_THREADS;
  B2_KERNEL(ctx, karrasNumbering ? "lbvh_fused_karras" : "lbvh_fused_apetrei");
  if (karrasNumbering)
    lbvh_fused_kernel<true><<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root);
  else
    lbvh_fused_kernel<false><<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root);
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
