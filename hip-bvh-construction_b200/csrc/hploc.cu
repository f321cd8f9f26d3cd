/*
 * hploc.cu — stage S7: H-PLOC (Benthin et al. 2024): the whole build in ONE launch; a warp walks the LBVH hierarchy
 * bottom-up and PLOC-merges at most 32 clusters in registers whenever a hierarchy node covers more than 16 leaves.
 *
 * Replaces SetupClusters (HplocKernel.h:39-56), HPloc (:257-315), findParent (:66-81), plocMerge (:220-255),
 * loadIndices / storeIndices (:192-218), findNearestNeighbours (:83-117), mergeClusters (:126-190) and the host side
 * HPLOC::build (Hploc.cpp:86-121).
 *
 * Semantics kept: hierarchy = radix tree over (key << 32 | index); siblings meet through one atomic
 * exchange; the second arriver owns the union range [L,R] with `split` = first index of the right part; if the range has
 * more than 16 leaves (or is the whole array) the <=16 leading cluster ids of each side are merged with radius-8
 * nearest-neighbour search on (area bits << 32 | lane) until <=16 (<=1 for the root) remain; the list is written back
 * to nodeIdx[L ..).
 * What is different:
 *   - the 32-slot cluster list lives in registers, one slot per lane: neighbour boxes come from shuffles, the
 *     nearest-neighbour key is exchanged with shuffles (the reference uses LDS arrays and 64-bit LDS atomics), and the
 *     compaction is a ballot + find-nth-set-bit gather — no same-address shared-memory writes (HplocKernel.h:183-185);
 *   - node numbering: no global atomicAdd (:162-168).  Every cluster carries one free node index (leaf g >= 1 starts with
 *     g-1); a merge of lanes l < p stores the new node at the index carried by p and keeps the one carried by l.  The
 *     numbering is deterministic and identical to the oracle's; the root is exchanged with the node at index 0 at the
 *     end so that the root index is 0 as in the reference;
 *   - launched with one lane per LEAF (the reference launches N-1 threads and loses leaf N-1 when (N-1) % 32 == 0).
 * Traffic per primitive: key 4 + (id, free index) 8 read/written ~2x + leaf box 28 + node written 32 + node box re-read ~32.
 */
#include "common.cuh"

#ifndef HP_THREADS
#define HP_THREADS 32  /* one warp per CTA: a CTA's slot is free again as soon as its warp has handed over its last range.  10 M uniform, merge stage in ms
                          (tools/variant_bench.py, gpurun r2i): 32 threads 2.28, 64 2.37, 128 2.40, 256 2.44; one fence per round instead of one per merge
                          call (HP_FENCE_ONCE) changes nothing (2.40 / 2.38 / 2.44): the kernel is bound by the ~7 us dependent chain of a merge call
                          (ids -> boxes -> 2-3 search rounds -> stores -> fence -> exchange) at ~24 resident warps per SM, not by the fences */
#endif
#define HP_R 8

/* scratch: u32 ctrl[64] | nodeIdx[n] | freeIdx[n] | meet[n] ;  ctrl[0] = merge calls, ctrl[1] = parent of node 0, ctrl[2] = side */
size_t b2_hploc_scratch_bytes(u32 n) { return 256 + 3 * (size_t)n * 4; }

__global__ void __launch_bounds__(256) hploc_setup_kernel(const b2bvh_aabb* __restrict__ triAabb, const u32* __restrict__ sortedVals, u32 n,
                                                          b2bvh_prim_ref* __restrict__ leaves, u32* __restrict__ nodeIdx, u32* __restrict__ freeIdx,
                                                          u32* __restrict__ meet, u32* ctrl) {
  __shared__ __align__(16) u32 sLeaf[256 * 7];
  const u32 t = threadIdx.x, g0 = blockIdx.x * 256, g = g0 + t;
  if (g < 8) ctrl[g] = (g == 1) ? B2_INVALID : 0u;
  if (g < n) {
    const u32 prim = __ldg(sortedVals + g);
    const float2* bp = reinterpret_cast<const float2*>(triAabb + prim); /* 24-byte boxes: 8-byte aligned */
    const float2 q0 = ldg_gather_f2(bp), q1 = ldg_gather_f2(bp + 1), q2 = ldg_gather_f2(bp + 2);
    u32* l = sLeaf + t * 7;
    l[0] = prim; l[1] = __float_as_uint(q0.x); l[2] = __float_as_uint(q0.y); l[3] = __float_as_uint(q1.x); l[4] = __float_as_uint(q1.y);
    l[5] = __float_as_uint(q2.x); l[6] = __float_as_uint(q2.y);
    nodeIdx[g] = g + (n - 1);
    freeIdx[g] = g ? g - 1 : B2_INVALID;
    meet[g] = B2_INVALID;
  }
  __syncthreads();
  cta_store_words(reinterpret_cast<u32*>(leaves + g0), sLeaf, min(256u, n - g0) * 7); /* 256 x 28 B = 7168 B per CTA: 16-byte aligned */
}

__device__ __forceinline__ Box shfl_box(const Box& b, int src) {
  return Box{__shfl_sync(B2_FULL, b.lx, src), __shfl_sync(B2_FULL, b.ly, src), __shfl_sync(B2_FULL, b.lz, src),
             __shfl_sync(B2_FULL, b.hx, src), __shfl_sync(B2_FULL, b.hy, src), __shfl_sync(B2_FULL, b.hz, src)};
}
__device__ __forceinline__ Box shfl_down_box(const Box& b, int d) {
  return Box{__shfl_down_sync(B2_FULL, b.lx, d), __shfl_down_sync(B2_FULL, b.ly, d), __shfl_down_sync(B2_FULL, b.lz, d),
             __shfl_down_sync(B2_FULL, b.hx, d), __shfl_down_sync(B2_FULL, b.hy, d), __shfl_down_sync(B2_FULL, b.hz, d)};
}

/* One plocMerge call, executed by a full warp; lane == list slot. */
__device__ void hploc_merge_warp(u32 L, u32 R, u32 split, bool fin, u32 n, b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* __restrict__ leaves,
                                 u32* nodeIdx, u32* freeIdx, u32* ctrl) {
  const u32 lane = lane_id();
  const u32 nInt = n - 1;
  /* ---- loadIndices: <=16 raw entries of the left part, then <=16 of the right part behind the left part's valid ones ---- */
  const u32 cL = min(split - L, 16u), cR = min(R + 1 - split, 16u);
  u32 a = B2_INVALID, af = B2_INVALID;
  if (lane < cL) { a = __ldcg(nodeIdx + L + lane); af = __ldcg(freeIdx + L + lane); }
  const u32 nLeft = __popc(__ballot_sync(B2_FULL, a != B2_INVALID));
  u32 cl = (lane < nLeft) ? a : B2_INVALID, fr = (lane < nLeft) ? af : B2_INVALID;
  if (lane >= nLeft && lane < nLeft + cR) { cl = __ldcg(nodeIdx + split + (lane - nLeft)); fr = __ldcg(freeIdx + split + (lane - nLeft)); }
  u32 np = __popc(__ballot_sync(B2_FULL, cl != B2_INVALID));
  const u32 stored = np;
  const u32 threshold = fin ? 1u : 16u;
  Box box = box_empty();
  if (cl != B2_INVALID) {
    if (cl >= nInt) {
      const float* f = reinterpret_cast<const float*>(leaves + (cl - nInt)) + 1;
      box = Box{__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5)};
    } else {
      box = load_node2_cg(nodes + cl).box;
    }
  }
  while (np > threshold) {
    /* ---- nearest neighbour inside the list: pairs (l, l+r), r = 1..8, evaluated once and offered to both lanes ---- */
    u64 nn = ~0ull;
#pragma unroll
    for (int r = 1; r <= HP_R; r++) {
      const Box o = shfl_down_box(box, r);
      const bool ok = lane + r < np; /* lane + r < 32 follows */
      const u32 ar = ok ? __float_as_uint(box_area(box_union(o, box))) : 0xFFFFFFFFu;
      if (ok) { const u64 k = ((u64)ar << 32) | (lane + r); nn = k < nn ? k : nn; }
      const u32 up = __shfl_up_sync(B2_FULL, ar, r);
      if (lane >= (u32)r && lane < np) { const u64 k = ((u64)up << 32) | (lane - r); nn = k < nn ? k : nn; }
    }
    const u32 p = (u32)nn & 31u;
    const u32 nnOfP = __shfl_sync(B2_FULL, (u32)nn, p);
    const bool valid = lane < np;
    const bool mutual = valid && nnOfP == lane;
    const bool keep = mutual && lane < p, removed = mutual && lane > p;
    /* ---- mergeClusters: the lower lane keeps the slot, the new node takes the partner's free index ---- */
    const u32 pcl = __shfl_sync(B2_FULL, cl, p), pfr = __shfl_sync(B2_FULL, fr, p);
    const Box pbox = shfl_box(box, p);
    if (keep) {
      box = box_union(box, pbox);
      store_node2(nodes + pfr, cl, pcl, box);
      if (cl == 0u || pcl == 0u) { ctrl[1] = pfr; ctrl[2] = (cl == 0u) ? 0u : 1u; } /* remembered for the final root exchange */
      cl = pfr;
    }
    /* ---- order-preserving compaction: lane d takes the d-th surviving slot ---- */
    const u32 survive = __ballot_sync(B2_FULL, valid && !removed);
    const u32 cnt = __popc(survive);
    const int src = (lane < cnt) ? (int)__fns(survive, 0, lane + 1) : 0;
    const u32 ncl = __shfl_sync(B2_FULL, cl, src), nfr = __shfl_sync(B2_FULL, fr, src);
    const Box nbox = shfl_box(box, src);
    if (lane < cnt) { cl = ncl; fr = nfr; box = nbox; } else { cl = B2_INVALID; fr = B2_INVALID; box = box_empty(); }
    np = cnt;
  }
  /* ---- storeIndices ---- */
  if (lane < stored) { __stcg(nodeIdx + L + lane, cl); __stcg(freeIdx + L + lane, fr); }
  if (lane == 0) atomicAdd(ctrl, 1u);
  if (fin) {
    /* root -> index 0: exchange with the node that was stored at index 0 and re-point that node's parent */
    __threadfence();
    __syncwarp();
    if (lane == 0 && cl != 0u) {
      const u32 root = cl;
      Node2 rn = load_node2_cg(nodes + root);
      const Node2 zn = load_node2_cg(nodes + 0);
      const u32 zp = __ldcg(ctrl + 1), zs = __ldcg(ctrl + 2);
      if (zp == root) { if (zs == 0u) rn.left = root; else rn.right = root; }
      else { u32* c = reinterpret_cast<u32*>(nodes + zp) + zs; *c = root; }
      store_node2(nodes + root, zn.left, zn.right, zn.box);
      store_node2(nodes + 0, rn.left, rn.right, rn.box);
    }
  }
}

/* findParent's comparison (HplocKernel.h:58-81): is the boundary right of leaf R deeper than the boundary left of leaf L? */
__device__ __forceinline__ bool hp_right_deeper(const u32* __restrict__ keys, u32 L, u32 R) {
  const u64 kR = ((u64)__ldg(keys + R) << 32) | R, kR1 = ((u64)__ldg(keys + R + 1) << 32) | (R + 1);
  const u64 kL = ((u64)__ldg(keys + L) << 32) | L, kL1 = ((u64)__ldg(keys + L - 1) << 32) | (L - 1);
  return (kR ^ kR1) < (kL1 ^ kL);
}
/* 64-bit keys (60-bit Morton variant): 96-bit augmented keys — key XORs first, index XORs on a tie */
__device__ __forceinline__ bool hp_right_deeper(const u64* __restrict__ keys, u32 L, u32 R) {
  const u64 xr = __ldg(keys + R) ^ __ldg(keys + R + 1), xl = __ldg(keys + L - 1) ^ __ldg(keys + L);
  if (xr != xl) return xr < xl;
  return (R ^ (R + 1)) < ((L - 1) ^ L);
}

template <typename K>
__global__ void __launch_bounds__(HP_THREADS) hploc_kernel(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes,
                                                          const b2bvh_prim_ref* __restrict__ leaves, u32* nodeIdx, u32* freeIdx, u32* meet, u32* ctrl) {
  const u32 i = blockIdx.x * HP_THREADS + threadIdx.x;
  bool active = i < n;
  u32 L = i, R = i;
  while (__any_sync(B2_FULL, active)) {
    u32 split = 0;
    bool fin = false, wantMerge = false;
    if (active) {
      /* findParent: this range becomes the LEFT child of split R, or the RIGHT child of split L-1 */
      bool isLeft;
      if (L == 0) isLeft = true;
      else if (R == n - 1) isLeft = false;
      else isLeft = hp_right_deeper(keys, L, R);
      const u32 parent = isLeft ? R : L - 1;
      /* Relaxed exchange + fences where they are needed: what a lane hands over was either written by an earlier launch
       * (leaves, initial lists) or is covered by the fence every lane executes after a merge call (below), so the first
       * arriver needs no fence of its own; the second arriver fences once — acquire for what it is about to read and, by
       * cumulativity, release for what it passes on at its next exchange.  (acq_rel on the atomic itself puts a fence on
       * both sides of every one of the 2n exchanges: 28 % of the kernel's stall samples, profiles/r01m.) */
      const u32 other = atom_exch_relaxed(meet + parent, isLeft ? L : R);
      if (other == B2_INVALID) active = false; /* first arriver stops; the sibling's lane continues */
      else {
        __threadfence();
        if (isLeft) { split = R + 1; R = other; } else { split = L; L = other; }
        fin = (L == 0 && R == n - 1);
        wantMerge = (R - L + 1 > 16u) || fin;
      }
    }
    u32 todo = __ballot_sync(B2_FULL, wantMerge);
    const bool merged = todo != 0u;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const u32 mL = __shfl_sync(B2_FULL, L, src), mR = __shfl_sync(B2_FULL, R, src), mS = __shfl_sync(B2_FULL, split, src);
      const bool mF = __shfl_sync(B2_FULL, fin ? 1 : 0, src) != 0;
      hploc_merge_warp(mL, mR, mS, mF, n, nodes, leaves, nodeIdx, freeIdx, ctrl);
#ifndef HP_FENCE_ONCE
      /* make this call's stores (done by all lanes) visible before lane `src` publishes the range further up */
      __threadfence();
#endif
      __syncwarp();
    }
#ifdef HP_FENCE_ONCE
    /* the merge calls of one round work on disjoint ranges and nothing is published before the next exchange: ONE fence makes the stores of
     * all of them (done by all lanes) visible before the lanes hand their ranges further up */
    if (merged) __threadfence();
    __syncwarp();
#else
    (void)merged;
#endif
    if (fin) active = false;
  }
}

int b2_launch_hploc(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const u32* d_sortedKeys, const u32* d_sortedVals, u32 n,
                    b2bvh_bvh2_node* d_nodes, b2bvh_prim_ref* d_leaves, void* d_scratch, u32* h_mergeCalls) {
  return b2_launch_hploc_keys(ctx, d_triAabb, d_sortedKeys, nullptr, d_sortedVals, n, d_nodes, d_leaves, d_scratch, h_mergeCalls);
}

/* d_sortedKeys64 != NULL: the walk compares 64-bit codes (60-bit Morton variant) */
int b2_launch_hploc_keys(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const u32* d_sortedKeys, const u64* d_sortedKeys64, const u32* d_sortedVals, u32 n,
                         b2bvh_bvh2_node* d_nodes, b2bvh_prim_ref* d_leaves, void* d_scratch, u32* h_mergeCalls) {
  u32* ctrl = reinterpret_cast<u32*>(d_scratch);
  u32* nodeIdx = ctrl + 64;
  u32* freeIdx = nodeIdx + n;
  u32* meet = freeIdx + n;
  B2_KERNEL(ctx, "hploc_setup");
  hploc_setup_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_triAabb, d_sortedVals, n, d_leaves, nodeIdx, freeIdx, meet, ctrl);
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "hploc");
  if (d_sortedKeys64)
    hploc_kernel<u64><<<(n + HP_THREADS - 1) / HP_THREADS, HP_THREADS, 0, ctx->stream>>>(d_sortedKeys64, n, d_nodes, d_leaves, nodeIdx, freeIdx, meet, ctrl);
  else
    hploc_kernel<u32><<<(n + HP_THREADS - 1) / HP_THREADS, HP_THREADS, 0, ctx->stream>>>(d_sortedKeys, n, d_nodes, d_leaves, nodeIdx, freeIdx, meet, ctrl);
  B2_LAUNCH_CHECK(ctx);
  (void)h_mergeCalls; /* ctrl[0] travels through the mailbox: b2_mailbox(ctx, B2_MB_HPLOC)[0] after the build's final synchronisation */
  B2_TRY(b2_fetch_words(ctx, ctrl, 1, B2_MB_HPLOC));
  return 0;
}
