/*
 * hploc.cu — stage S7: H-PLOC (Benthin et al. 2024): the whole build in ONE launch; a warp walks the LBVH hierarchy
 * bottom-up and PLOC-merges at most 32 clusters in registers whenever a hierarchy node covers more than 16 leaves.
 * (From 2^20 primitives the launch is hploc_tile_kernel: the same walk, with the part of the hierarchy that lies inside 128
 * consecutive leaves done in shared memory first — see the tile phase below.)
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
#include <stdlib.h>

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

/* One plocMerge call, executed by a full warp; lane == list slot.  SMEM: the lists (cluster id, free index, box per list position) live in the
 * CTA's shared memory — the tile phase of hploc_tile_kernel, L / R / split are then positions inside the tile — instead of in global memory. */
template <bool SMEM>
__device__ void hploc_merge_warp(u32 L, u32 R, u32 split, bool fin, u32 n, b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* __restrict__ leaves,
                                 u32* nodeIdx, u32* freeIdx, u32* ctrl, float (*sBox)[6] = nullptr) {
  const u32 lane = lane_id();
  const u32 nInt = n - 1;
  /* ---- loadIndices: <=16 raw entries of the left part, then <=16 of the right part behind the left part's valid ones ---- */
  const u32 cL = min(split - L, 16u), cR = min(R + 1 - split, 16u);
  u32 a = B2_INVALID, af = B2_INVALID;
  if (lane < cL) { a = SMEM ? nodeIdx[L + lane] : __ldcg(nodeIdx + L + lane); af = SMEM ? freeIdx[L + lane] : __ldcg(freeIdx + L + lane); }
  const u32 nLeft = __popc(__ballot_sync(B2_FULL, a != B2_INVALID));
  u32 cl = (lane < nLeft) ? a : B2_INVALID, fr = (lane < nLeft) ? af : B2_INVALID;
  u32 src = L + lane; /* list position this lane's cluster came from (SMEM: where its box sits) */
  if (lane >= nLeft && lane < nLeft + cR) {
    src = split + (lane - nLeft);
    cl = SMEM ? nodeIdx[src] : __ldcg(nodeIdx + src);
    fr = SMEM ? freeIdx[src] : __ldcg(freeIdx + src);
  }
  u32 np = __popc(__ballot_sync(B2_FULL, cl != B2_INVALID));
  const u32 stored = np;
  const u32 threshold = fin ? 1u : 16u;
  Box box = box_empty();
  if (cl != B2_INVALID) {
    if (SMEM) {
      const float* f = sBox[src];
      box = Box{f[0], f[1], f[2], f[3], f[4], f[5]};
    } else if (cl >= nInt) {
      const float* f = reinterpret_cast<const float*>(leaves + (cl - nInt)) + 1;
      box = Box{__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5)};
    } else {
      box = load_node2_cg(nodes + cl).box;
    }
  }
  if (SMEM) __syncwarp(); /* every lane has read its slot before the compacted list is written back over the same positions */
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
  if (SMEM) {
    if (lane < stored) {
      nodeIdx[L + lane] = cl; freeIdx[L + lane] = fr;
      float* f = sBox[L + lane];
      f[0] = box.lx; f[1] = box.ly; f[2] = box.lz; f[3] = box.hx; f[4] = box.hy; f[5] = box.hz;
    }
    return; /* the tile counts its merge calls itself; the root is never finished inside a tile */
  }
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

/* The walk up the hierarchy through global memory from the range [L, R] this lane holds (active) — whole warps call it; a lane leaves when it is
 * the first to arrive at a parent. */
template <typename K>
__device__ __forceinline__ void hploc_walk(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* __restrict__ leaves, u32* nodeIdx,
                                           u32* freeIdx, u32* meet, u32* ctrl, bool active, u32 L, u32 R) {
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
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const u32 mL = __shfl_sync(B2_FULL, L, src), mR = __shfl_sync(B2_FULL, R, src), mS = __shfl_sync(B2_FULL, split, src);
      const bool mF = __shfl_sync(B2_FULL, fin ? 1 : 0, src) != 0;
      hploc_merge_warp<false>(mL, mR, mS, mF, n, nodes, leaves, nodeIdx, freeIdx, ctrl);
      /* make this call's stores (done by all lanes) visible before lane `src` publishes the range further up */
      __threadfence();
      __syncwarp();
    }
    if (fin) active = false;
  }
}

template <typename K>
__global__ void __launch_bounds__(HP_THREADS) hploc_kernel(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes,
                                                          const b2bvh_prim_ref* __restrict__ leaves, u32* nodeIdx, u32* freeIdx, u32* meet, u32* ctrl) {
  const u32 i = blockIdx.x * HP_THREADS + threadIdx.x;
  hploc_walk(keys, n, nodes, leaves, nodeIdx, freeIdx, meet, ctrl, i < n, i, i);
}

/* ---------------------------------------------------------------- tile phase (round 2)
 * hploc_kernel sends every range of the hierarchy through global memory: an atomic exchange per node, and for every range with more than 16
 * leaves a merge call that loads its lists and boxes from global memory, stores them back and fences — a ~7 us dependent chain per call at
 * ~24 resident warps per SM (2.0 ms for 10 M primitives, 0.10 of the roofline).  Most of those ranges lie inside a few hundred consecutive leaves.
 * hploc_tile_kernel gives a CTA HT_TILE consecutive leaves and finds the ranges of the hierarchy inside them the way lbvh_tile_kernel does —
 * an ordered list of finished ranges, two neighbours form their parent exactly when the boundary between them is deeper than both
 * boundaries next to it, all such pairs in one ROUND — with the cluster lists (id, free index, box per list position) in shared memory.
 * The merge calls of a round work on disjoint positions: the warps of the CTA take them in turn, no atomics, no fences, no global loads.
 * What is left when no boundary inside the tile is a maximum (the ranges whose parents straddle tiles, ~15 per tile) is written back and
 * continues through hploc_walk.  Same merge calls on the same lists in an order the dependencies allow: identical output.
 * Measured (10 M uniform, merge stage): walk only 2.01 ms; tiles of 512 leaves 2.99 ms, 256 2.0 ms, 128 1.76 ms, 64 3.4 ms — a merge call is ~1000
 * dependent warp instructions (48 shuffles per search round) whether its operands come from shared or global memory, so what the tile saves is the
 * global round trips, and what it costs is the idle warps of a CTA whose upper rounds hold one or two calls: small tiles, many CTAs per SM. */
#ifndef HT_TILE
#define HT_TILE 128
#endif
#ifndef HT_MINB
#define HT_MINB 12
#endif
#define HT_D_SHIFT 16
struct HpTileSmem {
  u32 w[2][HT_TILE + 4];       /* range j of the ordered list = w[b][j + 2]: [22:16] depth + 1 of the boundary on its right, [9:0] first leaf - b0; sentinels as in lbvh_tile_kernel */
  u32 nodeIdx[HT_TILE], freeIdx[HT_TILE];
  float box[HT_TILE][6];       /* box of the cluster at list position p */
  uint4 tasks[HT_TILE / 2];    /* merge calls of the round: {L, R, split} (positions inside the tile) */
  u32 warpCount[2][HT_TILE / 32];
  u32 nTasks, finalCur, finalCount, calls;
};

__device__ __forceinline__ int hp_boundary_depth(u32 keyA, u32 keyB, u32 a /* b = a + 1 */) {
  return __clzll((long long)(((u64)(keyA ^ keyB) << 32) | (u64)(a ^ (a + 1u))));
}
__device__ __forceinline__ int hp_boundary_depth(u64 keyA, u64 keyB, u32 a) {
  const u64 kx = keyA ^ keyB;
  return kx ? __clzll((long long)kx) : 64 + __clz((int)(a ^ (a + 1u)));
}

template <typename K>
__global__ void __launch_bounds__(HT_TILE, HT_MINB) hploc_tile_kernel(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* __restrict__ leaves,
                                                                u32* nodeIdx, u32* freeIdx, u32* meet, u32* ctrl) {
  extern __shared__ __align__(16) unsigned char hpRaw[];
  HpTileSmem& S = *reinterpret_cast<HpTileSmem*>(hpRaw);
  constexpr u32 T = HT_TILE, NW = HT_TILE / 32;
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const u32 b0 = blockIdx.x * T, b1 = min(n, b0 + T), cnt0 = b1 - b0, nInt = n - 1;
  /* ---- leaves: initial lists (every leaf is a cluster with free index g - 1), boxes, boundary depths ---- */
  if (tid < cnt0) {
    const u32 g = b0 + tid;
    const K k0 = __ldg(keys + g);
    int d = -1;
    if (g + 1 < n) d = hp_boundary_depth(k0, __ldg(keys + g + 1), g);
    const float* f = reinterpret_cast<const float*>(leaves + g) + 1;
    float* bx = S.box[tid];
#pragma unroll
    for (int k = 0; k < 6; k++) bx[k] = __ldg(f + k);
    S.nodeIdx[tid] = nInt + g;
    S.freeIdx[tid] = g ? g - 1u : B2_INVALID;
    S.w[0][tid + 2] = ((u32)(d + 1) << HT_D_SHIFT) | tid;
  }
  if (tid == 0) {
    const int dl = b0 > 0 ? hp_boundary_depth(__ldg(keys + b0 - 1), __ldg(keys + b0), b0 - 1) : -1;
    S.w[0][1] = (u32)(dl + 1) << HT_D_SHIFT;
    S.w[0][cnt0 + 2] = cnt0;
    S.calls = 0;
  }
  __syncthreads();
  /* ---- rounds ---- */
  u32 cur = 0, count = cnt0;
  while (true) {
    const u32* W = S.w[cur];
    u32 x0 = 0, xp1 = 0;
    bool mrg = false, absorbed = false, task = false;
    u32 tL = 0, tS = 0, tE = 0;
    if (tid < count) {
      const u32 xm2 = W[tid], xm1 = W[tid + 1];
      x0 = W[tid + 2]; xp1 = W[tid + 3];
      const u32 dLL = xm2 >> HT_D_SHIFT, dL = xm1 >> HT_D_SHIFT, d0 = x0 >> HT_D_SHIFT, dR = xp1 >> HT_D_SHIFT;
      mrg = (tid + 1 < count) && d0 > dL && d0 > dR;
      absorbed = (tid >= 1) && dL > dLL && dL > d0;
      if (mrg) {
        tL = x0 & 0xFFFFu; tS = xp1 & 0xFFFFu; tE = W[tid + 4] & 0xFFFFu; /* the parent covers leaves [tL, tE), its right child starts at tS */
        task = tE - tL > 16u;
      }
    }
    const u32 balM = __ballot_sync(B2_FULL, mrg), balT = __ballot_sync(B2_FULL, task);
    if (lane == 0) { S.warpCount[0][warp] = __popc(balM); S.warpCount[1][warp] = __popc(balT); }
    __syncthreads();
    u32 beforeM = 0, totalM = 0, beforeT = 0, totalT = 0;
#pragma unroll
    for (u32 k = 0; k < NW; k++) {
      const u32 cm = S.warpCount[0][k], ct = S.warpCount[1][k];
      if (k < warp) { beforeM += cm; beforeT += ct; }
      totalM += cm; totalT += ct;
    }
    if (totalM == 0) break;
    u32* Wn = S.w[cur ^ 1u];
    if (tid < count && !absorbed) {
      const u32 k = tid - beforeM - __popc(balM & lanemask_lt());
      Wn[k + 2] = mrg ? ((xp1 & ~0xFFFFu) | (x0 & 0xFFFFu)) : x0;
    }
    if (task) S.tasks[beforeT + __popc(balT & lanemask_lt())] = make_uint4(tL, tE - 1u, tS, 0u);
    if (tid == 0) { Wn[1] = W[1]; Wn[count - totalM + 2] = W[count + 2]; S.calls += totalT; }
    __syncthreads();
    /* the merge calls of this round: disjoint positions, one warp each */
    for (u32 t = warp; t < totalT; t += NW) {
      const uint4 q = S.tasks[t];
      hploc_merge_warp<true>(q.x, q.y, q.z, false, n, nodes, leaves, S.nodeIdx, S.freeIdx, ctrl, S.box);
    }
    __syncthreads();
    cur ^= 1u;
    count -= totalM;
  }
  /* ---- lists back to global memory; the ranges that are left continue through the hierarchy above the tile ---- */
  if (tid < cnt0) { __stcg(nodeIdx + b0 + tid, S.nodeIdx[tid]); __stcg(freeIdx + b0 + tid, S.freeIdx[tid]); }
  if (tid == 0 && S.calls) atomicAdd(ctrl, S.calls);
  __threadfence(); /* lists and the nodes written by the merge calls, before any lane publishes a range */
  __syncthreads();
  if (warp * 32u >= count) return; /* whole warps without a range leave */
  const u32* W = S.w[cur];
  const bool active = tid < count;
  const u32 L = active ? b0 + (W[tid + 2] & 0xFFFFu) : 0u, R = active ? b0 + (W[tid + 3] & 0xFFFFu) - 1u : 0u;
  hploc_walk(keys, n, nodes, leaves, nodeIdx, freeIdx, meet, ctrl, active, L, R);
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
  static const bool walkOnly = getenv("B2BVH_HPLOC_WALK_ONLY") != nullptr; /* development switch: the all-global-memory kernel of round 1, identical output */
  /* tile phase from 2^20 primitives (merge stage in ms, tile / walk only, tools/hploc_ab.py: 10 M 1.76 / 2.01, 1 M 0.289 / 0.305, sponza 0.164 / 0.154,
   * bunny 0.138 / 0.117 — a small scene does not fill the machine and pays the tile's barriers); b2bvh_build_opts.lbvh_second_level = 1 / 2 forces it
   * on / off (tests).  Two tiles or more: the root is never finished inside a tile. */
  const bool wantTile = ctx->lbvh_second_level == 1 ? true : (ctx->lbvh_second_level == 2 ? false : n >= (1u << 20));
  if (n > HT_TILE && wantTile && !walkOnly) {
    if (!(ctx->once_mask & B2_ONCE_MISC)) {
      B2_CUDA(cudaFuncSetAttribute(hploc_tile_kernel<u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HpTileSmem)));
      B2_CUDA(cudaFuncSetAttribute(hploc_tile_kernel<u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HpTileSmem)));
      ctx->once_mask |= B2_ONCE_MISC;
    }
    const u32 grid = (n + HT_TILE - 1) / HT_TILE;
    if (d_sortedKeys64) hploc_tile_kernel<u64><<<grid, HT_TILE, sizeof(HpTileSmem), ctx->stream>>>(d_sortedKeys64, n, d_nodes, d_leaves, nodeIdx, freeIdx, meet, ctrl);
    else hploc_tile_kernel<u32><<<grid, HT_TILE, sizeof(HpTileSmem), ctx->stream>>>(d_sortedKeys, n, d_nodes, d_leaves, nodeIdx, freeIdx, meet, ctrl);
  } else if (d_sortedKeys64)
    hploc_kernel<u64><<<(n + HP_THREADS - 1) / HP_THREADS, HP_THREADS, 0, ctx->stream>>>(d_sortedKeys64, n, d_nodes, d_leaves, nodeIdx, freeIdx, meet, ctrl);
  else
    hploc_kernel<u32><<<(n + HP_THREADS - 1) / HP_THREADS, HP_THREADS, 0, ctx->stream>>>(d_sortedKeys, n, d_nodes, d_leaves, nodeIdx, freeIdx, meet, ctrl);
  B2_LAUNCH_CHECK(ctx);
  (void)h_mergeCalls; /* ctrl[0] travels through the mailbox: b2_mailbox(ctx, B2_MB_HPLOC)[0] after the build's final synchronisation */
  B2_TRY(b2_fetch_words(ctx, ctrl, 1, B2_MB_HPLOC));
  return 0;
}
