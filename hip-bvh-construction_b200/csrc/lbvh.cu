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
__device__ __forceinline__ bool right_boundary_deeper(const u32* __restrict__ keys, u32 lo, u32 hi) {
  const u64 xr = aug_key(keys, hi - 1) ^ aug_key(keys, hi);
  const u64 xl = aug_key(keys, lo - 1) ^ aug_key(keys, lo);
  return xr < xl;
}
/* 64-bit keys (60-bit Morton variant): the augmented key is 96 bits wide — compare the key XORs, then the index XORs */
__device__ __forceinline__ bool right_boundary_deeper(const u64* __restrict__ keys, u32 lo, u32 hi) {
  const u64 xr = __ldg(keys + hi - 1) ^ __ldg(keys + hi), xl = __ldg(keys + lo - 1) ^ __ldg(keys + lo);
  if (xr != xl) return xr < xl;
  return ((hi - 1) ^ hi) < ((lo - 1) ^ lo);
}
template <typename K>
__device__ __forceinline__ u32 choose_parent(const K* __restrict__ keys, u32 n, u32 lo, u32 hi, bool& isLeft) {
  if (lo == 0) { isLeft = true; return hi - 1; }
  if (hi == n) { isLeft = false; return lo - 1; }
  isLeft = right_boundary_deeper(keys, lo, hi);
  return isLeft ? hi - 1 : lo - 1;
}

/* The climb through GLOBAL memory, starting from a finished node `self` covering [lo, hi) with box `box` whose parent has
 * split `p` (this node being its left child iff isLeft).  Returns when the node is the first to arrive at some parent
 * (the sibling's thread takes over) or when the root has been written. */
template <bool KARRAS, typename K>
__device__ __forceinline__ void climb_global(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, u32* parents, u64* meet, u32* rootOut,
                                             u32 self, u32 lo, u32 hi, Box box, u32 p, bool isLeft) {
  const u32 nInt = n - 1;
  while (true) {
    /* one 64-bit exchange hands the sibling both what it cannot derive: the range bound and (Apetrei numbering) this node's index */
    const u64 other = atom_exch_acq_rel64(meet + p, ((u64)self << 32) | (u64)(isLeft ? lo : hi));
    if (other == ~0ull) return; /* first arriver: the sibling's thread finishes this node */
    st_relaxed64(meet + p, ~0ull); /* both children have been here: the word is as the next build expects it (b2_meet_acquire) */
    /* second arriver: the node with split p now has its full range */
    if (isLeft) hi = (u32)other; else lo = (u32)other;
    u32 sib;
    if (KARRAS) {
      /* derivable: left child = p (leaf: p + nInt), right child = p + 1 (leaf: p + 1 + nInt) */
      if (isLeft) sib = (p + 2 == hi) ? p + 1 + nInt : p + 1;
      else sib = (p == lo) ? p + nInt : p;
    } else {
      sib = (u32)(other >> 32);
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

template <bool KARRAS, typename K>
__global__ void __launch_bounds__(LBVH_THREADS) lbvh_fused_kernel(const K* __restrict__ keys, const u32* __restrict__ vals,
                                                                  const b2bvh_aabb* __restrict__ triAabb, u32 n, b2bvh_bvh2_node* nodes,
                                                                  u32* parents, u64* meet /* n-1 words, all ones */, u32* rootOut,
                                                                  const u32* __restrict__ refPrim /* early split: triangle of each reference, else NULL */,
                                                                  u32* __restrict__ leafPrim /* early split: triangle of leaf g, dense (for the collapse) */) {
  const u32 g = blockIdx.x * LBVH_THREADS + threadIdx.x;
  if (g >= n) return;
  const u32 nInt = n - 1;
  const u32 prim = __ldg(vals + g);
  const Box box = load_aabb(triAabb + prim);
  u32 leafId = prim;
  if (refPrim) { leafId = ldg_gather_u32(refPrim + prim); leafPrim[g] = leafId; }
  store_node2(nodes + nInt + g, leafId, B2_INVALID, box);
  if (n == 1) { if (rootOut) *rootOut = 0; return; }
  bool isLeft;
  const u32 p = choose_parent(keys, n, g, g + 1, isLeft);
  climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, nInt + g, g, g + 1, box, p, isLeft);
}

/* ---------------------------------------------------------------- CTA-local tile build in shared memory, then the global climb
 * A CTA owns TILE consecutive leaves [b0, b1) and keeps an ordered list of CLUSTERS (finished subtrees): range start, root
 * node index and the depth d of the boundary to the right neighbour (d = length of the common prefix of the two augmented
 * keys that meet there; -1 outside the key array).  A cluster picks the deeper of its two boundaries as its parent's split
 * (findParent, SinglePassLbvhKernel.h:64-86), so two neighbours form a node exactly when the boundary between them is
 * deeper than both boundaries next to it — a local maximum of d.  All such pairs merge in the same ROUND (no two
 * adjacent boundaries can both be maxima), the list is compacted with ballots and one shared-memory scan, and the next
 * round starts: about a third of the clusters disappear per round, every lane of a live warp holds a live cluster, and no
 * atomics or fences are involved.  (The former version climbed leaf by leaf with one shared-memory exchange per node:
 * 32 warp-instructions per leaf, half of them in warps with one live lane, profiles/r01c_ncu_summary.txt.)
 * Finished nodes are staged in shared memory — Apetrei and Karras indices of nodes finished inside a tile both lie in
 * [b0, b1) — and leave as whole 32-byte sectors: half-written sectors make the B200 L2 read the other half from DRAM.
 * Clusters left when no boundary inside the tile is a maximum (their parents straddle the tile) are handed to
 * lbvh_climb_kernel, which finishes the top of the tree with the global exchange protocol above. */
#define LBVH_TILE 256
#define LBVH_TILE_THREADS 256
#ifndef LBVH_EARLY_STOP
#define LBVH_EARLY_STOP 16   /* > 0: a tile whose list fits its slot stops merging once a round makes at most this many merges (second merge level only) */
#endif
#define LBVH_GROUP 32        /* tiles whose left-over clusters are merged further by one CTA of lbvh_group_kernel */
#define LBVH_TILE_CAP 32     /* left-over clusters a tile may park in its slot (typically ~14; at most 124: two monotone depth runs) */
#define LBVH_GROUP_CAP (LBVH_GROUP * LBVH_TILE_CAP)
#define LBVH_OVERFLOW 0xFFFFFFFFu
#define LBVH_PAR_UNSET 0xFFFFFFFEu
/* one cluster = one 32-bit word: [31:21] d + 1 (0 = outside the key array; at most 64 with 32-bit keys, 96 with 64-bit keys), [20:10] local index of the root node in the
 * staging buffer, [9:0] range start - b0 */
#define LW_D_SHIFT 21
#define LW_ID_SHIFT 10
#define LW_ID_MASK 0x7FFu
#define LW_LO_MASK 0x3FFu

/* A finished node that must continue through global memory: 48 bytes, written by lbvh_tile_kernel, consumed by lbvh_climb_kernel */
struct LbvhPending {
  u32 self, lo, hi, pSide; /* pSide = parent split | (isLeft << 31) */
  float box[6];
  u32 depthRight;          /* d + 1 of the boundary right of the cluster (lbvh_group_kernel) */
  u32 pad;
};
static_assert(sizeof(LbvhPending) == 48, "LbvhPending layout");

template <bool PARENTS>
struct LbvhTileSmem {
  uint4 stage[4 * LBVH_TILE];     /* node with local index i at stage[2i], stage[2i+1]; internal nodes [0,TILE), leaves [TILE,2*TILE) */
  u32 w[2][LBVH_TILE + 4];        /* w[b][j+2] = cluster j; w[b][1] = left sentinel (d of the boundary left of the tile);
                                     w[b][count+2] = right sentinel (range start = b1 - b0) */
  alignas(8) unsigned char chunkMerges[16]; /* merges per warp (32 clusters) */
  u32 finalCur, finalCount, pendBase;
  u32 par[PARENTS ? 2 * LBVH_TILE : 1];
};

__device__ __forceinline__ int boundary_depth(u32 keyA, u32 keyB, u32 a /* b = a + 1 */) {
  return __clzll((long long)(((u64)(keyA ^ keyB) << 32) | (u64)(a ^ (a + 1u))));
}
/* 64-bit keys (60-bit Morton variant): common prefix of the 96-bit augmented keys, 0..95 — still fits the depth field of a cluster word */
__device__ __forceinline__ int boundary_depth(u64 keyA, u64 keyB, u32 a) {
  const u64 kx = keyA ^ keyB;
  return kx ? __clzll((long long)kx) : 64 + __clz((int)(a ^ (a + 1u)));
}

template <bool KARRAS, typename K>
__global__ void __launch_bounds__(LBVH_TILE_THREADS, 8) lbvh_tile_kernel(const K* __restrict__ keys, const u32* __restrict__ vals,
                                                                      const b2bvh_aabb* __restrict__ triAabb, u32 n, b2bvh_bvh2_node* nodes, u32* parents,
                                                                      u64* meet, u32* rootOut, u32* pendingCount, LbvhPending* pending, u32 pendingCap,
                                                                      uint2* tileInfo, LbvhPending* tileBuf, const u32* __restrict__ refPrim, u32* __restrict__ leafPrim) {
  constexpr bool PARENTS = KARRAS; /* only TwoPassLbvh publishes d_parentIdxs */
  constexpr u32 T = LBVH_TILE;
  static_assert(LBVH_TILE == LBVH_TILE_THREADS, "one leaf per thread");
  extern __shared__ __align__(16) unsigned char smemRaw[];
  LbvhTileSmem<PARENTS>& S = *reinterpret_cast<LbvhTileSmem<PARENTS>*>(smemRaw);
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const u32 b0 = blockIdx.x * T, b1 = min(n, b0 + T), cnt0 = b1 - b0;
  const u32 nInt = n - 1;
  const bool single = gridDim.x == 1;
  auto globalId = [&](u32 local) -> u32 { return local < T ? b0 + local : nInt + b0 + (local - T); };

  /* ---- leaves: sorted value -> primitive box (gather), leaf node staged, boundary depths ---- */
  S.stage[2 * tid].x = B2_INVALID; /* internal node tid: not finished */
  if (PARENTS) { S.par[tid] = LBVH_PAR_UNSET; S.par[T + tid] = LBVH_PAR_UNSET; }
  if (tid < cnt0) {
    const u32 g = b0 + tid;
    const u32 prim = __ldg(vals + g);
    const K k0 = __ldg(keys + g);
    const float2* bp = reinterpret_cast<const float2*>(triAabb + prim); /* 24-byte boxes: 8-byte aligned */
    const float2 q0 = ldg_gather_f2(bp), q1 = ldg_gather_f2(bp + 1), q2 = ldg_gather_f2(bp + 2);
    int d = -1;
    if (g + 1 < n) d = boundary_depth(k0, __ldg(keys + g + 1), g);
    /* the leaf names the primitive; over early-split references that is the reference's triangle (InitBvhNodesPrimRef, TwoPassLbvhKernel.h:182) */
    u32 leafId = prim;
    if (refPrim) { leafId = ldg_gather_u32(refPrim + prim); leafPrim[g] = leafId; }
    S.stage[2 * (T + tid)] = make_uint4(leafId, B2_INVALID, __float_as_uint(q0.x), __float_as_uint(q0.y));
    S.stage[2 * (T + tid) + 1] = make_uint4(__float_as_uint(q1.x), __float_as_uint(q1.y), __float_as_uint(q2.x), __float_as_uint(q2.y));
    S.w[0][tid + 2] = ((u32)(d + 1) << LW_D_SHIFT) | ((T + tid) << LW_ID_SHIFT) | tid;
  }
  if (tid == 0) {
    const int dl = b0 > 0 ? boundary_depth(__ldg(keys + b0 - 1), __ldg(keys + b0), b0 - 1) : -1;
    S.w[0][1] = (u32)(dl + 1) << LW_D_SHIFT;
    S.w[0][cnt0 + 2] = cnt0;
  }
  __syncthreads();
  {
    /* two 16-byte pieces per leaf, 512 threads: exactly two steps (a strided loop here is unrolled into 150 instructions) */
    uint4* out = reinterpret_cast<uint4*>(nodes + nInt + b0);
    if (tid < 2 * cnt0) out[tid] = S.stage[2 * T + tid];
    if (tid + LBVH_TILE_THREADS < 2 * cnt0) out[tid + LBVH_TILE_THREADS] = S.stage[2 * T + tid + LBVH_TILE_THREADS];
  }
  if (n == 1) { if (tid == 0) { if (rootOut) *rootOut = 0; if (tileBuf) tileInfo[0] = make_uint2(0u, 0u); } return; }

  /* ---- rounds: only the warps that still hold clusters take part (named barrier over nW warps); the others wait below ---- */
  u32 cur = 0, count = cnt0;
  bool rootDone = false;
  while (true) {
    const u32 nW = (count + 31u) >> 5;
    if (warp >= nW) break; /* the list only shrinks: this warp is done for good */
    const u32* W = S.w[cur];
    u32 xm1 = 0, x0 = 0, xp1 = 0;
    bool mrg = false, absorbed = false;
    if (tid < count) {
      const u32 xm2 = W[tid];
      xm1 = W[tid + 1]; x0 = W[tid + 2]; xp1 = W[tid + 3];
      const u32 dLL = xm2 >> LW_D_SHIFT, dL = xm1 >> LW_D_SHIFT, d0 = x0 >> LW_D_SHIFT, dR = xp1 >> LW_D_SHIFT;
      mrg = (tid + 1 < count) && d0 > dL && d0 > dR;   /* boundary tid is deeper than both boundaries next to it */
      absorbed = (tid >= 1) && dL > dLL && dL > d0;    /* ... and so is boundary tid-1: this cluster joins its left neighbour */
    }
    const u32 bal = __ballot_sync(B2_FULL, mrg);
    if (lane == 0) S.chunkMerges[warp] = (unsigned char)__popc(bal);
    named_barrier(1, nW * 32u);
    /* exclusive prefix of the per-warp merge counts: byte-wise prefix sums by one multiplication (sums <= 128 per half) */
    u32 before, total;
    {
      const u64 ones = 0x0101010101010101ull;
      u64 lo8 = *reinterpret_cast<const u64*>(S.chunkMerges);
      if (nW < 8u) lo8 &= (1ull << (8u * nW)) - 1ull;
      const u64 preLo = lo8 * ones;
      const u32 totLo = (u32)(preLo >> 56);
      before = warp == 0 ? 0u : (u32)(preLo >> (8u * ((warp - 1u) & 7u))) & 0xFFu;
      total = totLo;
      if (nW > 8u) {
        u64 hi8 = *reinterpret_cast<const u64*>(S.chunkMerges + 8);
        if (nW < 16u) hi8 &= (1ull << (8u * (nW - 8u))) - 1ull;
        const u64 preHi = hi8 * ones;
        total += (u32)(preHi >> 56);
        if (warp >= 8u) before = totLo + (warp == 8u ? 0u : (u32)(preHi >> (8u * (warp - 9u))) & 0xFFu);
      }
    }
    if (total == 0) break;
    u32* Wn = S.w[cur ^ 1u];
    if (tid < count && !absorbed) {
      const u32 k = tid - before - __popc(bal & lanemask_lt());
      u32 outw = x0;
      if (mrg) {
        const u32 lL = (x0 >> LW_ID_SHIFT) & LW_ID_MASK, lR = (xp1 >> LW_ID_SHIFT) & LW_ID_MASK;
        const uint4 a0 = S.stage[2 * lL], a1 = S.stage[2 * lL + 1], c0 = S.stage[2 * lR], c1 = S.stage[2 * lR + 1];
        const Box box = box_union(Box{__uint_as_float(a0.z), __uint_as_float(a0.w), __uint_as_float(a1.x), __uint_as_float(a1.y), __uint_as_float(a1.z), __uint_as_float(a1.w)},
                                  Box{__uint_as_float(c0.z), __uint_as_float(c0.w), __uint_as_float(c1.x), __uint_as_float(c1.y), __uint_as_float(c1.z), __uint_as_float(c1.w)});
        const bool isRoot = single && count == 2u;
        u32 idLocal;
        if (KARRAS) {
          /* the merged cluster's own choice (its boundaries are tid-1 and tid+1) fixes its Karras index */
          const u32 hiRel = W[tid + 4] & LW_LO_MASK;
          idLocal = isRoot ? 0u : ((xp1 >> LW_D_SHIFT) > (xm1 >> LW_D_SHIFT) ? hiRel - 1u : (x0 & LW_LO_MASK));
        } else {
          idLocal = (xp1 & LW_LO_MASK) - 1u; /* Apetrei: the split position */
        }
        S.stage[2 * idLocal] = node2_lo(globalId(lL), globalId(lR), box);
        S.stage[2 * idLocal + 1] = node2_hi(box);
        if (PARENTS) { S.par[lL] = b0 + idLocal; S.par[lR] = b0 + idLocal; if (isRoot) S.par[idLocal] = B2_INVALID; }
        if (isRoot) { rootDone = true; if (rootOut) *rootOut = b0 + idLocal; }
        outw = (xp1 & ~((1u << LW_D_SHIFT) - 1u)) | (idLocal << LW_ID_SHIFT) | (x0 & LW_LO_MASK);
      }
      Wn[k + 2] = outw;
    }
    if (tid == 0) { Wn[1] = W[1]; Wn[count - total + 2] = W[count + 2]; }
    named_barrier(1, nW * 32u);
    cur ^= 1u;
    count -= total;
#if LBVH_EARLY_STOP > 0
    /* the tail of a tile is two monotone depth runs that merge one or two pairs per round with ONE warp at work while the other seven wait
     * at the barrier below (35 % of the kernel's stall samples, profiles/r02k_ncu_source_hotspots.txt).  With a second merge level the rest is
     * cheaper there: 32 tiles' left-overs merge side by side in one CTA of lbvh_group_kernel.  (count and total are the same in every
     * participating warp.) */
    if (tileBuf != nullptr && count <= LBVH_TILE_CAP && total <= LBVH_EARLY_STOP) break;
#endif
  }
  if (tid == 0) { S.finalCur = cur; S.finalCount = count; } /* warp 0 takes part in every round */
  rootDone = __syncthreads_or(rootDone);
  cur = S.finalCur; count = S.finalCount;

  /* ---- finished nodes (and parent indices) leave as whole sectors ---- */
  {
    uint4* out = reinterpret_cast<uint4*>(nodes + b0);
#pragma unroll
    for (u32 i = 0; i < 2; i++) {
      const u32 q = tid + i * LBVH_TILE_THREADS;
      if (q < 2 * cnt0 && S.stage[q & ~1u].x != B2_INVALID) out[q] = S.stage[q];
    }
    if (PARENTS && parents) {
      if (tid < cnt0) {
        const u32 a = S.par[tid], b = S.par[T + tid];
        if (a != LBVH_PAR_UNSET) parents[b0 + tid] = a;
        if (b != LBVH_PAR_UNSET) parents[nInt + b0 + tid] = b;
      }
    }
  }
  if (rootDone) { if (tid == 0 && tileBuf) tileInfo[blockIdx.x] = make_uint2(0u, 0u); return; } /* single tile: the whole tree was built here */

  /* ---- park the remaining clusters in the tile's slot for lbvh_group_kernel (or, if they do not fit, hand them to the global
   * climb directly) ---- */
  const bool parked = tileBuf != nullptr && count <= LBVH_TILE_CAP; /* tileBuf == nullptr: small input, no second level */
  if (tid == 0) {
    if (tileBuf) tileInfo[blockIdx.x] = make_uint2(parked ? count : LBVH_OVERFLOW, S.w[cur][1] >> LW_D_SHIFT);
    if (!parked) S.pendBase = atomicAdd(pendingCount, count);
  }
  __syncthreads();
  if (tid < count) {
    const u32* W = S.w[cur];
    const u32 xm1 = W[tid + 1], x0 = W[tid + 2], xp1 = W[tid + 3];
    const u32 l = (x0 >> LW_ID_SHIFT) & LW_ID_MASK;
    const u32 self = globalId(l), lo = b0 + (x0 & LW_LO_MASK), hi = b0 + (xp1 & LW_LO_MASK);
    const bool isLeft = (x0 >> LW_D_SHIFT) > (xm1 >> LW_D_SHIFT);
    const u32 p = isLeft ? hi - 1 : lo - 1;
    const Node2 me = node2_unpack(S.stage[2 * l], S.stage[2 * l + 1]);
    const u32 slotOut = parked ? blockIdx.x * LBVH_TILE_CAP + tid : S.pendBase + tid;
    if (parked || slotOut < pendingCap) {
      uint4* q = reinterpret_cast<uint4*>((parked ? tileBuf : pending) + slotOut);
      q[0] = make_uint4(self, lo, hi, p | (isLeft ? 0x80000000u : 0u));
      q[1] = make_uint4(__float_as_uint(me.box.lx), __float_as_uint(me.box.ly), __float_as_uint(me.box.lz), __float_as_uint(me.box.hx));
      q[2] = make_uint4(__float_as_uint(me.box.hy), __float_as_uint(me.box.hz), x0 >> LW_D_SHIFT, 0u);
    } else {
      climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, self, lo, hi, me.box, p, isLeft); /* overflow of the hand-over list */
    }
  }
}

/* ---------------------------------------------------------------- second level: the left-over clusters of LBVH_GROUP adjacent tiles
 * The clusters a tile could not merge (their parents straddle the tile) are, tile after tile, again an ordered list of
 * clusters with known boundary depths, so the same rounds apply; one CTA takes the ~450 clusters of 32 tiles (8192 leaves)
 * and leaves ~26.  Nodes finished here are stored directly (3 % of all nodes).  What is left goes to lbvh_climb_kernel.  A
 * group that holds a tile whose clusters did not fit its slot is forwarded unchanged. */
struct LbvhGroupSmem {
  u32 w[2][LBVH_GROUP_CAP + 4];   /* [27:21] d + 1, [10:0] slot; same sentinels as in the tile kernel */
  u32 lo[LBVH_GROUP_CAP + 1];     /* per slot: range start; slot LBVH_GROUP_CAP = end of the group's last cluster */
  u32 id[LBVH_GROUP_CAP];         /* per slot: root node */
  float box[LBVH_GROUP_CAP][6];
  u32 tileOff[LBVH_GROUP + 1];
  u32 chunkMerges[LBVH_GROUP_CAP / 32], chunkBefore[LBVH_GROUP_CAP / 32];
  u32 total, pendBase, forward;
};

template <bool KARRAS, typename K>
__global__ void __launch_bounds__(256) lbvh_group_kernel(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, u32* parents, u64* meet, u32* rootOut,
                                                         u32* pendingCount, LbvhPending* pending, u32 pendingCap, const uint2* __restrict__ tileInfo,
                                                         const LbvhPending* __restrict__ tileBuf, u32 nTiles) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  LbvhGroupSmem& S = *reinterpret_cast<LbvhGroupSmem*>(smemRaw);
  constexpr u32 SLOT_MASK = 0x7FFu, P = LBVH_GROUP_CAP / 256;
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const u32 t0 = blockIdx.x * LBVH_GROUP, t1 = min(nTiles, t0 + LBVH_GROUP);
  pdl_wait(); /* launched programmatically behind the tile kernel */
  if (warp == 0) {
    const u32 t = t0 + lane;
    const u32 c = t < t1 ? __ldg(&tileInfo[t].x) : 0u;
    const bool over = __any_sync(B2_FULL, c == LBVH_OVERFLOW);
    u32 incl = over ? 0u : c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)lane >= o) incl += v;
    }
    S.tileOff[lane] = incl - (over ? 0u : c);
    if (lane == 31) { S.tileOff[32] = incl; S.forward = over ? 1u : 0u; }
  }
  __syncthreads();
  u32 count = S.tileOff[LBVH_GROUP];
  if (S.forward) {
    /* some tile of the group went to the climb on its own: its neighbours cannot merge across it here, send them after it */
    for (u32 t = t0 + warp; t < t1; t += 8) {
      const u32 c = __ldg(&tileInfo[t].x);
      if (c == LBVH_OVERFLOW || c == 0) continue;
      u32 base = 0;
      if (lane == 0) base = atomicAdd(pendingCount, c);
      base = __shfl_sync(B2_FULL, base, 0);
      if (lane < c) {
        const uint4* q = reinterpret_cast<const uint4*>(tileBuf + (size_t)t * LBVH_TILE_CAP + lane);
        const uint4 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2);
        if (base + lane < pendingCap) {
          uint4* o = reinterpret_cast<uint4*>(pending + base + lane);
          o[0] = a; o[1] = b; o[2] = cc;
        } else {
          const Box box = Box{__uint_as_float(b.x), __uint_as_float(b.y), __uint_as_float(b.z), __uint_as_float(b.w), __uint_as_float(cc.x), __uint_as_float(cc.y)};
          climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, a.x, a.y, a.z, box, a.w & 0x7FFFFFFFu, (a.w >> 31) != 0u);
        }
      }
    }
    return;
  }
  if (count == 0) return;
  /* ---- gather the clusters in position order: warp per tile ---- */
  for (u32 t = t0 + warp; t < t1; t += 8) {
    const u32 off = S.tileOff[t - t0], c = S.tileOff[t - t0 + 1] - off;
    if (lane < c) {
      const uint4* q = reinterpret_cast<const uint4*>(tileBuf + (size_t)t * LBVH_TILE_CAP + lane);
      const uint4 a = __ldg(q), b = __ldg(q + 1), cc = __ldg(q + 2);
      const u32 s = off + lane;
      S.lo[s] = a.y; S.id[s] = a.x;
      float* bx = S.box[s];
      bx[0] = __uint_as_float(b.x); bx[1] = __uint_as_float(b.y); bx[2] = __uint_as_float(b.z); bx[3] = __uint_as_float(b.w);
      bx[4] = __uint_as_float(cc.x); bx[5] = __uint_as_float(cc.y);
      S.w[0][s + 2] = (cc.z << LW_D_SHIFT) | s;
      if (s + 1 == count) { S.lo[LBVH_GROUP_CAP] = a.z; S.w[0][count + 2] = LBVH_GROUP_CAP; }
    }
  }
  if (tid == 0) S.w[0][1] = __ldg(&tileInfo[t0].y) << LW_D_SHIFT;
  __syncthreads();

  u32 cur = 0;
  while (true) {
    const u32* W = S.w[cur];
    u32 xm1[P], x0[P], xp1[P], bal[P];
    bool mrg[P], absorbed[P];
#pragma unroll
    for (u32 i = 0; i < P; i++) {
      const u32 j = i * 256 + tid;
      mrg[i] = false; absorbed[i] = false; xm1[i] = x0[i] = xp1[i] = 0;
      if (j < count) {
        const u32 xm2 = W[j];
        xm1[i] = W[j + 1]; x0[i] = W[j + 2]; xp1[i] = W[j + 3];
        const u32 dLL = xm2 >> LW_D_SHIFT, dL = xm1[i] >> LW_D_SHIFT, d0 = x0[i] >> LW_D_SHIFT, dR = xp1[i] >> LW_D_SHIFT;
        mrg[i] = (j + 1 < count) && d0 > dL && d0 > dR;
        absorbed[i] = (j >= 1) && dL > dLL && dL > d0;
      }
      bal[i] = __ballot_sync(B2_FULL, mrg[i]);
      if (lane == 0) S.chunkMerges[i * 8 + warp] = __popc(bal[i]);
    }
    __syncthreads();
    if (warp == 0) {
      const u32 v = S.chunkMerges[lane]; /* chunks past the live clusters hold 0 */
      u32 incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 u = __shfl_up_sync(B2_FULL, incl, o);
        if ((int)lane >= o) incl += u;
      }
      S.chunkBefore[lane] = incl - v;
      if (lane == 31) S.total = incl;
    }
    __syncthreads();
    const u32 total = S.total;
    if (total == 0) break;
    u32* Wn = S.w[cur ^ 1u];
#pragma unroll
    for (u32 i = 0; i < P; i++) {
      const u32 j = i * 256 + tid;
      if (j < count && !absorbed[i]) {
        const u32 k = j - S.chunkBefore[i * 8 + warp] - __popc(bal[i] & lanemask_lt());
        u32 outw = x0[i];
        if (mrg[i]) {
          const u32 sL = x0[i] & SLOT_MASK, sR = xp1[i] & SLOT_MASK;
          float* bl = S.box[sL];
          const float* br = S.box[sR];
          const Box box = box_union(Box{bl[0], bl[1], bl[2], bl[3], bl[4], bl[5]}, Box{br[0], br[1], br[2], br[3], br[4], br[5]});
          const u32 loL = S.lo[sL], mid = S.lo[sR], hi = S.lo[W[j + 4] & SLOT_MASK];
          const bool isRoot = (loL == 0 && hi == n);
          /* the merged cluster's own choice (its boundaries are j-1 and j+1) fixes its Karras index; Apetrei: the split position */
          const u32 id = KARRAS ? (isRoot ? 0u : ((xp1[i] >> LW_D_SHIFT) > (xm1[i] >> LW_D_SHIFT) ? hi - 1u : loL)) : mid - 1u;
          const u32 idL = S.id[sL], idR = S.id[sR];
          store_node2(nodes + id, idL, idR, box);
          if (parents) { parents[idL] = id; parents[idR] = id; if (isRoot) parents[id] = B2_INVALID; }
          if (isRoot && rootOut) *rootOut = id;
          bl[0] = box.lx; bl[1] = box.ly; bl[2] = box.lz; bl[3] = box.hx; bl[4] = box.hy; bl[5] = box.hz;
          S.id[sL] = id;
          outw = (xp1[i] & ~((1u << LW_D_SHIFT) - 1u)) | sL;
        }
        Wn[k + 2] = outw;
      }
    }
    if (tid == 0) { Wn[1] = W[1]; Wn[count - total + 2] = W[count + 2]; }
    __syncthreads();
    cur ^= 1u;
    count -= total;
  }
  /* ---- what is left goes to the global climb ---- */
  const u32* W = S.w[cur];
  if (count == 1 && S.lo[W[2] & SLOT_MASK] == 0 && S.lo[LBVH_GROUP_CAP] == n) return; /* the root was finished here */
  if (tid == 0) S.pendBase = atomicAdd(pendingCount, count);
  __syncthreads();
  for (u32 j = tid; j < count; j += 256) {
    const u32 xm1 = W[j + 1], x0 = W[j + 2], xp1 = W[j + 3];
    const u32 s = x0 & SLOT_MASK;
    const u32 self = S.id[s], lo = S.lo[s], hi = S.lo[xp1 & SLOT_MASK];
    const bool isLeft = (x0 >> LW_D_SHIFT) > (xm1 >> LW_D_SHIFT);
    const u32 p = isLeft ? hi - 1 : lo - 1;
    const float* bx = S.box[s];
    const u32 slotOut = S.pendBase + j;
    if (slotOut < pendingCap) {
      uint4* q = reinterpret_cast<uint4*>(pending + slotOut);
      q[0] = make_uint4(self, lo, hi, p | (isLeft ? 0x80000000u : 0u));
      q[1] = make_uint4(__float_as_uint(bx[0]), __float_as_uint(bx[1]), __float_as_uint(bx[2]), __float_as_uint(bx[3]));
      q[2] = make_uint4(__float_as_uint(bx[4]), __float_as_uint(bx[5]), x0 >> LW_D_SHIFT, 0u);
    } else {
      climb_global<KARRAS>(keys, n, nodes, parents, meet, rootOut, self, lo, hi, Box{bx[0], bx[1], bx[2], bx[3], bx[4], bx[5]}, p, isLeft);
    }
  }
}

template <bool KARRAS, typename K>
__global__ void __launch_bounds__(LBVH_THREADS) lbvh_climb_kernel(const K* __restrict__ keys, u32 n, b2bvh_bvh2_node* nodes, u32* parents, u64* meet,
                                                                  u32* rootOut, const u32* __restrict__ pendingCount,
                                                                  const LbvhPending* __restrict__ pending, u32 pendingCap) {
  pdl_wait(); /* launched programmatically behind the tile / group kernel */
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
                                                                        u32* parents, const u32* __restrict__ refPrim, u32* __restrict__ leafPrim) {
  const u32 g = blockIdx.x * LBVH_THREADS + threadIdx.x;
  if (g >= n) return;
  const u32 nInt = n - 1;
  {
    const u32 prim = __ldg(vals + g);
    u32 leafId = prim;
    if (refPrim) { leafId = ldg_gather_u32(refPrim + prim); leafPrim[g] = leafId; }
    store_node2(nodes + nInt + g, leafId, B2_INVALID, load_aabb(triAabb + prim));
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

int b2_meet_acquire(b2bvh_ctx* ctx, u32 n, u64** out) {
  const size_t words = n > 1 ? (size_t)n - 1 : 1;
  b2bvh_ctx::Buf& b = ctx->bufs[SLOT_MEET];
  const void* pBefore = b.p;
  const size_t capBefore = b.cap;
  void* p = nullptr;
  B2_TRY(b2_reserve(ctx, SLOT_MEET, words * sizeof(u64), &p));
  if (p != pBefore || b.cap != capBefore) ctx->meet_clean = 0;
  if ((size_t)ctx->meet_clean < words) {
    B2_CUDA(cudaMemsetAsync(p, 0xFF, b.cap, ctx->stream));
    ctx->meet_clean = (u32)(b.cap / sizeof(u64)); /* n <= 2^30: fits */
  }
  *out = reinterpret_cast<u64*>(p);
  return 0;
}

size_t b2_lbvh_scratch_bytes(u32 n) {
  /* the larger of: fused path (meet words + hand-over list) and two-kernel path (2n-1 flags) */
  const size_t tiles = ((size_t)n + LBVH_TILE - 1) / LBVH_TILE;
  const size_t fused = (((size_t)n * 4 + 15) & ~(size_t)15) + 16 + ((size_t)n / 8 + 1024) * sizeof(LbvhPending) + ((tiles * sizeof(uint2) + 15) & ~(size_t)15) +
                       tiles * LBVH_TILE_CAP * sizeof(LbvhPending);
  const size_t two = (2 * (size_t)n - 1) * 4;
  return fused > two ? fused : two;
}

template <typename K>
static int launch_lbvh_fused_t(b2bvh_ctx* ctx, const K* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                         b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch, u32* d_root, int karrasNumbering) {
  /* the exchange words are NOT cleared per build (40 MB of writes at 10 M primitives): every word the climb uses sees exactly two arrivals and
   * the second one puts 0xFFFFFFFF back */
  u64* meet = nullptr;
  B2_TRY(b2_meet_acquire(ctx, n, &meet));
  B2_KERNEL(ctx, sizeof(K) == 8 ? (karrasNumbering ? "lbvh_fused64_karras" : "lbvh_fused64_apetrei") : (karrasNumbering ? "lbvh_fused_karras" : "lbvh_fused_apetrei"));
  static const bool globalOnly = getenv("B2BVH_LBVH_GLOBAL_ONLY") != nullptr; /* development switch: the all-global-memory variant */
  if (globalOnly) {
    const u32 grid = (n + LBVH_THREADS - 1) / LBVH_THREADS;
    if (karrasNumbering)
      lbvh_fused_kernel<true, K><<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, meet, d_root, ctx->ref_prim, ctx->ref_leaf_prim);
    else
      lbvh_fused_kernel<false, K><<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, meet, d_root, ctx->ref_prim, ctx->ref_leaf_prim);
  } else {
    /* scratch: meet[n-1] | (16-byte aligned) pendingCount | pending[cap] | tileInfo[tiles] | tileBuf[tiles][LBVH_TILE_CAP] */
    const size_t off = (((size_t)n * 4 + 15) & ~(size_t)15);
    u32* pendingCount = reinterpret_cast<u32*>(reinterpret_cast<unsigned char*>(d_scratch) + off);
    LbvhPending* pending = reinterpret_cast<LbvhPending*>(reinterpret_cast<unsigned char*>(d_scratch) + off + 16);
    const u32 cap = n / 8 + 1024;
    B2_CUDA(cudaMemsetAsync(pendingCount, 0, 4, ctx->stream));
    const u32 grid = (n + LBVH_TILE - 1) / LBVH_TILE;
    /* second level only where it pays (build stage in us without / with it: 1 M 104 / 119, 3 M 203 / 217, 7 M 379 / 399, 10 M 551 / 533,
     * tools/micro/lbvh_level_sweep.py): below ~8 M primitives the extra launch costs more than the shorter climb saves */
    /* (Round 2 tried more merge levels above the groups — 16 groups per CTA, repeated until one CTA finishes the root — in place of the climb:
     * three launches of 23 us each against 67 us of climbing at 10 M, no gain.  The left-over clusters are the two monotone depth runs of
     * every group, and a monotone run merges one pair per round, so a level costs ~40 dependent rounds whatever its size; gpurun r2h.) */
    const bool useGroups = ctx->lbvh_second_level == 1 ? true : (ctx->lbvh_second_level == 2 ? false : n >= (1u << 23));
    uint2* tileInfo = reinterpret_cast<uint2*>(pending + cap);
    LbvhPending* tileBuf = useGroups ? reinterpret_cast<LbvhPending*>(reinterpret_cast<unsigned char*>(tileInfo) + (((size_t)grid * sizeof(uint2) + 15) & ~(size_t)15)) : nullptr;
    const u32 onceBit = sizeof(K) == 8 ? B2_ONCE_LBVH64 : B2_ONCE_LBVH32;
    if (!(ctx->once_mask & onceBit)) {
      B2_CUDA(cudaFuncSetAttribute(lbvh_group_kernel<true, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbvhGroupSmem)));
      B2_CUDA(cudaFuncSetAttribute(lbvh_group_kernel<false, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbvhGroupSmem)));
      B2_CUDA(cudaFuncSetAttribute(lbvh_tile_kernel<true, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbvhTileSmem<true>)));
      B2_CUDA(cudaFuncSetAttribute(lbvh_tile_kernel<false, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbvhTileSmem<false>)));
      ctx->once_mask |= onceBit;
    }
    if (karrasNumbering)
      lbvh_tile_kernel<true, K><<<grid, LBVH_TILE_THREADS, sizeof(LbvhTileSmem<true>), ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, meet, d_root, pendingCount, pending, cap, tileInfo, tileBuf, ctx->ref_prim, ctx->ref_leaf_prim);
    else
      lbvh_tile_kernel<false, K><<<grid, LBVH_TILE_THREADS, sizeof(LbvhTileSmem<false>), ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, meet, d_root, pendingCount, pending, cap, tileInfo, tileBuf, ctx->ref_prim, ctx->ref_leaf_prim);
    B2_LAUNCH_CHECK(ctx);
    if (useGroups) {
      const u32 groups = (grid + LBVH_GROUP - 1) / LBVH_GROUP;
      B2_KERNEL(ctx, "lbvh_group");
      if (karrasNumbering)
        B2_LAUNCH_PDL((lbvh_group_kernel<true, K>), groups, 256, sizeof(LbvhGroupSmem), ctx->stream, d_sortedKeys, n, d_nodes, d_parents, meet, d_root, pendingCount, pending, cap, tileInfo, tileBuf, grid);
      else
        B2_LAUNCH_PDL((lbvh_group_kernel<false, K>), groups, 256, sizeof(LbvhGroupSmem), ctx->stream, d_sortedKeys, n, d_nodes, d_parents, meet, d_root, pendingCount, pending, cap, tileInfo, tileBuf, grid);
      B2_LAUNCH_CHECK(ctx);
    }
    u32 grid2 = (cap + LBVH_THREADS - 1) / LBVH_THREADS;
    const u32 cap2 = (u32)ctx->sm_count * 8u;
    if (grid2 > cap2) grid2 = cap2;
    B2_KERNEL(ctx, "lbvh_climb");
    if (karrasNumbering)
      B2_LAUNCH_PDL((lbvh_climb_kernel<true, K>), grid2, LBVH_THREADS, 0, ctx->stream, d_sortedKeys, n, d_nodes, d_parents, meet, d_root, pendingCount, pending, cap);
    else
      B2_LAUNCH_PDL((lbvh_climb_kernel<false, K>), grid2, LBVH_THREADS, 0, ctx->stream, d_sortedKeys, n, d_nodes, d_parents, meet, d_root, pendingCount, pending, cap);
  }
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2_launch_lbvh_fused(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                         b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch, u32* d_root, int karrasNumbering) {
  return launch_lbvh_fused_t<u32>(ctx, d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root, karrasNumbering);
}
/* 64-bit sorted keys (60-bit Morton variant): the same tile / group / climb kernels instantiated for the wider key — only the boundary
 * depth (common prefix of the 96-bit augmented keys) and the parent choice of the global climb look at key bits */
int b2_launch_lbvh_fused64(b2bvh_ctx* ctx, const u64* d_sortedKeys64, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                           b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch, u32* d_root, int karrasNumbering) {
  return launch_lbvh_fused_t<u64>(ctx, d_sortedKeys64, d_sortedVals, d_triAabb, n, d_nodes, d_parents, d_scratch, d_root, karrasNumbering);
}

int b2_launch_lbvh_karras_two_kernel(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                                     b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_flags) {
  const u32 grid = (n + LBVH_THREADS - 1) / LBVH_THREADS;
  B2_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)(2 * (size_t)n - 1) * sizeof(u32), ctx->stream));
  B2_KERNEL(ctx, "lbvh_karras_emit");
  lbvh_karras_emit_kernel<<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_sortedKeys, d_sortedVals, d_triAabb, n, d_nodes, d_parents, ctx->ref_prim, ctx->ref_leaf_prim);
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "lbvh_refit");
  lbvh_refit_kernel<<<grid, LBVH_THREADS, 0, ctx->stream>>>(d_nodes, d_parents, d_flags, n);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
