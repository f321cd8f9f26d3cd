/*
 * ploc.cu — stage S6: PLOC++ nearest-neighbour merging (Benthin et al. 2022) on the Morton-ordered primitive boxes.
 *
 * Replaces SetupClusters (Ploc++Kernel.h:39-55), Ploc (:211-362), SinglePassPloc (:98-209), binaryWarpPrefixSum /
 * binaryBlockPrefixSum (:57-96) and the host loop PLOCNew::build (PLOC++Bvh.cpp:101-152).
 *
 * Semantics kept: per iteration every cluster c looks at the clusters within +-8 positions, picks the
 * neighbour minimising (float bits of area(box_c U box_j) << 32 | j) (area without FMA), mutual pairs merge, the lower
 * index keeps the slot, the list is compacted in order; merged node index = C-2-rank, root ends at index 0.
 * What is different:
 *   - rank is the exclusive prefix count of merging clusters (CTA scan + decoupled look-back between tiles), not the
 *     arrival order of a global atomicAdd (:57-68,:311) -> node numbering is deterministic (canonical numbering of the oracle);
 *   - the same look-back word carries the count of removed clusters, which replaces the reference's serial inter-CTA
 *     chain (:341-347) for the compaction offsets;
 *   - cluster boxes live in a COMPACTED array next to the cluster ids (24 B per live cluster), so a tile's search window
 *     [base-16, base+TILE+16) is one contiguous range that a single TMA bulk copy (cp.async.bulk) stages into shared
 *     memory; the reference gathers boxes through the node index every iteration;
 *   - the cluster count stays on the device: the host launches iterations in batches and reads the count once per batch
 *     instead of a blocking D2H per iteration (PLOC++Bvh.cpp:150); tiles are claimed with tickets, so no co-residency
 *     or launch-order assumptions.
 * Per iteration and live cluster: 28 B read (id + box) + 28 B written (+32 B per merged node).
 */
#include <math.h>

#include "common.cuh"
#include "lookback.cuh"

#define PLOC_R 8
#define PLOC_THREADS 512
#define PLOC_TILE PLOC_THREADS
#define PLOC_HALO (2 * PLOC_R)
#define PLOC_WIN (PLOC_TILE + 2 * PLOC_HALO)
#define PLOC_SOA (PLOC_WIN + 6) /* component stride: 6 mod 32 keeps the AoS->SoA transpose (almost) bank-conflict free */
#define PLOC_TAIL 1024 /* clusters handled by the single-CTA tail kernel */
#define PLOC_TAIL_THREADS 1024

#define PL_FLAG_AGG (1ull << 62)
#define PL_FLAG_INC (2ull << 62)
#define PL_FLAG_MASK (3ull << 62)

struct PlocCtrl {
  u32 count[2];   /* count[it & 1] = live clusters at the start of iteration `it` */
  u32 ticket[2];
  u32 itersRun;   /* iterations that did work */
  u32 liveBuf;    /* which of the two cluster buffers holds the live list */
  u32 pad[2];
};

/* scratch: PlocCtrl (256 B) | ids[2][nPad] | boxes[2][nPad] (24 B each) | status[2][n/TILE + 2] (u64); nPad = n rounded up to 16
 * so that every region (and every tile window inside the box arrays) starts on a 16-byte boundary for the bulk copies */
static size_t ploc_tiles(u32 n) { return (size_t)n / PLOC_TILE + 2; }
static size_t ploc_pad(u32 n) { return ((size_t)n + 15) & ~(size_t)15; }
size_t b2_ploc_scratch_bytes(u32 n) {
  return 256 + 2 * ploc_pad(n) * 4 + 2 * ploc_pad(n) * sizeof(b2bvh_aabb) + 2 * ploc_tiles(n) * sizeof(u64);
}

__global__ void __launch_bounds__(256) ploc_setup_kernel(const b2bvh_aabb* __restrict__ triAabb, const u32* __restrict__ sortedVals, u32 n,
                                                         b2bvh_prim_ref* __restrict__ leaves, u32* __restrict__ ids, b2bvh_aabb* __restrict__ boxes,
                                                         PlocCtrl* ctrl) {
  __shared__ __align__(16) u32 sLeaf[256 * 7];
  __shared__ __align__(16) u32 sBox[256 * 6];
  const u32 t = threadIdx.x, g0 = blockIdx.x * 256, g = g0 + t;
  if (g == 0) { ctrl->count[0] = n; ctrl->count[1] = n; ctrl->ticket[0] = ctrl->ticket[1] = 0; ctrl->itersRun = 0; ctrl->liveBuf = 0; }
  if (g < n) {
    const u32 prim = __ldg(sortedVals + g);
    const float2* bp = reinterpret_cast<const float2*>(triAabb + prim); /* 24-byte boxes: 8-byte aligned */
    const float2 q0 = ldg_gather_f2(bp), q1 = ldg_gather_f2(bp + 1), q2 = ldg_gather_f2(bp + 2);
    u32* l = sLeaf + t * 7;
    l[0] = prim; l[1] = __float_as_uint(q0.x); l[2] = __float_as_uint(q0.y); l[3] = __float_as_uint(q1.x); l[4] = __float_as_uint(q1.y);
    l[5] = __float_as_uint(q2.x); l[6] = __float_as_uint(q2.y);
    u32* bx = sBox + t * 6;
    bx[0] = l[1]; bx[1] = l[2]; bx[2] = l[3]; bx[3] = l[4]; bx[4] = l[5]; bx[5] = l[6];
    ids[g] = g + (n - 1);
  }
  __syncthreads();
  const u32 cnt = min(256u, n - g0);
  cta_store_words(reinterpret_cast<u32*>(leaves + g0), sLeaf, cnt * 7); /* 256 x 28 B = 7168 B per CTA: 16-byte aligned */
  cta_store_words(reinterpret_cast<u32*>(boxes + g0), sBox, cnt * 6);
}

/* ---- TMA bulk copy helpers (same PTX as radix_sort.cu) ---- */
__device__ __forceinline__ u32 pl_smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pl_mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pl_smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void pl_mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pl_mbar_wait(u64* bar, u32 phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PL_DONE;\n"
      "bra PL_WAIT;\n"
      "PL_DONE:\n"
      "}\n" ::"r"(pl_smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void pl_tma_load_1d(void* smemDst, const void* gsrc, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pl_smem_u32(smemDst)), "l"(gsrc),
               "r"(bytes), "r"(pl_smem_u32(bar))
               : "memory");
}

/* nearest neighbour of window slot `s` (cluster index c = winBase + s) among clusters within +-R that exist;
 * boxes are read from the structure-of-arrays window. Returns the winning cluster's window slot, or -1. */
template <int STRIDE>
__device__ __forceinline__ int nearest_in_window(const float* __restrict__ soa, int s, int winBase, u32 count) {
  const Box me = Box{soa[s], soa[STRIDE + s], soa[2 * STRIDE + s], soa[3 * STRIDE + s], soa[4 * STRIDE + s], soa[5 * STRIDE + s]};
  u32 bestArea = 0xFFFFFFFFu;
  int best = -1;
#pragma unroll
  for (int r = -PLOC_R; r <= PLOC_R; r++) {
    if (r == 0) continue;
    const int t = s + r;
    const long long j = (long long)winBase + t;
    if (j < 0 || j >= (long long)count) continue;
    const Box o = Box{soa[t], soa[STRIDE + t], soa[2 * STRIDE + t], soa[3 * STRIDE + t], soa[4 * STRIDE + t], soa[5 * STRIDE + t]};
    const u32 a = __float_as_uint(box_area(box_union(o, me)));
    /* candidates are visited in increasing j: strict '<' keeps the smallest index among equal areas */
    if (a < bestArea) { bestArea = a; best = t; }
  }
  return best;
}

struct PlocSmem {
  alignas(16) float raw[PLOC_WIN * 6];   /* TMA destination: window boxes as stored (24 B records)       */
  float soa[6 * PLOC_SOA];               /* the same boxes, one array per component                      */
  u32 ids[PLOC_TILE + PLOC_HALO];        /* cluster ids of [base, base+TILE+HALO)                         */
  short nn[PLOC_TILE + 2 * PLOC_R];      /* window slot of the nearest neighbour, for slots [HALO-R, HALO+TILE+R) */
  u32 warpSum[PLOC_THREADS / 32];
  u32 tile;
  u32 exclKeep, exclRemoved;
  alignas(8) u64 bar;
};

__global__ void __launch_bounds__(PLOC_THREADS) ploc_iter_kernel(const u32* __restrict__ idsIn, const b2bvh_aabb* __restrict__ boxesIn,
                                                                u32* __restrict__ idsOut, b2bvh_aabb* __restrict__ boxesOut,
                                                                b2bvh_bvh2_node* __restrict__ nodes, PlocCtrl* ctrl, u64* statusCur, u64* statusNext,
                                                                u32 iter) {
  __shared__ PlocSmem S;
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  const u32 count = ctrl->count[iter & 1u];
  if (count <= PLOC_TAIL) { /* the tail kernel takes over; keep the count visible to whichever launch comes next */
    if (blockIdx.x == 0 && tid == 0) ctrl->count[(iter + 1) & 1u] = count;
    return;
  }
  const u32 nTiles = (count + PLOC_TILE - 1) / PLOC_TILE;
  if (tid == 0) { pl_mbar_init(&S.bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  u32 phase = 0;
  while (true) {
    __syncthreads();
    if (tid == 0) S.tile = atomicAdd(&ctrl->ticket[iter & 1u], 1u);
    __syncthreads();
    const u32 tile = S.tile;
    if (tile >= nTiles) return;
    const u32 base = tile * PLOC_TILE;
    const int winBase = (int)base - PLOC_HALO; /* cluster index of window slot 0 */
    if (tid == 0) statusNext[tile] = 0ull;     /* next iteration has at most as many tiles: hand it a clean word */

    /* ---- stage the window [base-16, base+TILE+16): one bulk copy when it lies inside the array ---- */
    const bool bulk = (winBase >= 0) && ((u32)winBase + PLOC_WIN <= count);
    if (bulk) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* earlier generic reads of raw[] vs. the async write */
        pl_mbar_expect_tx(&S.bar, PLOC_WIN * 24);
        pl_tma_load_1d(S.raw, boxesIn + winBase, PLOC_WIN * 24, &S.bar); /* (base-16)*24 B is a multiple of 16 */
      }
      pl_mbar_wait(&S.bar, phase);
      phase ^= 1u;
    } else {
      const float* src = reinterpret_cast<const float*>(boxesIn);
      for (u32 k = tid; k < PLOC_WIN * 6; k += PLOC_THREADS) {
        const long long e = (long long)winBase * 6 + k;
        S.raw[k] = (e >= 0 && e < (long long)count * 6) ? __ldg(src + e) : 0.0f;
      }
      __syncthreads();
    }
    for (u32 k = tid; k < PLOC_WIN * 6; k += PLOC_THREADS) S.soa[(k % 6) * PLOC_SOA + k / 6] = S.raw[k];
    for (u32 k = tid; k < PLOC_TILE + PLOC_HALO; k += PLOC_THREADS) S.ids[k] = (base + k < count) ? __ldg(idsIn + base + k) : B2_INVALID;
    __syncthreads();

    /* ---- nearest neighbours for the tile and R clusters on either side ---- */
    for (u32 k = tid; k < PLOC_TILE + 2 * PLOC_R; k += PLOC_THREADS) {
      const int s = (int)k + PLOC_HALO - PLOC_R;
      const long long c = (long long)winBase + s;
      S.nn[k] = (c >= 0 && c < (long long)count) ? (short)nearest_in_window<PLOC_SOA>(S.soa, s, winBase, count) : (short)-1;
    }
    __syncthreads();

    /* ---- merge decision for cluster c = base + tid ---- */
    const u32 c = base + tid;
    const int s = (int)tid + PLOC_HALO;
    bool keep = false, removed = false;
    int ps = -1;
    if (c < count) {
      ps = S.nn[s - (PLOC_HALO - PLOC_R)];
      const bool mutual = ps >= 0 && S.nn[ps - (PLOC_HALO - PLOC_R)] == s;
      keep = mutual && s < ps;
      removed = mutual && s > ps;
    }
    /* CTA scan of (keep | removed << 16) */
    const u32 packed = (keep ? 1u : 0u) | (removed ? 0x10000u : 0u);
    u32 incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (l == 31) S.warpSum[w] = incl;
    __syncthreads();
    u32 warpBase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < PLOC_THREADS / 32; k++) { const u32 t = S.warpSum[k]; if (k < (int)w) warpBase += t; total += t; }
    const u32 localExcl = warpBase + incl - packed;
    const u32 tileKeep = total & 0xFFFFu, tileRemoved = total >> 16;

    if (w == 0) { /* look-back: one word carries both prefix sums (merging clusters, removed clusters) */
      const u64 mine = ((u64)tileKeep << 31) | (u64)tileRemoved;
      if (l == 0) st_release64(statusCur + tile, (tile == 0 ? LB64_INC : LB64_AGG) | mine);
      const u64 excl = warp_lookback_u64(statusCur, tile);
      if (l == 0) {
        if (tile > 0) st_release64(statusCur + tile, LB64_INC | (excl + mine));
        S.exclKeep = (u32)(excl >> 31);
        S.exclRemoved = (u32)(excl & 0x7FFFFFFFu);
        if (tile == nTiles - 1) {
          ctrl->count[(iter + 1) & 1u] = count - (S.exclRemoved + tileRemoved);
          ctrl->ticket[(iter + 1) & 1u] = 0;
          ctrl->itersRun = iter + 1;
          ctrl->liveBuf = (iter + 1) & 1u;
        }
      }
    }
    __syncthreads();

    if (c < count && !removed) {
      const u32 outPos = c - (S.exclRemoved + (localExcl >> 16));
      Box b = Box{S.soa[s], S.soa[PLOC_SOA + s], S.soa[2 * PLOC_SOA + s], S.soa[3 * PLOC_SOA + s], S.soa[4 * PLOC_SOA + s], S.soa[5 * PLOC_SOA + s]};
      u32 id = S.ids[tid];
      if (keep) {
        const Box o = Box{S.soa[ps], S.soa[PLOC_SOA + ps], S.soa[2 * PLOC_SOA + ps], S.soa[3 * PLOC_SOA + ps], S.soa[4 * PLOC_SOA + ps], S.soa[5 * PLOC_SOA + ps]};
        b = box_union(b, o);
        const u32 m = count - 2u - (S.exclKeep + (localExcl & 0xFFFFu));
        store_node2(nodes + m, id, S.ids[ps - PLOC_HALO], b);
        id = m;
      }
      idsOut[outPos] = id;
      store_aabb(boxesOut + outPos, b);
    }
  }
}

/* ---- tail: <= PLOC_TAIL clusters, one CTA, all remaining iterations in shared memory ---- */
struct PlocTailSmem {
  float soa[2][6 * (PLOC_TAIL + 2 * PLOC_R)];
  u32 ids[2][PLOC_TAIL];
  short nn[PLOC_TAIL];
  u32 warpSum[PLOC_TAIL_THREADS / 32];
};

__global__ void __launch_bounds__(PLOC_TAIL_THREADS) ploc_tail_kernel(const u32* __restrict__ idsA, const b2bvh_aabb* __restrict__ boxesA,
                                                                     const u32* __restrict__ idsB, const b2bvh_aabb* __restrict__ boxesB,
                                                                     b2bvh_bvh2_node* __restrict__ nodes, PlocCtrl* ctrl, u32 iter) {
  extern __shared__ __align__(16) unsigned char tailRaw[];
  PlocTailSmem& S = *reinterpret_cast<PlocTailSmem*>(tailRaw);
  constexpr int STRIDE = PLOC_TAIL + 2 * PLOC_R;
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  u32 count = ctrl->count[iter & 1u];
  if (count > PLOC_TAIL) return; /* more batches of the tiled kernel are needed first */
  const u32* ids = ctrl->liveBuf ? idsB : idsA;
  const b2bvh_aabb* boxes = ctrl->liveBuf ? boxesB : boxesA;
  int cur = 0;
  if (tid < count) {
    const Box b = load_aabb(boxes + tid);
    float* d = S.soa[0] + PLOC_R + tid;
    d[0] = b.lx; d[STRIDE] = b.ly; d[2 * STRIDE] = b.lz; d[3 * STRIDE] = b.hx; d[4 * STRIDE] = b.hy; d[5 * STRIDE] = b.hz;
    S.ids[0][tid] = ids[tid];
  }
  __syncthreads();
  u32 iters = 0;
  while (count > 1) {
    const float* soa = S.soa[cur];
    /* slot s = R + tid holds cluster tid; winBase = -R */
    int ps = -1;
    if (tid < count) ps = nearest_in_window<STRIDE>(soa, (int)tid + PLOC_R, -PLOC_R, count);
    S.nn[tid] = (short)ps;
    __syncthreads();
    bool keep = false, removed = false;
    if (tid < count) {
      const bool mutual = ps >= 0 && S.nn[ps - PLOC_R] == (short)((int)tid + PLOC_R);
      keep = mutual && (int)tid + PLOC_R < ps;
      removed = mutual && (int)tid + PLOC_R > ps;
    }
    const u32 packed = (keep ? 1u : 0u) | (removed ? 0x10000u : 0u);
    u32 incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (l == 31) S.warpSum[w] = incl;
    __syncthreads();
    u32 warpBase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < PLOC_TAIL_THREADS / 32; k++) { const u32 t = S.warpSum[k]; if (k < (int)w) warpBase += t; total += t; }
    const u32 localExcl = warpBase + incl - packed;
    if (tid < count && !removed) {
      const int s = (int)tid + PLOC_R;
      Box b = Box{soa[s], soa[STRIDE + s], soa[2 * STRIDE + s], soa[3 * STRIDE + s], soa[4 * STRIDE + s], soa[5 * STRIDE + s]};
      u32 id = S.ids[cur][tid];
      if (keep) {
        const Box o = Box{soa[ps], soa[STRIDE + ps], soa[2 * STRIDE + ps], soa[3 * STRIDE + ps], soa[4 * STRIDE + ps], soa[5 * STRIDE + ps]};
        b = box_union(b, o);
        const u32 m = count - 2u - (localExcl & 0xFFFFu);
        store_node2(nodes + m, id, S.ids[cur][ps - PLOC_R], b);
        id = m;
      }
      const u32 outPos = tid - (localExcl >> 16);
      float* d = S.soa[cur ^ 1] + PLOC_R + outPos;
      d[0] = b.lx; d[STRIDE] = b.ly; d[2 * STRIDE] = b.lz; d[3 * STRIDE] = b.hx; d[4 * STRIDE] = b.hy; d[5 * STRIDE] = b.hz;
      S.ids[cur ^ 1][outPos] = id;
    }
    count -= total >> 16;
    cur ^= 1;
    iters++;
    __syncthreads();
  }
  if (tid == 0) { ctrl->count[0] = ctrl->count[1] = 1; ctrl->itersRun += iters; }
}

int b2_launch_ploc(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const u32* d_sortedVals, u32 n, b2bvh_bvh2_node* d_nodes,
                   b2bvh_prim_ref* d_leaves, void* d_scratch, u32* h_iterations) {
  unsigned char* p = reinterpret_cast<unsigned char*>(d_scratch);
  PlocCtrl* ctrl = reinterpret_cast<PlocCtrl*>(p);
  const size_t np = ploc_pad(n);
  u32* ids[2] = {reinterpret_cast<u32*>(p + 256), reinterpret_cast<u32*>(p + 256) + np};
  b2bvh_aabb* boxes[2];
  boxes[0] = reinterpret_cast<b2bvh_aabb*>(p + 256 + 2 * np * 4);
  boxes[1] = boxes[0] + np;
  const size_t off = 256 + 2 * np * 4 + 2 * np * sizeof(b2bvh_aabb);
  u64* status[2] = {reinterpret_cast<u64*>(p + off), reinterpret_cast<u64*>(p + off) + ploc_tiles(n)};
  B2_CUDA(cudaMemsetAsync(status[0], 0, 2 * ploc_tiles(n) * sizeof(u64), ctx->stream));
  B2_KERNEL(ctx, "ploc_setup");
  ploc_setup_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_triAabb, d_sortedVals, n, d_leaves, ids[0], boxes[0], ctrl);
  B2_LAUNCH_CHECK(ctx);

  static bool attrSet = false;
  if (!attrSet) {
    B2_CUDA(cudaFuncSetAttribute(ploc_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PlocTailSmem)));
    attrSet = true;
  }
  const u32 grid = (u32)ctx->sm_count * 4u;
  u32 iter = 0;
  PlocCtrl h;
  h.count[0] = n;
  h.itersRun = 0;
  u32 live = n;
  while (live > PLOC_TAIL) {
    /* batch size: clusters typically shrink by ~0.72x per iteration; never fewer than 4 launches per host round trip */
    u32 batch = (u32)ceil(log((double)live / PLOC_TAIL) / log(1.0 / 0.72)) + 1;
    if (batch < 4) batch = 4;
    if (batch > 64) batch = 64;
    for (u32 k = 0; k < batch; k++, iter++) {
      B2_KERNEL(ctx, "ploc_iter");
      ploc_iter_kernel<<<grid, PLOC_THREADS, 0, ctx->stream>>>(ids[iter & 1], boxes[iter & 1], ids[(iter + 1) & 1], boxes[(iter + 1) & 1], d_nodes, ctrl,
                                                               status[iter & 1], status[(iter + 1) & 1], iter);
      B2_LAUNCH_CHECK(ctx);
    }
    B2_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    live = h.count[iter & 1];
    if (iter > 100000) return b2_fail(B2BVH_ERR_INTERNAL, "ploc: no convergence after %u iterations (%u clusters left)", iter, live);
  }
  B2_KERNEL(ctx, "ploc_tail");
  ploc_tail_kernel<<<1, PLOC_TAIL_THREADS, sizeof(PlocTailSmem), ctx->stream>>>(ids[0], boxes[0], ids[1], boxes[1], d_nodes, ctrl, iter);
  B2_LAUNCH_CHECK(ctx);
  B2_CUDA(cudaMemcpyAsync(&h, ctrl, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h.count[0] != 1) return b2_fail(B2BVH_ERR_INTERNAL, "ploc: tail left %u clusters", h.count[0]);
  *h_iterations = h.itersRun;
  return 0;
}
