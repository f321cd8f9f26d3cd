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
 *   - rank is the exclusive prefix count of merging clusters in list order, not the arrival order of a global atomicAdd
 *     (:57-68,:311) -> node numbering is deterministic (canonical numbering of the oracle);
 *   - ONE cooperative launch runs every iteration down to 1024 clusters (ploc_merge_kernel: static chunks, posted chunk
 *     totals instead of the reference's serial inter-CTA chain :341-347, one grid barrier per iteration), a single-CTA kernel
 *     finishes in shared memory; the reference launches once per iteration and blocks on a D2H copy of the cluster count
 *     (PLOC++Bvh.cpp:150);
 *   - cluster boxes live in a COMPACTED array next to the cluster ids (24 B per live cluster), so a tile's search window
 *     [base-16, base+496) is one contiguous range that a single TMA bulk copy (cp.async.bulk) stages into shared memory;
 *     the reference gathers boxes through the node index every iteration;
 *   - every candidate pair is evaluated once (area(a U b) is symmetric) and boxes are read in place as 8-byte words.
 * Per iteration and live cluster: 24 B box read + 1 B decision written (pass 1); 28 B + 1 B read, 28 B written per survivor,
 * 32 B per merged node (pass 2).
 */
#include <stdio.h>

#include "common.cuh"

#define PLOC_R 8
#define PLOC_THREADS 512
#define PLOC_WIN PLOC_THREADS                  /* window slots of a tile = threads of the CTA */
#define PLOC_HALO (2 * PLOC_R)
#define PLOC_TILE (PLOC_WIN - 2 * PLOC_HALO)   /* clusters a window decides: 480 */
#define PLOC_TAIL 1024 /* clusters handled by the single-CTA tail kernel */
#define PLOC_TAIL_THREADS 1024
#define PLOC_MAX_GRID 4096

struct PlocCtrl {
  u32 bar;        /* grid barrier: arrivals so far */
  u32 count;      /* live clusters handed to the tail kernel; 1 when the tail is done */
  u32 itersRun;   /* iterations that did work */
  u32 liveBuf;    /* which of the two cluster buffers holds the live list */
  u32 arrive;     /* chunks that have posted their totals, all iterations so far */
  u32 pad[3];
  uint4 next[2];  /* {iteration, live clusters, -, -} valid after barrier b in next[b & 1] */
};

/* scratch: PlocCtrl (256 B) | ids[2][nPad] | boxes[2][nPad] (24 B each) | decisions[nPad] (1 B) | counts[PLOC_MAX_GRID] (u64);
 * nPad = n rounded up to 16 so that every region (and every tile window inside the box arrays) starts on a 16-byte boundary
 * for the bulk copies */
static size_t ploc_pad(u32 n) { return ((size_t)n + 15) & ~(size_t)15; }
size_t b2_ploc_scratch_bytes(u32 n) {
  return 256 + 2 * ploc_pad(n) * 4 + 2 * ploc_pad(n) * sizeof(b2bvh_aabb) + ploc_pad(n) + PLOC_MAX_GRID * sizeof(u64);
}

__global__ void __launch_bounds__(256) ploc_setup_kernel(const b2bvh_aabb* __restrict__ triAabb, const u32* __restrict__ sortedVals, u32 n,
                                                         b2bvh_prim_ref* __restrict__ leaves, u32* __restrict__ ids, b2bvh_aabb* __restrict__ boxes,
                                                         PlocCtrl* ctrl) {
  __shared__ __align__(16) u32 sLeaf[256 * 7];
  __shared__ __align__(16) u32 sBox[256 * 6];
  const u32 t = threadIdx.x, g0 = blockIdx.x * 256, g = g0 + t;
  if (g == 0) { ctrl->bar = 0; ctrl->arrive = 0; ctrl->count = n; ctrl->itersRun = 0; ctrl->liveBuf = 0; }
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

struct PlocSmem {
  alignas(128) float raw[2][PLOC_WIN * 6];   /* TMA destinations (double buffered): the window's boxes as stored, 24 B records */
  alignas(16) u32 ids[2][PLOC_WIN];          /* TMA destinations: the window's cluster ids (pass 2)                            */
  u32 dist[PLOC_R][PLOC_WIN + PLOC_R];       /* dist[r-1][R + s] = float bits of area(slot s U slot s+r)                       */
  short nn[PLOC_WIN];                        /* window slot of each slot's nearest neighbour                                  */
  unsigned char dec[PLOC_WIN];               /* decisions of a one-tile chunk, by window slot (kept on chip)                  */
  u32 red[4][PLOC_THREADS / 32];
  alignas(8) u64 bar[2];
};

/* A box record read in place: three 8-byte loads at a 24-byte lane stride are free of bank conflicts (the 16 lanes of a
 * half-warp start on 16 distinct even banks), so the records need no transposition into component arrays. */
__device__ __forceinline__ Box pl_slot_box(const float* raw, int s) {
  const float2* p = reinterpret_cast<const float2*>(raw + 6 * s);
  const float2 a = p[0], b = p[1], c = p[2];
  return Box{a.x, a.y, b.x, b.y, c.x, c.y};
}

/* Nearest neighbour of window slot s = threadIdx.x (cluster winBase + s) among the clusters within +-R that exist: the one
 * minimising (float bits of area(box_s U box_j), j) (Ploc++Kernel.h:253-268).  area(a U b) is symmetric, so every pair is
 * evaluated ONCE, by its left member, and handed to the right member through shared memory — half the arithmetic of each
 * cluster scanning all 16 candidates.  Candidates are compared in increasing j with a strict '<': the smallest index wins
 * among equal areas.  All threads that hold a slot call it (two barriers inside: `sync`); slots closer than R to a window edge get
 * incomplete answers that nobody reads.  CHECK: the window sticks out of [0, count). */
template <bool CHECK, int SLOTS, typename Sync>
__device__ __forceinline__ int pl_nearest(u32 (*dist)[SLOTS + PLOC_R], short* nn, const float* raw, int s, int winBase, u32 count, Sync sync) {
  const Box me = pl_slot_box(raw, s);
  const bool vs = !CHECK || (winBase + s >= 0 && winBase + s < (int)count);
  u32 d[PLOC_R];
#pragma unroll
  for (int r = 1; r <= PLOC_R; r++) {
    /* unchecked windows: nobody reads the answers of the slots >= SLOTS - R; they re-read the last record instead of running past the
     * buffer (the next buffer may be the destination of a bulk copy in flight: compute-sanitizer racecheck, profiles/r02_sanitizer.txt) */
    const int t = CHECK ? s + r : min(s + r, SLOTS - 1);
    bool ok = true;
    if (CHECK) ok = t < SLOTS && vs && winBase + t >= 0 && winBase + t < (int)count;
    u32 v = 0xFFFFFFFFu;
    if (ok) v = __float_as_uint(box_area(box_union(pl_slot_box(raw, t), me)));
    d[r - 1] = v;
    dist[r - 1][PLOC_R + s] = v;
    if (CHECK && s < PLOC_R) dist[r - 1][s] = 0xFFFFFFFFu; /* left of slot 0: nothing (unchecked windows never use slots < R) */
  }
  sync();
  u32 bestArea = 0xFFFFFFFFu;
  int best = -1;
#pragma unroll
  for (int r = PLOC_R; r >= 1; r--) {
    const u32 v = dist[r - 1][PLOC_R + s - r];
    if (v < bestArea) { bestArea = v; best = s - r; }
  }
#pragma unroll
  for (int r = 1; r <= PLOC_R; r++)
    if (d[r - 1] < bestArea) { bestArea = d[r - 1]; best = s + r; }
  nn[s] = (short)best;
  sync();
  return best;
}

/* ---- all iterations down to PLOC_TAIL clusters in ONE cooperative launch ----
 * Every CTA is resident.  In each iteration the live list [0, count) is cut into one contiguous chunk per CTA (a multiple of
 * PLOC_TILE clusters).  Pass 1 walks the chunk window by window — the window [base-16, base+496) arrives as one bulk copy,
 * the next one is in flight meanwhile — finds the nearest neighbours and records each cluster's decision (keeps and merges
 * with the partner at +k / is removed / stays) in one byte.  The CTA then posts its (merging, removed) totals, counts itself
 * in on an arrival counter that ONE thread per CTA watches, and sums the totals of the chunks before it: all chunks post at
 * the same moment, so there is no chain of dependent look-backs (the former tile-by-tile version spent 38 % of its time
 * waiting on chains as long as the number of resident CTAs, profiles/r01l; letting every thread spin on the posted words
 * instead cost as much: ~150 K polling threads starve the CTAs still at work, profiles/r01m).  Pass 2 streams the same
 * windows again (boxes and ids by bulk copy, decisions from L2) and writes the merged nodes, ids and boxes in place order;
 * a chunk of one window (<= 284 K clusters on a B200) keeps everything from pass 1 on chip.  One grid barrier separates
 * iterations; the host is not involved. */
#ifdef PLOC_TRACE /* development: per-iteration phase timestamps of CTA 0 and the last active CTA (B2BVH_LIB variant build) */
__device__ unsigned long long g_plocTrace[2][256][8];
__device__ __forceinline__ unsigned long long pl_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PL_TRACE(slot) do { if (threadIdx.x == 0 && iter < 256 && (c == 0 || c == nActive - 1)) g_plocTrace[c == 0 ? 0 : 1][iter][slot] = pl_now(); } while (0)
#else
#define PL_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ void pl_publish(PlocCtrl* ctrl, u32 barrier, u32 iter, u32 count) {
  u32* nx = reinterpret_cast<u32*>(&ctrl->next[barrier & 1u]);
  st_relaxed(nx, iter); st_relaxed(nx + 1, count);
}

__global__ void __launch_bounds__(PLOC_THREADS, 4) ploc_merge_kernel(u32* ids0, u32* ids1, b2bvh_aabb* boxes0, b2bvh_aabb* boxes1, unsigned char* dec,
                                                                    b2bvh_bvh2_node* __restrict__ nodes, PlocCtrl* ctrl, u64* counts, u32 n) {
  __shared__ PlocSmem S;
  const u32 G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  if (tid == 0) {
    pl_mbar_init(&S.bar[0], 1);
    pl_mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  u32 iter = 0, count = n, barriers = 0, phases = 0, arriveTarget = 0, barTarget = 0;
  while (count > PLOC_TAIL) {
    const u32* idsIn = (iter & 1u) ? ids1 : ids0;
    u32* idsOut = (iter & 1u) ? ids0 : ids1;
    const b2bvh_aabb* boxesIn = (iter & 1u) ? boxes1 : boxes0;
    b2bvh_aabb* boxesOut = (iter & 1u) ? boxes0 : boxes1;
    const u32 chunk = ((count + G - 1) / G + PLOC_TILE - 1) / PLOC_TILE * PLOC_TILE;
    const u32 nActive = (count + chunk - 1) / chunk;
    arriveTarget += nActive;
    /* once a chunk is a single window the number of chunks only falls: CTAs without one leave for good, and the barrier
     * (like the arrival counter) is among the CTAs that still work */
    const bool oneWindow = chunk == PLOC_TILE;
    if (oneWindow && c >= nActive) break;
    barTarget += oneWindow ? nActive : G;
    PL_TRACE(0);
    if (c < nActive) {
      const u32 cStart = c * chunk, cEnd = min(count, cStart + chunk);
      const bool single = chunk == PLOC_TILE;
      const u32 nT = (cEnd - cStart + PLOC_TILE - 1) / PLOC_TILE;
      auto bulk_ok = [&](u32 k) { /* the window of tile k lies inside the list */
        const int wb = (int)(cStart + k * PLOC_TILE) - PLOC_HALO;
        return wb >= 0 && (u32)wb + PLOC_WIN <= count;
      };
      auto stage = [&](u32 k, bool withIds) { /* thread 0 */
        const u32 b = k & 1u;
        const u32 wb = cStart + k * PLOC_TILE - PLOC_HALO; /* (base-16)*24 B and (base-16)*4 B are multiples of 16 */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* the buffer was read through the generic proxy */
        pl_mbar_expect_tx(&S.bar[b], PLOC_WIN * 24 + (withIds ? PLOC_WIN * 4 : 0));
        pl_tma_load_1d(S.raw[b], boxesIn + wb, PLOC_WIN * 24, &S.bar[b]);
        if (withIds) pl_tma_load_1d(S.ids[b], idsIn + wb, PLOC_WIN * 4, &S.bar[b]);
      };
      auto load_edge = [&](u32 k, bool withIds) { /* a window that sticks out of the list: every thread fetches its slot */
        const u32 b = k & 1u;
        const int ci = (int)(cStart + k * PLOC_TILE) - PLOC_HALO + (int)tid;
        float2 q0 = make_float2(0.f, 0.f), q1 = q0, q2 = q0;
        u32 id = B2_INVALID;
        if (ci >= 0 && ci < (int)count) {
          const float2* bp = reinterpret_cast<const float2*>(boxesIn + ci);
          q0 = __ldcg(bp); q1 = __ldcg(bp + 1); q2 = __ldcg(bp + 2);
          if (withIds) id = __ldcg(idsIn + ci);
        }
        float2* dst = reinterpret_cast<float2*>(S.raw[b] + 6 * tid);
        dst[0] = q0; dst[1] = q1; dst[2] = q2;
        if (withIds) S.ids[b][tid] = id;
        __syncthreads();
      };
      auto wait_tile = [&](u32 k) {
        const u32 b = k & 1u;
        pl_mbar_wait(&S.bar[b], (phases >> b) & 1u);
        phases ^= 1u << b;
      };
      /* ---- pass 1: decisions ---- */
      auto cta_sync = [] { __syncthreads(); };
      u32 myKeep = 0, myRemoved = 0;
      if (tid == 0 && bulk_ok(0)) stage(0, single);
      for (u32 k = 0; k < nT; k++) {
        if (tid == 0 && k + 1 < nT && bulk_ok(k + 1)) stage(k + 1, false); /* that buffer was last read two __syncthreads ago */
        const int winBase = (int)(cStart + k * PLOC_TILE) - PLOC_HALO;
        const float* raw = S.raw[k & 1u];
        int best;
        if (bulk_ok(k)) {
          wait_tile(k);
          best = pl_nearest<false, PLOC_WIN>(S.dist, S.nn, raw, (int)tid, winBase, count, cta_sync);
        } else {
          load_edge(k, single);
          best = pl_nearest<true, PLOC_WIN>(S.dist, S.nn, raw, (int)tid, winBase, count, cta_sync);
        }
        const u32 ci = (u32)(winBase + (int)tid);
        if (tid >= PLOC_HALO && tid < PLOC_HALO + PLOC_TILE && ci < cEnd) {
          const bool mutual = best >= 0 && S.nn[best] == (short)tid;
          const bool keep = mutual && (int)tid < best, removed = mutual && (int)tid > best;
          const unsigned char d = keep ? (unsigned char)(0x40 | (best - (int)tid)) : (removed ? (unsigned char)0x80 : (unsigned char)0);
          myKeep += keep ? 1u : 0u;
          myRemoved += removed ? 1u : 0u;
          if (single) S.dec[tid] = d; else dec[ci] = d;
        }
      }
      PL_TRACE(1);
      /* ---- totals of the chunk, posted; totals of the chunks before it ---- */
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { myKeep += __shfl_xor_sync(B2_FULL, myKeep, o); myRemoved += __shfl_xor_sync(B2_FULL, myRemoved, o); }
      if (l == 0) { S.red[0][w] = myKeep; S.red[1][w] = myRemoved; }
      __syncthreads();
      u32 chunkKeep = 0, chunkRemoved = 0;
#pragma unroll
      for (int q = 0; q < PLOC_THREADS / 32; q++) { chunkKeep += S.red[0][q]; chunkRemoved += S.red[1][q]; }
      if (tid == 0) {
        st_relaxed64(counts + c, ((u64)chunkKeep << 32) | (u64)chunkRemoved);
        __threadfence();
        atomicAdd(&ctrl->arrive, 1u);
        PL_TRACE(2);
        SpinGuard guard;
        while (ld_acquire(&ctrl->arrive) < arriveTarget) {
          __nanosleep(32);
          guard.tick();
        }
      }
      __syncthreads();
      u32 bk = 0, br = 0;
      for (u32 i = tid; i < c; i += PLOC_THREADS) {
        const u64 v = ld_relaxed64(counts + i);
        bk += (u32)(v >> 32);
        br += (u32)v;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { bk += __shfl_xor_sync(B2_FULL, bk, o); br += __shfl_xor_sync(B2_FULL, br, o); }
      if (l == 0) { S.red[2][w] = bk; S.red[3][w] = br; }
      __syncthreads();
      u32 runKeep = 0, runRemoved = 0;
#pragma unroll
      for (int q = 0; q < PLOC_THREADS / 32; q++) { runKeep += S.red[2][q]; runRemoved += S.red[3][q]; }
      if (c == nActive - 1 && tid == 0) pl_publish(ctrl, barriers + 1u, iter + 1u, count - (runRemoved + chunkRemoved));
      PL_TRACE(3);
      /* ---- pass 2: merged nodes, compacted ids and boxes ---- */
      if (!single && tid == 0 && bulk_ok(0)) stage(0, true);
      for (u32 k = 0; k < nT; k++) {
        const u32 b = k & 1u;
        const int winBase = (int)(cStart + k * PLOC_TILE) - PLOC_HALO;
        const u32 ci = (u32)(winBase + (int)tid);
        const bool owned = tid >= PLOC_HALO && tid < PLOC_HALO + PLOC_TILE && ci < cEnd;
        unsigned char d = 0;
        if (single) {
          if (owned) d = S.dec[tid]; /* boxes and ids of the window are still in buffer 0 */
        } else {
          if (tid == 0 && k + 1 < nT && bulk_ok(k + 1)) stage(k + 1, true);
          if (owned) d = __ldcg(dec + ci);
          if (bulk_ok(k)) wait_tile(k); else load_edge(k, true);
        }
        const bool keep = (d & 0x40) != 0, removed = (d & 0x80) != 0;
        const u32 packed = (keep ? 1u : 0u) | (removed ? 0x10000u : 0u);
        u32 incl = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_up_sync(B2_FULL, incl, o);
          if ((int)l >= o) incl += t;
        }
        if (l == 31) S.red[0][w] = incl;
        __syncthreads();
        u32 warpBase = 0, total = 0;
#pragma unroll
        for (int q = 0; q < PLOC_THREADS / 32; q++) { const u32 t = S.red[0][q]; if (q < (int)w) warpBase += t; total += t; }
        const u32 localExcl = warpBase + incl - packed;
        if (owned && !removed) {
          Box bx = pl_slot_box(S.raw[b], (int)tid);
          u32 id = S.ids[b][tid];
          if (keep) {
            const int ps = (int)tid + (d & 0xF);
            bx = box_union(bx, pl_slot_box(S.raw[b], ps));
            const u32 m = count - 2u - (runKeep + (localExcl & 0xFFFFu));
            store_node2(nodes + m, id, S.ids[b][ps], bx);
            id = m;
          }
          const u32 outPos = ci - (runRemoved + (localExcl >> 16));
          idsOut[outPos] = id;
          store_aabb(boxesOut + outPos, bx);
        }
        runKeep += total & 0xFFFFu;
        runRemoved += total >> 16;
        __syncthreads(); /* S.red[0] and the window buffers are reused */
      }
    }
    PL_TRACE(4);
#ifdef PLOC_TRACE
    if (tid == 0 && iter < 256 && (c == 0 || c == nActive - 1)) g_plocTrace[c == 0 ? 0 : 1][iter][6] = count;
#endif
    barriers++;
    grid_barrier(&ctrl->bar, barTarget);
    if (tid == 0) asm volatile("fence.proxy.async.global;" ::: "memory"); /* the lists were written through the generic proxy, the bulk copies read them through the async proxy */
    PL_TRACE(5);
    {
      const u32* nx = reinterpret_cast<const u32*>(&ctrl->next[barriers & 1u]);
      const u32 before = count;
      iter = ld_relaxed(nx); count = ld_relaxed(nx + 1);
      if (count >= before) break; /* cannot happen (the closest pair is always mutual); never spin on a defect: the host reports it */
    }
  }
  if (c == 0 && tid == 0) { ctrl->count = count; ctrl->itersRun = iter; ctrl->liveBuf = iter & 1u; }
}

/* ---- tail: <= PLOC_TAIL clusters, one CTA, all remaining iterations in shared memory ---- */
struct PlocTailSmem {
  alignas(16) float raw[2][PLOC_TAIL * 6]; /* 24-byte box records, double buffered */
  u32 dist[PLOC_R][PLOC_TAIL + PLOC_R];
  u32 ids[2][PLOC_TAIL];
  short nn[PLOC_TAIL];
  u32 warpSum[PLOC_TAIL_THREADS / 32];
};

__global__ void __launch_bounds__(PLOC_TAIL_THREADS) ploc_tail_kernel(const u32* __restrict__ idsA, const b2bvh_aabb* __restrict__ boxesA,
                                                                     const u32* __restrict__ idsB, const b2bvh_aabb* __restrict__ boxesB,
                                                                     b2bvh_bvh2_node* __restrict__ nodes, PlocCtrl* ctrl) {
  extern __shared__ __align__(16) unsigned char tailRaw[];
  PlocTailSmem& S = *reinterpret_cast<PlocTailSmem*>(tailRaw);
  static_assert(PLOC_TAIL == PLOC_TAIL_THREADS, "one cluster per thread");
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  u32 count = ctrl->count;
  if (count > PLOC_TAIL) return; /* the merge kernel did not get there: reported by the host */
  const u32* ids = ctrl->liveBuf ? idsB : idsA;
  const b2bvh_aabb* boxes = ctrl->liveBuf ? boxesB : boxesA;
  int cur = 0;
  if (tid < count) {
    const float2* bp = reinterpret_cast<const float2*>(boxes + tid);
    float2* d = reinterpret_cast<float2*>(S.raw[0] + 6 * tid);
    d[0] = __ldcg(bp); d[1] = __ldcg(bp + 1); d[2] = __ldcg(bp + 2);
    S.ids[0][tid] = __ldcg(ids + tid);
  }
  __syncthreads();
  u32 iters = 0;
  /* the list only shrinks: a warp whose slots are all past the end leaves for good, the others meet at a named barrier */
  while (count > 1) {
    const u32 nW = (count + 31u) >> 5;
    if (w >= nW) break;
    auto live_sync = [&] { named_barrier(1, nW * 32u); };
    const float* raw = S.raw[cur];
    const int ps = pl_nearest<true, PLOC_TAIL>(S.dist, S.nn, raw, (int)tid, 0, count, live_sync);
    bool keep = false, removed = false;
    if (tid < count) {
      const bool mutual = ps >= 0 && S.nn[ps] == (short)tid;
      keep = mutual && (int)tid < ps;
      removed = mutual && (int)tid > ps;
    }
    const u32 packed = (keep ? 1u : 0u) | (removed ? 0x10000u : 0u);
    u32 incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (l == 31) S.warpSum[w] = incl;
    live_sync();
    u32 warpBase = 0, total = 0;
    for (u32 k = 0; k < nW; k++) { const u32 t = S.warpSum[k]; if (k < w) warpBase += t; total += t; }
    const u32 localExcl = warpBase + incl - packed;
    if (tid < count && !removed) {
      Box b = pl_slot_box(raw, (int)tid);
      u32 id = S.ids[cur][tid];
      if (keep) {
        b = box_union(b, pl_slot_box(raw, ps));
        const u32 m = count - 2u - (localExcl & 0xFFFFu);
        store_node2(nodes + m, id, S.ids[cur][ps], b);
        id = m;
      }
      const u32 outPos = tid - (localExcl >> 16);
      float2* d = reinterpret_cast<float2*>(S.raw[cur ^ 1] + 6 * outPos);
      d[0] = make_float2(b.lx, b.ly); d[1] = make_float2(b.lz, b.hx); d[2] = make_float2(b.hy, b.hz);
      S.ids[cur ^ 1][outPos] = id;
    }
    cur ^= 1;
    iters++;
    live_sync();
    if ((total >> 16) == 0) break; /* cannot happen; reported by the host */
    count -= total >> 16;
  }
  if (tid == 0) { ctrl->count = count; ctrl->itersRun += iters; } /* warp 0 takes part in every iteration */
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
  unsigned char* dec = p + 256 + 2 * np * 4 + 2 * np * sizeof(b2bvh_aabb);
  u64* counts = reinterpret_cast<u64*>(dec + np);
  int& occ = ctx->occ[B2_OCC_PLOC];
  if (!(ctx->once_mask & B2_ONCE_PLOC)) {
    B2_CUDA(cudaFuncSetAttribute(ploc_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PlocTailSmem)));
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ploc_merge_kernel, PLOC_THREADS, 0));
    if (occ < 1) return b2_fail(B2BVH_ERR_INTERNAL, "ploc: kernel does not fit on an SM");
    ctx->once_mask |= B2_ONCE_PLOC;
  }
  B2_KERNEL(ctx, "ploc_setup");
  ploc_setup_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_triAabb, d_sortedVals, n, d_leaves, ids[0], boxes[0], ctrl);
  B2_LAUNCH_CHECK(ctx);
  if (n > PLOC_TAIL) {
    /* every CTA must be resident (grid barrier, posted counts): at most SMs x occupancy; small inputs use fewer CTAs */
    /* CTAs per SM by input size: every iteration ends in two grid-wide steps whose cost grows with the number of CTAs, the
     * window work per CTA shrinks with it (measured build stage in us with 1 / 2 / 3 / 4 CTAs per SM: 262 K 274 / 297 / 328 / 348,
     * 1 M 374 / 393 / 424 / 442, 2 M 548 / 544 / 573 / 598, 4 M 890 / 839 / 858 / 878, 10 M 1904 / 1668 / 1658 / 1684) */
    u32 perSm = n <= (2u << 20) ? 1u : (n <= (6u << 20) ? 2u : 3u);
    if (perSm > (u32)occ) perSm = (u32)occ;
    u32 grid = (u32)ctx->sm_count * perSm;
    const u32 want = (n + PLOC_TILE - 1) / PLOC_TILE;
    if (grid > want) grid = want;
    if (grid > PLOC_MAX_GRID) grid = PLOC_MAX_GRID;
    if (ctx->merge_max_ctas && grid > ctx->merge_max_ctas) grid = ctx->merge_max_ctas;
    B2_CUDA(cudaMemsetAsync(counts, 0, (size_t)grid * sizeof(u64), ctx->stream));
    B2_KERNEL(ctx, "ploc_merge");
    void* args[] = {(void*)&ids[0], (void*)&ids[1], (void*)&boxes[0], (void*)&boxes[1], (void*)&dec, (void*)&d_nodes, (void*)&ctrl, (void*)&counts, (void*)&n};
    B2_CUDA(cudaLaunchCooperativeKernel((const void*)ploc_merge_kernel, dim3(grid), dim3(PLOC_THREADS), args, 0, ctx->stream));
    B2_LAUNCH_CHECK(ctx);
  }
  B2_KERNEL(ctx, "ploc_tail");
  ploc_tail_kernel<<<1, PLOC_TAIL_THREADS, sizeof(PlocTailSmem), ctx->stream>>>(ids[0], boxes[0], ids[1], boxes[1], d_nodes, ctrl);
  B2_LAUNCH_CHECK(ctx);
  /* PlocCtrl::count (1 when the tree is complete) and ::itersRun come back through the mailbox, read by b2bvh_build after its
   * final synchronisation: b2_mailbox(ctx, B2_MB_PLOC)[1], [2] */
  (void)h_iterations;
  B2_TRY(b2_fetch_words(ctx, ctrl, 4, B2_MB_PLOC));
#ifdef PLOC_TRACE
  {
    cudaStreamSynchronize(ctx->stream);
    static unsigned long long tr[2][256][8];
    cudaMemcpyFromSymbol(tr, g_plocTrace, sizeof(tr));
    for (u32 i = 0; i < 256 && tr[0][i][0]; i++) {
      fprintf(stderr, "ploc it %3u count %9llu |", i, tr[0][i][6]);
      for (int k = 0; k < 2; k++) {
        fprintf(stderr, " cta%s:", k ? "L" : "0");
        for (int q = 1; q < 6; q++) fprintf(stderr, " %6.1f", (double)(tr[k][i][q] - tr[k][i][q - 1]) * 1e-3);
      }
      fprintf(stderr, "\n");
    }
  }
#endif
  return 0;
}
