/*
 * radix_sort.cu — stage S3: stable LSD radix sort of (Morton key, primitive index) u32 pairs.
 *
 * Replaces Oro::RadixSort::sort(KeyValueSoA, KeyValueSoA, n, startBit, endBit) (dependencies/Orochi/
 * ParallelPrimitives/RadixSort.cpp:291-318 and its kernels CountKernel / ParallelExclusiveScanAllWG /
 * SortKVKernel, RadixSortKernels.h:65,639,817), called from TwoPassLbvh.cpp:73-88, SinglePassLbvh.cpp:73-89,
 * PLOC++Bvh.cpp:63-79, Hploc.cpp:64-80.  Contract: result == std::stable_sort by key (the oracle Orochi's own
 * test uses, Test/RadixSort/main.cpp:130,239).
 *
 * Design.  8-bit digits, per digit three launches over a STATIC partition of the input into G contiguous chunks
 * (G = 2 CTAs per SM, so every CTA of a launch is resident and the partition is balanced to the element):
 *   radix_count_kernel    chunk -> 256 digit counts (16-byte loads, per-warp shared-memory bins)
 *   radix_scan_kernel     one warp per digit scans the G chunk counts (coalesced: counts are digit-major) -> exclusive
 *                         chunk offsets + digit totals
 *   radix_scatter_kernel  the CTA walks its chunk tile by tile (8192 pairs): keys and values arrive by TMA bulk copy
 *                         (cp.async.bulk + mbarrier, nothing staged in registers; the value tile is in flight while the
 *                         keys are ranked), warp-striped stable ranking with shared-memory atomics (see the kernel),
 *                         digit offsets carried in shared memory from tile to tile,
 *                         keys reordered in shared memory, values gathered through the tags, digit-contiguous stores.
 * No inter-CTA communication inside a launch: an earlier single-launch "onesweep" version with decoupled look-back spent
 * most of its time walking look-back chains as long as the number of resident CTAs (profiles/README.md, r01a) — on a
 * 148-SM part whose 126 MB L2 holds the whole key array of the 10 M benchmark, re-reading keys for the count costs far less
 * than the chains did.  The reference uses the same count/scan/scatter split per digit, but with spinning scan CTAs
 * (RadixSortKernels.h:606-637), LDS-atomic ranking and no vectorised or bulk loads.
 * Algorithmic traffic per pair: 4 passes x (4 count + 8 in + 8 out) = 80 B (the iota values of pass 0 are not read: 76 B).
 *
 * Same-box yardsticks (baseline/sort_compare.cu, profiles/r02a_sort_compare.json; 10 M / 100 M pairs, 30-bit keys): this sort 0.33 / 2.40 ms,
 * cub::DeviceRadixSort::SortPairs (onesweep, CUDA 12.9) 0.43 / 3.55 ms, the reference's own Orochi kernels compiled unmodified 1.12 / 7.42 ms.
 * Round-2 experiments that did NOT pay and were taken out again (profiles/r02_sort_experiments.txt):
 *   - three passes over 10-bit digits for the 30-bit Morton codes (+ a conditional pass for bits 30-31, which flat scenes use): 1024 bins cut
 *     the run a tile sends to one bin to 6 keys, every output sector becomes a partial write: 0.40 ms at 10 M, 5.1 ms at 100 M;
 *   - (key, value) pairs reordered in shared memory with one 8-byte store each instead of key + origin tag + value gather: 66.8 us per pass
 *     against 68.0 at 10 M and 2.7 ms against 2.4 at 100 M — the scatter kernel is not bound by shared-memory wavefronts but by its five
 *     barrier-separated phases per tile at two CTAs per SM.
 */
#include <stdlib.h>

#include "common.cuh"

#define RS_RADIX_BITS 8
#define RS_RADIX 256
#define RS_THREADS 512
#define RS_WARPS (RS_THREADS / 32)
#define RS_MAX_PASSES 4
#define RS_SCAN_WARPS 32
#define RS_SELF_MAX 136  /* chunks of 2048 pairs; sort of 144 K pairs 40.8 -> 34.8 us, 262 K (129 chunks) 40.9 -> 38.9, 390 K (191 chunks) 49.1 -> 53.2: the rows a CTA sums grow with the input */
#ifndef RS_ITEMS
#define RS_ITEMS 15   /* pairs per thread and tile of the scatter kernel (large inputs) */
#define RS_MINB 2     /* resident scatter CTAs per SM: 3 x 30 KB tile buffers + counters, twice */
#endif

/* scratch (u32 words): counts[256][Gpad] (digit-major; after the scan: exclusive chunk offsets) | totals[256] */
static inline u32 rs_grid(const b2bvh_ctx* ctx, u32 n, u32 tile) {
  const u32 tiles = (n + tile - 1) / tile;
  const u32 cap = (u32)ctx->sm_count * (tile == RS_THREADS * RS_ITEMS ? (u32)RS_MINB : 2u);
  return tiles < cap ? tiles : cap;
}
static inline u32 rs_gpad(u32 g) { return (g + 31u) & ~31u; }
size_t b2_sort_scratch_bytes(u32 n) {
  (void)n;
  return ((size_t)RS_RADIX * rs_gpad(1024) + RS_RADIX) * sizeof(u32); /* G <= 2 * SMs <= 1024 */
}

/* The twelve launches of a sort are one dependent chain of short kernels: each is launched programmatically (common.cuh), clears its shared
 * memory / initialises its barriers while its predecessor drains, and only then waits for it.  Sort stage, us, without / with: 10 M 303 / 297
 * (replayed graph) and 323 / 300 (stream launches); 1 M 71.6 / 67.5 and 93 / 73; 144 K 45.0 / 40.9 and 66.7 / 43.9. */
#define RS_PDL_PROLOGUE() pdl_wait()
#define RS_LAUNCH(kernel, grid, block, smem, stream, ...) B2_LAUNCH_PDL(kernel, grid, block, smem, stream, __VA_ARGS__)

/* ------------------------------------------------------------------------------------------------ count */
__global__ void __launch_bounds__(RS_THREADS) radix_count_kernel(const u32* __restrict__ keys, u32 n, u32 chunk, u32 shift, u32 mask,
                                                                 u32* __restrict__ counts, u32 gpad, u32 chunkMajor) {
  __shared__ u32 h[RS_WARPS][RS_RADIX];
  const u32 tid = threadIdx.x, w = tid >> 5;
  for (u32 k = tid; k < RS_WARPS * RS_RADIX; k += RS_THREADS) (&h[0][0])[k] = 0;
  RS_PDL_PROLOGUE();
  __syncthreads();
  const u32 begin = blockIdx.x * chunk, end = min(n, begin + chunk); /* chunk is a multiple of 4: 16-byte aligned loads */
  auto add4 = [&](const uint4& q) {
    atomicAdd(&h[w][(q.x >> shift) & mask], 1u);
    atomicAdd(&h[w][(q.y >> shift) & mask], 1u);
    atomicAdd(&h[w][(q.z >> shift) & mask], 1u);
    atomicAdd(&h[w][(q.w >> shift) & mask], 1u);
  };
  u32 i = begin + tid * 4;
  /* four independent 16-byte loads in flight per thread: the keys come from L2 (written by the previous pass) */
  for (; i + 3 * RS_THREADS * 4 + 4 <= end; i += 4 * RS_THREADS * 4) {
    uint4 q[4];
#pragma unroll
    for (int k = 0; k < 4; k++) q[k] = __ldg(reinterpret_cast<const uint4*>(keys + i + k * RS_THREADS * 4));
#pragma unroll
    for (int k = 0; k < 4; k++) add4(q[k]);
  }
  for (; i < end; i += RS_THREADS * 4) {
    if (i + 4 <= end) add4(__ldg(reinterpret_cast<const uint4*>(keys + i)));
    else
      for (u32 e = i; e < end; e++) atomicAdd(&h[w][(__ldg(keys + e) >> shift) & mask], 1u);
  }
  __syncthreads();
  if (tid < RS_RADIX) {
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) c += h[k][tid];
    if (chunkMajor) counts[(size_t)blockIdx.x * RS_RADIX + tid] = c; /* small inputs: no scan kernel, the scatter CTAs sum the rows themselves */
    else counts[(size_t)tid * gpad + blockIdx.x] = c;
  }
}

/* ------------------------------------------------------------------------------------------------ scan */
__global__ void __launch_bounds__(RS_SCAN_WARPS * 32) radix_scan_kernel(u32* __restrict__ counts, u32* __restrict__ totals, u32 g, u32 gpad) {
  /* one warp per digit walks its row 32 words at a time (a variant with one contiguous segment per lane and all loads
   * independent was slower: 12.5 vs 8.6 us, the strided accesses cost more than the dependent steps) */
  const u32 d = blockIdx.x * RS_SCAN_WARPS + (threadIdx.x >> 5), l = lane_id();
  RS_PDL_PROLOGUE();
  u32* row = counts + (size_t)d * gpad;
  u32 carry = 0;
  for (u32 base = 0; base < g; base += 32) {
    const u32 c = base + l;
    const u32 v = c < g ? row[c] : 0u;
    u32 incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (c < g) row[c] = carry + incl - v;
    carry += __shfl_sync(B2_FULL, incl, 31);
  }
  if (l == 0) totals[d] = carry;
}

/* ------------------------------------------------------------------------------------------------ scatter */
/* TMA bulk copy + mbarrier (PTX; cp.async.bulk -> SASS UBLKCP) */
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smemDst, const void* gsrc, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smemDst)), "l"(gsrc),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int ITEMS>
struct ScatterSmem {
  static constexpr int TILE = RS_THREADS * ITEMS;
  u32 keys[TILE];                 /* raw key tile (TMA destination), then the digit-ordered keys      */
  u32 vals[TILE];                 /* raw value tile (TMA destination), read through the tags           */
  u32 tags[TILE];                 /* per digit-ordered key: its position in the raw tile               */
  u32 warpHist[RS_WARPS][RS_RADIX]; /* per-warp digit counters, then the tile-local start of (warp, digit) */
  u32 digitBase[RS_RADIX];        /* global index where the next key of each digit goes (carried from tile to tile) */
  int globalBase[RS_RADIX];       /* digitBase - tile-local start of the digit                        */
  u32 warpTotals[8];
  alignas(8) u64 bar[2];
};

/* Ranking.  A warp owns 32*ITEMS consecutive keys of the tile, lane l taking keys l, l+32, ... (warp-striped), and counts
 * them on its private 256 counters in shared memory.  match.any on the 8-bit digit — the textbook way to learn which lanes
 * of an instruction share a digit — costs ~2 SM-cycles per DISTINCT value, 61 cycles for 8 random bits; eight ballots cost 26;
 * a shared-memory atomic add costs 4 (tools/micro/rank_micro.cu).  So every lane reads its counter, adds 1 atomically and
 * reads it again: the difference is the number of lanes of this instruction holding its digit.  A lane that is alone
 * (89 % of the lanes for random digits) has its rank — the first read.  The others run match.any among themselves only
 * (a couple of distinct values, a few cycles) and order themselves by lane: stable, whatever order the hardware served
 * the atomics in. */
template <int ITEMS, bool IOTA_VALUES>
__global__ void __launch_bounds__(RS_THREADS, ITEMS == RS_ITEMS ? RS_MINB : 2) radix_scatter_kernel(const u32* __restrict__ keysIn, const u32* __restrict__ valsIn,
                                                                      u32* __restrict__ keysOut, u32* __restrict__ valsOut,
                                                                      const u32* __restrict__ counts, const u32* __restrict__ totals, u32 n,
                                                                      u32 chunk, u32 shift, u32 mask, u32 gpad, u32 selfG) {
  using Smem = ScatterSmem<ITEMS>;
  constexpr u32 TILE = Smem::TILE;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  Smem& S = *reinterpret_cast<Smem*>(smemRaw);
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;

  if (tid == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  RS_PDL_PROLOGUE();
  /* small inputs (selfG = number of chunks, at most RS_SELF_MAX): the count kernel left its raw counts chunk-major and no scan kernel ran — every CTA sums
   * the rows itself (coalesced 1 KB rows, two half ranges per digit): digit totals and the counts of the chunks before this one.  One launch
   * less per pass where a launch is a third of the pass. */
  u32 selfTot = 0, selfBefore = 0;
  if (selfG) {
    static_assert(RS_THREADS == 2 * RS_RADIX, "two half ranges per digit");
    const u32 d = tid & (RS_RADIX - 1u), half = tid >> 8;
    const u32 gh = (selfG + 1u) >> 1, c0 = half * gh, c1 = min(selfG, c0 + gh);
    u32 tot = 0, bef = 0;
    for (u32 c = c0; c < c1; c++) {
      const u32 v = __ldg(counts + (size_t)c * RS_RADIX + d);
      tot += v;
      if (c < blockIdx.x) bef += v;
    }
    u32* tmp = reinterpret_cast<u32*>(S.keys); /* the key tile is not in use yet */
    tmp[tid] = tot; tmp[RS_THREADS + tid] = bef;
    __syncthreads();
    if (tid < RS_RADIX) { selfTot = tmp[tid] + tmp[tid + RS_RADIX]; selfBefore = tmp[RS_THREADS + tid] + tmp[RS_THREADS + tid + RS_RADIX]; }
    __syncthreads();
  }
  /* global start of every digit for this chunk = exclusive scan of the digit totals + this chunk's offset inside the digit */
  if (tid < RS_RADIX) {
    const u32 tot = selfG ? selfTot : __ldg(totals + tid);
    u32 incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (l == 31) S.warpTotals[w] = incl;
    S.globalBase[tid] = (int)(incl - tot);
  }
  __syncthreads();
  if (tid < RS_RADIX) {
    u32 add = 0;
    for (u32 k = 0; k < w; k++) add += S.warpTotals[k];
    S.digitBase[tid] = (u32)S.globalBase[tid] + add + (selfG ? selfBefore : __ldg(counts + (size_t)tid * gpad + blockIdx.x));
  }
  const u32 begin = blockIdx.x * chunk, end = min(n, begin + chunk);
  u32 phase = 0;

  for (u32 tileBase = begin; tileBase < end; tileBase += TILE) {
    const u32 valid = min(TILE, end - tileBase);
    const bool full = (valid == TILE);
    __syncthreads(); /* previous tile's shared-memory reads are done; digitBase is up to date */
    if (full && tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&S.bar[0], TILE * 4);
      tma_load_1d(S.keys, keysIn + tileBase, TILE * 4, &S.bar[0]);
      if (!IOTA_VALUES) {
        mbar_expect_tx(&S.bar[1], TILE * 4);
        tma_load_1d(S.vals, valsIn + tileBase, TILE * 4, &S.bar[1]);
      }
    }
#pragma unroll
    for (u32 k = 0; k < RS_RADIX / 32; k++) S.warpHist[w][l + 32 * k] = 0;

    /* ---- keys, warp-striped: item i of lane l of warp w is tile element w*32*ITEMS + i*32 + l ---- */
    u32 key[ITEMS];
    const u32 stripe = w * (32 * ITEMS) + l;
    if (full) {
      mbar_wait(&S.bar[0], phase);
#pragma unroll
      for (int i = 0; i < ITEMS; i++) key[i] = S.keys[stripe + i * 32];
    } else {
#pragma unroll
      for (int i = 0; i < ITEMS; i++) {
        const u32 e = stripe + i * 32;
        key[i] = e < valid ? __ldg(keysIn + tileBase + e) : 0xFFFFFFFFu; /* padding: last digit, last positions */
        if (!IOTA_VALUES && e < valid) S.vals[e] = __ldg(valsIn + tileBase + e);
      }
    }
    __syncwarp();
    /* ---- stable rank among the warp's keys of the same digit ---- */
    u32 rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const u32 d = (key[i] >> shift) & mask;
      volatile u32* h = &S.warpHist[w][d];
      const u32 old = *h;
      __syncwarp();
      atomicAdd(const_cast<u32*>(h), 1u);
      __syncwarp();
      const bool shared = (*h - old) > 1u; /* other lanes of this instruction hold the same digit */
      const u32 sm = __ballot_sync(B2_FULL, shared);
      u32 r = old;
      if (shared) r += __popc(__match_any_sync(sm, d) & lanemask_lt());
      rank[i] = r;
    }
    __syncthreads(); /* all raw keys are in registers, all warp histograms complete */

    /* ---- per digit: exclusive prefix over warps, tile count, tile-local start ---- */
    u32 count = 0;
    if (tid < RS_RADIX) {
#pragma unroll
      for (int k = 0; k < RS_WARPS; k++) count += S.warpHist[k][tid];
      u32 incl = count;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(B2_FULL, incl, o);
        if ((int)l >= o) incl += t;
      }
      if (l == 31) S.warpTotals[w] = incl;
      count = incl - count; /* exclusive inside the warp's 32 digits (the per-digit total is re-read below) */
    }
    __syncthreads();
    if (tid < RS_RADIX) {
      u32 start = count;
      for (u32 k = 0; k < w; k++) start += S.warpTotals[k];
      const u32 base = S.digitBase[tid];
      S.globalBase[tid] = (int)base - (int)start;
      u32 run = start;
#pragma unroll
      for (int k = 0; k < RS_WARPS; k++) { const u32 t = S.warpHist[k][tid]; S.warpHist[k][tid] = run; run += t; }
      S.digitBase[tid] = base + (run - start); /* padding keys of a ragged last tile only inflate the last digit of the last tile */
    }
    __syncthreads();

    /* ---- keys into digit order in shared memory (overwrites the raw tile), tagged with where they came from ---- */
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const u32 d = (key[i] >> shift) & mask;
      const u32 pos = S.warpHist[w][d] + rank[i];
      S.keys[pos] = key[i];
      S.tags[pos] = stripe + i * 32;
    }
    if (!IOTA_VALUES && full) mbar_wait(&S.bar[1], phase); /* the value tile landed long ago */
    __syncthreads();

    /* ---- out: element j of the digit-ordered tile goes to globalBase[digit] + (j corrected inside its group) ---- */
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
      const u32 j = tid + i * RS_THREADS;
      const u32 origin = S.tags[j];
      if (origin < valid) {
        const u32 k = S.keys[j];
        const int dst = S.globalBase[(k >> shift) & mask] + (int)j;
        keysOut[dst] = k;
        valsOut[dst] = IOTA_VALUES ? tileBase + origin : S.vals[origin];
      }
    }
    if (full) phase ^= 1u;
  }
}

template <int ITEMS>
static int launch_scatter(b2bvh_ctx* ctx, u32 grid, const u32* kin, const u32* vin, u32* kout, u32* vout, const u32* counts, const u32* totals, u32 n,
                          u32 chunk, u32 shift, u32 mask, u32 gpad, u32 selfG) {
  const size_t smem = sizeof(ScatterSmem<ITEMS>);
  const u32 onceBit = ITEMS == 4 ? B2_ONCE_SORT4 : B2_ONCE_SORT15;
  if (!(ctx->once_mask & onceBit)) {
    B2_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<ITEMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CUDA(cudaFuncSetAttribute(radix_scatter_kernel<ITEMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx->once_mask |= onceBit;
  }
  B2_KERNEL(ctx, "radix_scatter");
  if (vin == nullptr)
    RS_LAUNCH((radix_scatter_kernel<ITEMS, true>), grid, RS_THREADS, smem, ctx->stream, kin, (const u32*)nullptr, kout, vout, counts, totals, n, chunk, shift, mask, gpad, selfG);
  else
    RS_LAUNCH((radix_scatter_kernel<ITEMS, false>), grid, RS_THREADS, smem, ctx->stream, kin, vin, kout, vout, counts, totals, n, chunk, shift, mask, gpad, selfG);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2_launch_sort(b2bvh_ctx* ctx, const u32* d_keysIn, const u32* d_valsIn, u32* d_keysOut, u32* d_valsOut, u32* d_keysTmp, u32* d_valsTmp,
                   void* d_scratch, u32 n, u32 startBit, u32 endBit) {
  if (n == 0 || endBit <= startBit || endBit > 32) return b2_fail(B2BVH_ERR_INVALID, "sort: bad range n=%u bits [%u,%u)", n, startBit, endBit);
  if (n > 0x3FFFFFFFu) return b2_fail(B2BVH_ERR_INVALID, "sort: n=%u exceeds 2^30-1", n);
  const u32 nPasses = (endBit - startBit + RS_RADIX_BITS - 1) / RS_RADIX_BITS;
  /* small inputs use 2048-pair tiles so that more SMs take part */
  const bool small = n < (1u << 20);
  const u32 tile = RS_THREADS * (small ? 4u : (u32)RS_ITEMS);
  const u32 grid = rs_grid(ctx, n, tile);
  const u32 gpad = rs_gpad(grid);
  u32 chunk = (n + grid - 1) / grid;
  chunk = (chunk + 3u) & ~3u; /* 16-byte aligned chunk starts for the vector and bulk loads */
  /* up to RS_SELF_MAX chunks (~278 K pairs) the scatter CTAs do the scan themselves: two launches per pass instead of three */
  static const bool noSelf = getenv("B2BVH_SORT_NO_SELF_SCAN") != nullptr; /* development switch, identical output */
  const u32 selfG = (small && grid <= RS_SELF_MAX && !noSelf) ? grid : 0u;
  u32* counts = reinterpret_cast<u32*>(d_scratch);
  u32* totals = counts + (size_t)RS_RADIX * gpad;
  const u32* kin = d_keysIn;
  const u32* vin = d_valsIn;
  for (u32 p = 0; p < nPasses; p++) {
    /* ping-pong so that the last pass lands in the caller's output and the input is never overwritten */
    const bool toOut = ((nPasses - 1 - p) & 1u) == 0;
    u32* kout = toOut ? d_keysOut : d_keysTmp;
    u32* vout = toOut ? d_valsOut : d_valsTmp;
    const u32 shift = startBit + p * RS_RADIX_BITS;
    const u32 bits = (endBit - shift) < RS_RADIX_BITS ? (endBit - shift) : RS_RADIX_BITS;
    const u32 mask = (1u << bits) - 1u;
    B2_KERNEL(ctx, "radix_count");
    RS_LAUNCH(radix_count_kernel, grid, RS_THREADS, 0, ctx->stream, kin, n, chunk, shift, mask, counts, gpad, selfG ? 1u : 0u);
    B2_LAUNCH_CHECK(ctx);
    if (!selfG) {
      B2_KERNEL(ctx, "radix_scan");
      RS_LAUNCH(radix_scan_kernel, RS_RADIX / RS_SCAN_WARPS, RS_SCAN_WARPS * 32, 0, ctx->stream, counts, totals, grid, gpad);
      B2_LAUNCH_CHECK(ctx);
    }
    if (small) B2_TRY(launch_scatter<4>(ctx, grid, kin, vin, kout, vout, counts, totals, n, chunk, shift, mask, gpad, selfG));
    else B2_TRY(launch_scatter<RS_ITEMS>(ctx, grid, kin, vin, kout, vout, counts, totals, n, chunk, shift, mask, gpad, 0u));
    kin = kout;
    vin = vout;
  }
  return 0;
}
