/*
 * radix_sort.cu — stage S3: stable LSD radix sort of (Morton key, primitive index) u32 pairs.
 *
 * Replaces Oro::RadixSort::sort(KeyValueSoA, KeyValueSoA, n, startBit, endBit) (dependencies/Orochi/
 * ParallelPrimitives/RadixSort.cpp:291-318 and its kernels CountKernel / ParallelExclusiveScanAllWG /
 * SortKVKernel, RadixSortKernels.h:65,639,817), called from TwoPassLbvh.cpp:73-88, SinglePassLbvh.cpp:73-89,
 * PLOC++Bvh.cpp:63-79, Hploc.cpp:64-80.  Contract: result == std::stable_sort by key (the oracle Orochi's own
 * test uses, Test/RadixSort/main.cpp:130,239).
 *
 * Design (B200-first, not the reference's count/scan/scatter triple per digit):
 *   - ONE histogram launch for all digits (4 B read per key); its last CTA turns the 4x256 bins into
 *     exclusive offsets, so there is no separate scan launch.
 *   - per 8-bit digit ONE "onesweep" launch: each CTA claims a tile with an atomic ticket (forward
 *     progress without co-residency assumptions), ranks its keys with warp match-any (stable inside the
 *     warp-striped tile), gets the tile's global digit offsets by decoupled look-back over a
 *     flag|count status word per (tile, digit), reorders keys and values through shared memory and
 *     writes them in digit-contiguous runs.  8 B read + 8 B written per pair and pass.
 *   - full tiles arrive through the TMA engine: one elected thread issues cp.async.bulk copies of the
 *     key tile and the value tile into shared memory and the CTA waits on an mbarrier; nothing is staged
 *     in registers, and the value tile is in flight while the keys are being ranked.
 *   Algorithmic traffic: 4 + 4*16 = 68 B per pair for 32-bit keys (values of pass 0 are the iota and are
 *   not read: 64 B).
 */
#include "common.cuh"

#define RS_RADIX_BITS 8
#define RS_RADIX 256
#define RS_THREADS 512
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 16
#define RS_TILE (RS_THREADS * RS_ITEMS)
#define RS_MAX_PASSES 4

#define RS_FLAG_AGG 1u
#define RS_FLAG_INC 2u
#define RS_VAL_MASK 0x3FFFFFFFu /* n < 2^30 (historical limit of the flag|count word; kept as the documented maximum) */

#define HIST_THREADS 512
#define HIST_ITEMS 16

/* scratch layout (u32 words):
 *   [0, 1024)                 bins[pass][digit]  -> exclusive offsets after the histogram launch
 *   [1024, 1028)              tile tickets per pass
 *   [1028]                    finished-CTA counter of the histogram launch
 *   [1032 + pass*nTilesPad)   flags[pass][tile]: 0 / RS_FLAG_AGG / RS_FLAG_INC, ONE word per tile        (zeroed per sort)
 *   then counts[pass][tile][2][256]: the tile's digit aggregates and inclusive prefixes                  (never zeroed)
 * A tile's 256 counts are published with plain stores followed by one release-store of its flag; a reader polls 32 flags
 * per round trip (one warp load) and then sums the published rows with independent, pipelined loads. */
#define RS_OFF_TICKET 1024
#define RS_OFF_DONE 1028
#define RS_OFF_FLAGS 1032

static inline size_t rs_tiles(u32 n) { return ((size_t)n + RS_TILE - 1) / RS_TILE; }
static inline size_t rs_tiles_pad(u32 n) { return (rs_tiles(n) + 31) & ~(size_t)31; }
size_t b2_sort_scratch_bytes(u32 n) {
  return (RS_OFF_FLAGS + (size_t)RS_MAX_PASSES * rs_tiles_pad(n) + (size_t)RS_MAX_PASSES * rs_tiles(n) * 2 * RS_RADIX) * sizeof(u32);
}

__global__ void __launch_bounds__(HIST_THREADS) radix_hist_kernel(const u32* __restrict__ keys, u32 n, u32* __restrict__ scratch, u32 startBit,
                                                                  u32 endBit, u32 nPasses) {
  __shared__ u32 h[RS_MAX_PASSES * RS_RADIX];
  __shared__ u32 isLast;
  for (u32 k = threadIdx.x; k < RS_MAX_PASSES * RS_RADIX; k += HIST_THREADS) h[k] = 0;
  __syncthreads();
  const u32 chunk = HIST_THREADS * HIST_ITEMS;
  for (u32 base = blockIdx.x * chunk; base < n; base += gridDim.x * chunk) {
    /* 16-byte loads; n may end inside a word group */
    for (u32 k = 0; k < HIST_ITEMS / 4; k++) {
      const u32 i = base + (k * HIST_THREADS + threadIdx.x) * 4;
      u32 v[4];
      u32 cnt = 0;
      if (i + 4 <= n) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(keys + i));
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        cnt = 4;
      } else {
        for (; i + cnt < n && cnt < 4; cnt++) v[cnt] = __ldg(keys + i + cnt);
      }
      for (u32 e = 0; e < cnt; e++) {
#pragma unroll
        for (u32 p = 0; p < RS_MAX_PASSES; p++) {
          if (p < nPasses) {
            const u32 shift = startBit + p * RS_RADIX_BITS;
            const u32 bits = min(RS_RADIX_BITS, endBit - shift);
            atomicAdd(&h[p * RS_RADIX + ((v[e] >> shift) & ((1u << bits) - 1u))], 1u);
          }
        }
      }
    }
  }
  __syncthreads();
  for (u32 k = threadIdx.x; k < nPasses * RS_RADIX; k += HIST_THREADS)
    if (h[k]) atomicAdd(scratch + k, h[k]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) isLast = (atom_add_acq_rel(scratch + RS_OFF_DONE, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!isLast) return;
  /* last CTA: bins -> exclusive offsets, one warp per pass (warps 0..nPasses-1), 8 bins per lane */
  const u32 w = threadIdx.x >> 5, l = lane_id();
  if (w < nPasses) {
    u32 c[8];
    u32 s = 0;
    for (u32 k = 0; k < 8; k++) { c[k] = __ldcg(scratch + w * RS_RADIX + l * 8 + k); s += c[k]; }
    u32 incl = s;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    u32 run = incl - s;
    for (u32 k = 0; k < 8; k++) { scratch[w * RS_RADIX + l * 8 + k] = run; run += c[k]; }
  }
}

/* ---- TMA bulk copy + mbarrier (PTX; cp.async.bulk -> SASS UBLKCP) ---- */
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

struct OnesweepSmem {
  u32 keys[RS_TILE];              /* raw key tile (TMA destination), then the digit-ordered keys            */
  u32 vals[RS_TILE];              /* raw value tile (TMA destination), then the digit-ordered values        */
  u32 warpHist[RS_WARPS][RS_RADIX];
  u32 digitStart[RS_RADIX];       /* tile-local exclusive prefix per digit                                   */
  int globalBase[RS_RADIX];       /* global index of the digit's first key of this tile minus digitStart     */
  u32 warpTotals[8];
  u32 tile;
  alignas(8) u64 bar[2];
};

template <bool IOTA_VALUES>
__global__ void __launch_bounds__(RS_THREADS) onesweep_pass_kernel(const u32* __restrict__ keysIn, const u32* __restrict__ valsIn,
                                                                   u32* __restrict__ keysOut, u32* __restrict__ valsOut, u32* __restrict__ scratch,
                                                                   u32 n, u32 nTiles, u32 nTilesPad, u32 pass, u32 shift, u32 mask) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  OnesweepSmem& S = *reinterpret_cast<OnesweepSmem*>(smemRaw);
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;

  if (tid == 0) {
    S.tile = atomicAdd(scratch + RS_OFF_TICKET + pass, 1u);
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (u32 k = l; k < RS_RADIX; k += 32) S.warpHist[w][k] = 0;
  __syncthreads();
  const u32 tile = S.tile;
  const u32 tileBase = tile * RS_TILE;
  u32* flags = scratch + RS_OFF_FLAGS + (size_t)pass * nTilesPad;
  u32* counts = scratch + RS_OFF_FLAGS + (size_t)RS_MAX_PASSES * nTilesPad + (size_t)pass * nTiles * 2 * RS_RADIX; /* [tile][agg|inc][digit] */
  const u32 valid = min((u32)RS_TILE, n - tileBase);
  const bool full = (valid == RS_TILE);

  if (full && tid == 0) {
    mbar_expect_tx(&S.bar[0], RS_TILE * 4);
    tma_load_1d(S.keys, keysIn + tileBase, RS_TILE * 4, &S.bar[0]);
    if (!IOTA_VALUES) {
      mbar_expect_tx(&S.bar[1], RS_TILE * 4);
      tma_load_1d(S.vals, valsIn + tileBase, RS_TILE * 4, &S.bar[1]);
    }
  }

  /* ---- keys, warp-striped: item i of lane l of warp w is tile element w*32*ITEMS + i*32 + l ---- */
  u32 key[RS_ITEMS];
  const u32 stripe = w * (32 * RS_ITEMS) + l;
  if (full) {
    mbar_wait(&S.bar[0], 0);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) key[i] = S.keys[stripe + i * 32];
  } else {
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      const u32 e = stripe + i * 32;
      key[i] = e < valid ? __ldg(keysIn + tileBase + e) : 0xFFFFFFFFu;
    }
  }

  /* ---- stable rank inside the warp: match-any on the digit, per-warp running digit counters ---- */
  u32 pos[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const u32 d = (key[i] >> shift) & mask;
    const u32 peers = __match_any_sync(B2_FULL, d);
    const u32 leader = __ffs(peers) - 1;
    u32 old = 0;
    if (l == leader) { old = S.warpHist[w][d]; S.warpHist[w][d] = old + __popc(peers); }
    old = __shfl_sync(B2_FULL, old, leader);
    pos[i] = old + __popc(peers & lanemask_lt());
    __syncwarp();
  }
  __syncthreads(); /* all raw keys are in registers, all warp histograms complete */

  /* ---- per-digit: exclusive prefix over warps, CTA count ---- */
  u32 count = 0;
  if (tid < RS_RADIX) {
#pragma unroll
    for (int k = 0; k < RS_WARPS; k++) { const u32 t = S.warpHist[k][tid]; S.warpHist[k][tid] = count; count += t; }
    /* publish this tile's digit counts as early as possible (tile 0: they are already inclusive) */
    __stcg(counts + ((size_t)tile * 2 + (tile == 0 ? 1 : 0)) * RS_RADIX + tid, count);
    /* exclusive scan of count over the 256 digits */
    u32 incl = count;
    for (int o = 1; o < 32; o <<= 1) {
      const u32 t = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += t;
    }
    if (l == 31) S.warpTotals[w] = incl;
    S.digitStart[tid] = incl - count; /* completed below */
  }
  __syncthreads();
  if (tid == 0) st_release(flags + tile, tile == 0 ? RS_FLAG_INC : RS_FLAG_AGG); /* orders the 256 count stores of the CTA before the flag */
  if (tid < RS_RADIX) {
    u32 add = 0;
    for (u32 k = 0; k < w; k++) add += S.warpTotals[k];
    S.digitStart[tid] += add;
  }
  __syncthreads();

  /* ---- keys into digit order in shared memory (overwrites the raw tile) ---- */
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const u32 d = (key[i] >> shift) & mask;
    pos[i] += S.digitStart[d] + S.warpHist[w][d];
    S.keys[pos[i]] = key[i];
  }

  /* ---- decoupled look-back: warps 0..7 (one lane per digit) poll 32 tile flags per round trip, then add up the rows ---- */
  if (tid < RS_RADIX) {
    u32 excl = 0;
    int t = (int)tile; /* exclusive end of the window */
    while (t > 0) {
      const int idx = t - 1 - (int)l;
      u32 f, firstInc;
      while (true) {
        f = idx >= 0 ? ld_acquire(flags + idx) : RS_FLAG_INC; /* virtual tile -1 */
        const u32 incM = __ballot_sync(B2_FULL, f == RS_FLAG_INC), zeroM = __ballot_sync(B2_FULL, f == 0u);
        firstInc = incM ? (u32)__ffs(incM) - 1u : 32u;
        const u32 firstZero = zeroM ? (u32)__ffs(zeroM) - 1u : 32u;
        if (firstZero > min(firstInc, 31u)) break;
      }
      const int last = (int)min(firstInc, 31u);
      const u32 incBit = (firstInc < 32u) ? (1u << firstInc) : 0u;
#pragma unroll 8
      for (int k = 0; k <= last; k++) {
        const int tt = t - 1 - k;
        if (tt < 0) break;
        excl += __ldcg(counts + ((size_t)tt * 2 + ((incBit >> k) & 1u)) * RS_RADIX + tid);
      }
      if (firstInc < 32u) break;
      t -= 32;
    }
    if (tile > 0) __stcg(counts + ((size_t)tile * 2 + 1) * RS_RADIX + tid, excl + count);
    S.globalBase[tid] = (int)(__ldg(scratch + pass * RS_RADIX + tid) + excl) - (int)S.digitStart[tid];
  }
  __syncthreads();
  if (tid == 0 && tile > 0) st_release(flags + tile, RS_FLAG_INC);

  /* ---- keys out: element j of the digit-ordered tile goes to globalBase[digit] + j ---- */
  int dst[RS_ITEMS];
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const u32 j = tid + i * RS_THREADS;
    const u32 k = S.keys[j];
    dst[i] = S.globalBase[(k >> shift) & mask] + (int)j;
    if (j < valid) keysOut[dst[i]] = k;
  }

  /* ---- values: same permutation ---- */
  u32 val[RS_ITEMS];
  if (IOTA_VALUES) {
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) val[i] = tileBase + stripe + i * 32;
  } else if (full) {
    mbar_wait(&S.bar[1], 0);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) val[i] = S.vals[stripe + i * 32];
  } else {
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
      const u32 e = stripe + i * 32;
      val[i] = e < valid ? __ldg(valsIn + tileBase + e) : 0u;
    }
  }
  __syncthreads(); /* raw values are in registers */
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) S.vals[pos[i]] = val[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < RS_ITEMS; i++) {
    const u32 j = tid + i * RS_THREADS;
    if (j < valid) valsOut[dst[i]] = S.vals[j];
  }
}

int b2_launch_sort(b2bvh_ctx* ctx, const u32* d_keysIn, const u32* d_valsIn, u32* d_keysOut, u32* d_valsOut, u32* d_keysTmp, u32* d_valsTmp,
                   void* d_scratch, u32 n, u32 startBit, u32 endBit) {
  if (n == 0 || endBit <= startBit || endBit > 32) return b2_fail(B2BVH_ERR_INVALID, "sort: bad range n=%u bits [%u,%u)", n, startBit, endBit);
  if (n > RS_VAL_MASK) return b2_fail(B2BVH_ERR_INVALID, "sort: n=%u exceeds 2^30-1", n);
  const u32 nPasses = (endBit - startBit + RS_RADIX_BITS - 1) / RS_RADIX_BITS;
  const u32 nTiles = (n + RS_TILE - 1) / RS_TILE;
  u32* scratch = reinterpret_cast<u32*>(d_scratch);
  const u32 nTilesPad = (u32)rs_tiles_pad(n);
  B2_CUDA(cudaMemsetAsync(scratch, 0, (RS_OFF_FLAGS + (size_t)RS_MAX_PASSES * nTilesPad) * sizeof(u32), ctx->stream));
  {
    const u32 chunk = HIST_THREADS * HIST_ITEMS;
    u32 grid = (n + chunk - 1) / chunk;
    const u32 cap = (u32)ctx->sm_count * 4u;
    if (grid > cap) grid = cap;
    B2_KERNEL(ctx, "radix_hist");
    radix_hist_kernel<<<grid, HIST_THREADS, 0, ctx->stream>>>(d_keysIn, n, scratch, startBit, endBit, nPasses);
    B2_LAUNCH_CHECK(ctx);
  }
  static bool attrSet = false;
  const size_t smem = sizeof(OnesweepSmem);
  if (!attrSet) {
    B2_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2_CUDA(cudaFuncSetAttribute(onesweep_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attrSet = true;
  }
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
    B2_KERNEL(ctx, "onesweep_pass");
    if (vin == nullptr)
      onesweep_pass_kernel<true><<<nTiles, RS_THREADS, smem, ctx->stream>>>(kin, nullptr, kout, vout, scratch, n, nTiles, nTilesPad, p, shift, mask);
    else
      onesweep_pass_kernel<false><<<nTiles, RS_THREADS, smem, ctx->stream>>>(kin, vin, kout, vout, scratch, n, nTiles, nTilesPad, p, shift, mask);
    B2_LAUNCH_CHECK(ctx);
    kin = kout;
    vin = vout;
  }
  return 0;
}
