/*
 * lookback.cuh — decoupled look-back with a warp-wide window.
 *
 * Tiles of one launch are claimed in ticket order; tile t publishes a status word = flag | value, first its own
 * AGGREGATE (flag AGG), later its INCLUSIVE prefix (flag INC).  The exclusive prefix of tile t is the sum of the words
 * of tiles t-1, t-2, ... down to and including the first INC word.  A single thread walking that chain pays one L2 round
 * trip per predecessor, and when all SMs run in lock-step (every wave of CTAs publishes its aggregates at the same
 * moment) the chain is as long as half a wave.  Here a full warp reads 32 predecessors per round trip and reduces them
 * with ballots and shuffles, so the walk costs (chain length / 32) round trips.
 *
 * Replaces the serial inter-block chains of the reference: Ploc++Kernel.h:341-347 (compaction offsets) and the
 * gIsReady spin of Orochi's ParallelExclusiveScanAllWG (RadixSortKernels.h:606-637).
 */
#pragma once
#include "common.cuh"

/* 32-bit status words: [31:30] flags, [29:0] value.  Must be called by all 32 lanes of a warp. */
#define LB_AGG 0x40000000u
#define LB_INC 0x80000000u
#define LB_VAL 0x3FFFFFFFu

__device__ __forceinline__ u32 warp_lookback_u32(const u32* status, u32 tile) {
  const u32 lane = lane_id();
  u32 excl = 0;
  int t = (int)tile; /* exclusive end of the window */
  while (t > 0) {
    const int idx = t - 1 - (int)lane;
    u32 v, firstInc;
    while (true) {
      v = idx >= 0 ? ld_relaxed(status + idx) : LB_INC; /* virtual tile -1: inclusive prefix 0 */
      const u32 incM = __ballot_sync(B2_FULL, (v & LB_INC) != 0u), zeroM = __ballot_sync(B2_FULL, (v & (LB_AGG | LB_INC)) == 0u);
      firstInc = incM ? (u32)__ffs(incM) - 1u : 32u;
      const u32 firstZero = zeroM ? (u32)__ffs(zeroM) - 1u : 32u;
      if (firstZero > min(firstInc, 31u)) break; /* every word up to the first INC is published */
    }
    u32 c = (lane <= firstInc) ? (v & LB_VAL) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(B2_FULL, c, o);
    excl += c;
    if (firstInc < 32u) break;
    t -= 32;
  }
  return excl;
}

/* 64-bit status words: [63:62] flags, [61:0] two packed 31-bit counters (sums never carry between them). */
#define LB64_AGG (1ull << 62)
#define LB64_INC (2ull << 62)
#define LB64_FLAGS (3ull << 62)

__device__ __forceinline__ u64 warp_lookback_u64(const u64* status, u32 tile) {
  const u32 lane = lane_id();
  u64 excl = 0;
  int t = (int)tile;
  while (t > 0) {
    const int idx = t - 1 - (int)lane;
    u64 v;
    u32 firstInc;
    while (true) {
      v = idx >= 0 ? ld_acquire64(status + idx) : LB64_INC;
      const u32 incM = __ballot_sync(B2_FULL, (v & LB64_INC) != 0ull), zeroM = __ballot_sync(B2_FULL, (v & LB64_FLAGS) == 0ull);
      firstInc = incM ? (u32)__ffs(incM) - 1u : 32u;
      const u32 firstZero = zeroM ? (u32)__ffs(zeroM) - 1u : 32u;
      if (firstZero > min(firstInc, 31u)) break;
    }
    u64 c = (lane <= firstInc) ? (v & ~LB64_FLAGS) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(B2_FULL, c, o);
    excl += c;
    if (firstInc < 32u) break;
    t -= 32;
  }
  return excl;
}
