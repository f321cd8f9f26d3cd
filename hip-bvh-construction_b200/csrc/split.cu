/*
 * split.cu — early split clipping on the device: the step in front of the build path when the reference is compiled
 * with USE_PRIM_SPLITTING (TwoPassLbvh.cpp:23-28 → Utility::doEarlySplitClipping, Utility.cpp:456-538, a host FIFO).
 *
 * The reference pops one PrimRef at a time: area(box) <= saMax → emit, otherwise cut the box at its middle along the
 * largest extent and push both halves (same m_primIdx) to the back of the queue.  A FIFO visits the references
 * generation by generation, and inside a generation in queue order, so the emitted array is
 *     [accepted of generation 0 in order][accepted of generation 1 in order] ...
 * and generation g+1 is the list of (left half, right half) of the rejected references of generation g, in order.
 * That is one order-preserving two-way partition per generation: split_level_kernel does it in ONE pass over the
 * generation's list — classification, a CTA scan of the two flags, a decoupled look-back (lookback.cuh) over the packed
 * (accepted, rejected) tile totals, and the scatter of both outputs.  The host loop needs the totals of a generation to
 * size the next one, so there is one stream synchronisation per generation (a handful: log2(largest area / saMax)).
 *
 * Floating-point contract: area() and center() as the host code evaluates them — every operation rounded on its own
 * (common.cuh box_area; centre = (max + min) * 0.5f).  Output is compared byte for byte with the reference's function.
 */
#include "common.cuh"
#include "lookback.cuh"

#define SPLIT_THREADS 512
#ifndef SPLIT_MINB
#define SPLIT_MINB 2                                    /* resident CTAs per SM (2/3/4: 115/115/122 us per generation) */
#endif
#ifndef SPLIT_GROUP
#define SPLIT_GROUP 1                                   /* references per thread re-read together in pass 2 (measured 1/2/4/8: 145/148/164/176 us per generation) */
#endif
#ifndef SPLIT_ITEMS
#define SPLIT_ITEMS 8                                   /* references per thread */
#endif
#define SPLIT_TILE (SPLIT_THREADS * SPLIT_ITEMS)        /* 4096 references per tile.  The look-back costs a few L2 round trips PER TILE: with
                                                           512-reference tiles those round trips were the whole run time (458 us for 10 M boxes,
                                                           86 % of the warp stalls at the barrier behind the look-back, profiles/r01q); per
                                                           generation of the 10 M case: 4096 per tile 113 us, 8192: 123 us, 16384: 163 us
                                                           (too few tiles per SM to hide the two passes behind each other) */

/* 24-byte boxes are 8-byte aligned: three 8-byte accesses instead of six 4-byte ones halve the L2 requests */
__device__ __forceinline__ Box split_load_box(const b2bvh_aabb* p) {
  const float2* q = reinterpret_cast<const float2*>(p);
  const float2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  return Box{a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ void split_store_box(b2bvh_aabb* p, const Box& b) {
  float2* q = reinterpret_cast<float2*>(p);
  q[0] = make_float2(b.lx, b.ly); q[1] = make_float2(b.lz, b.hx); q[2] = make_float2(b.hy, b.hz);
}

struct SplitCtl {
  u32 ticket;   /* next tile */
  u32 pad;
  u32 totals[2]; /* accepted, rejected of the generation (written by the last tile) */
  u32 nonFinite; /* a rejected box has an infinite or NaN area: halving it never ends (the reference's loop would not terminate) */
};

/* warp w of a tile owns 32 x SPLIT_ITEMS consecutive references; reference (k, lane) = warpBase + 32 k + lane, so every load is
 * coalesced and the order inside the warp is k-major.  Pass 1 keeps only the two ballots per k; pass 2 reads the boxes again
 * (L2 hits: the tile was just streamed) and scatters them. */
__global__ void __launch_bounds__(SPLIT_THREADS, SPLIT_MINB) split_level_kernel(const b2bvh_aabb* __restrict__ inBox, const u32* __restrict__ inPrim /* NULL: identity */,
                                                                    u32 count, float saMax, b2bvh_aabb* outBox, u32* outPrim, u32 outBase,
                                                                    b2bvh_aabb* nextBox, u32* nextPrim, u64* status, SplitCtl* ctl) {
  __shared__ u32 sTile;
  __shared__ u32 sWarp[2][SPLIT_THREADS / 32];
  __shared__ u32 sBase[2];
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  if (tid == 0) sTile = atomicAdd(&ctl->ticket, 1u); /* ticket order == tile order: a tile only waits for tiles that already run */
  __syncthreads();
  const u32 tile = sTile;
  const u32 warpBase = tile * SPLIT_TILE + warp * (32u * SPLIT_ITEMS);
  u32 balA[SPLIT_ITEMS], balR[SPLIT_ITEMS];
  u32 wA = 0, wR = 0;
  bool bad = false;
#pragma unroll
  for (int k = 0; k < SPLIT_ITEMS; k++) {
    const u32 i = warpBase + 32u * k + lane;
    bool accept = false, reject = false;
    if (i < count) {
      const float area = box_area(split_load_box(inBox + i));
      accept = area <= saMax; /* Utility.cpp:477 (a NaN area is never accepted, as in the reference) */
      reject = !accept;
      bad |= reject && !(area < __int_as_float(0x7f800000));
    }
    balA[k] = __ballot_sync(B2_FULL, accept);
    balR[k] = __ballot_sync(B2_FULL, reject);
    wA += __popc(balA[k]);
    wR += __popc(balR[k]);
  }
  if (bad) ctl->nonFinite = 1u;
  if (lane == 0) { sWarp[0][warp] = wA; sWarp[1][warp] = wR; }
  __syncthreads();
  if (warp == 0) {
    /* exclusive scan of the 16 per-warp counts, both flags at once (lanes 0-15: accepted, 16-31: rejected) */
    const u32 w = lane & 15u, which = lane >> 4;
    const u32 own = sWarp[which][w];
    u32 inc = own;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const u32 v = __shfl_up_sync(B2_FULL, inc, o, 16);
      if (w >= (u32)o) inc += v;
    }
    sWarp[which][w] = inc - own;
    const u32 totA = __shfl_sync(B2_FULL, inc, 15), totR = __shfl_sync(B2_FULL, inc, 31);
    const u64 mine = ((u64)totR << 31) | (u64)totA;
    if (lane == 0) st_release64(status + tile, (tile == 0 ? LB64_INC : LB64_AGG) | mine);
    const u64 excl = warp_lookback_u64(status, tile);
    if (lane == 0) {
      if (tile != 0) st_release64(status + tile, LB64_INC | (excl + mine));
      sBase[0] = (u32)(excl & 0x7FFFFFFFull);
      sBase[1] = (u32)(excl >> 31);
      if (tile == gridDim.x - 1) { ctl->totals[0] = sBase[0] + totA; ctl->totals[1] = sBase[1] + totR; }
    }
  }
  __syncthreads();
  u32 runA = outBase + sBase[0] + sWarp[0][warp], runR = sBase[1] + sWarp[1][warp];
  const u32 lt = lanemask_lt();
  /* four references per thread in flight: the re-read is latency-bound when every load waits for the store before it */
#pragma unroll
  for (int k0 = 0; k0 < SPLIT_ITEMS; k0 += SPLIT_GROUP) {
    Box bx[SPLIT_GROUP];
    u32 pr[SPLIT_GROUP];
#pragma unroll
    for (int j = 0; j < SPLIT_GROUP; j++) {
      const u32 i = warpBase + 32u * (k0 + j) + lane;
      if (i < count) { bx[j] = split_load_box(inBox + i); pr[j] = inPrim ? __ldg(inPrim + i) : i; }
    }
#pragma unroll
    for (int j = 0; j < SPLIT_GROUP; j++) {
      const int k = k0 + j;
      const bool accept = (balA[k] >> lane) & 1u, reject = (balR[k] >> lane) & 1u;
      const Box b = bx[j];
      if (accept) {
        const u32 o = runA + __popc(balA[k] & lt);
        split_store_box(outBox + o, b);
        outPrim[o] = pr[j];
      } else if (reject) {
        const u32 o = 2u * (runR + __popc(balR[k] & lt));
        /* Aabb::maximumExtentDim / center (Common.h:347-359), Utility.cpp:483-527 */
        const float ex = __fsub_rn(b.hx, b.lx), ey = __fsub_rn(b.hy, b.ly), ez = __fsub_rn(b.hz, b.lz);
        const int dim = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
        Box L = b, R = b;
        if (dim == 0) { const float c = __fmul_rn(__fadd_rn(b.hx, b.lx), 0.5f); L.hx = c; R.lx = c; }
        if (dim == 1) { const float c = __fmul_rn(__fadd_rn(b.hy, b.ly), 0.5f); L.hy = c; R.ly = c; }
        if (dim == 2) { const float c = __fmul_rn(__fadd_rn(b.hz, b.lz), 0.5f); L.hz = c; R.lz = c; }
        split_store_box(nextBox + o, L);
        split_store_box(nextBox + o + 1, R);
        *reinterpret_cast<uint2*>(nextPrim + o) = make_uint2(pr[j], pr[j]);
      }
      runA += __popc(balA[k]);
      runR += __popc(balR[k]);
    }
  }
}

/* grow a build-owned buffer and KEEP its first `keep` bytes (b2_reserve discards the contents) */
static int split_grow(b2bvh_ctx* ctx, int slot, size_t bytes, size_t keep, void** out) {
  b2bvh_ctx::Buf& b = ctx->bufs[slot];
  if (b.cap >= bytes) { *out = b.p; return 0; }
  size_t want = b.cap + b.cap / 2;
  if (want < bytes) want = bytes;
  want = (want + 255) & ~(size_t)255;
  void* p = nullptr;
  B2_CUDA(cudaMalloc(&p, want));
  if (b.p) {
    if (keep) B2_CUDA(cudaMemcpyAsync(p, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    B2_CUDA(cudaFree(b.p));
  }
  b.p = p; b.cap = want;
  ctx->alloc_epoch++;
  *out = p;
  return 0;
}

int b2_launch_split(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, u32 n, float saMax, int slotOutBox, int slotOutPrim, int slotListA, int slotListB,
                    int slotStatus, b2bvh_aabb** d_refBox, u32** d_refPrim, u32* h_count, u32* h_levels) {
  cudaStream_t s = ctx->stream;
  const b2bvh_aabb* inBox = d_triAabb;
  const u32* inPrim = nullptr; /* generation 0: reference i is triangle i */
  u32 count = n, outCount = 0, level = 0;
  const int listSlot[2] = {slotListA, slotListB};
  int cur = 0; /* generation g+1 is written to listSlot[cur], generation g+2 to the other: the list being read is never the one reserved */
  void *outBox = nullptr, *outPrim = nullptr, *status = nullptr, *list = nullptr;
  while (count) {
    if (level >= 64)
      return b2_fail(B2BVH_ERR_INVALID, "early split: a box is still larger than saMax=%g after 64 generations (the reference would not terminate)", (double)saMax);
    if ((u64)outCount + count > 0x3FFFFFFFull)
      return b2_fail(B2BVH_ERR_INVALID, "early split: more than 2^30-1 references (saMax=%g is too small for this scene)", (double)saMax);
    /* this generation emits at most `count` references and hands at most 2*count to the next one */
    B2_TRY(split_grow(ctx, slotOutBox, ((size_t)outCount + count) * sizeof(b2bvh_aabb), (size_t)outCount * sizeof(b2bvh_aabb), &outBox));
    B2_TRY(split_grow(ctx, slotOutPrim, ((size_t)outCount + count) * 4, (size_t)outCount * 4, &outPrim));
    const size_t primOff = (2 * (size_t)count * sizeof(b2bvh_aabb) + 15) & ~(size_t)15;
    B2_TRY(b2_reserve(ctx, listSlot[cur], primOff + 2 * (size_t)count * 4, &list));
    b2bvh_aabb* nextBox = (b2bvh_aabb*)list;
    u32* nextPrim = (u32*)((unsigned char*)list + primOff);
    const u32 tiles = (count + SPLIT_TILE - 1) / SPLIT_TILE;
    B2_TRY(b2_reserve(ctx, slotStatus, (size_t)tiles * 8 + 64, &status));
    B2_CUDA(cudaMemsetAsync(status, 0, (size_t)tiles * 8 + 64, s));
    SplitCtl* ctl = (SplitCtl*)status;
    u64* st = (u64*)((unsigned char*)status + 64);
    B2_KERNEL(ctx, "split_level");
    split_level_kernel<<<tiles, SPLIT_THREADS, 0, s>>>(inBox, inPrim, count, saMax, (b2bvh_aabb*)outBox, (u32*)outPrim, outCount, nextBox, nextPrim, st, ctl);
    B2_LAUNCH_CHECK(ctx);
    B2_TRY(b2_fetch_words(ctx, ctl->totals, 3, B2_MB_SPLIT));
    B2_CUDA(cudaStreamSynchronize(s)); /* the next generation's size decides its launch and its buffers */
    const u32 acc = b2_mailbox(ctx, B2_MB_SPLIT)[0], rej = b2_mailbox(ctx, B2_MB_SPLIT)[1];
    if (b2_mailbox(ctx, B2_MB_SPLIT)[2])
      return b2_fail(B2BVH_ERR_INVALID, "early split: a primitive box has an infinite or NaN area (the reference's split loop would not terminate)");
    if (acc + rej != count) return b2_fail(B2BVH_ERR_INTERNAL, "early split: generation %u classified %u + %u of %u references", level, acc, rej, count);
    outCount += acc;
    if (rej > 0x3FFFFFFFu / 2) return b2_fail(B2BVH_ERR_INVALID, "early split: more than 2^30-1 references (saMax=%g is too small for this scene)", (double)saMax);
    count = 2 * rej;
    inBox = nextBox; inPrim = nextPrim;
    cur ^= 1;
    level++;
  }
  *d_refBox = (b2bvh_aabb*)outBox;
  *d_refPrim = (u32*)outPrim;
  *h_count = outCount;
  *h_levels = level;
  return 0;
}
