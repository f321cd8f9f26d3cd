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

struct SplitCtl {
  u32 ticket;   /* next tile */
  u32 pad;
  u32 totals[2]; /* accepted, rejected of the generation (written by the last tile) */
  u32 nonFinite; /* a rejected box has an infinite or NaN area: halving it never ends (the reference's loop would not terminate) */
};

__global__ void __launch_bounds__(SPLIT_THREADS) split_level_kernel(const b2bvh_aabb* __restrict__ inBox, const u32* __restrict__ inPrim /* NULL: identity */,
                                                                    u32 count, float saMax, b2bvh_aabb* outBox, u32* outPrim, u32 outBase,
                                                                    b2bvh_aabb* nextBox, u32* nextPrim, u64* status, SplitCtl* ctl) {
  __shared__ u32 sTile;
  __shared__ u32 sWarp[2][SPLIT_THREADS / 32];
  __shared__ u32 sBase[2];
  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  if (tid == 0) sTile = atomicAdd(&ctl->ticket, 1u); /* ticket order == tile order: a tile only waits for tiles that already run */
  __syncthreads();
  const u32 tile = sTile;
  const u32 i = tile * SPLIT_THREADS + tid;
  const bool valid = i < count;
  Box b = box_empty();
  u32 prim = i;
  bool accept = false;
  if (valid) {
    b = load_aabb(inBox + i);
    if (inPrim) prim = __ldg(inPrim + i);
    const float area = box_area(b);
    accept = area <= saMax; /* Utility.cpp:477 (a NaN area is never accepted, as in the reference) */
    if (!accept && !(area < __int_as_float(0x7f800000))) ctl->nonFinite = 1u;
  }
  const bool reject = valid && !accept;
  const u32 balA = __ballot_sync(B2_FULL, accept), balR = __ballot_sync(B2_FULL, reject);
  if (lane == 0) { sWarp[0][warp] = __popc(balA); sWarp[1][warp] = __popc(balR); }
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
  if (accept) {
    const u32 o = outBase + sBase[0] + sWarp[0][warp] + __popc(balA & lanemask_lt());
    store_aabb(outBox + o, b);
    outPrim[o] = prim;
  } else if (reject) {
    const u32 o = 2u * (sBase[1] + sWarp[1][warp] + __popc(balR & lanemask_lt()));
    /* Aabb::maximumExtentDim / center (Common.h:347-359), Utility.cpp:483-527 */
    const float ex = __fsub_rn(b.hx, b.lx), ey = __fsub_rn(b.hy, b.ly), ez = __fsub_rn(b.hz, b.lz);
    const int dim = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);
    Box L = b, R = b;
    if (dim == 0) { const float c = __fmul_rn(__fadd_rn(b.hx, b.lx), 0.5f); L.hx = c; R.lx = c; }
    if (dim == 1) { const float c = __fmul_rn(__fadd_rn(b.hy, b.ly), 0.5f); L.hy = c; R.ly = c; }
    if (dim == 2) { const float c = __fmul_rn(__fadd_rn(b.hz, b.lz), 0.5f); L.hz = c; R.lz = c; }
    store_aabb(nextBox + o, L);
    store_aabb(nextBox + o + 1, R);
    *reinterpret_cast<uint2*>(nextPrim + o) = make_uint2(prim, prim);
  }
}

/* after the build over split references: Bvh2 leaf g names reference sortedVals[g]; the reference's leaf carries the
 * TRIANGLE (InitBvhNodesPrimRef: node.m_leftChildIdx = primitives[idx].m_primIdx, TwoPassLbvhKernel.h:178-182).
 * leafPrim[g] is the same id as a dense array for the collapse (PrimNode.m_primIdx = leaf.m_leftChildIdx, :324). */
__global__ void __launch_bounds__(256) split_remap_kernel(const u32* __restrict__ sortedVals, const u32* __restrict__ refPrim, u32 n,
                                                          b2bvh_bvh2_node* nodes, u32* leafPrim) {
  const u32 g = blockIdx.x * 256u + threadIdx.x;
  if (g >= n) return;
  const u32 p = ldg_gather_u32(refPrim + __ldg(sortedVals + g));
  leafPrim[g] = p;
  nodes[(n - 1) + g].m_leftChildIdx = p;
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
    const u32 tiles = (count + SPLIT_THREADS - 1) / SPLIT_THREADS;
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

int b2_launch_split_remap(b2bvh_ctx* ctx, const u32* d_sortedVals, const u32* d_refPrim, u32 n, b2bvh_bvh2_node* d_nodes, u32* d_leafPrim) {
  B2_KERNEL(ctx, "split_remap");
  split_remap_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_sortedVals, d_refPrim, n, d_nodes, d_leafPrim);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
