/*
 * morton60.cu — the 60-bit Morton variant of stages S2 + S3 (north star "30/60-bit Morton coding").
 *
 * The reference has 30-bit codes only (computeExtendedMortonCode / computeMortonCode, CommonBlocksKernel.h:159-372): at 100 M
 * primitives neighbours share codes and the order below the code resolution is the input order.  The 60-bit code defined here
 * is the PLAIN interleave of computeMortonCode (:361-372) with 20 bits per axis instead of 10:
 *     q = (u32) min(max(p * 2^20, 0), 2^20 - 1)  per axis,   code = interleave(qx, qy, qz), x highest
 * which equals  interleave10(q >> 10) << 30 | interleave10(q & 1023):  two 30-bit words, HI and LO.
 *
 * The sort keeps the 32-bit radix sort (radix_sort.cu) untouched: an LSD sort over two 30-bit "digits" —
 *     (LO, iota) -> sort -> (., P1);   HI1[g] = HI[P1[g]];   (HI1, P1) -> stable sort -> (HI sorted, P)
 * — which is the stable order by (HI, LO, index); the 64-bit sorted keys the hierarchy needs are then HIs[g] << 30 | LO[P[g]].
 * 8 digit passes of 8 bits, the count a one-pass 64-bit sort would need for 60 bits, plus two 4-byte gathers.
 */
#include "common.cuh"
#include "morton.cuh"

#define M60_THREADS 256

__global__ void __launch_bounds__(M60_THREADS) morton60_kernel(const b2bvh_aabb* __restrict__ triAabb, const b2bvh_aabb* __restrict__ scene, u32 n,
                                                               u32* __restrict__ hi, u32* __restrict__ lo, u64* __restrict__ keys64) {
  const float* s = reinterpret_cast<const float*>(scene);
  const float mnx = __ldg(s), mny = __ldg(s + 1), mnz = __ldg(s + 2);
  const float ex = __fsub_rn(__ldg(s + 3), mnx), ey = __fsub_rn(__ldg(s + 4), mny), ez = __fsub_rn(__ldg(s + 5), mnz);
  for (u32 i = blockIdx.x * M60_THREADS + threadIdx.x; i < n; i += gridDim.x * M60_THREADS) {
    const float2* q = reinterpret_cast<const float2*>(triAabb + i);
    const float2 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2); /* (lx,ly) (lz,hx) (hy,hz) */
    const float px = __fdiv_rn(__fsub_rn(__fmul_rn(0.5f, __fadd_rn(b.y, a.x)), mnx), ex);
    const float py = __fdiv_rn(__fsub_rn(__fmul_rn(0.5f, __fadd_rn(d.x, a.y)), mny), ey);
    const float pz = __fdiv_rn(__fsub_rn(__fmul_rn(0.5f, __fadd_rn(d.y, b.x)), mnz), ez);
    /* 0/0 = NaN -> 0 through fmaxf, as in the 30-bit path */
    const u32 qx = (u32)fminf(fmaxf(__fmul_rn(px, 1048576.0f), 0.0f), 1048575.0f);
    const u32 qy = (u32)fminf(fmaxf(__fmul_rn(py, 1048576.0f), 0.0f), 1048575.0f);
    const u32 qz = (u32)fminf(fmaxf(__fmul_rn(pz, 1048576.0f), 0.0f), 1048575.0f);
    const u32 h = interleave3(qx >> 10) * 4u + interleave3(qy >> 10) * 2u + interleave3(qz >> 10);
    const u32 l = interleave3(qx & 1023u) * 4u + interleave3(qy & 1023u) * 2u + interleave3(qz & 1023u);
    hi[i] = h;
    lo[i] = l;
    keys64[i] = ((u64)h << 30) | l;
  }
}

__global__ void __launch_bounds__(M60_THREADS) gather_u32_kernel(const u32* __restrict__ src, const u32* __restrict__ idx, u32 n, u32* __restrict__ dst) {
  for (u32 g = blockIdx.x * M60_THREADS + threadIdx.x; g < n; g += gridDim.x * M60_THREADS) dst[g] = ldg_gather_u32(src + __ldg(idx + g));
}
__global__ void __launch_bounds__(M60_THREADS) combine60_kernel(const u32* __restrict__ hiSorted, const u32* __restrict__ lo, const u32* __restrict__ perm, u32 n,
                                                                u64* __restrict__ keys64Sorted) {
  for (u32 g = blockIdx.x * M60_THREADS + threadIdx.x; g < n; g += gridDim.x * M60_THREADS)
    keys64Sorted[g] = ((u64)__ldg(hiSorted + g) << 30) | ldg_gather_u32(lo + __ldg(perm + g));
}

static u32 m60_grid(const b2bvh_ctx* ctx, u32 n) {
  u32 grid = (n + M60_THREADS - 1) / M60_THREADS;
  const u32 cap = (u32)ctx->sm_count * 16u;
  return grid > cap ? cap : grid;
}

int b2_launch_morton60(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const b2bvh_aabb* d_scene, u32 n, u32* d_hi, u32* d_lo, u64* d_keys64) {
  B2_KERNEL(ctx, "morton60");
  morton60_kernel<<<m60_grid(ctx, n), M60_THREADS, 0, ctx->stream>>>(d_triAabb, d_scene, n, d_hi, d_lo, d_keys64);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

/* d_a / d_aVals: two n-word work arrays; d_hiSorted / d_valsSorted / d_keys64Sorted: the outputs */
int b2_launch_sort60(b2bvh_ctx* ctx, const u32* d_hi, const u32* d_lo, u32 n, u32* d_a, u32* d_aVals, u32* d_hiSorted, u32* d_valsSorted, u64* d_keys64Sorted,
                     u32* d_keysTmp, u32* d_valsTmp, void* d_sortScratch) {
  /* digit 0: LO, values = iota (not read) */
  B2_TRY(b2_launch_sort(ctx, d_lo, nullptr, d_a, d_aVals, d_keysTmp, d_valsTmp, d_sortScratch, n, 0, 30)); /* both words of the code have 30 bits */
  B2_KERNEL(ctx, "morton60_gather_hi");
  gather_u32_kernel<<<m60_grid(ctx, n), M60_THREADS, 0, ctx->stream>>>(d_hi, d_aVals, n, d_a); /* the sorted LO words are not needed again */
  B2_LAUNCH_CHECK(ctx);
  /* digit 1: HI in LO order, stable */
  B2_TRY(b2_launch_sort(ctx, d_a, d_aVals, d_hiSorted, d_valsSorted, d_keysTmp, d_valsTmp, d_sortScratch, n, 0, 30)); /* both words of the code have 30 bits */
  B2_KERNEL(ctx, "morton60_combine");
  combine60_kernel<<<m60_grid(ctx, n), M60_THREADS, 0, ctx->stream>>>(d_hiSorted, d_lo, d_valsSorted, n, d_keys64Sorted);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
