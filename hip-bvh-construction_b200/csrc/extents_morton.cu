/*
 * extents_morton.cu — stage S1 (primitive boxes + scene box) and stage S2 (extended 30-bit Morton codes).
 *
 * S1 replaces  Utility::doEarlySplitClipping (host, no-split path, Utility.cpp:456-476)
 *            + CalculateSceneExtents / CalculatePrimRefExtents (CommonBlocksKernel.h:92-137)
 *            + Aabb::atomicGrow (Common.h:400-408).
 *    One pass over the 64-byte triangles: two 16-byte loads + one 4-byte load per triangle, boxes staged in
 *    shared memory and written back as full 16-byte words, scene box reduced warp -> CTA -> 6 ordered-int
 *    atomics per CTA; the last CTA to finish decodes the box (no host round trip, no float atomics).
 *    Algorithmic traffic: 64 B read + 24 B written per primitive.
 * S2 replaces  computeExtendedMortonCode + CalculateMortonCodes[PrimRef] (CommonBlocksKernel.h:159-398).
 *    The bit allocation depends only on the scene extent: it is derived once per CTA by one thread and
 *    broadcast through shared memory.  Traffic: 24 B read + 8 B written per primitive.
 */
#include "common.cuh"

#define EXT_THREADS 256

/* scratch8 layout (u32): [0..2] = ~ordered(min.xyz), [3..5] = ordered(max.xyz), [6] = finished-CTA counter.
 * All-zero is the neutral element for atomicMax on every slot; the last CTA restores it. */
__global__ void __launch_bounds__(EXT_THREADS) primref_extents_kernel(const b2bvh_triangle* __restrict__ tris, u32 n,
                                                                      b2bvh_aabb* __restrict__ triAabb, b2bvh_aabb* __restrict__ scene,
                                                                      u32* __restrict__ scratch8, float* __restrict__ negminMax6) {
  __shared__ __align__(16) float stage[EXT_THREADS * 6];
  __shared__ float red[6][EXT_THREADS / 32];
  __shared__ u32 isLast;
  Box acc = box_empty();
  const u32 nTiles = (n + EXT_THREADS - 1) / EXT_THREADS;
  for (u32 tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const u32 base = tile * EXT_THREADS;
    const u32 i = base + threadIdx.x;
    Box b = box_empty();
    if (i < n) {
      const float4* p = reinterpret_cast<const float4*>(tris + i);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      const float v3z = __ldg(reinterpret_cast<const float*>(p + 2));
      /* v1 = (a.x,a.y,a.z)  v2 = (a.w,c.x,c.y)  v3 = (c.z,c.w,v3z) */
      b.lx = fminf(fminf(fminf(b.lx, a.x), a.w), c.z); b.hx = fmaxf(fmaxf(fmaxf(b.hx, a.x), a.w), c.z);
      b.ly = fminf(fminf(fminf(b.ly, a.y), c.x), c.w); b.hy = fmaxf(fmaxf(fmaxf(b.hy, a.y), c.x), c.w);
      b.lz = fminf(fminf(fminf(b.lz, a.z), c.y), v3z); b.hz = fmaxf(fmaxf(fmaxf(b.hz, a.z), c.y), v3z);
      acc = box_union(acc, b);
    }
    if (base + EXT_THREADS <= n) {
      float* s = stage + threadIdx.x * 6;
      s[0] = b.lx; s[1] = b.ly; s[2] = b.lz; s[3] = b.hx; s[4] = b.hy; s[5] = b.hz;
      __syncthreads();
      float4* dst = reinterpret_cast<float4*>(triAabb + base); /* base*24 B is a multiple of 16 */
      const float4* src = reinterpret_cast<const float4*>(stage);
      for (u32 k = threadIdx.x; k < EXT_THREADS * 6 / 4; k += EXT_THREADS) dst[k] = src[k];
      __syncthreads();
    } else if (i < n) {
      store_aabb(triAabb + i, b);
    }
  }
  /* warp -> CTA reduction */
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.lx = fminf(acc.lx, __shfl_xor_sync(B2_FULL, acc.lx, o)); acc.ly = fminf(acc.ly, __shfl_xor_sync(B2_FULL, acc.ly, o));
    acc.lz = fminf(acc.lz, __shfl_xor_sync(B2_FULL, acc.lz, o)); acc.hx = fmaxf(acc.hx, __shfl_xor_sync(B2_FULL, acc.hx, o));
    acc.hy = fmaxf(acc.hy, __shfl_xor_sync(B2_FULL, acc.hy, o)); acc.hz = fmaxf(acc.hz, __shfl_xor_sync(B2_FULL, acc.hz, o));
  }
  const u32 w = threadIdx.x >> 5;
  if (lane_id() == 0) { red[0][w] = acc.lx; red[1][w] = acc.ly; red[2][w] = acc.lz; red[3][w] = acc.hx; red[4][w] = acc.hy; red[5][w] = acc.hz; }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = red[threadIdx.x][0];
    for (u32 k = 1; k < EXT_THREADS / 32; k++) v = threadIdx.x < 3 ? fminf(v, red[threadIdx.x][k]) : fmaxf(v, red[threadIdx.x][k]);
    const u32 key = threadIdx.x < 3 ? ~float_to_ordered(v) : float_to_ordered(v);
    atomicMax(scratch8 + threadIdx.x, key);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) isLast = (atom_add_acq_rel(scratch8 + 6, 1u) == gridDim.x - 1);
  __syncthreads();
  if (isLast && threadIdx.x < 6) {
    const u32 key = atomicExch(scratch8 + threadIdx.x, 0u); /* read and restore the neutral element */
    const float v = threadIdx.x < 3 ? ordered_to_float(~key) : ordered_to_float(key);
    if (scene) reinterpret_cast<float*>(scene)[threadIdx.x] = v;
    if (negminMax6) negminMax6[threadIdx.x] = threadIdx.x < 3 ? -v : v;
    if (threadIdx.x == 0) scratch8[6] = 0u;
  }
}

/* ------------------------------------------------------------------------------------------- Morton */
struct MortonCfg {
  int axis0, axis1, axis2; /* axis0 has the largest extent */
  int preX, preY;          /* leading bits given to axis0 alone / to axis0+axis1 pairs */
  int swap, sum;
  int nbX, nbY, nbZ;
};

/* (int)log2f(a/b): float log2 (taken as the double-precision log2 rounded to float, which is what a correctly
 * rounded log2f returns), then the hardware float->int conversion, which saturates (NaN -> 0, +-inf -> INT_MAX/MIN)
 * exactly like the reference's GPU path; the integer arithmetic that follows wraps (done in unsigned). */
__device__ int ilog2_ratio(float a, float b) {
  const float r = a / b;
  const float l = (float)log2((double)r);
  return __float2int_rz(l);
}
__device__ __forceinline__ int wadd(int a, int b) { return (int)((u32)a + (u32)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((u32)a - (u32)b); }
__device__ __forceinline__ int wmul2(int a) { return (int)((u32)a * 2u); }

__device__ void morton_make_cfg(const float ex, const float ey, const float ez, MortonCfg& c) {
  /* order axes by extent, strict '<' so ties resolve as in CommonBlocksKernel.h:167-250 */
  const int k = (ex < ey ? 4 : 0) | (ex < ez ? 2 : 0) | (ey < ez ? 1 : 0);
  /* packed permutations, 2 bits per axis slot, indexed by k */
  const int tab0 = (0 << 0) | (0 << 2) | (0 << 4) | (2 << 6) | (1 << 8) | (1 << 10) | (1 << 12) | (2 << 14);
  const int tab1 = (1 << 0) | (2 << 2) | (1 << 4) | (0 << 6) | (0 << 8) | (0 << 10) | (2 << 12) | (1 << 14);
  const int tab2 = (2 << 0) | (1 << 2) | (2 << 4) | (1 << 6) | (2 << 8) | (2 << 10) | (0 << 12) | (0 << 14);
  c.axis0 = (tab0 >> (2 * k)) & 3; c.axis1 = (tab1 >> (2 * k)) & 3; c.axis2 = (tab2 >> (2 * k)) & 3;
  const float e[3] = {ex, ey, ez};
  const float e0 = e[c.axis0], e1 = e[c.axis1], e2 = e[c.axis2];
  int px = ilog2_ratio(e0, e1), py = ilog2_ratio(e1, e2);
  const int pz = ilog2_ratio(e0, e2);
  int swap = wsub(pz, wadd(px, py));
  px = min(px, 30);
  py = min(wmul2(py), wsub(30, px)) / 2;
  int sum = wadd(px, wmul2(py));
  if (sum != 30) sum = wadd(sum, swap); else swap = 0;
  const int nbz = (e2 != 0.0f) ? max(0, wsub(30, sum) / 3) : 0;
  int nbx, nby;
  if (swap > 0) { nbx = max(0, wadd(wadd(wadd(wsub(wsub(30, nbz), sum) / 2, py), px), 1)); nby = wsub(wsub(30, nbx), nbz); }
  else { nby = max(0, wadd(wsub(wsub(30, nbz), sum) / 2, py)); nbx = wsub(wsub(30, nby), nbz); }
  c.preX = px; c.preY = py; c.swap = swap; c.sum = sum; c.nbX = nbx; c.nbY = nby; c.nbZ = nbz;
}

__device__ __forceinline__ u32 shl_s(u32 v, int n) { return (n < 0 || n > 31) ? 0u : (v << n); }
__device__ __forceinline__ u32 shr_s(u32 v, int n) { return (n < 0 || n > 31) ? 0u : (v >> n); }
__device__ __forceinline__ u32 interleave2(u32 v) { /* 16 bits -> every other bit */
  v &= 0x0000ffffu;
  v = (v ^ (v << 8)) & 0x00ff00ffu;
  v = (v ^ (v << 4)) & 0x0f0f0f0fu;
  v = (v ^ (v << 2)) & 0x33333333u;
  v = (v ^ (v << 1)) & 0x55555555u;
  return v;
}
__device__ __forceinline__ u32 interleave3(u32 x) { /* 10 bits -> every third bit, multiply-mask form */
  x = (x * 0x00010001u) & 0xFF0000FFu;
  x = (x * 0x00000101u) & 0x0F00F00Fu;
  x = (x * 0x00000011u) & 0xC30C30C3u;
  x = (x * 0x00000005u) & 0x49249249u;
  return x;
}
__device__ __forceinline__ u32 quantize(float p, int nb) {
  const u32 one = shl_s(1u, nb);
  const float v = fmaxf(p * (float)one, 0.0f);
  const u32 q = (u32)v; /* cvt.rzi.u32.f32 saturates */
  return min(q, one - 1u);
}

__device__ u32 morton_code(const float p[3], const MortonCfg& c) {
  int nbx = c.nbX, nby = c.nbY;
  const int nbz = c.nbZ;
  u32 ax = quantize(p[c.axis0], nbx), ay = quantize(p[c.axis1], nby), az = quantize(p[c.axis2], nbz);
  u32 code = 0;
  int d0 = 0, d1 = 0;
  if (c.sum > 0) {
    nbx -= c.preX;
    code = shr_s(ax & shl_s(shl_s(1u, c.preX) - 1u, nbx), nbx);
    code = shl_s(code, c.preY * 2);
    nbx -= c.preY; nby -= c.preY;
    const u32 t0 = interleave2(shr_s(ax & shl_s(shl_s(1u, c.preY) - 1u, nbx), nbx));
    const u32 t1 = interleave2(shr_s(ay & shl_s(shl_s(1u, c.preY) - 1u, nby), nby));
    code |= t0 * 2 + t1;
    if (c.swap > 0) {
      code = shl_s(code, 1);
      nbx -= 1;
      code |= shr_s(ax & shl_s(1u, nbx), nbx);
    }
    code = shl_s(code, nbx + nby + nbz);
    ax &= shl_s(1u, nbx) - 1u;
    ay &= shl_s(1u, nby) - 1u;
    if (c.swap > 0) { d0 = nby - nbx; ax = shl_s(ax, d0); d1 = nby - nbz; az = shl_s(az, d1); }
    else { d0 = nbx - nby; ay = shl_s(ay, d0); d1 = nbx - nbz; az = shl_s(az, d1); }
  }
  if (nbz == 0) {
    code |= interleave2(ax) * 2 + interleave2(ay);
  } else {
    const u32 mx = ax ? interleave3(ax) : 0u, my = ay ? interleave3(ay) : 0u, mz = az ? interleave3(az) : 0u;
    if (c.swap > 0) code |= shr_s(my * 4 + mx * 2 + mz, d0 + d1);
    else code |= shr_s(mx * 4 + my * 2 + mz, d0 + d1);
  }
  return code;
}

#define MORTON_THREADS 256
__global__ void __launch_bounds__(MORTON_THREADS) morton30_kernel(const b2bvh_aabb* __restrict__ triAabb, const b2bvh_aabb* __restrict__ scene,
                                                                  u32 n, u32* __restrict__ keys, u32* __restrict__ vals) {
  __shared__ MortonCfg cfg;
  __shared__ float smin[3], sext[3];
  if (threadIdx.x == 0) {
    const float* s = reinterpret_cast<const float*>(scene);
    const float ex = s[3] - s[0], ey = s[4] - s[1], ez = s[5] - s[2];
    smin[0] = s[0]; smin[1] = s[1]; smin[2] = s[2];
    sext[0] = ex; sext[1] = ey; sext[2] = ez;
    morton_make_cfg(ex, ey, ez, cfg);
  }
  __syncthreads();
  const MortonCfg c = cfg;
  const float mnx = smin[0], mny = smin[1], mnz = smin[2], ex = sext[0], ey = sext[1], ez = sext[2];
  for (u32 i = blockIdx.x * MORTON_THREADS + threadIdx.x; i < n; i += gridDim.x * MORTON_THREADS) {
    const float2* q = reinterpret_cast<const float2*>(triAabb + i); /* 24-byte stride: 8-byte aligned */
    const float2 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2);   /* (lx,ly) (lz,hx) (hy,hz) */
    float p[3];
    p[0] = (0.5f * (b.y + a.x) - mnx) / ex;
    p[1] = (0.5f * (d.x + a.y) - mny) / ey;
    p[2] = (0.5f * (d.y + b.x) - mnz) / ez;
    keys[i] = morton_code(p, c);
    vals[i] = i;
  }
}

int b2_launch_extents(b2bvh_ctx* ctx, const b2bvh_triangle* d_tris, u32 n, b2bvh_aabb* d_triAabb, b2bvh_aabb* d_scene, u32* d_scratch8,
                      float* d_negmin_max6) {
  const u32 nTiles = (n + EXT_THREADS - 1) / EXT_THREADS;
  u32 grid = (u32)ctx->sm_count * 8u;
  if (grid > nTiles) grid = nTiles;
  B2_KERNEL(ctx, "primref_extents");
  primref_extents_kernel<<<grid, EXT_THREADS, 0, ctx->stream>>>(d_tris, n, d_triAabb, d_scene, d_scratch8, d_negmin_max6);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2_launch_morton(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const b2bvh_aabb* d_scene, u32 n, u32* d_keys, u32* d_vals) {
  u32 grid = (n + MORTON_THREADS - 1) / MORTON_THREADS;
  const u32 cap = (u32)ctx->sm_count * 16u;
  if (grid > cap) grid = cap;
  B2_KERNEL(ctx, "morton30");
  morton30_kernel<<<grid, MORTON_THREADS, 0, ctx->stream>>>(d_triAabb, d_scene, n, d_keys, d_vals);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
