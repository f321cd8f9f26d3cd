/*
 * morton.cuh — extent-adaptive 30-bit Morton code, device functions shared by morton30_kernel and the top-level builder.
 * Restates computeExtendedMortonCode (CommonBlocksKernel.h:159-359); see extents_morton.cu.
 */
#pragma once
#include "common.cuh"

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
__device__ inline int ilog2_ratio(float a, float b) {
  const float r = a / b;
  const float l = (float)log2((double)r);
  return __float2int_rz(l);
}
__device__ __forceinline__ int wadd(int a, int b) { return (int)((u32)a + (u32)b); }
__device__ __forceinline__ int wsub(int a, int b) { return (int)((u32)a - (u32)b); }
__device__ __forceinline__ int wmul2(int a) { return (int)((u32)a * 2u); }

__device__ inline void morton_make_cfg(const float ex, const float ey, const float ez, MortonCfg& c) {
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

__device__ inline u32 morton_code(const float p[3], const MortonCfg& c) {
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

