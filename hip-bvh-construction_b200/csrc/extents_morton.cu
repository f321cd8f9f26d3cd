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
#include "morton.cuh"

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

/* global scene box of a sharded build: the all-reduced {-min, max} vector (b2bvh_shard_extents) back into an Aabb, on the device */
__global__ void scene_from_negmin_max_kernel(const float* __restrict__ v, b2bvh_aabb* scene) {
  if (threadIdx.x == 0) {
    scene->m_min.x = -v[0]; scene->m_min.y = -v[1]; scene->m_min.z = -v[2];
    scene->m_max.x = v[3]; scene->m_max.y = v[4]; scene->m_max.z = v[5];
  }
}
int b2_launch_scene_from_negmin_max(b2bvh_ctx* ctx, const float* d_negmin_max6, b2bvh_aabb* d_scene) {
  B2_KERNEL(ctx, "scene_from_negmin_max");
  scene_from_negmin_max_kernel<<<1, 32, 0, ctx->stream>>>(d_negmin_max6, d_scene);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
