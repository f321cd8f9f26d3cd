/*
 * misc.cu — benchmark input generator and the per-launch profiler of the C ABI.
 */
#include <stdio.h>
#include <string.h>

#include "common.cuh"

/* synth_uniform_v1: RNG restates tea<16> / lcg / randf (CommonBlocksKernel.h:401-430);
 * every float operation is individually rounded (no FMA) so host and device generate identical bits. */
__device__ __forceinline__ u32 tea16(u32 v0, u32 v1) {
  u32 s0 = 0;
#pragma unroll
  for (int r = 0; r < 16; r++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
__device__ __forceinline__ float rand01(u32& s) {
  s = 1103515245u * s + 12345u;
  return __fdiv_rn((float)(s & 0x00FFFFFFu), 16777216.0f);
}

__global__ void __launch_bounds__(256) synth_uniform_kernel(u64 first, u32 count, u32 seed, float half, b2bvh_triangle* __restrict__ out) {
  const u32 k = blockIdx.x * 256 + threadIdx.x;
  if (k >= count) return;
  u32 s = tea16((u32)(first + k), seed);
  float r[12];
#pragma unroll
  for (int j = 0; j < 12; j++) r[j] = rand01(s);
  float c[3], v[9];
  const float twoH = __fmul_rn(2.0f, half);
#pragma unroll
  for (int j = 0; j < 3; j++) c[j] = __fadd_rn(-1000.0f, __fmul_rn(2000.0f, r[j]));
#pragma unroll
  for (int j = 0; j < 9; j++) v[j] = __fadd_rn(c[j % 3], __fmul_rn(__fsub_rn(r[3 + j], 0.5f), twoH));
  float4* q = reinterpret_cast<float4*>(out + k);
  q[0] = make_float4(v[0], v[1], v[2], v[3]);
  q[1] = make_float4(v[4], v[5], v[6], v[7]);
  q[2] = make_float4(v[8], 0.f, 0.f, 0.f);
  q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

/* synth_clustered_v1 (second distribution: many primitives per Morton cell): triangle i draws a cluster
 * j = floor(r0 * 4096); the cluster's centre comes from its own stream tea16(j, seed ^ 0xC1) (three draws, the uniform map); the
 * triangle's centre is that + 20 * (r - 0.5) per axis; vertices as in synth_uniform_v1.  16 draws: r0, three for the centre, nine. */
__global__ void __launch_bounds__(256) synth_clustered_kernel(u64 first, u32 count, u32 seed, float half, b2bvh_triangle* __restrict__ out) {
  const u32 k = blockIdx.x * 256 + threadIdx.x;
  if (k >= count) return;
  u32 s = tea16((u32)(first + k), seed);
  const u32 j = (u32)__fmul_rn(rand01(s), 4096.0f);
  u32 sj = tea16(j, seed ^ 0xC1u);
  float c[3], v[9];
  const float twoH = __fmul_rn(2.0f, half);
#pragma unroll
  for (int a = 0; a < 3; a++) c[a] = __fadd_rn(-1000.0f, __fmul_rn(2000.0f, rand01(sj)));
#pragma unroll
  for (int a = 0; a < 3; a++) c[a] = __fadd_rn(c[a], __fmul_rn(20.0f, __fsub_rn(rand01(s), 0.5f)));
#pragma unroll
  for (int a = 0; a < 9; a++) v[a] = __fadd_rn(c[a % 3], __fmul_rn(__fsub_rn(rand01(s), 0.5f), twoH));
  float4* q = reinterpret_cast<float4*>(out + k);
  q[0] = make_float4(v[0], v[1], v[2], v[3]);
  q[1] = make_float4(v[4], v[5], v[6], v[7]);
  q[2] = make_float4(v[8], 0.f, 0.f, 0.f);
  q[3] = make_float4(0.f, 0.f, 0.f, 0.f);
}

int b2_prof_begin(b2bvh_ctx* ctx, const char* name) {
  if (ctx->prof_n >= 512) return 0;
  b2bvh_ctx::Prof& p = ctx->prof[ctx->prof_n];
  if (ctx->prof_n >= ctx->prof_events) {
    B2_CUDA(cudaEventCreate(&p.a));
    B2_CUDA(cudaEventCreate(&p.b));
    ctx->prof_events = ctx->prof_n + 1;
  }
  p.name = name;
  B2_CUDA(cudaEventRecord(p.a, ctx->stream));
  return 0;
}
int b2_prof_end(b2bvh_ctx* ctx) {
  if (ctx->prof_n >= 512) return 0;
  B2_CUDA(cudaEventRecord(ctx->prof[ctx->prof_n].b, ctx->stream));
  ctx->prof_n++;
  return 0;
}

extern "C" {

int b2bvh_synth_uniform(b2bvh_ctx* ctx, uint64_t first, uint32_t count, uint32_t seed, float half, b2bvh_triangle* d_tris) {
  if (!ctx || !d_tris || count == 0) return b2_fail(B2BVH_ERR_INVALID, "synth_uniform: bad argument");
  B2_KERNEL(ctx, "synth_uniform");
  synth_uniform_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(first, count, seed, half, d_tris);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2bvh_synth_clustered(b2bvh_ctx* ctx, uint64_t first, uint32_t count, uint32_t seed, float half, b2bvh_triangle* d_tris) {
  if (!ctx || !d_tris || count == 0) return b2_fail(B2BVH_ERR_INVALID, "synth_clustered: bad argument");
  B2_KERNEL(ctx, "synth_clustered");
  synth_clustered_kernel<<<(count + 255) / 256, 256, 0, ctx->stream>>>(first, count, seed, half, d_tris);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2bvh_profile_enable(b2bvh_ctx* ctx, int on) {
  if (!ctx) return b2_fail(B2BVH_ERR_INVALID, "profile_enable: bad argument");
  ctx->prof_on = on != 0;
  ctx->prof_n = 0;
  return 0;
}
int b2bvh_profile_count(b2bvh_ctx* ctx, int* count) {
  if (!ctx || !count) return b2_fail(B2BVH_ERR_INVALID, "profile_count: bad argument");
  *count = ctx->prof_n;
  return 0;
}
int b2bvh_profile_entry(b2bvh_ctx* ctx, int index, char* name, size_t cap, float* ms) {
  if (!ctx || !ms || index < 0 || index >= ctx->prof_n) return b2_fail(B2BVH_ERR_INVALID, "profile_entry: bad argument");
  B2_CUDA(cudaEventSynchronize(ctx->prof[index].b));
  B2_CUDA(cudaEventElapsedTime(ms, ctx->prof[index].a, ctx->prof[index].b));
  if (name && cap) snprintf(name, cap, "%s", ctx->prof[index].name);
  return 0;
}
}

/* root box of the build that just ran, device to device, without the host knowing the root index (sharded build: the all-gather of
 * the sub-tree roots can be enqueued before the build's host synchronisation) */
__global__ void root_box_kernel(const b2bvh_bvh2_node* __restrict__ nodes, const u32* __restrict__ rootIdx, float* __restrict__ box6) {
  if (threadIdx.x < 6) box6[threadIdx.x] = reinterpret_cast<const float*>(nodes + *rootIdx)[2 + threadIdx.x];
}
int b2_launch_root_box(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const u32* d_rootIdx, float* d_box6) {
  B2_KERNEL(ctx, "root_box");
  root_box_kernel<<<1, 32, 0, ctx->stream>>>(d_nodes, d_rootIdx, d_box6);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

/* ---- small device -> host results through mapped pinned memory (b2_fetch_words, common.cuh) ---- */
__global__ void fetch_words_kernel(const u32* __restrict__ src, u32 words, u32* dst) {
  if (threadIdx.x < words) dst[threadIdx.x] = src[threadIdx.x];
}
int b2_fetch_words(b2bvh_ctx* ctx, const void* d_src, u32 words, int slot) {
  if (words > 16 || slot < 0 || slot >= B2_MAILBOX_SLOTS) return b2_fail(B2BVH_ERR_INTERNAL, "fetch_words: bad request");
  B2_KERNEL(ctx, "fetch_words");
  fetch_words_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const u32*>(d_src), words, ctx->mailbox_dev + slot * 16);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
