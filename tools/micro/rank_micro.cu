// Micro-benchmark (development tool): cost of the warp-level "which lanes hold my digit" primitives on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rank_micro rank_micro.cu ; run: ./rank_micro
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
#define ITERS 512
__device__ __forceinline__ u32 hash(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

template <int BITS> __global__ void k_match(u32* out) {
  u32 x = hash(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
  for (int i = 0; i < ITERS; i++) { x = x * 1664525u + 1013904223u; const u32 d = (x >> 11) & ((1u << BITS) - 1u); acc += __popc(__match_any_sync(0xffffffffu, d) & lanemask_lt()); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int BITS> __global__ void k_ballot(u32* out) {
  u32 x = hash(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
  for (int i = 0; i < ITERS; i++) {
    x = x * 1664525u + 1013904223u; const u32 d = (x >> 11) & ((1u << BITS) - 1u);
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < BITS; b++) { const bool bit = (d >> b) & 1u; const u32 m = __ballot_sync(0xffffffffu, bit); peers &= bit ? m : ~m; }
    acc += __popc(peers & lanemask_lt());
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// full ranking step: peers (ballots) + per-warp shared histogram read/update, 8-bit digits
template <int MODE> __global__ void k_rank(u32* out) {
  __shared__ u32 hist[16][256];
  const u32 w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int k = l; k < 256; k += 32) hist[w][k] = 0;
  __syncwarp();
  u32 x = hash(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
  for (int i = 0; i < ITERS; i++) {
    x = x * 1664525u + 1013904223u; const u32 d = (x >> 11) & 255u;
    u32 peers;
    if (MODE == 0) peers = __match_any_sync(0xffffffffu, d);
    else {
      peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; b++) { const bool bit = (d >> b) & 1u; const u32 m = __ballot_sync(0xffffffffu, bit); peers &= bit ? m : ~m; }
    }
    const u32 lt = peers & lanemask_lt();
    const u32 old = hist[w][d];
    __syncwarp();
    if (lt == 0) hist[w][d] = old + __popc(peers);
    __syncwarp();
    acc += old + __popc(lt);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// shared-memory atomic with return, random 8-bit digit (NOT stable; cost reference only)
__global__ void k_atoms(u32* out) {
  __shared__ u32 hist[16][256];
  const u32 w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int k = l; k < 256; k += 32) hist[w][k] = 0;
  __syncwarp();
  u32 x = hash(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
  for (int i = 0; i < ITERS; i++) { x = x * 1664525u + 1013904223u; const u32 d = (x >> 11) & 255u; acc += atomicAdd(&hist[w][d], 1u); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <typename F> static void run(const char* name, F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double warpIters = 148.0 * 4 * 16 * ITERS;  // per launch
  printf("%-28s %8.3f ms  -> %.1f SM-cycles per warp-iteration per SM (at 1.965 GHz), %.2f Gitems/s\n", name, ms, ms * 1e-3 * 1.965e9 / (warpIters / 148.0), warpIters * 32 / ms / 1e6);
}
int main() {
  u32* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
  const int G = 148 * 4, T = 512;
  run("match.any 2 bits", [&] { k_match<2><<<G, T>>>(out); });
  run("match.any 4 bits", [&] { k_match<4><<<G, T>>>(out); });
  run("match.any 8 bits", [&] { k_match<8><<<G, T>>>(out); });
  run("ballot x4", [&] { k_ballot<4><<<G, T>>>(out); });
  run("ballot x8", [&] { k_ballot<8><<<G, T>>>(out); });
  run("rank step, match.any", [&] { k_rank<0><<<G, T>>>(out); });
  run("rank step, 8 ballots", [&] { k_rank<1><<<G, T>>>(out); });
  run("atoms.add return (unstable)", [&] { k_atoms<<<G, T>>>(out); });
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
