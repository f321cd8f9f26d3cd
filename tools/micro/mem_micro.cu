// Micro-benchmark (development tool): DRAM traffic of access patterns that matter for the BVH kernels, read with
//   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ./mem_micro
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
__device__ __forceinline__ u32 hash(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
// N elements of 32 B; each thread gathers one random element
__global__ void k_gather32(const uint4* a, u32 n, uint4* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const u32 j = hash(i) % n; const uint4 x = __ldg(a + 2 * (size_t)j), y = __ldg(a + 2 * (size_t)j + 1);
  if (x.x == 0x12345 && y.y == 0x777) out[0] = x;
}
__device__ __forceinline__ uint4 ld_nc_l2_64(const uint4* p) { uint4 v; asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ uint4 ld_nc_noalloc(const uint4* p) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__global__ void k_gather32_l64(const uint4* a, u32 n, uint4* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const u32 j = hash(i) % n; const uint4 x = ld_nc_l2_64(a + 2 * (size_t)j), y = ld_nc_l2_64(a + 2 * (size_t)j + 1);
  if (x.x == 0x12345 && y.y == 0x777) out[0] = x;
}
__global__ void k_gather32_cs(const uint4* a, u32 n, uint4* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const u32 j = hash(i) % n; const uint4 x = __ldcs(a + 2 * (size_t)j), y = __ldcs(a + 2 * (size_t)j + 1);
  if (x.x == 0x12345 && y.y == 0x777) out[0] = x;
}
__global__ void k_gather32_na(const uint4* a, u32 n, uint4* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const u32 j = hash(i) % n; const uint4 x = ld_nc_noalloc(a + 2 * (size_t)j), y = ld_nc_noalloc(a + 2 * (size_t)j + 1);
  if (x.x == 0x12345 && y.y == 0x777) out[0] = x;
}
// gather 16 B out of a 16 B-element array
__global__ void k_gather16(const uint4* a, u32 n, uint4* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const u32 j = hash(i) % n; const uint4 x = __ldg(a + j);
  if (x.x == 0x12345 && x.y == 0x777) out[0] = x;
}
// gather 24 B boxes (3 x 8 B) out of a 24 B-element array
__global__ void k_gather24(const float2* a, u32 n, uint4* out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  const u32 j = hash(i) % n; const float2 x = __ldg(a + 3 * (size_t)j), y = __ldg(a + 3 * (size_t)j + 1), z = __ldg(a + 3 * (size_t)j + 2);
  if (x.x == 1.5f && y.y == 2.5f && z.x == 3.f) out[0] = make_uint4(1, 2, 3, 4);
}
// streaming, fully coalesced 16 B per thread
__global__ void k_write_coalesced(uint4* a, u32 n16) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n16) return;
  a[i] = make_uint4(i, 1, 2, 3);
}
// each thread writes one 32 B element as two 16 B stores (half sectors per instruction)
__global__ void k_write_halves(uint4* a, u32 n32) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n32) return;
  a[2 * (size_t)i] = make_uint4(i, 1, 2, 3); a[2 * (size_t)i + 1] = make_uint4(i, 4, 5, 6);
}
// each thread writes one 128 B element as eight 16 B stores
__global__ void k_write_128(uint4* a, u32 n128) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n128) return;
#pragma unroll
  for (int k = 0; k < 8; k++) a[8 * (size_t)i + k] = make_uint4(i, k, 2, 3);
}
// scattered 8 B writes (random slot)
__global__ void k_write_scatter8(uint2* a, u32 n8) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n8) return;
  a[hash(i) % n8] = make_uint2(i, 1);
}
// scattered 32 B writes (two 16 B stores to a random 32 B slot)
__global__ void k_write_scatter32(uint4* a, u32 n32) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n32) return;
  const u32 j = hash(i) % n32; a[2 * (size_t)j] = make_uint4(i, 1, 2, 3); a[2 * (size_t)j + 1] = make_uint4(i, 4, 5, 6);
}
int main(int argc, char** argv) {
  if (argc > 1) { size_t g = atoi(argv[1]); cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("set L2 fetch granularity %zu: %s, now %zu\n", g, cudaGetErrorString(e), got); }
  else { size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n", got); }
  const size_t bytes = 1ull << 30;
  void *a, *o; cudaMalloc(&a, bytes); cudaMalloc(&o, 256); cudaMemset(a, 0, bytes);
  const u32 T = 256;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
#define RUN(name, elems, ...) { cudaDeviceSynchronize(); cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("%-20s %8.3f ms  (%u elements)\n", name, ms, (u32)(elems)); }
  const u32 n32 = bytes / 32, n16 = bytes / 16, n24 = bytes / 24, n128 = bytes / 128, n8 = bytes / 8;
  for (int rep = 0; rep < 2; rep++) {
    RUN("gather32", n32 / 2, k_gather32<<<(n32 / 2 + T - 1) / T, T>>>((const uint4*)a, n32, (uint4*)o));
    RUN("gather32_l64", n32 / 2, k_gather32_l64<<<(n32 / 2 + T - 1) / T, T>>>((const uint4*)a, n32, (uint4*)o));
    RUN("gather32_cs", n32 / 2, k_gather32_cs<<<(n32 / 2 + T - 1) / T, T>>>((const uint4*)a, n32, (uint4*)o));
    RUN("gather32_na", n32 / 2, k_gather32_na<<<(n32 / 2 + T - 1) / T, T>>>((const uint4*)a, n32, (uint4*)o));
    RUN("gather16", n16 / 4, k_gather16<<<(n16 / 4 + T - 1) / T, T>>>((const uint4*)a, n16, (uint4*)o));
    RUN("gather24", n24 / 3, k_gather24<<<(n24 / 3 + T - 1) / T, T>>>((const float2*)a, n24, (uint4*)o));
    RUN("write_coalesced", n16, k_write_coalesced<<<(n16 + T - 1) / T, T>>>((uint4*)a, n16));
    RUN("write_halves", n32, k_write_halves<<<(n32 + T - 1) / T, T>>>((uint4*)a, n32));
    RUN("write_128", n128, k_write_128<<<(n128 + T - 1) / T, T>>>((uint4*)a, n128));
    RUN("write_scatter8", n8 / 8, k_write_scatter8<<<(n8 / 8 + T - 1) / T, T>>>((uint2*)a, n8));
    RUN("write_scatter32", n32 / 2, k_write_scatter32<<<(n32 / 2 + T - 1) / T, T>>>((uint4*)a, n32));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
