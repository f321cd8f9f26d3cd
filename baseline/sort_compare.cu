/*
 * sort_compare.cu — same-box comparators for stage S3 (stable u32 key/value radix sort), a BENCH tool, not product code:
 *
 *   orochi  the reference's own sort: Oro::RadixSort's kernels, UNMODIFIED, compiled for sm_100a from where they lie under
 *           /root/reference (dependencies/Orochi/ParallelPrimitives/RadixSortKernels.h -> baseline/_ref/oro_radixsort.cubin, flags of
 *           RadixSort.cpp:189-213 as an NVIDIA device gets them) and driven by this file in the launch order of
 *           RadixSort::sort / sort1pass (RadixSort.cpp:291-318, RadixSort.inl:77-160): per 8-bit digit CountKernel,
 *           ParallelExclusiveScanAllWG, SortKVKernel over SMs x occupancy work-groups of 1024 threads
 *   cub     cub::DeviceRadixSort::SortPairs of the CUDA toolkit (library code: allowed as a yardstick only)
 *   b2bvh   b2bvh_sort_pairs of libb2bvh.so (the product), loaded with dlopen
 *
 * All three sort the same keys (iota values) and must produce the same arrays (a stable sort has one answer).  One JSON line.
 * Usage: sort_compare <libb2bvh.so> <oro_radixsort.cubin> [n ...]
 */
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <string>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } \
  } while (0)
#define CU(x)                                                                     \
  do {                                                                            \
    CUresult r_ = (x);                                                            \
    if (r_ != CUDA_SUCCESS) { const char* s_; cuGetErrorString(r_, &s_); fprintf(stderr, "%s: %s\n", #x, s_); exit(2); } \
  } while (0)

__global__ void fill_keys(uint32_t* k, uint32_t* v, uint32_t n, uint32_t bits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = i * 0x9E3779B9u + 0x7F4A7C15u; /* a 32-bit mix: uniform digits, as Morton codes of a uniform soup are */
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  k[i] = bits >= 32 ? x : (x >> (32 - bits));
  v[i] = i;
}
__global__ void compare_arrays(const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* diff) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && a[i] != b[i]) atomicAdd(diff, 1u);
}

static float median(std::vector<float> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }

template <class F>
static float time_ms(cudaStream_t s, int warm, int reps, F f) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < warm; i++) f();
  std::vector<float> t;
  for (int i = 0; i < reps; i++) {
    CK(cudaEventRecord(a, s)); f(); CK(cudaEventRecord(b, s)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); t.push_back(ms);
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return median(t);
}

typedef struct b2bvh_ctx b2bvh_ctx;
typedef int (*ctx_create_t)(int, void*, b2bvh_ctx**);
typedef int (*sort_pairs_t)(b2bvh_ctx*, const uint32_t*, const uint32_t*, uint32_t*, uint32_t*, uint32_t, uint32_t, uint32_t);
typedef const char* (*last_error_t)(void);

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s libb2bvh.so oro_radixsort.cubin [n ...]\n", argv[0]); return 2; }
  std::vector<uint32_t> sizes;
  for (int i = 3; i < argc; i++) sizes.push_back((uint32_t)strtoul(argv[i], nullptr, 10));
  if (sizes.empty()) { sizes.push_back(10000000u); sizes.push_back(100000000u); }
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  cudaStream_t s;
  CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

  /* ---- the product ---- */
  void* lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen %s: %s\n", argv[1], dlerror()); return 2; }
  ctx_create_t ctx_create = (ctx_create_t)dlsym(lib, "b2bvh_ctx_create");
  sort_pairs_t sort_pairs = (sort_pairs_t)dlsym(lib, "b2bvh_sort_pairs");
  last_error_t last_error = (last_error_t)dlsym(lib, "b2bvh_last_error");
  b2bvh_ctx* ctx = nullptr;
  if (ctx_create(0, (void*)s, &ctx)) { fprintf(stderr, "b2bvh_ctx_create: %s\n", last_error()); return 2; }

  /* ---- the reference's kernels (driver API, as Orochi loads them) ---- */
  CUmodule mod;
  CUfunction fCount, fScan, fSortKV;
  bool haveOro = cuModuleLoad(&mod, argv[2]) == CUDA_SUCCESS;
  if (haveOro) {
    CU(cuModuleGetFunction(&fCount, mod, "CountKernel"));
    CU(cuModuleGetFunction(&fScan, mod, "ParallelExclusiveScanAllWG"));
    CU(cuModuleGetFunction(&fSortKV, mod, "SortKVKernel"));
  }
  /* RadixSort::calculateWGsToExecute (RadixSort.cpp:236-262) with 1024-thread work-groups */
  const int WG = 1024, BIN = 256;
  int occupancy = std::max(1, (prop.maxThreadsPerMultiProcessor / 32) / (WG / 32));
  int nBlocks = prop.multiProcessorCount * occupancy;
  nBlocks = (nBlocks / (WG / BIN)) * (WG / BIN);
  const int tmpSize = BIN * nBlocks, nScanBlocks = tmpSize / WG;
  int *dTmp, *dPartial;
  bool* dReady;
  CK(cudaMalloc(&dTmp, tmpSize * sizeof(int)));
  CK(cudaMalloc(&dPartial, nScanBlocks * sizeof(int)));
  CK(cudaMalloc(&dReady, nScanBlocks));
  CK(cudaMemset(dReady, 0, nScanBlocks));

  printf("{\"device\": \"%s\", \"orochi_work_groups\": %d, \"cub_version\": %d, \"sizes\": [", prop.name, nBlocks, CUB_VERSION);
  for (size_t si = 0; si < sizes.size(); si++) {
    const uint32_t n = sizes[si];
    uint32_t *kIn, *vIn, *kA, *vA, *kB, *vB, *kRef, *vRef, *dDiff;
    CK(cudaMalloc(&kIn, (size_t)n * 4)); CK(cudaMalloc(&vIn, (size_t)n * 4));
    CK(cudaMalloc(&kA, (size_t)n * 4)); CK(cudaMalloc(&vA, (size_t)n * 4));
    CK(cudaMalloc(&kB, (size_t)n * 4)); CK(cudaMalloc(&vB, (size_t)n * 4));
    CK(cudaMalloc(&kRef, (size_t)n * 4)); CK(cudaMalloc(&vRef, (size_t)n * 4));
    CK(cudaMalloc(&dDiff, 4));
    fill_keys<<<(n + 255) / 256, 256, 0, s>>>(kIn, vIn, n, 30);
    CK(cudaStreamSynchronize(s));
    auto differs = [&](const uint32_t* k, const uint32_t* v) {
      CK(cudaMemsetAsync(dDiff, 0, 4, s));
      compare_arrays<<<(n + 255) / 256, 256, 0, s>>>(k, kRef, n, dDiff);
      compare_arrays<<<(n + 255) / 256, 256, 0, s>>>(v, vRef, n, dDiff);
      uint32_t h = 0;
      CK(cudaMemcpyAsync(&h, dDiff, 4, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      return h;
    };
    /* ---- CUB (also the answer the other two are compared with) ---- */
    size_t cubBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, kIn, kRef, vIn, vRef, (int)n, 0, 32, s);
    void* cubTmp;
    CK(cudaMalloc(&cubTmp, cubBytes));
    const float msCub32 = time_ms(s, 3, 15, [&] { cub::DeviceRadixSort::SortPairs(cubTmp, cubBytes, kIn, kRef, vIn, vRef, (int)n, 0, 32, s); });
    const float msCub30 = time_ms(s, 3, 15, [&] { cub::DeviceRadixSort::SortPairs(cubTmp, cubBytes, kIn, kRef, vIn, vRef, (int)n, 0, 30, s); });
    /* ---- the product ---- */
    int st = 0;
    const float msB2 = time_ms(s, 3, 15, [&] { st |= sort_pairs(ctx, kIn, vIn, kA, vA, n, 0, 32); });
    if (st) { fprintf(stderr, "b2bvh_sort_pairs: %s\n", last_error()); return 2; }
    const uint32_t diffB2 = differs(kA, vA);
    const float msB2iota = time_ms(s, 3, 15, [&] { st |= sort_pairs(ctx, kIn, nullptr, kA, vA, n, 0, 32); }); /* what the build calls: values = iota, not read */
    const float msB230 = time_ms(s, 3, 15, [&] { st |= sort_pairs(ctx, kIn, nullptr, kA, vA, n, 0, 30); });
    const uint32_t diffB230 = differs(kA, vA);
    /* ---- Orochi: RadixSort::sort(KeyValueSoA, KeyValueSoA, n, 0, 32) ---- */
    float msOro = -1.0f;
    uint32_t diffOro = 0;
    if (haveOro) {
      auto oroSort = [&] {
        uint32_t *sk = kIn, *sv = vIn, *dk = kA, *dv = vA;
        int nn = (int)n, nItemPerWG = (nn + nBlocks - 1) / nBlocks, nWG = nBlocks;
        for (int startBit = 0; startBit < 32; startBit += 8) {
          { void* a[] = {&sk, &dTmp, &nn, &nItemPerWG, &startBit, &nWG}; CU(cuLaunchKernel(fCount, nBlocks, 1, 1, WG, 1, 1, 0, s, a, nullptr)); }
          { void* a[] = {&dTmp, &dTmp, &dPartial, &dReady}; CU(cuLaunchKernel(fScan, nScanBlocks, 1, 1, WG, 1, 1, 0, s, a, nullptr)); }
          { void* a[] = {&sk, &sv, &dk, &dv, &dTmp, &nn, &nItemPerWG, &startBit, &nWG}; CU(cuLaunchKernel(fSortKV, nBlocks, 1, 1, WG, 1, 1, 0, s, a, nullptr)); }
          /* ping-pong: the input arrays are never written (first pass reads them, later passes alternate kA/kB) */
          uint32_t *nk = (dk == kA) ? kB : kA, *nv = (dv == vA) ? vB : vA;
          sk = dk; sv = dv; dk = nk; dv = nv;
        }
      };
      msOro = time_ms(s, 2, 7, oroSort);
      diffOro = differs(kB, vB); /* four passes: A, B, A, B */
    }
    printf("%s{\"n\": %u, \"key_bits\": 30, \"cub_ms\": %.4f, \"cub_ms_bits_0_30\": %.4f, \"b2bvh_ms\": %.4f, \"b2bvh_ms_iota_values\": %.4f, \"b2bvh_ms_iota_values_bits_0_30\": %.4f, "
           "\"orochi_ms\": %s, \"b2bvh_differs_from_cub\": %u, \"b2bvh_30_differs_from_cub\": %u, \"orochi_differs_from_cub\": %u}",
           si ? ", " : "", n, msCub32, msCub30, msB2, msB2iota, msB230, haveOro ? std::to_string(msOro).c_str() : "null", diffB2, diffB230, diffOro);
    cudaFree(kIn); cudaFree(vIn); cudaFree(kA); cudaFree(vA); cudaFree(kB); cudaFree(vB); cudaFree(kRef); cudaFree(vRef); cudaFree(dDiff); cudaFree(cubTmp);
  }
  printf("]}\n");
  return 0;
}
