/*
 * cuda_emul.h — TEST INFRASTRUCTURE ONLY (used to build oracle/_ref/libref_emul.so).
 *
 * A force-included prelude (-include) that lets g++ compile the reference's
 * HIP/CUDA kernel headers (src/*Kernel.h) UNMODIFIED, from where they lie under
 * /root/reference, as ordinary C++ and run one GPU thread at a time:
 *   - vector types + component-wise operators the device compilers provide natively
 *     (Common.h only defines them for the host when __KERNELCC__ is not set),
 *   - threadIdx/blockIdx/blockDim as thread-local globals set by the launcher,
 *   - atomics / fences / warp votes with single-thread semantics.
 * Only kernels whose result does not depend on inter-thread timing are executed
 * (see ref_emul_*.cpp); the rest merely have to compile.
 */
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __host__
#define __device__
#define __global__
#ifdef B2_EMUL_MT
#define __shared__ static /* one block runs at a time: a static is the block's shared memory (cuda_emul_mt.h) */
#else
#define __shared__
#endif
#define __forceinline__ inline

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct int3 { int x, y, z; };
struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct dim3e { uint32_t x, y, z; };

extern thread_local dim3e threadIdx, blockIdx, blockDim, gridDim;

inline bool operator==(const float3& a, const float3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline float3 operator+(const float3& a, const float3& b) { return float3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float4 operator+(const float4& a, const float4& b) { return float4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline float3 operator-(const float3& a, const float3& b) { return float3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float4 operator-(const float4& a, const float4& b) { return float4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline float2 operator-(const float2& a, const float2& b) { return float2{a.x - b.x, a.y - b.y}; }
inline float3 operator-(const float3& a) { return float3{-a.x, -a.y, -a.z}; }
inline float4 operator-(const float4& a) { return float4{-a.x, -a.y, -a.z, -a.w}; }
inline float3 operator/(const float3& a, const float3& b) { return float3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator/(const float3& a, float b) { return float3{a.x / b, a.y / b, a.z / b}; }
inline float3 operator/(float b, const float3& a) { return float3{b / a.x, b / a.y, b / a.z}; }
inline float4 operator/(const float4& a, float b) { return float4{a.x / b, a.y / b, a.z / b, a.w / b}; }
inline float3& operator*=(float3& a, float c) { a.x *= c; a.y *= c; a.z *= c; return a; }
inline float3 operator*(float c, const float3& a) { return float3{c * a.x, c * a.y, c * a.z}; }
inline float3 operator*(const float3& a, float c) { return float3{c * a.x, c * a.y, c * a.z}; }
inline float3 operator*(const float3& a, const float3& b) { return float3{a.x * b.x, a.y * b.y, a.z * b.z}; }

/* scalar min/max overloads the device headers provide */
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline int min(int a, uint32_t b) { return (uint32_t)a < b ? a : (int)b; }
inline uint32_t min(uint32_t a, int b) { return a < (uint32_t)b ? a : (uint32_t)b; }
inline uint32_t max(int a, uint32_t b) { return (uint32_t)a > b ? (uint32_t)a : b; }
inline uint32_t max(uint32_t a, int b) { return a > (uint32_t)b ? a : (uint32_t)b; }

inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline int __clzll(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __popcll(uint64_t v) { return __builtin_popcountll(v); }
inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
inline int __ffsll(unsigned long long v) { return __builtin_ffsll((long long)v); }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline uint32_t __float_as_uint(float f) { uint32_t i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(uint32_t i) { float f; memcpy(&f, &i, 4); return f; }

#ifdef B2_EMUL_MT
#include "cuda_emul_mt.h" /* atomics, __syncthreads and warp collectives for one block run by real threads */
#else
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline uint32_t atomicAdd(uint32_t* p, int v) { uint32_t o = *p; *p = o + (uint32_t)v; return o; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
inline void __threadfence() {}
inline void __syncthreads() {}
/* one-lane warp: votes/shuffles see only the calling thread */
template <typename T> inline T __shfl(T v, int) { return v; }
inline uint64_t __ballot(bool p) { return p ? 1ull : 0ull; }
inline bool __any(bool p) { return p; }
#endif /* B2_EMUL_MT */
