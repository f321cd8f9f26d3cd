/*
 * cuda_emul_mt.h — TEST INFRASTRUCTURE ONLY (included by cuda_emul.h when B2_EMUL_MT is defined; builds
 * oracle/_ref/libref_emul_mt.so).
 *
 * Block-level emulation for the reference kernels whose result depends on cooperation inside a block (Ploc,
 * SinglePassPloc: shared memory, __syncthreads, 64-bit shared atomics, __ballot/__shfl over the ACTIVE lanes).  One block
 * is run by blockDim.x cooperative fibers (ucontext), blocks of a launch run one after the other in index order (the
 * inter-block spin of Ploc, Ploc++Kernel.h:341-347, waits for lower block indices only).
 *   __shared__      -> static (one block at a time)
 *   __syncthreads   -> the fiber yields; all fibers that have not returned continue once every one of them waits
 *   __ballot/__shfl -> the fiber yields; the collective completes when every lane of the warp that is neither waiting at
 *                      __syncthreads nor returned has arrived — what a lock-step wavefront executes when some lanes skipped
 *                      a branch and sit at the reconvergence barrier
 *   atomics         -> plain read-modify-write (one fiber runs at a time)
 * Scheduling is deterministic: between two synchronisation points the runnable fibers run one at a time, HIGHEST thread
 * index first.  The reference reuses the cache of binaryBlockPrefixSum without a barrier (every warp reads
 * blockCache[warpIndex] right after the last __syncthreads, Ploc++Kernel.h:95, and warp 0 overwrites the same words a few
 * statements later, :190-192 / :316 via the node-index array); lock-step warps read before they overwrite, and so does
 * this order.
 */
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>

namespace b2emul {
constexpr int kWarp = 32; /* Common.h:104 (WarpSize for non-gfx9 targets) */
enum State { READY, AT_BARRIER, AT_COLLECTIVE, EXITED };
struct Fiber {
  ucontext_t ctx;
  State state;
  bool pred;
  uint64_t val, resVal;
  uint64_t resBallot;
  int src;
};
struct Block {
  int nThreads;
  Fiber* fibers;
  ucontext_t main;
  int current;
};
extern Block g_blk;
void yield_to_scheduler(); /* ref_emul_ploc_mt.cpp */

inline void sync_threads() {
  g_blk.fibers[g_blk.current].state = AT_BARRIER;
  yield_to_scheduler();
}
inline uint64_t warp_collective(bool p, uint64_t v, int src, uint64_t* out) {
  Fiber& f = g_blk.fibers[g_blk.current];
  f.pred = p; f.val = v; f.src = src & (kWarp - 1);
  f.state = AT_COLLECTIVE;
  yield_to_scheduler();
  Fiber& g = g_blk.fibers[g_blk.current];
  if (out) *out = g.resVal;
  return g.resBallot;
}
}  // namespace b2emul

inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
inline uint32_t atomicAdd(uint32_t* p, int v) { uint32_t o = *p; *p = o + (uint32_t)v; return o; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
template <typename T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
inline void __threadfence() {}
inline void __syncthreads() { b2emul::sync_threads(); }
template <typename T> inline T __shfl(T v, int src) {
  uint64_t in = 0, out = 0;
  memcpy(&in, &v, sizeof(T));
  b2emul::warp_collective(false, in, src, &out);
  T r;
  memcpy(&r, &out, sizeof(T));
  return r;
}
inline uint64_t __ballot(bool p) { return b2emul::warp_collective(p, 0, 0, nullptr); }
inline bool __any(bool p) { return b2emul::warp_collective(p, 0, 0, nullptr) != 0; }
