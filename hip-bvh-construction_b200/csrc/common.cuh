/*
 * common.cuh — device-side vocabulary shared by every kernel of libb2bvh: box math with a fixed
 * floating-point contract, cache-hinted loads/stores, small warp helpers, and the context struct.
 *
 * Floating-point contract: every +,-,*,/ is rounded individually — no FMA
 * contraction anywhere results are compared against the oracle.  The whole library is compiled
 * with -fmad=false; the area formula additionally spells the roundings out so the contract
 * survives a flag change.  Reference math being restated: src/Common.h:310-416 (Aabb).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/b2bvh.h"

typedef uint32_t u32;
typedef uint64_t u64;

#define B2_INVALID 0xFFFFFFFFu
#define B2_FLT_MAX 3.402823466e+38f
#define B2_FULL 0xFFFFFFFFu

struct Box {
  float lx, ly, lz, hx, hy, hz;
};

__device__ __forceinline__ Box box_empty() { return Box{B2_FLT_MAX, B2_FLT_MAX, B2_FLT_MAX, -B2_FLT_MAX, -B2_FLT_MAX, -B2_FLT_MAX}; }
__device__ __forceinline__ Box box_union(const Box& a, const Box& b) {
  return Box{fminf(a.lx, b.lx), fminf(a.ly, b.ly), fminf(a.lz, b.lz), fmaxf(a.hx, b.hx), fmaxf(a.hy, b.hy), fmaxf(a.hz, b.hz)};
}
/* Aabb::area(), Common.h:361-365: 2*(ex*ey + ex*ez + ey*ez), left to right, each op rounded. */
__device__ __forceinline__ float box_area(const Box& b) {
  const float ex = __fsub_rn(b.hx, b.lx), ey = __fsub_rn(b.hy, b.ly), ez = __fsub_rn(b.hz, b.lz);
  const float xy = __fmul_rn(ex, ey), xz = __fmul_rn(ex, ez), yz = __fmul_rn(ey, ez);
  return __fmul_rn(2.0f, __fadd_rn(__fadd_rn(xy, xz), yz));
}

/* 24-byte Aabb (align 4): three 8-byte loads when the address allows it (even element index), scalars otherwise. */
__device__ __forceinline__ Box load_aabb(const b2bvh_aabb* p) {
  const float* f = reinterpret_cast<const float*>(p);
  return Box{__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5)};
}
__device__ __forceinline__ void store_aabb(b2bvh_aabb* p, const Box& b) {
  float* f = reinterpret_cast<float*>(p);
  f[0] = b.lx; f[1] = b.ly; f[2] = b.lz; f[3] = b.hx; f[4] = b.hy; f[5] = b.hz;
}

/* 32-byte Bvh2Node as two 16-byte words: {left, right, lx, ly} {lz, hx, hy, hz}. */
struct Node2 {
  u32 left, right;
  Box box;
};
__device__ __forceinline__ uint4 node2_lo(u32 l, u32 r, const Box& b) { return make_uint4(l, r, __float_as_uint(b.lx), __float_as_uint(b.ly)); }
__device__ __forceinline__ uint4 node2_hi(const Box& b) {
  return make_uint4(__float_as_uint(b.lz), __float_as_uint(b.hx), __float_as_uint(b.hy), __float_as_uint(b.hz));
}
__device__ __forceinline__ Node2 node2_unpack(uint4 a, uint4 c) {
  Node2 n;
  n.left = a.x; n.right = a.y;
  n.box = Box{__uint_as_float(a.z), __uint_as_float(a.w), __uint_as_float(c.x), __uint_as_float(c.y), __uint_as_float(c.z), __uint_as_float(c.w)};
  return n;
}
/* plain store (write-back in L2) */
__device__ __forceinline__ void store_node2(b2bvh_bvh2_node* p, u32 l, u32 r, const Box& b) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = node2_lo(l, r, b);
  q[1] = node2_hi(b);
}
/* L2-coherent load (ld.global.cg): for nodes written by other CTAs during the same launch */
__device__ __forceinline__ Node2 load_node2_cg(const b2bvh_bvh2_node* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  return node2_unpack(__ldcg(q), __ldcg(q + 1));
}
/* Random gathers: a plain load miss makes the B200 L2 fetch the whole 128-byte line from DRAM (measured: 118 B of DRAM
 * traffic per random 32-byte read, tools/micro/mem_micro.cu); the L2::64B qualifier halves that. */
__device__ __forceinline__ uint4 ldg_gather_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_gather_f2(const float2* p) {
  float2 v;
  asm volatile("ld.global.nc.L2::64B.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ u32 ldg_gather_u32(const u32* p) {
  u32 v;
  asm volatile("ld.global.nc.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
/* read-only path: for nodes produced by an earlier launch */
__device__ __forceinline__ Node2 load_node2_ro(const b2bvh_bvh2_node* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  return node2_unpack(__ldg(q), __ldg(q + 1));
}
__device__ __forceinline__ Node2 load_node2_gather(const b2bvh_bvh2_node* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  return node2_unpack(ldg_gather_u4(q), ldg_gather_u4(q + 1));
}

/* acq_rel / release / acquire primitives for the bottom-up passes (replace the reference's
 * __threadfence() + atomicAdd pairs, TwoPassLbvhKernel.h:225-233, SinglePassLbvhKernel.h:100-124). */
__device__ __forceinline__ u32 atom_add_acq_rel(u32* p, u32 v) {
  u32 old;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ u32 atom_exch_acq_rel(u32* p, u32 v) {
  u32 old;
  asm volatile("atom.exch.acq_rel.gpu.global.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ u64 atom_exch_acq_rel64(u64* p, u64 v) {
  u64 old;
  asm volatile("atom.exch.acq_rel.gpu.global.b64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
__device__ __forceinline__ u32 atom_exch_relaxed(u32* p, u32 v) {
  u32 old;
  asm volatile("atom.exch.relaxed.gpu.global.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ u32 ld_acquire(const u32* p) {
  u32 v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(u32* p, u32 v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acquire64(const u64* p) {
  u64 v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release64(u64* p, u64 v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u32 ld_relaxed(const u32* p) {
  u32 v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_relaxed64(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed64(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed(u32* p, u32 v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

/* Barrier among the first `threads` threads of the CTA (a multiple of 32; whole warps arrive). */
__device__ __forceinline__ void named_barrier(u32 id, u32 threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

/* Watchdog of every spin loop in the library: a wait that lasts longer than 20 s of wall clock (%globaltimer) traps — a CTA is missing
 * (not a cooperative launch?) or a chunk never posted; fail, never hang the device.  Time based, not a poll count: under compute-sanitizer
 * the CTAs being waited for run two orders of magnitude slower while a poll costs the same. */
struct SpinGuard {
  u64 t0;
  u32 polls;
  __device__ __forceinline__ SpinGuard() : t0(0), polls(0) {}
  __device__ __forceinline__ void tick() {
    if ((++polls & 0xFFFu) == 0) {
      u64 now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) __trap();
    }
  }
};

/* Grid-wide barrier of a cooperative launch (every CTA resident): `target` = arrivals expected so far (the counter is never
 * reset: barrier number b of a launch of G CTAs waits for b * G). */
__device__ __forceinline__ void grid_barrier(u32* bar, u32 target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    SpinGuard guard;
    while (ld_acquire(bar) < target) guard.tick();
  }
  __syncthreads();
}

/* order-preserving float <-> uint map: unsigned compare of the image == float compare (-0 < +0) */
__device__ __forceinline__ u32 float_to_ordered(float f) {
  const u32 b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(u32 k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt() {
  u32 m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

/* A CTA's contiguous output staged as 32-bit words in shared memory leaves as 16-byte stores covering whole sectors
 * (records of 24 or 28 bytes written field by field from one thread each cost one partial-sector store per field).
 * `dst` must be 16-byte aligned; all threads of the CTA call it after a __syncthreads(). */
__device__ __forceinline__ void cta_store_words(u32* dst, const u32* smemWords, u32 nWords) {
  const u32 n4 = nWords >> 2;
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const uint4* s4 = reinterpret_cast<const uint4*>(smemWords);
  for (u32 q = threadIdx.x; q < n4; q += blockDim.x) d4[q] = s4[q];
  for (u32 q = (n4 << 2) + threadIdx.x; q < nWords; q += blockDim.x) dst[q] = smemWords[q];
}

/* ------------------------------------------------------------------ host side */
struct b2bvh_ctx {
  int device;
  int sm_count;
  cudaStream_t stream;
  bool own_stream;
  cudaStream_t dl_stream; /* device -> host copies run here, ordered after `stream` by dl_event: a stream that has carried a D2H copy
                             no longer overlaps its H2D copies with another context's D2H (measured, tools/micro/e2e_pipeline_probe.py) */
  cudaEvent_t dl_event;
  cudaEvent_t ev[16];
  char name[256];
  /* build-owned device buffers, grown on demand and reused across builds */
  struct Buf {
    void* p;
    size_t cap;
  } bufs[48];
  u32* mailbox;          /* pinned + mapped host memory, B2_MAILBOX_SLOTS x 16 words: small results come back through a one-warp
                            kernel instead of the copy engine, where a 4-byte read would queue behind another context's bulk copy */
  u32* mailbox_dev;      /* the same memory as the device sees it */
  /* one cached CUDA graph of a whole build (b2bvh_build_opts.use_graph) */
  struct GraphCache {
    cudaGraphExec_t exec;
    int algo;
    u32 n, launches, epoch;
    const void* tris;
    b2bvh_build_opts opts;
  } graph;
  struct Pending { /* a build enqueued with defer_sync, waiting for b2bvh_build_finish */
    bool active, collapse, host_tris, split;
    int algo;
  } pending;
  /* one-time per-context (= per-device) kernel setup: cudaFuncSetAttribute applies to the CURRENT device only, so the opt-ins and the
   * occupancy answers live here and not in process-wide statics (a second context on another device needs its own) */
  u32 once_mask;         /* B2_ONCE_* bits already done on this context's device */
  int occ[8];            /* B2_OCC_* occupancy answers */
  u32 meet_clean;        /* leading 64-bit words of SLOT_MEET known to hold all ones (b2_meet_acquire): the LBVH climb leaves its exchange words as it found them */
  u32 alloc_epoch;       /* bumped whenever a build-owned buffer is (re)allocated: a cached graph holds the old pointers */
  u32 launches;
  u32 lbvh_second_level; /* b2bvh_build_opts.lbvh_second_level of the running build */
  u32 merge_max_ctas;    /* b2bvh_build_opts.merge_max_ctas of the running build */
  u32* ref_leaf_prim;    /* early split: triangle of Bvh2 leaf g, written by the LBVH kernels for the collapse */
  const u32* ref_prim;   /* early split (b2bvh_build_opts.split_sa_max): triangle of every reference of the running build, else NULL */
  /* optional per-launch profiler (b2bvh_profile_*): one CUDA event pair per kernel launch */
  bool prof_on;
  int prof_n;
  struct Prof {
    const char* name;
    cudaEvent_t a, b;
  } prof[512];
  int prof_events; /* number of event pairs created so far */
};

enum { B2_ONCE_SORT4 = 1u << 0, B2_ONCE_SORT15 = 1u << 1, B2_ONCE_LBVH32 = 1u << 2, B2_ONCE_LBVH64 = 1u << 3, B2_ONCE_COLLAPSE = 1u << 4, B2_ONCE_PLOC = 1u << 5,
       B2_ONCE_SORT3 = 1u << 6, B2_ONCE_MISC = 1u << 7 };
enum { B2_OCC_COLLAPSE_LARGE = 0, B2_OCC_COLLAPSE_SMALL = 1, B2_OCC_PLOC = 2 };

int b2_fail(int code, const char* fmt, ...);
int b2_check(cudaError_t e, const char* what);
#define B2_CUDA(x)                                 \
  do {                                             \
    int _s = b2_check((x), #x);                    \
    if (_s) return _s;                             \
  } while (0)
#define B2_TRY(x)              \
  do {                         \
    int _s = (x);              \
    if (_s) return _s;         \
  } while (0)
int b2_prof_begin(b2bvh_ctx* ctx, const char* name);
int b2_prof_end(b2bvh_ctx* ctx);
/* Programmatic dependent launch (sm_90+): a kernel launched through b2_launch_pdl may be made resident while its predecessor in the stream is
 * still running; it must not touch global memory before pdl_wait() (which returns once the predecessor has completed and its writes are
 * visible).  Without an explicit pdl_launch_dependents() in the predecessor the successor comes in as the predecessor's CTAs exit: what
 * overlaps is the launch latency, the prologue (shared-memory clearing, barrier initialisation) and the drain of the short kernels of a
 * chain.  (Letting the successor in from the START of the predecessor made the 10 M sort 9 % slower: resident, waiting CTAs of the next
 * kernel get in the way of the tail of this one; profiles/r02_experiments.txt.)  Launched the ordinary way, both instructions do nothing. */
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t b2_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define B2_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) B2_CUDA(b2_launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__))

/* bracket of every kernel launch: B2_KERNEL(ctx, "name"); kernel<<<...>>>(...); B2_LAUNCH_CHECK(ctx); */
#define B2_KERNEL(ctx, name)                          \
  do {                                                \
    if ((ctx)->prof_on) B2_TRY(b2_prof_begin((ctx), (name))); \
  } while (0)
#define B2_LAUNCH_CHECK(ctx)                          \
  do {                                                \
    (ctx)->launches++;                                \
    B2_CUDA(cudaGetLastError());                      \
    if ((ctx)->prof_on) B2_TRY(b2_prof_end((ctx)));   \
  } while (0)

/* context-owned device buffers (b2bvh_ctx::bufs) */
enum {
  SLOT_TRIS = 0, SLOT_AABB, SLOT_CTL, SLOT_KEYS, SLOT_VALS, SLOT_SKEYS, SLOT_SVALS, SLOT_TKEYS, SLOT_TVALS, SLOT_SORT, SLOT_NODES,
  SLOT_PARENTS, SLOT_LBVH, SLOT_WIDE, SLOT_WLEAVES, SLOT_COLLAPSE, SLOT_LEAVES, SLOT_PLOC, SLOT_HPLOC, SLOT_MISC,
  SLOT_SPLIT_BOX, SLOT_SPLIT_PRIM, SLOT_SPLIT_LIST_A, SLOT_SPLIT_LIST_B, SLOT_SPLIT_STATUS, SLOT_SPLIT_LEAFPRIM,
  SLOT_BATCH_NODES, SLOT_BATCH_LEAVES, SLOT_BATCH_ROOTS, SLOT_BATCH_SCENES, SLOT_BATCH_OFFSETS,
  SLOT_KEYS_LO, SLOT_KEYS64, SLOT_SKEYS64, SLOT_M60_KEYS, SLOT_M60_VALS, SLOT_MEET, SLOT_COUNT
};
/* SLOT_CTL (256 B): [0..23] scene box, [32..63] extents scratch8, [64..87] {-min,max}, [96] root index, [128] range-extract count, [160] traversal overflow flag, [192] range-extract root, [224..247] root box of a sharded build */
int b2_reserve(b2bvh_ctx* ctx, int slot, size_t bytes, void** out);
/* the n-1 64-bit exchange words of the LBVH climb (SLOT_MEET), all ones: filled when the buffer is new or grew, never again (lbvh.cu).  Allocates:
 * call it before a stream capture begins; the launchers call it again, which is then free. */
int b2_meet_acquire(b2bvh_ctx* ctx, u32 n, u64** out);
/* words <= 16 from device memory into mailbox slot `slot`; readable at b2_mailbox(ctx, slot) after the next stream synchronisation */
#define B2_MAILBOX_SLOTS 8
enum { B2_MB_COLLAPSE = 0, B2_MB_PLOC = 1, B2_MB_HPLOC = 2, B2_MB_ROOT = 3, B2_MB_SPLIT = 4, B2_MB_RANGE = 5, B2_MB_TRAVERSE = 6, B2_MB_GLOBAL = 7 };
int b2_fetch_words(b2bvh_ctx* ctx, const void* d_src, u32 words, int slot);
static inline const u32* b2_mailbox(const b2bvh_ctx* ctx, int slot) { return ctx->mailbox + slot * 16; }

/* stage launchers (one per .cu file) */
int b2_launch_extents(b2bvh_ctx* ctx, const b2bvh_triangle* d_tris, u32 n, b2bvh_aabb* d_triAabb, b2bvh_aabb* d_scene, u32* d_scratch8,
                      float* d_negmin_max6);
int b2_launch_scene_from_negmin_max(b2bvh_ctx* ctx, const float* d_negmin_max6, b2bvh_aabb* d_scene);
int b2_launch_morton(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const b2bvh_aabb* d_scene, u32 n, u32* d_keys, u32* d_vals);
size_t b2_sort_scratch_bytes(u32 n);
int b2_launch_sort(b2bvh_ctx* ctx, const u32* d_keysIn, const u32* d_valsIn, u32* d_keysOut, u32* d_valsOut, u32* d_keysTmp, u32* d_valsTmp,
                   void* d_scratch, u32 n, u32 startBit, u32 endBit);
size_t b2_lbvh_scratch_bytes(u32 n);
int b2_launch_lbvh_fused(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                         b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch /* b2_lbvh_scratch_bytes(n) */, u32* d_root, int karrasNumbering);
int b2_launch_lbvh_karras_two_kernel(b2bvh_ctx* ctx, const u32* d_sortedKeys, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                                     b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_flags);
int b2_launch_lbvh_fused64(b2bvh_ctx* ctx, const u64* d_sortedKeys64, const u32* d_sortedVals, const b2bvh_aabb* d_triAabb, u32 n,
                           b2bvh_bvh2_node* d_nodes, u32* d_parents, u32* d_scratch, u32* d_root, int karrasNumbering);
/* 60-bit Morton variant (morton60.cu) */
int b2_launch_morton60(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const b2bvh_aabb* d_scene, u32 n, u32* d_hi, u32* d_lo, u64* d_keys64);
int b2_launch_sort60(b2bvh_ctx* ctx, const u32* d_hi, const u32* d_lo, u32 n, u32* d_a, u32* d_aVals, u32* d_hiSorted, u32* d_valsSorted, u64* d_keys64Sorted,
                     u32* d_keysTmp, u32* d_valsTmp, void* d_sortScratch);
int b2_launch_root_box(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const u32* d_rootIdx, float* d_box6);
int b2_launch_range_extract(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_local, u32 m, const u32* d_root, int karras, u32 ghostL, u32 ghostR, u32 firstPos, u32 nGlobal,
                            unsigned char* d_flags, b2bvh_bvh2_node* d_out, b2bvh_cluster* d_clusters, u32* d_count);
size_t b2_collapse_scratch_bytes(u32 n);
int b2_launch_collapse(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const b2bvh_prim_ref* d_leaves, const u32* d_sortedVals, const u32* d_rootIdx,
                       u32 n,
                       b2bvh_bvh4_node* d_wide, b2bvh_prim_node* d_wideLeaves, void* d_scratch, u32* h_nWide);
size_t b2_ploc_scratch_bytes(u32 n);
int b2_launch_ploc(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const u32* d_sortedVals, u32 n, b2bvh_bvh2_node* d_nodes,
                   b2bvh_prim_ref* d_leaves, void* d_scratch, u32* h_iterations);
/* early split clipping (split.cu): references of all generations in emission order; synchronises the stream once per generation */
int b2_launch_split(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, u32 n, float saMax, int slotOutBox, int slotOutPrim, int slotListA, int slotListB,
                    int slotStatus, b2bvh_aabb** d_refBox, u32** d_refPrim, u32* h_count, u32* h_levels);
int b2_launch_batched(b2bvh_ctx* ctx, const b2bvh_triangle* d_tris, const u32* d_leafOff, const u32* d_nodeOff, u32 nItems, b2bvh_bvh2_node* d_nodes,
                      b2bvh_prim_ref* d_leaves, u32* d_roots, b2bvh_aabb* d_scenes);
int b2_launch_hploc_keys(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const u32* d_sortedKeys, const u64* d_sortedKeys64, const u32* d_sortedVals, u32 n,
                         b2bvh_bvh2_node* d_nodes, b2bvh_prim_ref* d_leaves, void* d_scratch, u32* h_mergeCalls);
size_t b2_hploc_scratch_bytes(u32 n);
int b2_launch_hploc(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const u32* d_sortedKeys, const u32* d_sortedVals, u32 n,
                    b2bvh_bvh2_node* d_nodes, b2bvh_prim_ref* d_leaves, void* d_scratch, u32* h_mergeCalls);
