/*
 * api.cu — the C ABI of libb2bvh.so (include/b2bvh.h): context, memory, the build orchestration that replaces the
 * bodies of TwoPassLbvh::build / SinglePassLbvh::build / PLOCNew::build / HPLOC::build (src/*.cpp), and the host-side
 * SAH cost reporting (Utility::calculatebvh4Cost / calculateLbvhCost, Utility.cpp:317-396).
 *
 * Differences from the reference's host flow that do not change results:
 *   - kernels are compiled ahead of time for sm_100a (no RTC per launch, Kernel.cpp:52-122), the sort object and all
 *     device buffers live in the context and are reused across builds (TwoPassLbvh.cpp:73 constructs RadixSort per build);
 *   - no blocking event-sync after every launch (Timer.h:48-56): stage boundaries are CUDA events on one stream and
 *     are read once at the end;
 *   - primitive boxes are computed on the device (the reference's TwoPass path does it on the host in
 *     doEarlySplitClipping with splitting disabled).
 */
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

int b2_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int b2_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return b2_fail(e == cudaErrorMemoryAllocation ? B2BVH_ERR_OOM : B2BVH_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int b2_reserve(b2bvh_ctx* ctx, int slot, size_t bytes, void** out) {
  b2bvh_ctx::Buf& b = ctx->bufs[slot];
  if (bytes == 0) bytes = 16;
  if (b.cap < bytes) {
    if (b.p) { B2_CUDA(cudaStreamSynchronize(ctx->stream)); B2_CUDA(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    const size_t want = (bytes + 255) & ~(size_t)255;
    B2_CUDA(cudaMalloc(&b.p, want));
    b.cap = want;
    ctx->alloc_epoch++;
  }
  *out = b.p;
  return 0;
}


extern "C" {

uint32_t b2bvh_abi_version(void) { return B2BVH_ABI_VERSION; }
const char* b2bvh_last_error(void) { return g_err; }

int b2bvh_ctx_destroy(b2bvh_ctx* ctx);
int b2bvh_ctx_create(int device, void* cuda_stream, b2bvh_ctx** out) {
  if (!out) return b2_fail(B2BVH_ERR_INVALID, "ctx_create: out is null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return b2_fail(B2BVH_ERR_CUDA, "ctx_create: no CUDA device (%s); libb2bvh has no CPU fallback", e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return b2_fail(B2BVH_ERR_INVALID, "ctx_create: device %d out of range [0,%d)", device, count);
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0)
    return b2_fail(B2BVH_ERR_CUDA, "ctx_create: device '%s' is sm_%d%d; libb2bvh is built for sm_100a only", prop.name, prop.major, prop.minor);
  b2bvh_ctx* c = new b2bvh_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  snprintf(c->name, sizeof(c->name), "%s", prop.name);
  /* a failure half way must not leak what was created so far: everything below goes through `st` and one exit */
  int st = 0;
  auto cu = [&](cudaError_t e, const char* what) { if (!st) st = b2_check(e, what); };
  if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
  else { cu(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate"); c->own_stream = st == 0; }
  for (int i = 0; i < 16 && !st; i++) cu(cudaEventCreate(&c->ev[i]), "cudaEventCreate");
  if (!st) cu(cudaStreamCreateWithFlags(&c->dl_stream, cudaStreamNonBlocking), "cudaStreamCreate (download)");
  if (!st) cu(cudaEventCreateWithFlags(&c->dl_event, cudaEventDisableTiming), "cudaEventCreate (download)");
  if (!st) cu(cudaHostAlloc(reinterpret_cast<void**>(&c->mailbox), B2_MAILBOX_SLOTS * 16 * sizeof(u32), cudaHostAllocMapped), "cudaHostAlloc (mailbox)");
  if (!st) cu(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->mailbox_dev), c->mailbox, 0), "cudaHostGetDevicePointer");
  void* ctl = nullptr;
  if (!st) st = b2_reserve(c, SLOT_CTL, 256, &ctl);
  if (!st) cu(cudaMemsetAsync(ctl, 0, 256, c->stream), "cudaMemsetAsync");
  if (!st) cu(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
  if (st) {
    char keep[sizeof(g_err)];
    memcpy(keep, g_err, sizeof(keep)); /* the message of the first failure, not of the clean-up */
    b2bvh_ctx_destroy(c);
    memcpy(g_err, keep, sizeof(keep));
    return st;
  }
  *out = c;
  return 0;
}

int b2bvh_ctx_destroy(b2bvh_ctx* ctx) {
  if (!ctx) return 0;
  /* also the clean-up path of a context that b2bvh_ctx_create could not finish: every member may still be null */
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 48; i++) if (ctx->bufs[i].p) cudaFree(ctx->bufs[i].p);
  for (int i = 0; i < 16; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < ctx->prof_events; i++) { cudaEventDestroy(ctx->prof[i].a); cudaEventDestroy(ctx->prof[i].b); }
  if (ctx->graph.exec) cudaGraphExecDestroy(ctx->graph.exec);
  if (ctx->mailbox) cudaFreeHost(ctx->mailbox);
  if (ctx->dl_stream) { cudaStreamSynchronize(ctx->dl_stream); cudaStreamDestroy(ctx->dl_stream); }
  if (ctx->dl_event) cudaEventDestroy(ctx->dl_event);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
  return 0;
}

int b2bvh_device_name(b2bvh_ctx* ctx, char* buf, size_t cap) {
  if (!ctx || !buf || cap == 0) return b2_fail(B2BVH_ERR_INVALID, "device_name: bad argument");
  snprintf(buf, cap, "%s", ctx->name);
  return 0;
}
int b2bvh_device_sm_count(b2bvh_ctx* ctx, int* out) {
  if (!ctx || !out) return b2_fail(B2BVH_ERR_INVALID, "device_sm_count: bad argument");
  *out = ctx->sm_count;
  return 0;
}
int b2bvh_alloc(b2bvh_ctx* ctx, size_t bytes, void** dptr) {
  if (!ctx || !dptr) return b2_fail(B2BVH_ERR_INVALID, "alloc: bad argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
  return 0;
}
int b2bvh_free(b2bvh_ctx* ctx, void* dptr) {
  if (!ctx) return b2_fail(B2BVH_ERR_INVALID, "free: bad argument");
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  B2_CUDA(cudaFree(dptr));
  return 0;
}
int b2bvh_memset(b2bvh_ctx* ctx, void* dptr, int value, size_t bytes) {
  if (!ctx || !dptr) return b2_fail(B2BVH_ERR_INVALID, "memset: bad argument");
  B2_CUDA(cudaMemsetAsync(dptr, value, bytes, ctx->stream));
  return 0;
}
int b2bvh_h2d(b2bvh_ctx* ctx, void* dptr, const void* hptr, size_t bytes) {
  if (!ctx || (bytes && (!dptr || !hptr))) return b2_fail(B2BVH_ERR_INVALID, "h2d: bad argument");
  B2_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
/* device -> host copies of a context run on its download stream, after everything enqueued on the main stream so far */
static int b2_download(b2bvh_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  B2_CUDA(cudaEventRecord(ctx->dl_event, ctx->stream));
  B2_CUDA(cudaStreamWaitEvent(ctx->dl_stream, ctx->dl_event, 0));
  B2_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, ctx->dl_stream));
  return 0;
}
int b2bvh_d2h(b2bvh_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  if (!ctx || (bytes && (!dptr || !hptr))) return b2_fail(B2BVH_ERR_INVALID, "d2h: bad argument");
  B2_TRY(b2_download(ctx, hptr, dptr, bytes));
  B2_CUDA(cudaStreamSynchronize(ctx->dl_stream));
  return 0;
}
int b2bvh_h2d_async(b2bvh_ctx* ctx, void* dptr, const void* hptr, size_t bytes) {
  if (!ctx || (bytes && (!dptr || !hptr))) return b2_fail(B2BVH_ERR_INVALID, "h2d_async: bad argument");
  B2_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return 0;
}
int b2bvh_d2h_async(b2bvh_ctx* ctx, void* hptr, const void* dptr, size_t bytes) {
  if (!ctx || (bytes && (!dptr || !hptr))) return b2_fail(B2BVH_ERR_INVALID, "d2h_async: bad argument");
  return b2_download(ctx, hptr, dptr, bytes);
}
int b2bvh_d2d(b2bvh_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx || (bytes && (!dst || !src))) return b2_fail(B2BVH_ERR_INVALID, "d2d: bad argument");
  B2_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
int b2bvh_sync(b2bvh_ctx* ctx) {
  if (!ctx) return b2_fail(B2BVH_ERR_INVALID, "sync: bad argument");
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->dl_stream));
  return 0;
}
int b2bvh_host_alloc_pinned(size_t bytes, void** hptr) {
  if (!hptr) return b2_fail(B2BVH_ERR_INVALID, "host_alloc_pinned: bad argument");
  B2_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}
int b2bvh_host_free_pinned(void* hptr) {
  B2_CUDA(cudaFreeHost(hptr));
  return 0;
}

/* ------------------------------------------------------------------ individually callable stages */
int b2bvh_scene_extents(b2bvh_ctx* ctx, const b2bvh_triangle* d_tris, uint32_t n, b2bvh_aabb* d_triAabb, b2bvh_aabb* d_scene) {
  if (!ctx || !d_tris || !d_triAabb || !d_scene || n == 0) return b2_fail(B2BVH_ERR_INVALID, "scene_extents: bad argument");
  unsigned char* ctl = (unsigned char*)ctx->bufs[SLOT_CTL].p;
  return b2_launch_extents(ctx, d_tris, n, d_triAabb, d_scene, (u32*)(ctl + 32), nullptr);
}
int b2bvh_morton_codes(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const b2bvh_aabb* d_scene, uint32_t n, uint32_t* d_keys, uint32_t* d_vals) {
  if (!ctx || !d_triAabb || !d_scene || !d_keys || !d_vals || n == 0) return b2_fail(B2BVH_ERR_INVALID, "morton_codes: bad argument");
  return b2_launch_morton(ctx, d_triAabb, d_scene, n, d_keys, d_vals);
}
int b2bvh_sort_pairs(b2bvh_ctx* ctx, const uint32_t* d_keysIn, const uint32_t* d_valsIn, uint32_t* d_keysOut, uint32_t* d_valsOut, uint32_t n,
                     uint32_t startBit, uint32_t endBit) {
  if (!ctx || !d_keysIn || !d_keysOut || !d_valsOut) return b2_fail(B2BVH_ERR_INVALID, "sort_pairs: bad argument");
  /* the passes ping-pong between the context's temporaries and the caller's OUTPUT arrays, so the outputs are read back with 16-byte
   * vector loads and bulk copies as well */
  if (((uintptr_t)d_keysIn | (uintptr_t)d_valsIn | (uintptr_t)d_keysOut | (uintptr_t)d_valsOut) & 15)
    return b2_fail(B2BVH_ERR_INVALID, "sort_pairs: inputs and outputs must be 16-byte aligned");
  void *tk, *tv, *sc;
  B2_TRY(b2_reserve(ctx, SLOT_TKEYS, (size_t)n * 4, &tk));
  B2_TRY(b2_reserve(ctx, SLOT_TVALS, (size_t)n * 4, &tv));
  B2_TRY(b2_reserve(ctx, SLOT_SORT, b2_sort_scratch_bytes(n), &sc));
  return b2_launch_sort(ctx, d_keysIn, d_valsIn, d_keysOut, d_valsOut, (u32*)tk, (u32*)tv, sc, n, startBit, endBit);
}

/* the hierarchy stage alone, over caller-provided SORTED 64-bit keys: the building block of the globally sorted multi-GPU build
 * (DESIGN.md section 9), where a rank runs it over its range of the global order with keys widened to (code << 32 | global position) */
int b2bvh_lbvh_from_sorted64(b2bvh_ctx* ctx, const uint64_t* d_sortedKeys64, const uint32_t* d_sortedVals, const b2bvh_aabb* d_primAabb, uint32_t n,
                             int karras, b2bvh_bvh2_node* d_nodes, uint32_t* d_parents, uint32_t* root) {
  if (!ctx || !d_sortedKeys64 || !d_sortedVals || !d_primAabb || !d_nodes || !root) return b2_fail(B2BVH_ERR_INVALID, "lbvh_from_sorted64: null argument");
  if (n < 2 || n > 0x3FFFFFFFu) return b2_fail(B2BVH_ERR_INVALID, "lbvh_from_sorted64: n=%u out of range", n);
  if (karras && !d_parents) return b2_fail(B2BVH_ERR_INVALID, "lbvh_from_sorted64: the Karras numbering writes parent indices (2n-1 words)");
  B2_CUDA(cudaSetDevice(ctx->device));
  void* dLbvh;
  B2_TRY(b2_reserve(ctx, SLOT_LBVH, b2_lbvh_scratch_bytes(n), &dLbvh));
  u32* dRoot = (u32*)((unsigned char*)ctx->bufs[SLOT_CTL].p + 96);
  ctx->ref_prim = nullptr; ctx->ref_leaf_prim = nullptr; ctx->lbvh_second_level = 0;
  B2_TRY(b2_launch_lbvh_fused64(ctx, (const u64*)d_sortedKeys64, d_sortedVals, d_primAabb, n, d_nodes, karras ? d_parents : nullptr, (u32*)dLbvh, dRoot, karras ? 1 : 0));
  if (karras) B2_CUDA(cudaMemsetAsync(dRoot, 0, 4, ctx->stream));
  B2_TRY(b2_fetch_words(ctx, dRoot, 1, B2_MB_ROOT));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  *root = b2_mailbox(ctx, B2_MB_ROOT)[0];
  return 0;
}

int b2bvh_range_extract(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_local, uint32_t m, uint32_t local_root, int karras, uint32_t ghost_left,
                        uint32_t ghost_right, uint32_t first_pos, uint32_t n_global, b2bvh_bvh2_node* d_out, b2bvh_cluster* d_clusters, uint32_t* count) {
  if (!ctx || !d_local || !d_out || !d_clusters || !count) return b2_fail(B2BVH_ERR_INVALID, "range_extract: null argument");
  if (m < 2 || (uint64_t)first_pos + m > n_global) return b2_fail(B2BVH_ERR_INVALID, "range_extract: range [%u, %u + %u) does not fit %u positions", first_pos, first_pos, m, n_global);
  B2_CUDA(cudaSetDevice(ctx->device));
  void* dFlags;
  B2_TRY(b2_reserve(ctx, SLOT_MISC, (size_t)m + 64, &dFlags));
  u32* dCount = (u32*)((unsigned char*)ctx->bufs[SLOT_CTL].p + 128);
  u32* dRootIn = (u32*)((unsigned char*)ctx->bufs[SLOT_CTL].p + 192);
  B2_CUDA(cudaMemcpyAsync(dRootIn, &local_root, sizeof(u32), cudaMemcpyHostToDevice, ctx->stream)); /* pageable source: staged before the call returns */
  B2_TRY(b2_launch_range_extract(ctx, d_local, m, dRootIn, karras, ghost_left, ghost_right, first_pos, n_global, (unsigned char*)dFlags, d_out, d_clusters, dCount));
  B2_TRY(b2_fetch_words(ctx, dCount, 1, B2_MB_RANGE));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  *count = b2_mailbox(ctx, B2_MB_RANGE)[0];
  if (*count > 256u) return b2_fail(B2BVH_ERR_INTERNAL, "range_extract: %u left-over clusters (at most 256 expected)", *count);
  return 0;
}

/* ------------------------------------------------------------------ the build */
int b2bvh_build(b2bvh_ctx* ctx, int algo, const b2bvh_triangle* tris, uint32_t n, const b2bvh_build_opts* optsIn, b2bvh_tree* out) {
  if (!ctx || !tris || !out) return b2_fail(B2BVH_ERR_INVALID, "build: null argument");
  if (algo < B2BVH_TWO_PASS_LBVH || algo > B2BVH_HPLOC) return b2_fail(B2BVH_ERR_INVALID, "build: unknown algo %d", algo);
  if (n < 2) return b2_fail(B2BVH_ERR_INVALID, "build: need at least 2 primitives, got %u", n);
  if (n > 0x3FFFFFFFu) return b2_fail(B2BVH_ERR_INVALID, "build: n=%u exceeds 2^30-1", n);
  b2bvh_build_opts opts;
  memset(&opts, 0, sizeof(opts));
  if (optsIn) opts = *optsIn; else opts.collapse = 1;
  B2_CUDA(cudaSetDevice(ctx->device));
  memset(out, 0, sizeof(*out));
  ctx->lbvh_second_level = opts.lbvh_second_level;
  ctx->merge_max_ctas = opts.merge_max_ctas;
  ctx->ref_prim = nullptr;
  ctx->ref_leaf_prim = nullptr;
  const bool separate = (algo == B2BVH_PLOCPP || algo == B2BVH_HPLOC);
  const u32 launches0 = ctx->launches;
  cudaStream_t s = ctx->stream;
  /* early split clipping (USE_PRIM_SPLITTING): only TwoPassLbvh builds over PrimRefs end to end — SinglePassLbvh with the macro
   * indexes the triangle array with reference indices (InitBvhNodes, SinglePassLbvh.cpp:112), PLOC++/H-PLOC never split */
  const bool split = opts.split_sa_max > 0.0f;
  if (split && algo != B2BVH_TWO_PASS_LBVH) return b2_fail(B2BVH_ERR_INVALID, "build: split_sa_max needs B2BVH_TWO_PASS_LBVH (algo %d)", algo);
  if (split && (opts.boxes_ready || opts.use_scene_box || opts.d_scene_negmin_max))
    return b2_fail(B2BVH_ERR_INVALID, "build: split_sa_max cannot be combined with the sharded-build options");
  const u32 nTris = n;
  const bool m60 = opts.morton_bits == 60;
  if (opts.morton_bits != 0 && opts.morton_bits != 30 && opts.morton_bits != 60)
    return b2_fail(B2BVH_ERR_INVALID, "build: morton_bits must be 0/30 (the reference's extended 30-bit code) or 60, got %u", opts.morton_bits);
  if (m60 && opts.karras_two_kernel) return b2_fail(B2BVH_ERR_INVALID, "build: morton_bits=60 cannot be combined with karras_two_kernel");

  /* ---- buffers (grown on demand, reused) ---- */
  void *dTris = nullptr, *dAabb, *dKeys, *dVals, *dSKeys, *dSVals, *dTKeys, *dTVals, *dSort, *dNodes, *dParents = nullptr, *dLbvh = nullptr,
       *dWide = nullptr, *dWLeaves = nullptr, *dCollapse = nullptr, *dLeaves = nullptr, *dMerge = nullptr;
  unsigned char* ctl = (unsigned char*)ctx->bufs[SLOT_CTL].p;
  b2bvh_aabb* dScene = (b2bvh_aabb*)ctl;
  u32* dScratch8 = (u32*)(ctl + 32);
  u32* dRoot = (u32*)(ctl + 96);
  if (!opts.tris_on_device) B2_TRY(b2_reserve(ctx, SLOT_TRIS, (size_t)n * sizeof(b2bvh_triangle), &dTris));
  B2_TRY(b2_reserve(ctx, SLOT_AABB, (size_t)n * sizeof(b2bvh_aabb), &dAabb));
  const b2bvh_triangle* dT = tris;
  if (!opts.tris_on_device) dT = (const b2bvh_triangle*)dTris;
  u32* dRefPrim = nullptr;
  u32 splitLevels = 0;
  if (split) {
    /* upload + S1 + the split run ahead of everything whose size depends on the reference count (one sync per generation) */
    if (!opts.tris_on_device) {
      B2_CUDA(cudaEventRecord(ctx->ev[8], s));
      B2_CUDA(cudaMemcpyAsync(dTris, tris, (size_t)n * sizeof(b2bvh_triangle), cudaMemcpyHostToDevice, s));
      B2_CUDA(cudaEventRecord(ctx->ev[9], s));
    }
    B2_CUDA(cudaEventRecord(ctx->ev[0], s));
    B2_TRY(b2_launch_extents(ctx, dT, n, (b2bvh_aabb*)dAabb, dScene, dScratch8, nullptr));
    B2_CUDA(cudaEventRecord(ctx->ev[10], s));
    b2bvh_aabb* refBox = nullptr;
    u32 m = 0;
    B2_TRY(b2_launch_split(ctx, (const b2bvh_aabb*)dAabb, n, opts.split_sa_max, SLOT_SPLIT_BOX, SLOT_SPLIT_PRIM, SLOT_SPLIT_LIST_A, SLOT_SPLIT_LIST_B,
                           SLOT_SPLIT_STATUS, &refBox, &dRefPrim, &m, &splitLevels));
    B2_CUDA(cudaEventRecord(ctx->ev[11], s));
    if (m < 2) return b2_fail(B2BVH_ERR_INTERNAL, "early split produced %u references", m);
    n = m;               /* from here on the primitives of the build are the references */
    void* dLeafPrim = nullptr;
    B2_TRY(b2_reserve(ctx, SLOT_SPLIT_LEAFPRIM, (size_t)n * 4, &dLeafPrim));
    ctx->ref_leaf_prim = (u32*)dLeafPrim;
    ctx->ref_prim = dRefPrim; /* leaves and PrimNodes name the reference's TRIANGLE (InitBvhNodesPrimRef, TwoPassLbvhKernel.h:178-182, :324) */
    dAabb = refBox;
  }
  B2_TRY(b2_reserve(ctx, SLOT_KEYS, (size_t)n * 4, &dKeys));
  B2_TRY(b2_reserve(ctx, SLOT_VALS, (size_t)n * 4, &dVals));
  B2_TRY(b2_reserve(ctx, SLOT_SKEYS, (size_t)n * 4, &dSKeys));
  B2_TRY(b2_reserve(ctx, SLOT_SVALS, (size_t)n * 4, &dSVals));
  B2_TRY(b2_reserve(ctx, SLOT_TKEYS, (size_t)n * 4, &dTKeys));
  B2_TRY(b2_reserve(ctx, SLOT_TVALS, (size_t)n * 4, &dTVals));
  B2_TRY(b2_reserve(ctx, SLOT_SORT, b2_sort_scratch_bytes(n), &dSort));
  void *dKeysLo = nullptr, *dKeys64 = nullptr, *dSKeys64 = nullptr, *dM60K = nullptr, *dM60V = nullptr;
  if (m60) {
    B2_TRY(b2_reserve(ctx, SLOT_KEYS_LO, (size_t)n * 4, &dKeysLo));
    B2_TRY(b2_reserve(ctx, SLOT_KEYS64, (size_t)n * 8, &dKeys64));
    B2_TRY(b2_reserve(ctx, SLOT_SKEYS64, (size_t)n * 8, &dSKeys64));
    B2_TRY(b2_reserve(ctx, SLOT_M60_KEYS, (size_t)n * 4, &dM60K));
    B2_TRY(b2_reserve(ctx, SLOT_M60_VALS, (size_t)n * 4, &dM60V));
  }
  B2_TRY(b2_reserve(ctx, SLOT_NODES, (size_t)(separate ? n - 1 : 2 * (size_t)n - 1) * sizeof(b2bvh_bvh2_node), &dNodes));
  if (algo == B2BVH_TWO_PASS_LBVH) B2_TRY(b2_reserve(ctx, SLOT_PARENTS, (2 * (size_t)n - 1) * 4, &dParents));
  if (!separate) B2_TRY(b2_reserve(ctx, SLOT_LBVH, b2_lbvh_scratch_bytes(n), &dLbvh));
  if (!separate) { u64* dMeet = nullptr; B2_TRY(b2_meet_acquire(ctx, n, &dMeet)); } /* allocates and fills on first use: not inside a capture */
  if (separate) B2_TRY(b2_reserve(ctx, SLOT_LEAVES, (size_t)n * sizeof(b2bvh_prim_ref), &dLeaves));
  if (algo == B2BVH_PLOCPP) B2_TRY(b2_reserve(ctx, SLOT_PLOC, b2_ploc_scratch_bytes(n), &dMerge));
  if (algo == B2BVH_HPLOC) B2_TRY(b2_reserve(ctx, SLOT_HPLOC, b2_hploc_scratch_bytes(n), &dMerge));
  if (opts.collapse) {
    B2_TRY(b2_reserve(ctx, SLOT_WIDE, (size_t)n * sizeof(b2bvh_bvh4_node), &dWide));
    B2_TRY(b2_reserve(ctx, SLOT_WLEAVES, (size_t)n * sizeof(b2bvh_prim_node), &dWLeaves));
    B2_TRY(b2_reserve(ctx, SLOT_COLLAPSE, b2_collapse_scratch_bytes(n), &dCollapse));
  }

  /* ---- the launch sequence: enqueued directly, or captured once into a CUDA graph and replayed (opts.use_graph) ---- */
  u32 iterations = 0, nWide = 0;
  bool capturing = false;
  /* stage events: inside a capture they must become event-record NODES (cudaEventRecordExternal), or they cannot be timed */
  /* (the four event nodes between the stages cost 7-8 us of a 1.6 ms replay at 10 M primitives: measured with them left out, gpurun r2y) */
  auto record = [&](int i) { return capturing ? cudaEventRecordWithFlags(ctx->ev[i], s, cudaEventRecordExternal) : cudaEventRecord(ctx->ev[i], s); };
  auto enqueue = [&]() -> int {
  /* ---- upload (TwoPassLbvh.cpp:19-20) ---- */
  if (!opts.tris_on_device && !split) {
    B2_CUDA(record(8));
    /* boxes_ready: b2bvh_shard_extents already uploaded these triangles into the same buffer */
    if (!opts.boxes_ready) B2_CUDA(cudaMemcpyAsync(dTris, tris, (size_t)n * sizeof(b2bvh_triangle), cudaMemcpyHostToDevice, s));
    B2_CUDA(record(9));
  }

  /* ---- S1 extents ---- */
  if (!split) B2_CUDA(record(0));
  if (!opts.boxes_ready && !split) B2_TRY(b2_launch_extents(ctx, dT, n, (b2bvh_aabb*)dAabb, dScene, dScratch8, nullptr));
  if (opts.d_scene_negmin_max) B2_TRY(b2_launch_scene_from_negmin_max(ctx, opts.d_scene_negmin_max, dScene));
  else if (opts.use_scene_box) B2_CUDA(cudaMemcpyAsync(dScene, &opts.scene_box, sizeof(b2bvh_aabb), cudaMemcpyHostToDevice, s));
  B2_CUDA(record(1));
  /* ---- S2 Morton (+ SetupClusters for PLOC/HPLOC is inside their launchers but is accounted under BUILD here) ---- */
  if (m60) B2_TRY(b2_launch_morton60(ctx, (const b2bvh_aabb*)dAabb, dScene, n, (u32*)dKeys /* upper 30 bits */, (u32*)dKeysLo, (u64*)dKeys64));
  else B2_TRY(b2_launch_morton(ctx, (const b2bvh_aabb*)dAabb, dScene, n, (u32*)dKeys, (u32*)dVals));
  B2_CUDA(record(2));
  /* ---- S3 sort (values of pass 0 are the iota written by S2: not re-read) ---- */
  if (m60)
    B2_TRY(b2_launch_sort60(ctx, (const u32*)dKeys, (const u32*)dKeysLo, n, (u32*)dM60K, (u32*)dM60V, (u32*)dSKeys, (u32*)dSVals, (u64*)dSKeys64, (u32*)dTKeys,
                            (u32*)dTVals, dSort));
  else
    B2_TRY(b2_launch_sort(ctx, (const u32*)dKeys, nullptr, (u32*)dSKeys, (u32*)dSVals, (u32*)dTKeys, (u32*)dTVals, dSort, n, 0, 32)); /* as the reference: bits 0..32; three 10-bit passes + a conditional one for bits 30-31 (radix_sort.cu) */
  B2_CUDA(record(3));
  /* ---- S4 / S6 / S7 hierarchy ---- */
  switch (algo) {
    case B2BVH_TWO_PASS_LBVH:
      if (m60)
        B2_TRY(b2_launch_lbvh_fused64(ctx, (const u64*)dSKeys64, (const u32*)dSVals, (const b2bvh_aabb*)dAabb, n, (b2bvh_bvh2_node*)dNodes, (u32*)dParents,
                                      (u32*)dLbvh, dRoot, 1));
      else if (opts.karras_two_kernel)
        B2_TRY(b2_launch_lbvh_karras_two_kernel(ctx, (const u32*)dSKeys, (const u32*)dSVals, (const b2bvh_aabb*)dAabb, n, (b2bvh_bvh2_node*)dNodes,
                                                (u32*)dParents, (u32*)dLbvh));
      else
        B2_TRY(b2_launch_lbvh_fused(ctx, (const u32*)dSKeys, (const u32*)dSVals, (const b2bvh_aabb*)dAabb, n, (b2bvh_bvh2_node*)dNodes,
                                    (u32*)dParents, (u32*)dLbvh, dRoot, 1));
      B2_CUDA(cudaMemsetAsync(dRoot, 0, 4, s));
      break;
    case B2BVH_SINGLE_PASS_LBVH:
      if (m60)
        B2_TRY(b2_launch_lbvh_fused64(ctx, (const u64*)dSKeys64, (const u32*)dSVals, (const b2bvh_aabb*)dAabb, n, (b2bvh_bvh2_node*)dNodes, nullptr, (u32*)dLbvh,
                                      dRoot, 0));
      else
      B2_TRY(b2_launch_lbvh_fused(ctx, (const u32*)dSKeys, (const u32*)dSVals, (const b2bvh_aabb*)dAabb, n, (b2bvh_bvh2_node*)dNodes, nullptr,
                                  (u32*)dLbvh, dRoot, 0));
      break;
    case B2BVH_PLOCPP:
      B2_TRY(b2_launch_ploc(ctx, (const b2bvh_aabb*)dAabb, (const u32*)dSVals, n, (b2bvh_bvh2_node*)dNodes, (b2bvh_prim_ref*)dLeaves, dMerge,
                            &iterations));
      B2_CUDA(cudaMemsetAsync(dRoot, 0, 4, s));
      break;
    case B2BVH_HPLOC:
      B2_TRY(b2_launch_hploc_keys(ctx, (const b2bvh_aabb*)dAabb, (const u32*)dSKeys, m60 ? (const u64*)dSKeys64 : nullptr, (const u32*)dSVals, n,
                                  (b2bvh_bvh2_node*)dNodes, (b2bvh_prim_ref*)dLeaves, dMerge, &iterations));
      B2_CUDA(cudaMemsetAsync(dRoot, 0, 4, s));
      break;
  }
  B2_CUDA(record(4));
  /* ---- S5 collapse ---- */
  if (opts.collapse)
    B2_TRY(b2_launch_collapse(ctx, (const b2bvh_bvh2_node*)dNodes, (const b2bvh_prim_ref*)dLeaves, split ? (const u32*)ctx->ref_leaf_prim : (const u32*)dSVals, dRoot, n, (b2bvh_bvh4_node*)dWide,
                              (b2bvh_prim_node*)dWLeaves, dCollapse, &nWide));
  B2_CUDA(record(5));
  B2_TRY(b2_fetch_words(ctx, dRoot, 1, B2_MB_ROOT));
  if (opts.d_root_box_out) B2_TRY(b2_launch_root_box(ctx, (const b2bvh_bvh2_node*)dNodes, dRoot, opts.d_root_box_out));
  return 0;
  };
  bool wantGraph = opts.use_graph && !split && !ctx->prof_on && !(opts.use_scene_box && !opts.d_scene_negmin_max);
  if (wantGraph && !opts.tris_on_device) { /* an upload can only be part of a graph when it reads pinned memory */
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, tris) != cudaSuccess || pa.type != cudaMemoryTypeHost) { cudaGetLastError(); wantGraph = false; }
  }
  b2bvh_ctx::GraphCache& G = ctx->graph;
  if (wantGraph && G.exec && G.algo == algo && G.n == n && G.tris == (const void*)tris && G.epoch == ctx->alloc_epoch &&
      memcmp(&G.opts, &opts, sizeof(opts)) == 0) {
    B2_CUDA(cudaGraphLaunch(G.exec, s));
    ctx->launches += G.launches;
  } else if (wantGraph) {
    if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
    const u32 before = ctx->launches;
    B2_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    capturing = true;
    const int st = enqueue();
    capturing = false;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (st) { if (graph) cudaGraphDestroy(graph); return st; }
    B2_CUDA(ce);
    const cudaError_t ie = cudaGraphInstantiate(&G.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { G.exec = nullptr; return b2_check(ie, "cudaGraphInstantiate"); }
    G.algo = algo; G.n = n; G.tris = tris; G.epoch = ctx->alloc_epoch; G.opts = opts; G.launches = ctx->launches - before;
    B2_CUDA(cudaGraphLaunch(G.exec, s));
  } else {
    B2_TRY(enqueue());
  }
  /* everything that does not need the device's answer */
  out->algo = (u32)algo;
  out->n_prims = n;
  out->n_internal = n - 1;
  out->leaves_separate = separate ? 1u : 0u;
  out->d_triangleBuff = dT;
  out->d_triangleAabb = (const b2bvh_aabb*)dAabb;
  out->d_sceneExtents = dScene;
  out->d_mortonCodeKeys = (const u32*)dKeys;
  out->d_mortonCodeValues = (const u32*)dVals;
  out->d_sortedMortonCodeKeys = (const u32*)dSKeys;
  out->d_sortedMortonCodeValues = (const u32*)dSVals;
  out->d_bvhNodes = (const b2bvh_bvh2_node*)dNodes;
  out->d_parentIdxs = (const u32*)dParents;
  out->d_leafNodes = (const b2bvh_prim_ref*)dLeaves;
  out->d_wideBvhNodes = (const b2bvh_bvh4_node*)dWide;
  out->d_wideLeafNodes = (const b2bvh_prim_node*)dWLeaves;
  out->n_launches = ctx->launches - launches0;
  out->n_triangles = nTris;
  out->d_primRefIdx = dRefPrim;
  out->morton_bits = m60 ? 60u : 30u;
  out->d_mortonCodeKeys64 = (const uint64_t*)dKeys64;
  out->d_sortedMortonCodeKeys64 = (const uint64_t*)dSKeys64;
  out->n_split_levels = splitLevels;
  ctx->pending.active = true; ctx->pending.algo = algo; ctx->pending.collapse = opts.collapse != 0; ctx->pending.host_tris = !opts.tris_on_device;
  ctx->pending.split = split;
  if (opts.defer_sync) return 0; /* b2bvh_build_finish synchronises and fills in the rest */
  return b2bvh_build_finish(ctx, out);
}

/* the build's single host synchronisation and everything read after it: root index, wide-node count, iteration counts, stage times */
int b2bvh_build_finish(b2bvh_ctx* ctx, b2bvh_tree* out) {
  if (!ctx || !out) return b2_fail(B2BVH_ERR_INVALID, "build_finish: null argument");
  if (!ctx->pending.active) return b2_fail(B2BVH_ERR_INVALID, "build_finish: no build is waiting on this context");
  ctx->pending.active = false;
  const int algo = ctx->pending.algo;
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_CUDA(cudaStreamSynchronize(ctx->stream)); /* the only host synchronisation of a build */
  out->root = b2_mailbox(ctx, B2_MB_ROOT)[0];
  out->n_wide = ctx->pending.collapse ? b2_mailbox(ctx, B2_MB_COLLAPSE)[1] : 0u;
  if (out->n_wide == B2BVH_INVALID) {
    out->n_wide = 0;
    return b2_fail(B2BVH_ERR_INTERNAL, "collapse: the Bvh2 nodes do not form a tree (more wide-node tasks than internal nodes)");
  }
  u32 iterations = 0;
  if (algo == B2BVH_PLOCPP) {
    iterations = b2_mailbox(ctx, B2_MB_PLOC)[2];
    if (b2_mailbox(ctx, B2_MB_PLOC)[1] != 1u)
      return b2_fail(B2BVH_ERR_INTERNAL, "ploc: %u clusters left after %u iterations", b2_mailbox(ctx, B2_MB_PLOC)[1], iterations);
  }
  float ms = 0;
  B2_CUDA(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[5]));
  out->build_ms = ms;
  B2_CUDA(cudaEventElapsedTime(&out->stage_ms[B2BVH_T_EXTENTS], ctx->ev[0], ctx->ev[1]));
  B2_CUDA(cudaEventElapsedTime(&out->stage_ms[B2BVH_T_MORTON], ctx->ev[1], ctx->ev[2]));
  B2_CUDA(cudaEventElapsedTime(&out->stage_ms[B2BVH_T_SORT], ctx->ev[2], ctx->ev[3]));
  B2_CUDA(cudaEventElapsedTime(&out->stage_ms[B2BVH_T_BUILD], ctx->ev[3], ctx->ev[4]));
  B2_CUDA(cudaEventElapsedTime(&out->stage_ms[B2BVH_T_COLLAPSE], ctx->ev[4], ctx->ev[5]));
  if (ctx->pending.host_tris) B2_CUDA(cudaEventElapsedTime(&out->h2d_ms, ctx->ev[8], ctx->ev[9]));
  out->n_iterations = algo == B2BVH_HPLOC ? b2_mailbox(ctx, B2_MB_HPLOC)[0] : iterations;
  if (ctx->pending.split) B2_CUDA(cudaEventElapsedTime(&out->split_ms, ctx->ev[10], ctx->ev[11]));
  return 0;
}

/* ------------------------------------------------------------------ batched builder (BatchedBvhBuilder::build, BatchedBuilder.cpp:16-77) */
int b2bvh_build_batched(b2bvh_ctx* ctx, const b2bvh_triangle* tris, uint32_t tris_on_device, const uint32_t* counts, uint32_t n_items, b2bvh_batch* out) {
  static_assert(SLOT_COUNT <= 48, "b2bvh_ctx::bufs has 48 slots");
  if (!ctx || !tris || !counts || !out) return b2_fail(B2BVH_ERR_INVALID, "build_batched: null argument");
  if (n_items == 0) return b2_fail(B2BVH_ERR_INVALID, "build_batched: no items");
  B2_CUDA(cudaSetDevice(ctx->device));
  memset(out, 0, sizeof(*out));
  std::vector<u32> off(2 * ((size_t)n_items + 1));
  u32* leafOff = off.data();
  u32* nodeOff = off.data() + n_items + 1;
  u64 total = 0;
  for (u32 i = 0; i < n_items; i++) {
    if (counts[i] == 0 || counts[i] > 32u)
      return b2_fail(B2BVH_ERR_INVALID, "build_batched: item %u has %u primitives; the batched builder takes 1..32 (MaxBatchedBlockSize, Common.h:597)", i, counts[i]);
    leafOff[i] = (u32)total; nodeOff[i] = (u32)(total - i);
    total += counts[i];
    if (total > 0x7FFFFFFFull) return b2_fail(B2BVH_ERR_INVALID, "build_batched: more than 2^31-1 primitives in one batch");
  }
  leafOff[n_items] = (u32)total; nodeOff[n_items] = (u32)(total - n_items);
  const u32 nPrims = (u32)total, nNodes = nPrims - n_items;
  cudaStream_t s = ctx->stream;
  void *dTris = nullptr, *dNodes, *dLeaves, *dRoots, *dScenes, *dOff;
  if (!tris_on_device) B2_TRY(b2_reserve(ctx, SLOT_TRIS, (size_t)nPrims * sizeof(b2bvh_triangle), &dTris));
  B2_TRY(b2_reserve(ctx, SLOT_BATCH_NODES, (size_t)nNodes * sizeof(b2bvh_bvh2_node), &dNodes));
  B2_TRY(b2_reserve(ctx, SLOT_BATCH_LEAVES, (size_t)nPrims * sizeof(b2bvh_prim_ref), &dLeaves));
  B2_TRY(b2_reserve(ctx, SLOT_BATCH_ROOTS, (size_t)n_items * 4, &dRoots));
  B2_TRY(b2_reserve(ctx, SLOT_BATCH_SCENES, (size_t)n_items * sizeof(b2bvh_aabb), &dScenes));
  B2_TRY(b2_reserve(ctx, SLOT_BATCH_OFFSETS, off.size() * 4, &dOff));
  const b2bvh_triangle* dT = tris;
  B2_CUDA(cudaEventRecord(ctx->ev[14], s));
  if (!tris_on_device) {
    B2_CUDA(cudaMemcpyAsync(dTris, tris, (size_t)nPrims * sizeof(b2bvh_triangle), cudaMemcpyHostToDevice, s));
    dT = (const b2bvh_triangle*)dTris;
  }
  B2_CUDA(cudaMemcpyAsync(dOff, off.data(), off.size() * 4, cudaMemcpyHostToDevice, s));
  B2_CUDA(cudaEventRecord(ctx->ev[15], s));
  B2_CUDA(cudaEventRecord(ctx->ev[12], s));
  B2_TRY(b2_launch_batched(ctx, dT, (const u32*)dOff, (const u32*)dOff + n_items + 1, n_items, (b2bvh_bvh2_node*)dNodes, (b2bvh_prim_ref*)dLeaves, (u32*)dRoots,
                           (b2bvh_aabb*)dScenes));
  B2_CUDA(cudaEventRecord(ctx->ev[13], s));
  B2_CUDA(cudaStreamSynchronize(s)); /* `off` is pageable host memory: the copy must have left it before it goes out of scope */
  out->n_items = n_items;
  out->n_prims_total = nPrims;
  out->n_nodes_total = nNodes;
  out->d_triangles = dT;
  out->d_bvhNodes = (const b2bvh_bvh2_node*)dNodes;
  out->d_primRefs = (const b2bvh_prim_ref*)dLeaves;
  out->d_rootNodes = (const u32*)dRoots;
  out->d_sceneExtents = (const b2bvh_aabb*)dScenes;
  out->d_leafOffsets = (const u32*)dOff;
  out->d_nodeOffsets = (const u32*)dOff + n_items + 1;
  B2_CUDA(cudaEventElapsedTime(&out->build_ms, ctx->ev[12], ctx->ev[13]));
  B2_CUDA(cudaEventElapsedTime(&out->h2d_ms, ctx->ev[14], ctx->ev[15]));
  return 0;
}

/* ------------------------------------------------------------------ sharded helpers */
int b2bvh_shard_extents(b2bvh_ctx* ctx, const b2bvh_triangle* tris, uint32_t n, uint32_t tris_on_device, float* d_negmin_max6) {
  if (!ctx || !tris || !d_negmin_max6 || n == 0) return b2_fail(B2BVH_ERR_INVALID, "shard_extents: bad argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  void *dTris = nullptr, *dAabb;
  const b2bvh_triangle* dT = tris;
  if (!tris_on_device) {
    B2_TRY(b2_reserve(ctx, SLOT_TRIS, (size_t)n * sizeof(b2bvh_triangle), &dTris));
    B2_CUDA(cudaMemcpyAsync(dTris, tris, (size_t)n * sizeof(b2bvh_triangle), cudaMemcpyHostToDevice, ctx->stream));
    dT = (const b2bvh_triangle*)dTris;
  }
  B2_TRY(b2_reserve(ctx, SLOT_AABB, (size_t)n * sizeof(b2bvh_aabb), &dAabb));
  unsigned char* ctl = (unsigned char*)ctx->bufs[SLOT_CTL].p;
  return b2_launch_extents(ctx, dT, n, (b2bvh_aabb*)dAabb, (b2bvh_aabb*)ctl, (u32*)(ctl + 32), d_negmin_max6);
}

/* ------------------------------------------------------------------ host-side SAH cost reporting */
static inline float h_area(const b2bvh_aabb& b) {
  const float ex = b.m_max.x - b.m_min.x, ey = b.m_max.y - b.m_min.y, ez = b.m_max.z - b.m_min.z;
  const float xy = ex * ey, xz = ex * ez, yz = ey * ez;
  float s = xy + xz;
  s = s + yz;
  return 2 * s;
}
static inline float h_min(float a, float b) { return (b < a) ? b : a; }
static inline float h_max(float a, float b) { return (b > a) ? b : a; }

/* The whole primitive-range sharded build from ONE host thread over G contexts (SURVEY.md §8(b) proposal; the reference is single-device,
 * Context.cpp:11): local boxes on every device -> the 6-float {-min,max} vectors meet on the host (24 bytes per device: the exchange is
 * latency, not bandwidth; ranks in separate processes use NCCL for it, b2bvh/sharded.py) -> every device builds its shard in the global
 * frame, all enqueued before the first is waited for -> root boxes -> top-level tree on the first context.  Contexts may share a device. */
int b2bvh_build_sharded(b2bvh_ctx* const* ctxs, uint32_t n_gpus, int algo, const b2bvh_triangle* const* tris, const uint32_t* counts,
                        const b2bvh_build_opts* optsIn, b2bvh_tree* trees, b2bvh_aabb* h_scene, b2bvh_bvh2_node* h_topNodes) {
  if (!ctxs || !tris || !counts || !trees || !h_topNodes || n_gpus == 0 || n_gpus > 256u) return b2_fail(B2BVH_ERR_INVALID, "build_sharded: bad argument (1 <= n_gpus <= 256)");
  for (u32 g = 0; g < n_gpus; g++)
    if (!ctxs[g] || !tris[g] || counts[g] < 2) return b2_fail(B2BVH_ERR_INVALID, "build_sharded: shard %u needs a context and at least 2 primitives", g);
  b2bvh_build_opts base;
  memset(&base, 0, sizeof(base));
  if (optsIn) base = *optsIn; else base.collapse = 1;
  if (base.split_sa_max > 0.0f || base.use_scene_box || base.d_scene_negmin_max || base.boxes_ready || base.d_root_box_out)
    return b2_fail(B2BVH_ERR_INVALID, "build_sharded: the scene-box, boxes_ready, root-box and split options are set by the sharded build itself");
  std::vector<float> local(6 * (size_t)n_gpus);
  auto ctl = [&](u32 g, size_t off) { return (float*)((unsigned char*)ctxs[g]->bufs[SLOT_CTL].p + off); };
  for (u32 g = 0; g < n_gpus; g++) B2_TRY(b2bvh_shard_extents(ctxs[g], tris[g], counts[g], base.tris_on_device, ctl(g, 64)));
  for (u32 g = 0; g < n_gpus; g++) {
    B2_CUDA(cudaSetDevice(ctxs[g]->device));
    B2_CUDA(cudaMemcpyAsync(&local[6 * g], ctl(g, 64), 24, cudaMemcpyDeviceToHost, ctxs[g]->stream));
  }
  float global6[6] = {-B2BVH_FLT_MAX, -B2BVH_FLT_MAX, -B2BVH_FLT_MAX, -B2BVH_FLT_MAX, -B2BVH_FLT_MAX, -B2BVH_FLT_MAX};
  for (u32 g = 0; g < n_gpus; g++) {
    B2_CUDA(cudaSetDevice(ctxs[g]->device));
    B2_CUDA(cudaStreamSynchronize(ctxs[g]->stream));
    for (int k = 0; k < 6; k++) global6[k] = h_max(global6[k], local[6 * g + k]); /* the all-reduce(MAX) of {-min, max} */
  }
  if (h_scene) { h_scene->m_min = {-global6[0], -global6[1], -global6[2]}; h_scene->m_max = {global6[3], global6[4], global6[5]}; }
  for (u32 g = 0; g < n_gpus; g++) {
    B2_CUDA(cudaSetDevice(ctxs[g]->device));
    B2_CUDA(cudaMemcpyAsync(ctl(g, 64), global6, 24, cudaMemcpyHostToDevice, ctxs[g]->stream));
    b2bvh_build_opts o = base;
    o.boxes_ready = 1;
    o.d_scene_negmin_max = ctl(g, 64);
    o.d_root_box_out = ctl(g, 224);
    o.defer_sync = 1; /* every shard is enqueued before the first one is waited for */
    B2_TRY(b2bvh_build(ctxs[g], algo, tris[g], counts[g], &o, &trees[g]));
  }
  std::vector<b2bvh_aabb> roots(n_gpus);
  for (u32 g = 0; g < n_gpus; g++) {
    B2_TRY(b2bvh_build_finish(ctxs[g], &trees[g]));
    B2_CUDA(cudaSetDevice(ctxs[g]->device));
    B2_CUDA(cudaMemcpyAsync(&roots[g], ctl(g, 224), sizeof(b2bvh_aabb), cudaMemcpyDeviceToHost, ctxs[g]->stream));
    B2_CUDA(cudaStreamSynchronize(ctxs[g]->stream));
  }
  b2bvh_ctx* c0 = ctxs[0];
  B2_CUDA(cudaSetDevice(c0->device));
  void* dTop;
  const size_t rootBytes = ((size_t)n_gpus * sizeof(b2bvh_aabb) + 255) & ~(size_t)255;
  B2_TRY(b2_reserve(c0, SLOT_MISC, rootBytes + (2 * (size_t)n_gpus - 1) * sizeof(b2bvh_bvh2_node), &dTop));
  B2_CUDA(cudaMemcpyAsync(dTop, roots.data(), n_gpus * sizeof(b2bvh_aabb), cudaMemcpyHostToDevice, c0->stream));
  b2bvh_bvh2_node* dNodes = (b2bvh_bvh2_node*)((unsigned char*)dTop + rootBytes);
  B2_TRY(b2bvh_top_level(c0, (const b2bvh_aabb*)dTop, n_gpus, dNodes));
  B2_CUDA(cudaMemcpyAsync(h_topNodes, dNodes, (2 * (size_t)n_gpus - 1) * sizeof(b2bvh_bvh2_node), cudaMemcpyDeviceToHost, c0->stream));
  B2_CUDA(cudaStreamSynchronize(c0->stream));
  return 0;
}


/* Utility::calculatebvh4Cost (Utility.cpp:351-396): root box = union of the root's child boxes, then a float
 * running sum in index order: 1 + sum over wide nodes of internal-child areas / rootArea + sum over leaf slots of
 * primitive areas / rootArea. */
float b2bvh_cost_bvh4(const b2bvh_bvh4_node* wide, const b2bvh_prim_node* wideLeaves, const b2bvh_aabb* primAabbs, uint32_t root, uint32_t n_wide,
                      uint32_t n_internal) {
  b2bvh_aabb rb = {{B2BVH_FLT_MAX, B2BVH_FLT_MAX, B2BVH_FLT_MAX}, {-B2BVH_FLT_MAX, -B2BVH_FLT_MAX, -B2BVH_FLT_MAX}};
  for (int k = 0; k < 4; k++) {
    if (wide[root].m_child[k] == B2BVH_INVALID) continue;
    const b2bvh_aabb& c = wide[root].m_aabb[k];
    rb.m_min.x = h_min(rb.m_min.x, c.m_min.x); rb.m_min.y = h_min(rb.m_min.y, c.m_min.y); rb.m_min.z = h_min(rb.m_min.z, c.m_min.z);
    rb.m_max.x = h_max(rb.m_max.x, c.m_max.x); rb.m_max.y = h_max(rb.m_max.y, c.m_max.y); rb.m_max.z = h_max(rb.m_max.z, c.m_max.z);
  }
  const float inv = 1.0f / h_area(rb);
  float cost = 1.0f;
  for (uint32_t i = 0; i < n_wide; i++)
    for (int k = 0; k < 4; k++) {
      const uint32_t c = wide[i].m_child[k];
      if (c != B2BVH_INVALID && c < n_internal) cost += h_area(wide[i].m_aabb[k]) * inv;
    }
  for (uint32_t i = 0; i < n_internal + 1; i++) cost += h_area(primAabbs[wideLeaves[i].m_primIdx]) * inv;
  return cost;
}

/* Utility::calculateLbvhCost (Utility.cpp:317-349). */
float b2bvh_cost_lbvh(const b2bvh_bvh2_node* nodes, uint32_t root, uint32_t n_leaf, uint32_t n_internal) {
  const float inv = 1.0f / h_area(nodes[root].m_aabb);
  float cost = 1.0f;
  for (uint32_t i = 0; i < n_internal; i++) {
    if (nodes[i].m_leftChildIdx != B2BVH_INVALID) cost += h_area(nodes[nodes[i].m_leftChildIdx].m_aabb) * inv;
    if (nodes[i].m_rightChildIdx != B2BVH_INVALID) cost += h_area(nodes[nodes[i].m_rightChildIdx].m_aabb) * inv;
  }
  for (uint32_t i = n_internal; i < n_leaf + n_internal; i++)
    if (nodes[i].m_leftChildIdx != B2BVH_INVALID) cost += h_area(nodes[i].m_aabb) * inv;
  return cost;
}

int b2bvh_tree_cost(b2bvh_ctx* ctx, const b2bvh_tree* tree, float* cost) {
  if (!ctx || !tree || !cost) return b2_fail(B2BVH_ERR_INVALID, "tree_cost: bad argument");
  if (tree->n_wide == 0) return b2_fail(B2BVH_ERR_INVALID, "tree_cost: tree has no wide nodes (collapse was off)");
  std::vector<b2bvh_bvh4_node> w(tree->n_wide);
  std::vector<b2bvh_prim_node> wl(tree->n_prims);
  std::vector<b2bvh_aabb> pb(tree->n_prims);
  B2_TRY(b2bvh_d2h(ctx, w.data(), tree->d_wideBvhNodes, w.size() * sizeof(b2bvh_bvh4_node)));
  B2_TRY(b2bvh_d2h(ctx, wl.data(), tree->d_wideLeafNodes, wl.size() * sizeof(b2bvh_prim_node)));
  B2_TRY(b2bvh_d2h(ctx, pb.data(), tree->d_triangleAabb, pb.size() * sizeof(b2bvh_aabb)));
  if (tree->d_primRefIdx) {
    /* split references: the reference fills triangleAabb[ref.m_primIdx] = ref.m_aabb in reference order, so the LAST fragment of a
     * triangle is the one the cost counts, and the table has one (otherwise empty) entry per reference (TwoPassLbvh.cpp:188-193) */
    std::vector<u32> rp(tree->n_prims);
    B2_TRY(b2bvh_d2h(ctx, rp.data(), tree->d_primRefIdx, rp.size() * 4));
    const b2bvh_aabb empty = {{B2BVH_FLT_MAX, B2BVH_FLT_MAX, B2BVH_FLT_MAX}, {-B2BVH_FLT_MAX, -B2BVH_FLT_MAX, -B2BVH_FLT_MAX}};
    std::vector<b2bvh_aabb> table(tree->n_prims, empty);
    for (size_t i = 0; i < rp.size(); i++) table[rp[i]] = pb[i];
    pb.swap(table);
  }
  *cost = b2bvh_cost_bvh4(w.data(), wl.data(), pb.data(), 0, tree->n_wide, tree->n_internal);
  return 0;
}

} /* extern "C" */
