/*
 * BvhConstruction.h — the reference's C++ host API on top of the C ABI of libb2bvh.so.
 *
 * Same namespace, class names, method signatures and public members as the reference's builder objects:
 *   BvhConstruction::Context          src/Context.h:8-22
 *   BvhConstruction::Timer            src/Timer.h:11-97       (getTimeRecord(TimerCodes))
 *   BvhConstruction::TwoPassLbvh      src/TwoPassLbvh.h:12-31
 *   BvhConstruction::SinglePassLbvh   src/SinglePassLbvh.h:12-31
 *   BvhConstruction::PLOCNew          src/PLOC++Bvh.h:12-32
 *   BvhConstruction::HPLOC            src/Hploc.h:12-32
 * so that a caller written against the reference (src/main.cpp:52-77: `TwoPassLbvh bvh; bvh.build(ctx, tris);
 * bvh.traverseBvh(ctx);`) compiles unchanged.  Where the reference binds HIP/CUDA at run time through Orochi
 * (oroInitialize -> dlopen), this header binds libb2bvh.so with dlopen and resolves exactly the symbols of include/b2bvh.h.
 * GpuMemory<T> members are views of the buffers the context owns (ptr(), size(), getData()).
 *
 * Differences (all additive): the Bvh4 result survives build() (d_wideBvhNodes, d_wideLeafNodes, m_wideNodeCount — in the
 * reference they are locals of build(), TwoPassLbvh.cpp:154-155); traverseBvh() keeps the hit buffer (m_hits) and does not
 * write test.png; errors throw std::runtime_error with b2bvh_last_error() (the reference prints and continues, Error.cpp:9-24).
 */
#pragma once
#include <dlfcn.h>
#include <math.h>

#include <iostream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b2bvh.h"

namespace BvhConstruction {

using u32 = uint32_t;
using u8 = uint8_t;
using u64 = uint64_t;
using float3 = b2bvh_float3;
using float4 = b2bvh_float4;
using Triangle = b2bvh_triangle;
using Aabb = b2bvh_aabb;
using Bvh2Node = b2bvh_bvh2_node;
using Bvh4Node = b2bvh_bvh4_node;
using PrimRef = b2bvh_prim_ref;
using PrimNode = b2bvh_prim_node;
using Ray = b2bvh_ray;
using HitInfo = b2bvh_hit;
using Transformation = b2bvh_transform;
using Camera = b2bvh_camera;
constexpr float Pi = 3.14159265358979323846f;
constexpr u32 INVALID_NODE_IDX = 0xFFFFFFFF;

enum TimerCodes { CalculateCentroidExtentsTime, CalculateMortonCodesTime, SortingTime, BvhBuildTime, TraversalTime, CollapseBvhTime, RayGenTime };

/* qtRotation / qtGetIdentity (Common.h:461-481), host side */
inline float4 qtGetIdentity() { return float4{0.0f, 0.0f, 0.0f, 1.0f}; }
inline float4 qtRotation(float4 axisAngle) {
  const float len = sqrtf(axisAngle.x * axisAngle.x + axisAngle.y * axisAngle.y + axisAngle.z * axisAngle.z);
  const float ax = axisAngle.x / len, ay = axisAngle.y / len, az = axisAngle.z / len, ang = axisAngle.w;
  return float4{ax * sinf(ang / 2.0f), ay * sinf(ang / 2.0f), az * sinf(ang / 2.0f), cosf(ang / 2.0f)};
}

/* ---- run-time binding of libb2bvh.so (the role Orochi's oroInitialize plays in the reference) ---- */
struct Api {
#define B2_SYM(name) decltype(&::name) name = nullptr;
  B2_SYM(b2bvh_ctx_create) B2_SYM(b2bvh_ctx_destroy) B2_SYM(b2bvh_device_name) B2_SYM(b2bvh_device_sm_count) B2_SYM(b2bvh_alloc)
  B2_SYM(b2bvh_free) B2_SYM(b2bvh_memset) B2_SYM(b2bvh_h2d) B2_SYM(b2bvh_d2h) B2_SYM(b2bvh_sync) B2_SYM(b2bvh_last_error)
  B2_SYM(b2bvh_build) B2_SYM(b2bvh_build_batched) B2_SYM(b2bvh_generate_rays) B2_SYM(b2bvh_traverse) B2_SYM(b2bvh_traverse_ex) B2_SYM(b2bvh_heat_map) B2_SYM(b2bvh_tree_cost) B2_SYM(b2bvh_cost_bvh4)
  B2_SYM(b2bvh_cost_lbvh) B2_SYM(b2bvh_synth_uniform) B2_SYM(b2bvh_abi_version) B2_SYM(b2bvh_build_sharded)
#undef B2_SYM
  void* handle = nullptr;
  static Api& get(const char* path = nullptr) {
    static Api api;
    if (api.handle) return api;
    const char* env = getenv("B2BVH_LIB");
    const std::string lib = path ? path : (env ? env : "libb2bvh.so");
    api.handle = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!api.handle) throw std::runtime_error(std::string("cannot load ") + lib + ": " + dlerror() + " (there is no CPU fallback)");
#define B2_SYM(name)                                                                     \
  api.name = reinterpret_cast<decltype(api.name)>(dlsym(api.handle, #name));             \
  if (!api.name) throw std::runtime_error(std::string("libb2bvh.so lacks symbol ") + #name);
    B2_SYM(b2bvh_ctx_create) B2_SYM(b2bvh_ctx_destroy) B2_SYM(b2bvh_device_name) B2_SYM(b2bvh_device_sm_count) B2_SYM(b2bvh_alloc)
    B2_SYM(b2bvh_free) B2_SYM(b2bvh_memset) B2_SYM(b2bvh_h2d) B2_SYM(b2bvh_d2h) B2_SYM(b2bvh_sync) B2_SYM(b2bvh_last_error)
    B2_SYM(b2bvh_build) B2_SYM(b2bvh_build_batched) B2_SYM(b2bvh_generate_rays) B2_SYM(b2bvh_traverse) B2_SYM(b2bvh_traverse_ex) B2_SYM(b2bvh_heat_map) B2_SYM(b2bvh_tree_cost) B2_SYM(b2bvh_cost_bvh4)
    B2_SYM(b2bvh_cost_lbvh) B2_SYM(b2bvh_synth_uniform) B2_SYM(b2bvh_abi_version) B2_SYM(b2bvh_build_sharded)
#undef B2_SYM
    if (api.b2bvh_abi_version() != B2BVH_ABI_VERSION)
      throw std::runtime_error("libb2bvh.so speaks ABI version " + std::to_string(api.b2bvh_abi_version()) + ", this header was written for " +
                               std::to_string(B2BVH_ABI_VERSION) + " (struct layouts differ): rebuild one of them");
    return api;
  }
};
inline void checkStatus(int status, const char* what) {
  if (status != 0) throw std::runtime_error(std::string(what) + ": " + Api::get().b2bvh_last_error());
}

class Context {
 public:
  explicit Context(int device = 0) { /* device 0 by default, as Context.cpp:11; the sharded build makes one per device */
    checkStatus(Api::get().b2bvh_ctx_create(device, nullptr, &m_ctx), "b2bvh_ctx_create");
    char name[256];
    Api::get().b2bvh_device_name(m_ctx, name, sizeof(name));
    std::cout << "Executing on '" << name << "'" << std::endl; /* Context.cpp:14 */
  }
  ~Context() { Api::get().b2bvh_ctx_destroy(m_ctx); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  u32 getMaxGridSize() const { int sm = 0; Api::get().b2bvh_device_sm_count(m_ctx, &sm); return (u32)sm; }
  b2bvh_ctx* m_ctx = nullptr;
};

/* Timer: stage times come from CUDA events inside b2bvh_build (no host sync per launch, unlike Timer.h:48-56) */
class Timer {
 public:
  using TokenType = int;
  using TimeUnit = float;
  bool EnableTimer = true;
  TimeUnit getTimeRecord(const TokenType token) const noexcept {
    auto it = timeRecord.find(token);
    return it == timeRecord.end() ? TimeUnit{} : it->second;
  }
  void reset(const TokenType token) noexcept { timeRecord.erase(token); }
  void clear() noexcept { timeRecord.clear(); }
  std::unordered_map<TokenType, TimeUnit> timeRecord;
};

/* View of a device buffer owned by the context; the subset of Oro::GpuMemory<T> the reference's callers use. */
template <typename T>
class GpuMemory {
 public:
  T* ptr() const noexcept { return m_ptr; }
  size_t size() const noexcept { return m_size; }
  std::vector<T> getData() const {
    std::vector<T> out(m_size);
    if (m_size) checkStatus(Api::get().b2bvh_d2h(m_ctx, out.data(), m_ptr, m_size * sizeof(T)), "b2bvh_d2h");
    return out;
  }
  void bind(b2bvh_ctx* ctx, const T* p, size_t n) { m_ctx = ctx; m_ptr = const_cast<T*>(p); m_size = p ? n : 0; }

 private:
  b2bvh_ctx* m_ctx = nullptr;
  T* m_ptr = nullptr;
  size_t m_size = 0;
};

namespace detail {
struct BuilderBase {
  GpuMemory<Triangle> d_triangleBuff;
  GpuMemory<Aabb> d_triangleAabb;
  GpuMemory<Aabb> d_sceneExtents;
  GpuMemory<u32> d_mortonCodeKeys;
  GpuMemory<u32> d_mortonCodeValues;
  GpuMemory<u32> d_sortedMortonCodeKeys;
  GpuMemory<u32> d_sortedMortonCodeValues;
  GpuMemory<Bvh2Node> d_bvhNodes;
  GpuMemory<Bvh4Node> d_wideBvhNodes;
  GpuMemory<PrimNode> d_wideLeafNodes;
  u32 m_rootNodeIdx = 0;
  Timer m_timer;
  u32 m_nInternalNodes = 0;
  u32 m_wideNodeCount = 0;
  float m_cost = 0.0f;
  b2bvh_tree m_tree{};
  std::vector<HitInfo> m_hits;
  std::vector<u32> m_rayCounter; /* triangle tests per ray (if-if, restart trail, Bvh4); Utility::generateTraversalHeatMap input */
  Transformation m_transform{};
  Camera m_camera{};
  u32 m_width = 512, m_height = 512; /* TwoPassLbvh.cpp:221-222 */
  bool m_useGraph = false;            /* replay repeated builds of the same size from a CUDA graph (b2bvh_build_opts.use_graph) */
  float m_saMax = 0.0f;               /* TwoPassLbvh only: > 0 builds over early-split references, the reference compiled with USE_PRIM_SPLITTING
                                         (`float saMax = 10.0f`, TwoPassLbvh.cpp:23-28); 0 = the reference's default (no splitting) */
  GpuMemory<u32> d_primRefIdx;        /* PrimRef::m_primIdx of every reference when m_saMax > 0 (d_triangleAabb then holds PrimRef::m_aabb) */
  int m_traversalKernel = B2BVH_TRAVERSE_SPECULATIVE_WHILE; /* WHILEWHILE is defined at TwoPassLbvh.cpp:12 */

  BuilderBase() {
    /* cornell-box preset, the one compiled into the reference (TwoPassLbvh.cpp:202-216) */
    m_transform.m_translation = float3{0.0f, 0.0f, -5.0f};
    m_transform.m_scale = float3{1.0f, 1.0f, 1.0f};
    m_transform.m_quat = qtGetIdentity();
    m_camera.m_eye = float4{0.0f, 2.5f, 5.8f, 0.0f};
    m_camera.m_quat = qtRotation(float4{0.0f, 0.0f, 1.0f, -1.57f});
    m_camera.m_fov = 45.0f * Pi / 180.f;
    m_camera.m_near = 0.0f;
    m_camera.m_far = 100000.0f;
  }

  void buildWith(int algo, Context& context, std::vector<Triangle>& primitives) {
    Api& api = Api::get();
    b2bvh_build_opts opts{};
    opts.collapse = 1;
    opts.stage_timing = 1;
    opts.split_sa_max = algo == B2BVH_TWO_PASS_LBVH ? m_saMax : 0.0f;
    opts.use_graph = m_useGraph ? 1u : 0u; /* takes effect for device or pinned triangles; a std::vector upload is enqueued plainly */
    checkStatus(api.b2bvh_build(context.m_ctx, algo, primitives.data(), (u32)primitives.size(), &opts, &m_tree), "b2bvh_build");
    const b2bvh_tree& t = m_tree;
    b2bvh_ctx* c = context.m_ctx;
    const size_t n = t.n_prims;
    d_triangleBuff.bind(c, t.d_triangleBuff, t.n_triangles);
    d_primRefIdx.bind(c, t.d_primRefIdx, n);
    d_triangleAabb.bind(c, t.d_triangleAabb, n);
    d_sceneExtents.bind(c, t.d_sceneExtents, 1);
    d_mortonCodeKeys.bind(c, t.d_mortonCodeKeys, n);
    d_mortonCodeValues.bind(c, t.d_mortonCodeValues, n);
    d_sortedMortonCodeKeys.bind(c, t.d_sortedMortonCodeKeys, n);
    d_sortedMortonCodeValues.bind(c, t.d_sortedMortonCodeValues, n);
    d_bvhNodes.bind(c, t.d_bvhNodes, t.leaves_separate ? n - 1 : 2 * n - 1);
    d_wideBvhNodes.bind(c, t.d_wideBvhNodes, t.n_wide);
    d_wideLeafNodes.bind(c, t.d_wideLeafNodes, n);
    m_rootNodeIdx = t.root;
    m_nInternalNodes = t.n_internal;
    m_wideNodeCount = t.n_wide;
    m_timer.timeRecord[CalculateCentroidExtentsTime] += t.stage_ms[B2BVH_T_EXTENTS];
    m_timer.timeRecord[CalculateMortonCodesTime] += t.stage_ms[B2BVH_T_MORTON];
    m_timer.timeRecord[SortingTime] += t.stage_ms[B2BVH_T_SORT];
    m_timer.timeRecord[BvhBuildTime] += t.stage_ms[B2BVH_T_BUILD];
    m_timer.timeRecord[CollapseBvhTime] += t.stage_ms[B2BVH_T_COLLAPSE];
    checkStatus(api.b2bvh_tree_cost(c, &t, &m_cost), "b2bvh_tree_cost"); /* TwoPassLbvh.cpp:185-196 */
  }

  void printPerf() const { /* wording of TwoPassLbvh.cpp:300-310 */
    std::cout << "==========================Perf Times==========================" << std::endl;
    std::cout << "CalculateCentroidExtentsTime :" << m_timer.getTimeRecord(CalculateCentroidExtentsTime) << "ms" << std::endl;
    std::cout << "CalculateMortonCodesTime :" << m_timer.getTimeRecord(CalculateMortonCodesTime) << "ms" << std::endl;
    std::cout << "SortingTime : " << m_timer.getTimeRecord(SortingTime) << "ms" << std::endl;
    std::cout << "BvhBuildTime : " << m_timer.getTimeRecord(BvhBuildTime) << "ms" << std::endl;
    std::cout << "TraversalTime : " << m_timer.getTimeRecord(TraversalTime) << "ms" << std::endl;
    std::cout << "CollapseTime : " << m_timer.getTimeRecord(CollapseBvhTime) << "ms" << std::endl;
    std::cout << "Bvh Cost : " << m_cost << std::endl;
    std::cout << "Total Time : "
              << m_timer.getTimeRecord(CalculateCentroidExtentsTime) + m_timer.getTimeRecord(CalculateMortonCodesTime) +
                     m_timer.getTimeRecord(SortingTime) + m_timer.getTimeRecord(BvhBuildTime)
              << "ms" << std::endl;
    std::cout << "==============================================================" << std::endl;
  }

  /* GenerateRays + traversal kernel + perf table (TwoPassLbvh.cpp:199-311) */
  void traceAndReport(Context& context, bool trace) {
    if (trace) {
      Api& api = Api::get();
      void* dRays = nullptr;
      void* dHits = nullptr;
      const u32 nRays = m_width * m_height;
      checkStatus(api.b2bvh_alloc(context.m_ctx, (size_t)nRays * sizeof(Ray), &dRays), "b2bvh_alloc");
      checkStatus(api.b2bvh_alloc(context.m_ctx, (size_t)nRays * sizeof(HitInfo), &dHits), "b2bvh_alloc");
      float ms = 0;
      checkStatus(api.b2bvh_generate_rays(context.m_ctx, &m_camera, m_width, m_height, (Ray*)dRays, &ms), "b2bvh_generate_rays");
      m_timer.timeRecord[RayGenTime] += ms;
      /* the if-if / restart-trail kernels of the reference (and the Bvh4 kernel) also count the triangle tests per ray:
       * d_rayCounterBuffer, TwoPassLbvh.cpp:224,267,272 */
      const bool counted = m_traversalKernel >= B2BVH_TRAVERSE_IFIF;
      void* dCounter = nullptr;
      if (counted) checkStatus(api.b2bvh_alloc(context.m_ctx, (size_t)nRays * sizeof(u32), &dCounter), "b2bvh_alloc");
      checkStatus(api.b2bvh_traverse_ex(context.m_ctx, &m_tree, (const Ray*)dRays, nRays, &m_transform, m_traversalKernel, (HitInfo*)dHits, nullptr,
                                        (u32*)dCounter, &ms),
                  "b2bvh_traverse_ex");
      m_timer.timeRecord[TraversalTime] += ms;
      m_hits.resize(nRays);
      checkStatus(api.b2bvh_d2h(context.m_ctx, m_hits.data(), dHits, (size_t)nRays * sizeof(HitInfo)), "b2bvh_d2h");
      m_rayCounter.clear();
      if (counted) {
        m_rayCounter.resize(nRays);
        checkStatus(api.b2bvh_d2h(context.m_ctx, m_rayCounter.data(), dCounter, (size_t)nRays * sizeof(u32)), "b2bvh_d2h");
        api.b2bvh_free(context.m_ctx, dCounter);
      }
      api.b2bvh_free(context.m_ctx, dRays);
      api.b2bvh_free(context.m_ctx, dHits);
    }
    printPerf();
  }
};
}  // namespace detail

class TwoPassLbvh : public detail::BuilderBase {
 public:
  void build(Context& context, std::vector<Triangle>& primitives) { buildWith(B2BVH_TWO_PASS_LBVH, context, primitives); d_flags.bind(context.m_ctx, m_tree.d_parentIdxs, 2 * (size_t)m_tree.n_prims - 1); }
  void traverseBvh(Context& context) { traceAndReport(context, true); }
  GpuMemory<u32> d_flags; /* the reference exposes its refit flags here; this build exposes the parent indices in their place */
};

class SinglePassLbvh : public detail::BuilderBase {
 public:
  SinglePassLbvh() { m_traversalKernel = B2BVH_TRAVERSE_WHILE; }
  void build(Context& context, std::vector<Triangle>& primitives) { buildWith(B2BVH_SINGLE_PASS_LBVH, context, primitives); }
  void traverseBvh(Context& context) { traceAndReport(context, true); }
};

/* PLOC++ / H-PLOC: traverseBvh() of the reference only prints the timings (PLOC++Bvh.cpp:198-211, Hploc.cpp:167-180);
 * traceBvh() additionally traces the separate-leaf tree. */
class PLOCNew : public detail::BuilderBase {
 public:
  void build(Context& context, std::vector<Triangle>& primitives) { buildWith(B2BVH_PLOCPP, context, primitives); d_leafNodes.bind(context.m_ctx, m_tree.d_leafNodes, m_tree.n_prims); }
  void traverseBvh(Context& context) { traceAndReport(context, false); }
  void traceBvh(Context& context) { traceAndReport(context, true); }
  GpuMemory<PrimRef> d_leafNodes;
};

class HPLOC : public detail::BuilderBase {
 public:
  void build(Context& context, std::vector<Triangle>& primitives) { buildWith(B2BVH_HPLOC, context, primitives); d_leafNodes.bind(context.m_ctx, m_tree.d_leafNodes, m_tree.n_prims); }
  void traverseBvh(Context& context) { traceAndReport(context, false); }
  void traceBvh(Context& context) { traceAndReport(context, true); }
  GpuMemory<PrimRef> d_leafNodes;
};

/* BatchedBvhBuilder (src/BatchedBuilder.h:12-30): many small BVHs (<= MaxBatchedBlockSize = 32 triangles each) in one launch.
 * Same members as the reference; item i's internal nodes start at d_bvhNodes[m_nodeOffsets[i]], its leaves (PrimRefs in sorted order)
 * at d_primRefs[m_leafOffsets[i]], child indices are local to the item (leaf g = (n-1) + g), d_rootNodes[i] is its local root. */
struct BatchedBuildInput {
  std::vector<Triangle> m_primitives;
};

class BatchedBvhBuilder {
 public:
  void build(Context& context, std::vector<BatchedBuildInput>& batch) {
    std::vector<u32> counts(batch.size());
    size_t total = 0;
    for (size_t i = 0; i < batch.size(); i++) { counts[i] = (u32)batch[i].m_primitives.size(); total += counts[i]; }
    std::vector<Triangle> flat;
    flat.reserve(total);
    for (const BatchedBuildInput& b : batch) flat.insert(flat.end(), b.m_primitives.begin(), b.m_primitives.end());
    checkStatus(Api::get().b2bvh_build_batched(context.m_ctx, flat.data(), 0, counts.data(), (u32)counts.size(), &m_batch), "b2bvh_build_batched");
    b2bvh_ctx* c = context.m_ctx;
    d_bvhNodes.bind(c, m_batch.d_bvhNodes, m_batch.n_nodes_total);
    d_primRefs.bind(c, m_batch.d_primRefs, m_batch.n_prims_total);
    d_rootNodes.bind(c, m_batch.d_rootNodes, m_batch.n_items);
    d_sceneExtents.bind(c, m_batch.d_sceneExtents, m_batch.n_items);
    m_leafOffsets.assign(counts.size() + 1, 0);
    m_nodeOffsets.assign(counts.size() + 1, 0);
    for (size_t i = 0; i < counts.size(); i++) { m_leafOffsets[i + 1] = m_leafOffsets[i] + counts[i]; m_nodeOffsets[i + 1] = m_nodeOffsets[i] + counts[i] - 1; }
    m_timer.timeRecord[BvhBuildTime] += m_batch.build_ms;
    /* wording of BatchedBuilder.cpp:72-76 */
    std::cout << "==========================Perf Times==========================" << std::endl;
    std::cout << "BatchSize : " << batch.size() << std::endl;
    std::cout << "BvhBuildTime : " << m_timer.getTimeRecord(BvhBuildTime) << "ms" << std::endl;
    std::cout << "==============================================================" << std::endl;
  }
  void traverseBvh(Context&) {} /* empty in the reference as well (BatchedBuilder.cpp:79-82) */

  GpuMemory<Bvh2Node> d_bvhNodes;
  GpuMemory<PrimRef> d_primRefs;
  GpuMemory<u32> d_rootNodes;
  GpuMemory<Aabb> d_sceneExtents;
  std::vector<u32> m_leafOffsets, m_nodeOffsets;
  u32 m_rootNodeIdx = 0;
  Timer m_timer;
  u32 m_nInternalNodes = 0;
  float m_cost = 0.0f;
  b2bvh_batch m_batch{};
};

/* Primitive-range sharded build over several contexts, one per GPU (north star: large meshes shard by primitive range across the GPUs
 * of a node; the reference is single-device, Context.cpp:11 — no counterpart).  Shard g = primitives [g*N/G, (g+1)*N/G), built by the
 * chosen LBVH builder in the frame of the global scene box; m_topNodes = the top-level tree over the shard roots (leaf m_leftChildIdx =
 * shard).  One host thread: b2bvh_build_sharded enqueues every shard before it waits for the first. */
class ShardedLbvh {
 public:
  void build(std::vector<Context*>& contexts, std::vector<Triangle>& primitives, int algo = B2BVH_SINGLE_PASS_LBVH) {
    const u32 G = (u32)contexts.size();
    const size_t N = primitives.size();
    std::vector<b2bvh_ctx*> ctxs(G);
    std::vector<const Triangle*> ptrs(G);
    std::vector<u32> counts(G);
    m_firstPrim.assign(G + 1, 0);
    for (u32 g = 0; g <= G; g++) m_firstPrim[g] = (u32)((N * g) / G);
    for (u32 g = 0; g < G; g++) { ctxs[g] = contexts[g]->m_ctx; ptrs[g] = primitives.data() + m_firstPrim[g]; counts[g] = m_firstPrim[g + 1] - m_firstPrim[g]; }
    m_trees.assign(G, b2bvh_tree{});
    m_topNodes.assign(2 * (size_t)G - 1, Bvh2Node{});
    checkStatus(Api::get().b2bvh_build_sharded(ctxs.data(), G, algo, ptrs.data(), counts.data(), nullptr, m_trees.data(), &m_sceneExtents, m_topNodes.data()),
                "b2bvh_build_sharded");
    float slowest = 0.0f;
    std::cout << "==========================Perf Times==========================" << std::endl;
    for (u32 g = 0; g < G; g++) {
      std::cout << "Shard " << g << " : " << counts[g] << " primitives, BuildTime " << m_trees[g].build_ms << "ms, wide nodes " << m_trees[g].n_wide << std::endl;
      slowest = m_trees[g].build_ms > slowest ? m_trees[g].build_ms : slowest;
    }
    std::cout << "Total Time (slowest shard) : " << slowest << "ms" << std::endl;
    std::cout << "==============================================================" << std::endl;
  }
  std::vector<b2bvh_tree> m_trees;   /* device buffers of shard g live in contexts[g] */
  std::vector<Bvh2Node> m_topNodes;  /* host copy: 2G-1 nodes, root 0 */
  std::vector<u32> m_firstPrim;      /* shard g holds primitives [m_firstPrim[g], m_firstPrim[g+1]) */
  Aabb m_sceneExtents{};
};

}  // namespace BvhConstruction
