/*
 * demo_main.cpp — the reference's driver (src/main.cpp:26-85) against the B200 library: pick a builder, load triangles,
 * build(), traverseBvh().  The reference chooses the builder with #defines (main.cpp:18-22) and hard-codes an OBJ path; here
 * both are command-line arguments and the mesh is a raw triangle file (float32 x 9 per triangle, the format written by
 * the mesh staging script of the test infrastructure with the reference's own OBJ loader) or a synthetic stream.
 *
 *   b2bvh_demo <twopass|singlepass|ploc|hploc> <mesh.tri | synth:N> [expected_cost]
 *   b2bvh_demo twopass-split:<saMax> <mesh.tri | synth:N> [expected_cost]   TwoPassLbvh compiled with USE_PRIM_SPLITTING (TwoPassLbvh.cpp:23-28)
 *   b2bvh_demo batched <mesh.tri | synth:N>      the USE_BATCHED_BUILDER branch (main.cpp:38-52): 4096 items; a mesh of <= 32
 *                                                 triangles is one item repeated (the reference's cornell box), a larger one is cut
 *                                                 into items of 32 consecutive triangles
 *   b2bvh_demo sharded:<G> <mesh.tri | synth:N>  primitive-range sharded single-pass LBVH over G contexts (device g % device count): prints the
 *                                                 per-shard table, checks that the top-level root box is the global scene box
 * exits 0 when the build succeeded (and the cost matches expected_cost to 1e-5 relative, when given).
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>

#include "BvhConstruction.h"

using namespace BvhConstruction;

static std::vector<Triangle> loadTriangles(Context& ctx, const std::string& arg) {
  std::vector<Triangle> tris;
  if (arg.rfind("synth:", 0) == 0) {
    const u32 n = (u32)atoll(arg.c_str() + 6);
    void* d = nullptr;
    checkStatus(Api::get().b2bvh_alloc(ctx.m_ctx, (size_t)n * sizeof(Triangle), &d), "b2bvh_alloc");
    const float half = (float)(1000.0 * pow((double)n, -1.0 / 3.0));
    checkStatus(Api::get().b2bvh_synth_uniform(ctx.m_ctx, 0, n, 0x00B20010u, half, (Triangle*)d), "b2bvh_synth_uniform");
    tris.resize(n);
    checkStatus(Api::get().b2bvh_d2h(ctx.m_ctx, tris.data(), d, (size_t)n * sizeof(Triangle)), "b2bvh_d2h");
    Api::get().b2bvh_free(ctx.m_ctx, d);
    return tris;
  }
  std::ifstream f(arg, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + arg);
  const size_t bytes = (size_t)f.tellg();
  f.seekg(0);
  std::vector<float> raw(bytes / 4);
  f.read(reinterpret_cast<char*>(raw.data()), (std::streamsize)bytes);
  tris.resize(raw.size() / 9);
  for (size_t i = 0; i < tris.size(); i++) {
    memset(&tris[i], 0, sizeof(Triangle));
    tris[i].v1 = float3{raw[9 * i + 0], raw[9 * i + 1], raw[9 * i + 2]};
    tris[i].v2 = float3{raw[9 * i + 3], raw[9 * i + 4], raw[9 * i + 5]};
    tris[i].v3 = float3{raw[9 * i + 6], raw[9 * i + 7], raw[9 * i + 8]};
  }
  return tris;
}

template <class Builder>
static float run(Context& ctx, std::vector<Triangle>& tris, float saMax = 0.0f) {
  Builder bvh;
  bvh.m_saMax = saMax;
  bvh.build(ctx, tris);
  if (saMax > 0.0f) std::cout << "references : " << bvh.d_primRefIdx.size() << " of " << bvh.d_triangleBuff.size() << " triangles" << std::endl;
  bvh.traverseBvh(ctx);
  std::cout << "wide nodes : " << bvh.m_wideNodeCount << "  root : " << bvh.m_rootNodeIdx << "  internal nodes : " << bvh.m_nInternalNodes << std::endl;
  size_t hits = 0;
  for (const HitInfo& h : bvh.m_hits) hits += h.m_primIdx != INVALID_NODE_IDX;
  if (!bvh.m_hits.empty()) std::cout << "primary rays : " << bvh.m_hits.size() << "  hits : " << hits << std::endl;
  return bvh.m_cost;
}

int main(int argc, char* argv[]) {
  try {
    if (argc < 3) { fprintf(stderr, "usage: %s <twopass|singlepass|ploc|hploc> <mesh.tri | synth:N> [expected_cost]\n", argv[0]); return 2; }
    Context context;
    std::vector<Triangle> triangles = loadTriangles(context, argv[2]);
    std::cout << "triangles : " << triangles.size() << std::endl;
    const std::string which = argv[1];
    float cost;
    if (which == "batched") {
      constexpr u32 batchSize = 2048 * 2; /* main.cpp:42 */
      std::vector<BatchedBuildInput> batches(batchSize);
      for (u32 i = 0; i < batchSize; i++) {
        if (triangles.size() <= 32) batches[i].m_primitives = triangles;
        else {
          const size_t first = ((size_t)i * 32) % (triangles.size() - 31);
          batches[i].m_primitives.assign(triangles.begin() + first, triangles.begin() + first + 32);
        }
      }
      BatchedBvhBuilder bvh;
      bvh.build(context, batches);
      bvh.traverseBvh(context);
      const std::vector<u32> roots = bvh.d_rootNodes.getData();
      const std::vector<Bvh2Node> nodes = bvh.d_bvhNodes.getData();
      const std::vector<Aabb> scenes = bvh.d_sceneExtents.getData();
      /* every item's root box is its scene box */
      for (u32 i = 0; i < batchSize; i++) {
        if (batches[i].m_primitives.size() < 2) continue;
        const Bvh2Node& r = nodes[bvh.m_nodeOffsets[i] + roots[i]];
        if (memcmp(&r.m_aabb, &scenes[i], sizeof(Aabb)) != 0) { fprintf(stderr, "item %u: root box differs from the item's scene box\n", i); return 1; }
      }
      std::cout << "items : " << batchSize << "  nodes : " << nodes.size() << "  root of item 0 : " << roots[0] << std::endl;
      return 0;
    }
    if (which.rfind("sharded:", 0) == 0) {
      const int G = atoi(which.c_str() + 8);
      if (G < 1 || G > 64) throw std::runtime_error("sharded:<G> needs 1 <= G <= 64");
      int devices = 1;
      if (const char* e = getenv("B2BVH_DEVICES")) devices = atoi(e) > 0 ? atoi(e) : 1;
      std::vector<std::unique_ptr<Context>> owned;
      std::vector<Context*> ctxs;
      for (int g = 0; g < G; g++) { owned.emplace_back(new Context(g % devices)); ctxs.push_back(owned.back().get()); }
      ShardedLbvh bvh;
      bvh.build(ctxs, triangles);
      const Bvh2Node& root = bvh.m_topNodes[0];
      if (memcmp(&root.m_aabb, &bvh.m_sceneExtents, sizeof(Aabb)) != 0) { fprintf(stderr, "top-level root box differs from the global scene box\n"); return 1; }
      std::cout << "shards : " << G << "  top-level nodes : " << bvh.m_topNodes.size() << "  scene min : " << bvh.m_sceneExtents.m_min.x << " " << bvh.m_sceneExtents.m_min.y
                << " " << bvh.m_sceneExtents.m_min.z << std::endl;
      return 0;
    }
    if (which == "twopass") cost = run<TwoPassLbvh>(context, triangles);
    else if (which.rfind("twopass-split:", 0) == 0) cost = run<TwoPassLbvh>(context, triangles, (float)atof(which.c_str() + 14));
    else if (which == "singlepass") cost = run<SinglePassLbvh>(context, triangles);
    else if (which == "ploc") cost = run<PLOCNew>(context, triangles);
    else if (which == "hploc") cost = run<HPLOC>(context, triangles);
    else throw std::runtime_error("unknown builder " + which);
    if (argc > 3) {
      const float want = (float)atof(argv[3]);
      if (fabsf(cost - want) > 1e-5f * fabsf(want)) { fprintf(stderr, "cost %.7g differs from expected %.7g\n", cost, want); return 1; }
    }
  } catch (std::exception& e) { /* main.cpp:80-84 */
    std::cerr << e.what();
    return -1;
  }
  return 0;
}
