/* TEST INFRASTRUCTURE ONLY — C wrappers around the reference's own host utilities
 * (src/Utility.cpp, compiled as is next to this file by oracle/Makefile). */
#include <src/Utility.h>
#include <cstring>
using namespace BvhConstruction;
extern "C" {
float ref_cost_bvh4(const Bvh4Node* w, const PrimNode* wl, Aabb* prim, uint32_t root, uint32_t total, uint32_t nInt) {
  return Utility::calculatebvh4Cost(w, wl, prim, root, total, nInt);
}
float ref_cost_lbvh(const Bvh2Node* nodes, uint32_t root, uint32_t nLeaf, uint32_t nInt) { return Utility::calculateLbvhCost(nodes, root, nLeaf, nInt); }
float ref_cost_sah(const SahBvhNode* nodes, uint32_t root, uint32_t total) { return Utility::calculateBinnedSahBvhCost(nodes, root, total); }
int ref_check_root_aabb(const Bvh2Node* nodes, uint32_t root, uint32_t nLeaf, uint32_t nInt) { return Utility::checkLbvhRootAabb(nodes, root, nLeaf, nInt); }
int ref_check_bvh4(const Bvh4Node* w, const PrimNode* wl, uint32_t root, uint32_t nInt) { return Utility::checkLBvh4Correctness(w, wl, root, nInt); }
/* doEarlySplitClipping with the default saMax (no splitting) */
uint32_t ref_early_split(const Triangle* tris, uint32_t n, PrimRef* out) {
  std::vector<Triangle> in(tris, tris + n);
  std::vector<PrimRef> refs;
  Utility::doEarlySplitClipping(in, refs);
  memcpy(out, refs.data(), refs.size() * sizeof(PrimRef));
  return (uint32_t)refs.size();
}
/* doEarlySplitClipping with a finite saMax (USE_PRIM_SPLITTING, TwoPassLbvh.cpp:23-28); returns the count, writes <= cap */
uint32_t ref_early_split_sa(const Triangle* tris, uint32_t n, float saMax, PrimRef* out, uint32_t cap) {
  std::vector<Triangle> in(tris, tris + n);
  std::vector<PrimRef> refs;
  Utility::doEarlySplitClipping(in, refs, saMax);
  if (out) memcpy(out, refs.data(), std::min<size_t>(cap, refs.size()) * sizeof(PrimRef));
  return (uint32_t)refs.size();
}
/* OBJ -> triangles with the reference's loader; returns the count (call with out == nullptr to size) */
uint32_t ref_load_obj(const char* file, const char* mtlDir, Triangle* out, uint32_t cap) {
  static std::vector<Triangle> cache; static std::string cached;
  if (cached != file) { cache.clear(); MeshLoader::loadScene(file, mtlDir, cache); cached = file; }
  if (out) memcpy(out, cache.data(), sizeof(Triangle) * std::min<size_t>(cap, cache.size()));
  return (uint32_t)cache.size();
}
/* Utility::TraversalLbvhCPU writes grey levels (t/30*255) for hit pixels into dst (RGBA8) */
void ref_traverse_cpu(const Ray* rays, const Bvh2Node* nodes, uint32_t nNodes, const Triangle* tris, uint32_t nTris,
                      Transformation* t, uint8_t* dst, uint32_t width, uint32_t height, uint32_t nInt) {
  std::vector<Ray> r(rays, rays + (size_t)width * height);
  std::vector<Bvh2Node> nd(nodes, nodes + nNodes);
  std::vector<Triangle> tr(tris, tris + nTris);
  Utility::TraversalLbvhCPU(r, nd, tr, *t, dst, width, height, nInt);
}
}
