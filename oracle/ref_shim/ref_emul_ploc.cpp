/* TEST INFRASTRUCTURE ONLY — reference Ploc++Kernel.h: SetupClusters and the PLOC-layout collapse are
 * executed thread by thread; Ploc / SinglePassPloc need real block-level execution and only compile. */
#include <src/Common.h>
#define SetupClusters PL_SetupClusters
#define CollapseToWide4Bvh PL_CollapseToWide4Bvh
namespace {
#include <src/Ploc++Kernel.h>
}
template <class F> static void run1d(uint32_t n, F f) {
  blockDim = {256, 1, 1};
  for (uint32_t g = 0; g < n; g++) { blockIdx.x = g / 256; threadIdx.x = g % 256; f(); }
}
extern "C" {
void ref_ploc_setup(Bvh2Node* nodes, PrimRef* leaves, uint32_t* sortedVals, Aabb* triAabb, int* nodeIdx, uint32_t n) {
  run1d(n, [&] { PL_SetupClusters(nodes, leaves, sortedVals, triAabb, nodeIdx, n); });
}
uint32_t ref_collapse_ploc(Bvh2Node* nodes, PrimRef* leaves, uint32_t root, uint32_t n, Bvh4Node* wide, PrimNode* wideLeaves) {
  const uint32_t nInt = n - 1;
  uint2* taskQ = new uint2[n];
  for (uint32_t i = 0; i < n; i++) taskQ[i] = uint2{INVALID_NODE_IDX, INVALID_NODE_IDX};
  taskQ[0] = uint2{root, INVALID_NODE_IDX};
  uint32_t taskCount = 0, offset = 1;
  blockDim = {256, 1, 1};
  for (uint32_t g = 0; g < offset && taskCount < n; g++) {
    blockIdx.x = g / 256; threadIdx.x = g % 256;
    PL_CollapseToWide4Bvh(nodes, leaves, wide, wideLeaves, taskQ, &taskCount, &offset, nInt, n);
  }
  delete[] taskQ;
  return offset;
}
}
