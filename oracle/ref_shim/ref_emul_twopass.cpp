/* TEST INFRASTRUCTURE ONLY — reference TwoPassLbvhKernel.h executed thread by thread. */
#include <src/Common.h>
#define InitBvhNodes TP_InitBvhNodes
#define CollapseToWide4Bvh TP_CollapseToWide4Bvh
namespace {
#include <src/TwoPassLbvhKernel.h>
}
template <class F> static void run1d(uint32_t n, F f) {
  blockDim = {256, 1, 1};
  for (uint32_t g = 0; g < n; g++) { blockIdx.x = g / 256; threadIdx.x = g % 256; f(); }
}
extern "C" {
/* launch order of TwoPassLbvh.cpp:99-143 */
void ref_twopass_build(const PrimRef* refs, const uint32_t* keys, const uint32_t* vals, uint32_t n,
                       Bvh2Node* nodes, uint32_t* parents, uint32_t* flags) {
  const uint32_t nInt = n - 1;
  run1d(n, [&] { InitBvhNodesPrimRef(refs, nodes, parents, vals, n, nInt); });
  run1d(nInt, [&] { BvhBuild(nodes, parents, keys, n, nInt); });
  memset(flags, 0, sizeof(uint32_t) * (2 * (size_t)n - 1));
  run1d(n, [&] { FitBvhNodes(nodes, parents, flags, n, nInt); });
}
/* TwoPassLbvh.cpp:154-183; threads are run in task order, each once its task exists */
uint32_t ref_collapse_lbvh(Bvh2Node* nodes, uint32_t root, uint32_t n, Bvh4Node* wide, PrimNode* wideLeaves) {
  const uint32_t nInt = n - 1;
  uint2* taskQ = new uint2[n];
  for (uint32_t i = 0; i < n; i++) taskQ[i] = uint2{INVALID_NODE_IDX, INVALID_NODE_IDX};
  taskQ[0] = uint2{root, INVALID_NODE_IDX};
  uint32_t taskCount = 0, offset = 1;
  blockDim = {256, 1, 1};
  for (uint32_t g = 0; g < offset && taskCount < n; g++) {
    blockIdx.x = g / 256; threadIdx.x = g % 256;
    TP_CollapseToWide4Bvh(nodes, wide, wideLeaves, taskQ, &taskCount, &offset, nInt, n);
  }
  delete[] taskQ;
  return offset;
}
}
