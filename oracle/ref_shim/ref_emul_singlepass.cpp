/* TEST INFRASTRUCTURE ONLY — reference SinglePassLbvhKernel.h executed thread by thread. */
#include <src/Common.h>
#define InitBvhNodes SP_InitBvhNodes
namespace {
#include <src/SinglePassLbvhKernel.h>
}
template <class F> static void run1d(uint32_t n, F f) {
  blockDim = {256, 1, 1};
  for (uint32_t g = 0; g < n; g++) { blockIdx.x = g / 256; threadIdx.x = g % 256; f(); }
}
extern "C" {
/* launch order of SinglePassLbvh.cpp:99-131; returns bvhNodeCounter[nLeafNodes-1] */
uint32_t ref_singlepass_build(const Triangle* tris, const uint32_t* keys, const uint32_t* vals, uint32_t n, Bvh2Node* nodes) {
  const uint32_t nInt = n - 1;
  run1d(n, [&] { SP_InitBvhNodes(tris, nodes, vals, nInt, n); });
  uint2* spans = new uint2[n]();
  int* counter = new int[n]();
  run1d(n, [&] { BvhBuildAndFit(nodes, counter, spans, keys, n, nInt); });
  uint32_t root = (uint32_t)counter[n - 1];
  delete[] spans; delete[] counter;
  return root;
}
}
