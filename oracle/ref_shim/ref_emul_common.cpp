/* TEST INFRASTRUCTURE ONLY — runs the reference's own kernels (CommonBlocksKernel.h), unmodified,
 * one emulated GPU thread at a time.  Sources are included from /root/reference in place. */
#include <src/Common.h>
#include <src/CommonBlocksKernel.h>

thread_local dim3e threadIdx, blockIdx, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

template <class F> static void run1d(uint32_t n, F f) {
  blockDim = {256, 1, 1};
  for (uint32_t g = 0; g < n; g++) { blockIdx.x = g / 256; threadIdx.x = g % 256; f(); }
}

extern "C" {
uint32_t ref_morton_code(const float p[3], const float ext[3]) {
  return computeExtendedMortonCode(float3{p[0], p[1], p[2]}, float3{ext[0], ext[1], ext[2]});
}
void ref_init_primrefs(PrimRef* refs, const Triangle* tris, uint32_t n) { run1d(n, [&] { InitPrimRefs(refs, tris, n); }); }
void ref_morton_primref(const PrimRef* refs, const Aabb* scene, uint32_t* keys, uint32_t* vals, uint32_t n) {
  run1d(n, [&] { CalculateMortonCodesPrimRef(refs, scene, keys, vals, n); });
}
void ref_morton_aabb(const Aabb* boxes, const Aabb* scene, uint32_t* keys, uint32_t* vals, uint32_t n) {
  run1d(n, [&] { CalculateMortonCodes(boxes, scene, keys, vals, n); });
}
uint32_t ref_morton_plain(const float p[3]) { return computeMortonCode(float3{p[0], p[1], p[2]}, float3{1.0f, 1.0f, 1.0f}); } /* CommonBlocksKernel.h:361-372 */
uint32_t ref_tea16(uint32_t a, uint32_t b) { return tea<16>(a, b).x; }
float ref_randf(uint32_t* seed) { return randf(*seed); }
}
