#include "common.cuh"
size_t b2_hploc_scratch_bytes(u32 n) { return 16; }
int b2_launch_hploc(b2bvh_ctx*, const b2bvh_aabb*, const u32*, const u32*, u32, b2bvh_bvh2_node*, b2bvh_prim_ref*, void*, u32*) { return b2_fail(B2BVH_ERR_INTERNAL, "hploc: not built yet"); }
extern "C" {
int b2bvh_generate_rays(b2bvh_ctx*, const b2bvh_camera*, uint32_t, uint32_t, b2bvh_ray*, float*) { return b2_fail(B2BVH_ERR_INTERNAL, "not built yet"); }
int b2bvh_traverse(b2bvh_ctx*, const b2bvh_tree*, const b2bvh_ray*, uint32_t, const b2bvh_transform*, int, b2bvh_hit*, uint8_t*, float*) { return b2_fail(B2BVH_ERR_INTERNAL, "not built yet"); }
int b2bvh_top_level(b2bvh_ctx*, const b2bvh_aabb*, uint32_t, b2bvh_bvh2_node*) { return b2_fail(B2BVH_ERR_INTERNAL, "not built yet"); }
}
