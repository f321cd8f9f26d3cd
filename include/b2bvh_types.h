/*
 * b2bvh_types.h — POD layouts shared by the C ABI, the CUDA kernels, the C++ host
 * classes and the CPU oracle.  Every struct is byte-for-byte the layout of the
 * reference type it replaces (reference: src/Common.h), checked by the
 * static asserts at the bottom (sizes/offsets from SURVEY.md Appendix A).
 *
 *   b2bvh_triangle   <- Triangle        Common.h:429-434   64 B, align 64
 *   b2bvh_aabb       <- Aabb            Common.h:310-416   24 B, align 4
 *   b2bvh_bvh2_node  <- Bvh2Node        Common.h:436-441   32 B, align 32
 *   b2bvh_bvh4_node  <- Bvh4Node        Common.h:560-566  128 B, align 128
 *   b2bvh_sah_node   <- SahBvhNode      Common.h:443-453   32 B, align 32
 *   b2bvh_prim_ref   <- PrimRef         Common.h:574-578   28 B, align 4
 *   b2bvh_prim_node  <- PrimNode        Common.h:568-572    8 B
 *   b2bvh_ray        <- Ray             Common.h:533-539   32 B, align 32
 *   b2bvh_hit        <- HitInfo         Common.h:580-585   32 B, align 32
 *   b2bvh_transform  <- Transformation  Common.h:541-548   64 B, align 64
 *   b2bvh_camera     <- Camera          Common.h:550-558   64 B, align 64
 */
#ifndef B2BVH_TYPES_H
#define B2BVH_TYPES_H

#include <stddef.h>
#include <stdint.h>

#if defined(__cplusplus)
#define B2BVH_ALIGNAS(n) alignas(n)
#else
#define B2BVH_ALIGNAS(n) __attribute__((aligned(n))) /* C: attribute form is accepted after `struct` */
#endif

#define B2BVH_INVALID 0xFFFFFFFFu          /* INVALID_NODE_IDX / INVALID_PRIM_IDX, Common.h:90-92 */
#define B2BVH_FLT_MAX 3.402823466e+38f     /* FltMax, Common.h:86 */
#define B2BVH_PLOC_RADIUS 8                /* PlocRadius, Common.h:595 */
#define B2BVH_PLOC_BLOCK 1024              /* PlocBlockSize, Common.h:593 */

typedef struct { float x, y, z; } b2bvh_float3;
typedef struct { float x, y, z, w; } b2bvh_float4;

typedef struct { b2bvh_float3 m_min, m_max; } b2bvh_aabb;

typedef struct B2BVH_ALIGNAS(64) { b2bvh_float3 v1, v2, v3; } b2bvh_triangle;

typedef struct B2BVH_ALIGNAS(32) {
  uint32_t m_leftChildIdx;
  uint32_t m_rightChildIdx;
  b2bvh_aabb m_aabb;
} b2bvh_bvh2_node;

typedef struct B2BVH_ALIGNAS(128) {
  b2bvh_aabb m_aabb[4];
  uint32_t m_child[4];
  uint32_t m_parent;
  uint32_t m_childCount;
  uint32_t m_pad[2]; /* bytes 120..127: indeterminate in the reference, zero here */
} b2bvh_bvh4_node;

typedef struct B2BVH_ALIGNAS(32) {
  b2bvh_aabb m_aabb;
  uint32_t m_firstChildIdx;
  uint32_t m_primCount;
} b2bvh_sah_node;

typedef struct { uint32_t m_primIdx; b2bvh_aabb m_aabb; } b2bvh_prim_ref;
typedef struct { uint32_t m_primIdx; uint32_t m_parent; } b2bvh_prim_node;

typedef struct B2BVH_ALIGNAS(32) {
  b2bvh_float3 m_origin;
  b2bvh_float3 m_direction;
  float m_tMin;
  float m_tMax;
} b2bvh_ray;

typedef struct B2BVH_ALIGNAS(32) {
  uint32_t m_primIdx;
  float m_t;
  float m_u, m_v; /* HitInfo::m_uv */
} b2bvh_hit;

typedef struct B2BVH_ALIGNAS(64) {
  b2bvh_float3 m_translation; float m_pad;
  b2bvh_float3 m_scale;       float m_pad1;
  b2bvh_float4 m_quat;
} b2bvh_transform;

typedef struct B2BVH_ALIGNAS(64) {
  b2bvh_float4 m_eye;
  b2bvh_float4 m_quat;
  float m_fov, m_near, m_far, m_pad;
} b2bvh_camera;

/* Stage tokens, same order as TimerCodes (Common.h:418-427). */
enum b2bvh_stage {
  B2BVH_T_EXTENTS = 0, /* CalculateCentroidExtentsTime */
  B2BVH_T_MORTON = 1,  /* CalculateMortonCodesTime (PLOC/HPLOC: + SetupClusters, PLOC++Bvh.cpp:111) */
  B2BVH_T_SORT = 2,    /* SortingTime */
  B2BVH_T_BUILD = 3,   /* BvhBuildTime */
  B2BVH_T_TRAVERSAL = 4,
  B2BVH_T_COLLAPSE = 5,
  B2BVH_T_RAYGEN = 6,
  B2BVH_T_COUNT = 7
};

#if defined(__cplusplus)
static_assert(sizeof(b2bvh_triangle) == 64 && alignof(b2bvh_triangle) == 64, "Triangle layout");
static_assert(offsetof(b2bvh_triangle, v2) == 12 && offsetof(b2bvh_triangle, v3) == 24, "Triangle layout");
static_assert(sizeof(b2bvh_aabb) == 24 && alignof(b2bvh_aabb) == 4, "Aabb layout");
static_assert(sizeof(b2bvh_bvh2_node) == 32 && offsetof(b2bvh_bvh2_node, m_aabb) == 8, "Bvh2Node layout");
static_assert(sizeof(b2bvh_bvh4_node) == 128 && offsetof(b2bvh_bvh4_node, m_child) == 96 &&
              offsetof(b2bvh_bvh4_node, m_parent) == 112 && offsetof(b2bvh_bvh4_node, m_childCount) == 116, "Bvh4Node layout");
static_assert(sizeof(b2bvh_sah_node) == 32 && offsetof(b2bvh_sah_node, m_firstChildIdx) == 24, "SahBvhNode layout");
static_assert(sizeof(b2bvh_prim_ref) == 28 && offsetof(b2bvh_prim_ref, m_aabb) == 4, "PrimRef layout");
static_assert(sizeof(b2bvh_prim_node) == 8, "PrimNode layout");
static_assert(sizeof(b2bvh_ray) == 32 && offsetof(b2bvh_ray, m_direction) == 12 && offsetof(b2bvh_ray, m_tMin) == 24, "Ray layout");
static_assert(sizeof(b2bvh_hit) == 32 && offsetof(b2bvh_hit, m_t) == 4 && offsetof(b2bvh_hit, m_u) == 8, "HitInfo layout");
static_assert(sizeof(b2bvh_transform) == 64 && offsetof(b2bvh_transform, m_scale) == 16 && offsetof(b2bvh_transform, m_quat) == 32, "Transformation layout");
static_assert(sizeof(b2bvh_camera) == 64 && offsetof(b2bvh_camera, m_fov) == 32, "Camera layout");
#endif

#endif /* B2BVH_TYPES_H */
