/*
 * b2bvh.h — C ABI of libb2bvh.so, the B200 (sm_100a) BVH-build hot path.
 *
 * This is the drop-in boundary.  The reference (Niravaana/HIP-BVH-Construction) has no FFI
 * layer of its own: its builder classes talk to the GPU through Orochi (oro* calls, dlopen'd
 * HIP or CUDA driver API).  Every entry point below names the reference interface it stands
 * in for; the C++ builder classes in hip-bvh-construction_b200/host/ (same names and members as
 * the reference's) and the Python mirror in hip-bvh-construction_b200/b2bvh/ bind exactly these
 * symbols at run time (dlopen / ctypes), the way the reference binds Orochi.
 *
 * Conventions: extern "C", plain pointers and sizes, every call returns 0 on success and a
 * non-zero b2bvh_status otherwise (b2bvh_last_error() holds the text); no exceptions cross
 * the boundary; one context = one device + one stream; calls on one context are serialised
 * by the caller (the reference is single-threaded, Timer.h:48-56).
 * There is NO CPU fallback: without a CUDA device b2bvh_ctx_create fails with B2BVH_ERR_CUDA.
 */
#ifndef B2BVH_H
#define B2BVH_H

#include "b2bvh_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum b2bvh_status {
  B2BVH_OK = 0,
  B2BVH_ERR_INVALID = 1, /* bad argument (null pointer, n == 0, n too large, unknown enum) */
  B2BVH_ERR_CUDA = 2,    /* CUDA runtime error, or no device */
  B2BVH_ERR_OOM = 3,
  B2BVH_ERR_INTERNAL = 4 /* device-side consistency check failed */
} b2bvh_status;

/* Which builder object of the reference the build reproduces. */
typedef enum b2bvh_algo {
  B2BVH_TWO_PASS_LBVH = 0,    /* TwoPassLbvh::build    src/TwoPassLbvh.cpp:17-197    (Karras numbering, root 0)        */
  B2BVH_SINGLE_PASS_LBVH = 1, /* SinglePassLbvh::build src/SinglePassLbvh.cpp:17-188 (Apetrei numbering, root reported) */
  B2BVH_PLOCPP = 2,           /* PLOCNew::build        src/PLOC++Bvh.cpp:16-196      (canonical numbering, root 0)     */
  B2BVH_HPLOC = 3             /* HPLOC::build          src/Hploc.cpp:16-165          (canonical numbering, root 0)     */
} b2bvh_algo;

typedef enum b2bvh_traversal {
  B2BVH_TRAVERSE_WHILE = 0,            /* BvhTraversalWhile            src/TraversalKernel.h:238 */
  B2BVH_TRAVERSE_SPECULATIVE_WHILE = 1, /* BvhTraversalSpeculativeWhile src/TraversalKernel.h:337 */
  B2BVH_TRAVERSE_IFIF = 2,              /* BvhTraversalifif             src/TraversalKernel.h:148 (TwoPassLbvh.cpp:250-269 under IFIF) */
  B2BVH_TRAVERSE_RESTART_TRAIL = 3,     /* BvhTraversalRestartTrail     src/TraversalKernel.h:49  (stackless)                         */
  B2BVH_TRAVERSE_WIDE4 = 4              /* closest hit through the Bvh4 of the build (no reference kernel: the reference never walks it) */
} b2bvh_traversal;

typedef struct b2bvh_ctx b2bvh_ctx; /* opaque: device + stream + scratch arena; replaces Context (src/Context.cpp:7-15) */

typedef struct b2bvh_build_opts {
  uint32_t collapse;        /* 1: also build the 4-wide tree (CollapseToWide4Bvh, TwoPassLbvh.cpp:154-183)           */
  uint32_t tris_on_device;  /* 0: `tris` is a host pointer (copied H2D like TwoPassLbvh.cpp:19-20); 1: device pointer */
  uint32_t use_scene_box;   /* 1: use `scene_box` (global box of a sharded build) instead of the local union          */
  b2bvh_aabb scene_box;
  uint32_t stage_timing;    /* 1: CUDA-event time per stage (the reference's Timer tokens); 0: whole-build time only  */
  uint32_t karras_two_kernel; /* TWO_PASS only. 1: emit (determineRange/findSplit) + separate refit kernel, as the
                                 reference's launch order; 0: one fused bottom-up pass that yields the same numbering */
  uint32_t boxes_ready;     /* 1: b2bvh_shard_extents ran on the same context and triangles: primitive boxes are in place, skip S1 */
  const float* d_scene_negmin_max; /* device, 6 floats {-min.xyz, max.xyz} (the all-reduced output of b2bvh_shard_extents): the
                                      global scene box without a host round trip; overrides scene_box when not NULL     */
  uint32_t lbvh_second_level; /* LBVH builders: 0 = automatic (second merge level from 2^23 primitives), 1 = always, 2 = never; same output.
                                 H-PLOC: the same switch for its tile phase (automatic from 2^20 primitives) */
  uint32_t merge_max_ctas;  /* PLOC++: cap on the CTAs of the cooperative merge launch, 0 = every resident CTA; same output (tests use it to
                               reach many-tile chunks with small inputs) */
  uint32_t use_graph;       /* 1: capture the launch sequence of this build in a CUDA graph and REPLAY it when the next build on the context
                               has the same algorithm, size, options and triangle pointer (rebuilds of an animated mesh, benchmark loops):
                               one graph launch instead of ~20 kernel launches, which is what bounds builds of a few 100 K primitives.
                               Same output; ignored while the per-launch profiler is on or when scene_box comes from the host */
  float split_sa_max;       /* TWO_PASS only. > 0: early split clipping on the device before the build — TwoPassLbvh compiled with
                               USE_PRIM_SPLITTING (TwoPassLbvh.cpp:23-28; Utility::doEarlySplitClipping(prims, refs, saMax), Utility.cpp:456-538):
                               every primitive box with area > saMax is halved along its largest extent until all pieces fit; the build
                               then runs over the REFERENCES (b2bvh_tree.n_prims = their count, d_primRefIdx = their triangles).
                               0: off (the reference's default, saMax = FltMax) */
  uint32_t morton_bits;     /* 0 or 30: the reference's extended 30-bit code (computeExtendedMortonCode, CommonBlocksKernel.h:159-359).
                               60: plain 60-bit code, 20 bits per axis — computeMortonCode (:361-372) with 2^20 instead of 2^10 cells; no reference
                               counterpart (its codes stop at 30 bits), defined by the oracle.  All four builders (PLOC++ only needs the order,
                               the LBVH kernels and the H-PLOC walk are instantiated for 64-bit keys).  b2bvh_tree.d_mortonCodeKeys64 / d_sortedMortonCodeKeys64 hold the codes, the 32-bit key
                               arrays their upper 30 bits, d_mortonCodeValues is not written (the values are the iota) */
  uint32_t defer_sync;      /* 1: b2bvh_build enqueues the build and returns WITHOUT its host synchronisation; the b2bvh_tree then holds the
                               device pointers and sizes, and b2bvh_build_finish fills in root, n_wide, iteration counts and times.  Lets a caller
                               enqueue dependent work (the root all-gather of a sharded build) before the host waits */
  float* d_root_box_out;    /* device, 6 floats, or NULL: the root box {min.xyz, max.xyz} is written there at the end of the build (stream-ordered) */
} b2bvh_build_opts;

/* Everything a build leaves on the device.  Pointers are DEVICE pointers owned by the context
 * (valid until the next build on the same context or b2bvh_ctx_destroy).  Field names follow the
 * public GpuMemory<> members of the reference builders (src/TwoPassLbvh.h:19-30, src/Hploc.h:19-32). */
typedef struct b2bvh_tree {
  uint32_t algo;
  uint32_t n_prims;          /* N                                                     */
  uint32_t n_internal;       /* m_nInternalNodes = N-1                                */
  uint32_t root;             /* m_rootNodeIdx (Bvh2); the Bvh4 root is always 0       */
  uint32_t n_wide;           /* wide-node count (internalNodeOffset, TwoPassLbvh.cpp:187); 0 when collapse == 0 */
  uint32_t leaves_separate;  /* 0: LBVH layout, nodes[2N-1], leaf i = nodes[N-1+i]; 1: PLOC/HPLOC layout, nodes[N-1] + leaf_nodes[N] */
  const b2bvh_triangle* d_triangleBuff;       /* N                  */
  const b2bvh_aabb* d_triangleAabb;           /* N   primitive boxes */
  const b2bvh_aabb* d_sceneExtents;           /* 1                  */
  const uint32_t* d_mortonCodeKeys;           /* N                  */
  const uint32_t* d_mortonCodeValues;         /* N                  */
  const uint32_t* d_sortedMortonCodeKeys;     /* N                  */
  const uint32_t* d_sortedMortonCodeValues;   /* N                  */
  const b2bvh_bvh2_node* d_bvhNodes;          /* 2N-1 or N-1        */
  const uint32_t* d_parentIdxs;               /* 2N-1, TWO_PASS only (TwoPassLbvh.cpp:96), else NULL */
  const b2bvh_prim_ref* d_leafNodes;          /* N, PLOC/HPLOC only (PLOC++Bvh.h), else NULL         */
  const b2bvh_bvh4_node* d_wideBvhNodes;      /* n_wide             */
  const b2bvh_prim_node* d_wideLeafNodes;     /* N                  */
  float stage_ms[B2BVH_T_COUNT];              /* indexed by b2bvh_stage == TimerCodes; filled when stage_timing */
  float build_ms;                             /* device time of the whole build (extents..collapse), H2D excluded */
  float h2d_ms;                               /* device time of the triangle upload when tris_on_device == 0      */
  uint32_t n_iterations;                      /* PLOC++: iterations run; HPLOC: merge calls; else 0               */
  uint32_t n_launches;                        /* kernels launched by this build                                   */
  /* early split clipping (split_sa_max > 0): the build's primitives are PrimRefs; n_prims counts them, d_triangleAabb holds their
   * boxes (PrimRef::m_aabb) and d_primRefIdx their triangles (PrimRef::m_primIdx), both in the emission order of
   * doEarlySplitClipping; Bvh2 leaves and PrimNodes name TRIANGLES (InitBvhNodesPrimRef, TwoPassLbvhKernel.h:178-182) */
  uint32_t n_triangles;                       /* triangles uploaded (== n_prims without splitting)                */
  uint32_t n_split_levels;                    /* generations the split ran (0 without splitting)                  */
  float split_ms;                             /* device time of the split (inside stage_ms[extents])              */
  uint32_t reserved;
  const uint32_t* d_primRefIdx;               /* n_prims, or NULL without splitting                               */
  /* 60-bit Morton variant (morton_bits = 60) */
  const uint64_t* d_mortonCodeKeys64;         /* n_prims, or NULL                                                 */
  const uint64_t* d_sortedMortonCodeKeys64;   /* n_prims, or NULL                                                 */
  uint32_t morton_bits;                       /* 30 or 60                                                         */
  uint32_t reserved4;
} b2bvh_tree;

/* ---- context / memory: replaces Context (src/Context.cpp:7-22) and OrochiUtils malloc/copy helpers
 * (dependencies/Orochi/Orochi/OrochiUtils.h:84-161) that back GpuMemory<T>. ---- */
int b2bvh_ctx_create(int device, void* cuda_stream /* cudaStream_t or NULL for a private stream */, b2bvh_ctx** out);
int b2bvh_ctx_destroy(b2bvh_ctx* ctx);
int b2bvh_device_name(b2bvh_ctx* ctx, char* buf, size_t cap);          /* "Executing on '<device>'", Context.cpp:14 */
int b2bvh_device_sm_count(b2bvh_ctx* ctx, int* out);
int b2bvh_alloc(b2bvh_ctx* ctx, size_t bytes, void** dptr);            /* oroMalloc                                  */
int b2bvh_free(b2bvh_ctx* ctx, void* dptr);                            /* oroFree                                    */
int b2bvh_memset(b2bvh_ctx* ctx, void* dptr, int value, size_t bytes); /* GpuMemory::reset                           */
int b2bvh_h2d(b2bvh_ctx* ctx, void* dptr, const void* hptr, size_t bytes); /* OrochiUtils::copyHtoD                  */
int b2bvh_d2h(b2bvh_ctx* ctx, void* hptr, const void* dptr, size_t bytes); /* GpuMemory::getData                     */
int b2bvh_d2d(b2bvh_ctx* ctx, void* dst, const void* src, size_t bytes); /* OrochiUtils::copyDtoDAsync (stream-ordered)  */
/* stream-ordered copies that do not wait (pinned host memory; b2bvh_sync ends them): with two contexts the read-back of one
 * build overlaps the upload and build of the next */
int b2bvh_h2d_async(b2bvh_ctx* ctx, void* dptr, const void* hptr, size_t bytes); /* OrochiUtils::copyHtoDAsync, OrochiUtils.h:137 */
int b2bvh_d2h_async(b2bvh_ctx* ctx, void* hptr, const void* dptr, size_t bytes); /* OrochiUtils::copyDtoHAsync, OrochiUtils.h:144 */
int b2bvh_sync(b2bvh_ctx* ctx);                                        /* OrochiUtils::waitForCompletion             */
int b2bvh_host_alloc_pinned(size_t bytes, void** hptr);
int b2bvh_host_free_pinned(void* hptr);
const char* b2bvh_last_error(void);

/* ---- the build: replaces the body of <Builder>::build(Context&, std::vector<Triangle>&). ---- */
int b2bvh_build(b2bvh_ctx* ctx, int algo, const b2bvh_triangle* tris, uint32_t n, const b2bvh_build_opts* opts, b2bvh_tree* out);

/* ---- the batched builder: replaces BatchedBvhBuilder::build(Context&, std::vector<BatchedBuildInput>&) (src/BatchedBuilder.cpp:16-77,
 * kernel BatchedBuildKernelLbvh, src/BatchedBuildKernel.h:218-312): one small BVH per item, 1..32 triangles each (MaxBatchedBlockSize,
 * Common.h:597), all items in one launch.  `tris` = the items' triangles back to back (the reference's D_BatchedBuildInputs array of
 * pointers, Common.h:587-591, flattened), host or device; `counts` = HOST array, triangles per item.  Per item: scene box, plain
 * 10/10/10 Morton codes, stable sort, Apetrei hierarchy; leaves in sorted order; node / leaf indices LOCAL to the item (internal k in
 * [0, n-1), leaf g = (n-1)+g), item i's nodes at d_bvhNodes + d_nodeOffsets[i], its leaves at d_primRefs + d_leafOffsets[i]. ---- */
typedef struct b2bvh_batch {
  uint32_t n_items;
  uint32_t n_prims_total;                 /* sum of counts                                            */
  uint32_t n_nodes_total;                 /* sum of (count - 1)                                       */
  uint32_t reserved;
  const b2bvh_triangle* d_triangles;      /* n_prims_total                                            */
  const b2bvh_bvh2_node* d_bvhNodes;      /* BatchedBvhBuilder::d_bvhNodes, n_nodes_total             */
  const b2bvh_prim_ref* d_primRefs;       /* BatchedBvhBuilder::d_primRefs, n_prims_total             */
  const uint32_t* d_rootNodes;            /* BatchedBvhBuilder::d_rootNodes, n_items (local index)    */
  const b2bvh_aabb* d_sceneExtents;       /* d_sceneExtent of the kernel, n_items                     */
  const uint32_t* d_leafOffsets;          /* n_items + 1                                              */
  const uint32_t* d_nodeOffsets;          /* n_items + 1                                              */
  float build_ms;                         /* BvhBuildTime (BatchedBuilder.cpp:61)                     */
  float h2d_ms;
} b2bvh_batch;
int b2bvh_build_batched(b2bvh_ctx* ctx, const b2bvh_triangle* tris, uint32_t tris_on_device, const uint32_t* counts, uint32_t n_items,
                        b2bvh_batch* out);

/* completes a build enqueued with defer_sync = 1 (the build's single host synchronisation) */
int b2bvh_build_finish(b2bvh_ctx* ctx, b2bvh_tree* tree);

/* Stages, individually callable (all pointers are device pointers).  Each names the reference kernel(s) it replaces. */
int b2bvh_scene_extents(b2bvh_ctx* ctx, const b2bvh_triangle* d_tris, uint32_t n, b2bvh_aabb* d_triAabb,
                        b2bvh_aabb* d_scene);             /* CalculateSceneExtents / CalculatePrimRefExtents, CommonBlocksKernel.h:92,116 */
int b2bvh_morton_codes(b2bvh_ctx* ctx, const b2bvh_aabb* d_triAabb, const b2bvh_aabb* d_scene, uint32_t n, uint32_t* d_keys,
                       uint32_t* d_vals);                 /* CalculateMortonCodes[PrimRef], CommonBlocksKernel.h:374,387                  */
int b2bvh_sort_pairs(b2bvh_ctx* ctx, const uint32_t* d_keysIn, const uint32_t* d_valsIn, uint32_t* d_keysOut, uint32_t* d_valsOut,
                     uint32_t n, uint32_t startBit, uint32_t endBit); /* Oro::RadixSort::sort(KeyValueSoA...), RadixSort.cpp:291-318   */

/* The hierarchy stage alone (InitBvhNodes + BvhBuildAndFit / BvhBuild + FitBvhNodes over an already sorted sequence), with 64-bit keys:
 * d_sortedKeys64 non-decreasing, d_sortedVals[g] indexes d_primAabb; nodes in the LBVH layout (2n-1), Karras (root 0, parents written) or
 * Apetrei numbering (root returned).  Building block of the globally sorted multi-GPU build: a rank runs it over its range of the global
 * order with keys widened to (code << 32 | global sorted position), see DESIGN.md section 9. */
int b2bvh_lbvh_from_sorted64(b2bvh_ctx* ctx, const uint64_t* d_sortedKeys64, const uint32_t* d_sortedVals, const b2bvh_aabb* d_primAabb, uint32_t n,
                             int karras, b2bvh_bvh2_node* d_nodes, uint32_t* d_parents /* 2n-1 words, Karras only */, uint32_t* root);

/* A finished sub-tree of a rank whose parent straddles a rank boundary (globally sorted multi-GPU build): global sorted positions
 * [lo, hi), global node index, box.  48 bytes. */
typedef struct b2bvh_cluster {
  uint32_t lo, hi, node, pad;
  b2bvh_aabb box;
  uint32_t pad2[2];
} b2bvh_cluster;
/* Second building block of the globally sorted build: d_local = the tree b2bvh_lbvh_from_sorted64 built over the
 * rank's range of the global order extended by a ghost leaf on the left (ghost_left) and / or right (ghost_right) edge, local leaf 0 at
 * global sorted position first_pos.  d_out (2m-1 nodes) receives the ghost-free nodes with GLOBAL child indices (internal i -> first_pos
 * + i, leaf slot g -> (n_global - 1) + first_pos + g), artefact nodes as {INVALID, INVALID, empty box}, leaves unchanged; d_clusters
 * (room for 256) the left-over clusters in position order, *count their number. */
int b2bvh_range_extract(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_local, uint32_t m, uint32_t local_root, int karras, uint32_t ghost_left,
                        uint32_t ghost_right, uint32_t first_pos, uint32_t n_global, b2bvh_bvh2_node* d_out, b2bvh_cluster* d_clusters, uint32_t* count);

/* ---- The globally sorted multi-GPU build as device steps: G ranks produce the nodes of the ONE-GPU LBVH over all
 * triangles, distributed by sorted position.  Rank-local kernels only, all asynchronous on the context's stream; the collectives between
 * them (scene box all-reduce, sample all-gather, all-to-all of counts and payload, all-gather of edge codes and clusters) belong to the
 * caller's communicator (b2bvh/sharded.py GlobalBuildDevice).  New work specified by the north star; no reference counterpart. ---- */
/* A node of the top of the tree, whose range straddles rank borders: global node index, global child indices, box.  48 bytes. */
typedef struct b2bvh_top_node {
  uint32_t index, left, right, pad;
  b2bvh_aabb box;
  uint32_t pad2[2];
} b2bvh_top_node;
/* destination rank of every primitive = number of splitters <= its code (equal codes never split); stable partition into send order:
 * d_sendCodes / d_sendGids (= first_gid + local index) / d_sendBoxes hold destination 0's primitives first, then 1's, ...; d_sendCounts[world] */
int b2bvh_global_partition(b2bvh_ctx* ctx, const uint32_t* d_codes, const b2bvh_aabb* d_boxes, uint32_t n, uint32_t first_gid,
                           const uint32_t* d_splitters /* world - 1, non-decreasing */, uint32_t world, uint32_t* d_sendCodes, uint32_t* d_sendGids,
                           b2bvh_aabb* d_sendBoxes, uint32_t* d_sendCounts);
/* stable sort of the received codes (pieces in source-rank order); d_perm[g] = received position of sorted leaf g; d_edge2 = {first, last} code */
int b2bvh_global_sort(b2bvh_ctx* ctx, const uint32_t* d_codes, uint32_t cnt, uint32_t* d_sortedCodes, uint32_t* d_perm, uint32_t* d_edge2);
/* hierarchy over the rank's range [first_pos, first_pos + cnt) of the global order, one ghost leaf per inner edge (prev_rank / next_rank: the
 * nearest non-empty neighbour whose edge code in d_allEdges[2 * rank + {0,1}] supplies it, or -1): d_nodesOut (2m-1 nodes, m = cnt + ghosts) =
 * ghost-free nodes with GLOBAL indices (internal i -> first_pos - ghostL + i; leaves name the global primitive d_gids[...]), artefacts as
 * {INVALID, INVALID, empty}; d_clusters (256) = left-over clusters in position order, pad = depth + 1 of the boundary on the right (0: none),
 * pad2[0] = 1 where valid; *d_clusterCount their number */
int b2bvh_global_tree(b2bvh_ctx* ctx, const uint32_t* d_sortedCodes, const uint32_t* d_perm, const uint32_t* d_gids, const b2bvh_aabb* d_boxes,
                      uint32_t cnt, const uint32_t* d_allEdges, int prev_rank, int next_rank, uint32_t first_pos, uint32_t n_total, int karras,
                      b2bvh_bvh2_node* d_nodesOut, b2bvh_cluster* d_clusters, uint32_t* d_clusterCount);
/* the nodes above the ranks from the all-gathered clusters (world x 256 records, world counts): d_topNodes (room for world * 256),
 * d_result3 = {number of top nodes, root node index, status (0 ok, 1 inconsistent depths, 2 a rank reported more than 256 clusters)} */
int b2bvh_global_top(b2bvh_ctx* ctx, const b2bvh_cluster* d_allClusters, const uint32_t* d_allCounts, uint32_t world, uint32_t n_total,
                     int karras, b2bvh_top_node* d_topNodes, uint32_t* d_result3);

/* the WHOLE tree on every rank: d_pieceNodes / d_pieceLeaves = every rank's b2bvh_global_tree output all-gathered, padded to max_nodes / max_leaves
 * records per rank; h_layout4 (HOST, world x 4) = {first global node index, node count, first global leaf position, leaf count} per rank; writes the
 * one-GPU node array d_fullNodes (2 n_total - 1, LBVH layout; artefact nodes skipped, the nodes above the ranks from d_top) and d_leafPrim (n_total) */
int b2bvh_global_assemble(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_pieceNodes, const b2bvh_bvh2_node* d_pieceLeaves, uint32_t world,
                          uint32_t max_nodes, uint32_t max_leaves, const uint32_t* h_layout4, const b2bvh_top_node* d_top, const uint32_t* d_result3,
                          uint32_t n_total, b2bvh_bvh2_node* d_fullNodes, uint32_t* d_leafPrim);
/* CollapseToWide4Bvh (TwoPassLbvhKernel.h:237-337) as an individually callable stage over a Bvh2 in the LBVH layout (2n-1 nodes): d_root = DEVICE
 * pointer of the root index, d_leafPrim[slot] = primitive of leaf slot; d_wide / d_wideLeaves have room for n records; synchronises once */
int b2bvh_collapse_bvh2(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const uint32_t* d_leafPrim, const uint32_t* d_root, uint32_t n,
                        b2bvh_bvh4_node* d_wide, b2bvh_prim_node* d_wideLeaves, uint32_t* n_wide);

/* ---- traversal: replaces the body of TwoPassLbvh::traverseBvh (src/TwoPassLbvh.cpp:199-311). ---- */
int b2bvh_generate_rays(b2bvh_ctx* ctx, const b2bvh_camera* cam, uint32_t width, uint32_t height, b2bvh_ray* d_rays,
                        float* ms);                       /* GenerateRays, CommonBlocksKernel.h:432-463 */
int b2bvh_traverse(b2bvh_ctx* ctx, const b2bvh_tree* tree, const b2bvh_ray* d_rays, uint32_t n_rays, const b2bvh_transform* xform,
                   int kernel, b2bvh_hit* d_hits /* may be NULL */, uint8_t* d_rgba /* may be NULL */, float* ms);

/* same, plus the per-ray count of triangle tests the reference's if-if / restart-trail kernels keep (rayCounter, u32 per ray,
 * TwoPassLbvh.cpp:224,267; NULL: not wanted; not available for the two while-while kernels, as in the reference) */
int b2bvh_traverse_ex(b2bvh_ctx* ctx, const b2bvh_tree* tree, const b2bvh_ray* d_rays, uint32_t n_rays, const b2bvh_transform* xform,
                      int kernel, b2bvh_hit* d_hits, uint8_t* d_rgba, uint32_t* d_rayCounter, float* ms);
/* Utility::generateTraversalHeatMap (Utility.cpp:424-454) on HOST buffers, without the PNG write: rgba = (c/max*150, c/max*255, 255, 255) */
int b2bvh_heat_map(const uint32_t* rayCounter, uint32_t count, uint8_t* rgba);

/* ---- sharded (multi-GPU) build helpers; new work specified by the north star, no reference counterpart.
 * rank-local steps only — the 6-float all-reduce and the root all-gather between them belong to the caller's
 * communicator (torch.distributed / NCCL). ---- */
int b2bvh_shard_extents(b2bvh_ctx* ctx, const b2bvh_triangle* tris, uint32_t n, uint32_t tris_on_device,
                        float* d_negmin_max6 /* device, 6 floats: {-min.xyz, max.xyz}, ready for one all-reduce(MAX) */);
int b2bvh_top_level(b2bvh_ctx* ctx, const b2bvh_aabb* d_rootBoxes, uint32_t g, b2bvh_bvh2_node* d_topNodes /* 2g-1 */);
/* The whole sharded build from ONE host thread over n_gpus contexts (one per device; contexts may also share a device): shard g =
 * tris[g][0 .. counts[g]) (host or device pointers per opts->tris_on_device), built in the frame of the union of all shards' boxes;
 * trees[g] = shard g's tree (as b2bvh_build fills it), *h_scene = the global scene box (may be NULL), h_topNodes (host, 2*n_gpus-1 nodes) =
 * the top-level tree over the shard roots (leaf m_leftChildIdx = shard).  The 24-byte exchanges go through the host; ranks in separate
 * processes use b2bvh_shard_extents / b2bvh_build / b2bvh_top_level with their own communicator instead (b2bvh/sharded.py). */
int b2bvh_build_sharded(b2bvh_ctx* const* ctxs, uint32_t n_gpus, int algo, const b2bvh_triangle* const* tris, const uint32_t* counts,
                        const b2bvh_build_opts* opts, b2bvh_tree* trees, b2bvh_aabb* h_scene, b2bvh_bvh2_node* h_topNodes);

/* ---- SAH-cost reporting on the host (pure functions on host copies; Utility.cpp:317-396). ---- */
float b2bvh_cost_bvh4(const b2bvh_bvh4_node* wide, const b2bvh_prim_node* wideLeaves, const b2bvh_aabb* primAabbs, uint32_t root,
                      uint32_t n_wide, uint32_t n_internal);                     /* Utility::calculatebvh4Cost */
float b2bvh_cost_lbvh(const b2bvh_bvh2_node* nodes, uint32_t root, uint32_t n_leaf, uint32_t n_internal); /* Utility::calculateLbvhCost */
/* convenience: D2H the wide tree and return calculatebvh4Cost, i.e. the builders' m_cost (TwoPassLbvh.cpp:185-196) */
int b2bvh_tree_cost(b2bvh_ctx* ctx, const b2bvh_tree* tree, float* cost);

/* ---- synthetic input of the benchmark configs (BASELINE.json configs 4/5; definition synth_uniform_v1 in SURVEY.md §8d):
 * triangles [first, first+count) of the stream with `seed`, written to DEVICE memory; RNG = the reference's tea<16>/lcg/randf
 * (CommonBlocksKernel.h:401-430).  `half` = 1000 * Ntotal^(-1/3) rounded to float by the caller. ---- */
int b2bvh_synth_uniform(b2bvh_ctx* ctx, uint64_t first, uint32_t count, uint32_t seed, float half, b2bvh_triangle* d_tris);
/* synth_clustered_v1 (optional second distribution): the same stream, but every triangle sits within +-10 of one of 4096
 * cluster centres (cluster = floor(first draw * 4096), centre from tea<16>(cluster, seed ^ 0xC1)): many primitives per Morton cell */
int b2bvh_synth_clustered(b2bvh_ctx* ctx, uint64_t first, uint32_t count, uint32_t seed, float half, b2bvh_triangle* d_tris);

/* ---- per-launch profiler: when enabled every kernel launch of the context is bracketed by a CUDA event pair; entries are
 * read after a sync.  Replaces Timer::measure (src/Timer.h:31-73) at kernel granularity, without its per-launch host sync. ---- */
int b2bvh_profile_enable(b2bvh_ctx* ctx, int on);  /* also clears the recorded entries */
int b2bvh_profile_count(b2bvh_ctx* ctx, int* count);
int b2bvh_profile_entry(b2bvh_ctx* ctx, int index, char* name, size_t cap, float* ms);

/* the layout of b2bvh_build_opts / b2bvh_tree / b2bvh_batch this header describes; a binding compares it with the library's answer
 * at load time and refuses a mismatch (the host classes and the Python mirror do) */
#define B2BVH_ABI_VERSION 6u
uint32_t b2bvh_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B2BVH_H */
