/*
 * global_build.cu — the rank-local steps of the globally sorted multi-GPU build (no reference
 * counterpart: the reference is single-device, Context.cpp:11), every one a kernel sequence behind the C ABI with no host round trip:
 *
 *   b2bvh_global_partition   destination of every primitive = the rank whose Morton interval holds its code (search in <= 255 splitters),
 *                            stable partition by destination (one 8-bit pass of the radix sort over the destinations), payload gathered into
 *                            send order: code, global id, 24-byte box (32 B per primitive on the wire); per-destination counts
 *   b2bvh_global_sort        the received pieces (source-rank order = global index order among equal codes) -> local stable sort;
 *                            the rank's first and last code for the neighbours' ghost leaves
 *   b2bvh_global_tree        keys widened to (code << 32 | GLOBAL sorted position) + one ghost leaf per inner edge -> the ordinary hierarchy
 *                            stage (lbvh_tile / group / climb over 64-bit keys) -> ghost-free nodes moved to global indices, the rank's
 *                            left-over clusters with the depth of the boundary on their right
 *   b2bvh_global_top         all ranks' left-over clusters (<= 256 each, all-gathered) -> the few nodes whose ranges straddle rank borders,
 *                            by the usual rule: two neighbours form a node when the boundary between them is deeper than both next to it
 *
 * The collectives in between (all-reduce of the scene box, all-gather of the sample, all-to-all of counts and payload, all-gather of edge
 * codes and clusters) belong to the caller's communicator (b2bvh/sharded.py GlobalBuildDevice over torch.distributed / NCCL).  The one
 * host synchronisation of a build is the count matrix of the all-to-all (NCCL takes send and receive sizes from the host).
 * The CPU restatement of the procedure lives with the test infrastructure (global_sort_by_exchange, lbvh_by_ranges, stitch_leftovers).
 */
#include "common.cuh"

#define GB_THREADS 256
#define GB_MAX_CLUSTERS 256u

/* ---- 1. partition ---- */
__global__ void __launch_bounds__(GB_THREADS) gb_dest_kernel(const u32* __restrict__ codes, u32 n, const u32* __restrict__ splitters, u32 world,
                                                             u32* __restrict__ dest) {
  __shared__ u32 sp[256];
  if (threadIdx.x < world - 1u) sp[threadIdx.x] = __ldg(splitters + threadIdx.x);
  __syncthreads();
  const u32 i = blockIdx.x * GB_THREADS + threadIdx.x;
  if (i >= n) return;
  const u32 c = __ldg(codes + i);
  /* number of splitters <= code: a code equal to a splitter goes right, so equal codes never split */
  u32 lo = 0, hi = world - 1u;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    if (sp[mid] <= c) lo = mid + 1u; else hi = mid;
  }
  dest[i] = lo;
}

__global__ void __launch_bounds__(GB_THREADS) gb_gather_kernel(const u32* __restrict__ order, const u32* __restrict__ codes,
                                                               const b2bvh_aabb* __restrict__ boxes, u32 n, u32 firstGid, u32* __restrict__ outCodes,
                                                               u32* __restrict__ outGids, b2bvh_aabb* __restrict__ outBoxes) {
  const u32 i = blockIdx.x * GB_THREADS + threadIdx.x;
  if (i >= n) return;
  const u32 o = __ldg(order + i); /* the partition is stable: every destination's run reads the local arrays front to back */
  outCodes[i] = __ldg(codes + o);
  outGids[i] = firstGid + o;
  const float2* s = reinterpret_cast<const float2*>(boxes + o);
  float2* d = reinterpret_cast<float2*>(outBoxes + i);
  const float2 a = __ldg(s), b = __ldg(s + 1), c = __ldg(s + 2);
  d[0] = a; d[1] = b; d[2] = c;
}

__global__ void gb_counts_kernel(const u32* __restrict__ sortedDest, u32 n, u32 world, u32* __restrict__ counts) {
  const u32 d = threadIdx.x;
  if (d >= world) return;
  auto lower = [&](u32 v) { /* first index with sortedDest[i] >= v */
    u32 lo = 0, hi = n;
    while (lo < hi) {
      const u32 mid = (lo + hi) >> 1;
      if (__ldg(sortedDest + mid) < v) lo = mid + 1u; else hi = mid;
    }
    return lo;
  };
  counts[d] = lower(d + 1u) - lower(d);
}

/* ---- 2. edges of the locally sorted range ---- */
__global__ void gb_edges_kernel(const u32* __restrict__ sorted, u32 cnt, u32* __restrict__ edge2) {
  if (threadIdx.x == 0) { edge2[0] = cnt ? sorted[0] : 0u; edge2[1] = cnt ? sorted[cnt - 1u] : 0u; }
}

/* ---- 3. widened keys with ghosts ---- */
__global__ void __launch_bounds__(GB_THREADS) gb_keys_kernel(const u32* __restrict__ sorted, const u32* __restrict__ perm, u32 cnt, u32 ghostL, u32 ghostR,
                                                             const u32* __restrict__ allEdges, u32 prevRank, u32 nextRank, u32 firstPos /* of local leaf 0 */,
                                                             u64* __restrict__ k64, u32* __restrict__ vals) {
  const u32 m = cnt + ghostL + ghostR;
  const u32 j = blockIdx.x * GB_THREADS + threadIdx.x;
  if (j >= m) return;
  u32 code, v = 0;
  if (ghostL && j == 0u) code = __ldg(allEdges + 2u * prevRank + 1u);       /* the left neighbour's last code */
  else if (ghostR && j == m - 1u) code = __ldg(allEdges + 2u * nextRank);    /* the right neighbour's first code */
  else { code = __ldg(sorted + j - ghostL); v = __ldg(perm + j - ghostL); }
  k64[j] = ((u64)code << 32) | (u64)(firstPos + j);
  vals[j] = v;
}

/* depth (+1; 0 = no boundary: the end of the global order) of the boundary on the right of every left-over cluster, from the widened keys */
__global__ void gb_cluster_depth_kernel(b2bvh_cluster* clusters, const u32* __restrict__ count, const u64* __restrict__ k64, u32 m, u32 firstPos) {
  const u32 c = threadIdx.x;
  if (c >= GB_MAX_CLUSTERS) return;
  if (c >= *count) { clusters[c].pad2[0] = 0u; return; }
  const u32 hiLocal = clusters[c].hi - firstPos;
  u32 d1 = 0;
  if (hiLocal < m) d1 = (u32)__clzll((long long)(k64[hiLocal - 1u] ^ k64[hiLocal])) + 1u;
  clusters[c].pad = d1;
  clusters[c].pad2[0] = 1u; /* valid */
}

/* ---- 4. the top of the tree over all ranks' left-over clusters: one thread (a few hundred clusters) ---- */
template <bool KARRAS>
__global__ void gb_top_kernel(const b2bvh_cluster* __restrict__ all, const u32* __restrict__ counts, u32 world, u32 nTotal, b2bvh_cluster* work /* world*256 x 2 */,
                              b2bvh_top_node* __restrict__ top, u32* __restrict__ result /* [0] top-node count, [1] root index, [2] status */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  b2bvh_cluster* cur = work;
  b2bvh_cluster* nxt = work + (size_t)world * GB_MAX_CLUSTERS;
  u32 len = 0;
  for (u32 r = 0; r < world; r++) {
    const u32 c = counts[r];
    if (c > GB_MAX_CLUSTERS) { result[0] = 0; result[1] = B2_INVALID; result[2] = 2u; return; }
    for (u32 k = 0; k < c; k++) cur[len++] = all[(size_t)r * GB_MAX_CLUSTERS + k];
  }
  u32 nTop = 0, root = len == 1u ? cur[0].node : B2_INVALID;
  while (len > 1u) {
    u32 out = 0, i = 0;
    bool merged = false;
    while (i < len) {
      if (i + 1u < len) {
        const b2bvh_cluster &a = cur[i], &b = cur[i + 1u];
        const u32 dl = i > 0u ? cur[i - 1u].pad : 0u, d0 = a.pad, dr = b.pad;
        if (d0 > dl && d0 > dr) {
          const u32 lo = a.lo, mid = a.hi, hi = b.hi;
          const bool isRoot = lo == 0u && hi == nTotal;
          const u32 nid = KARRAS ? (isRoot ? 0u : (dr > dl ? hi - 1u : lo)) : mid - 1u;
          b2bvh_cluster mclu;
          mclu.lo = lo; mclu.hi = hi; mclu.node = nid; mclu.pad = dr; mclu.pad2[0] = 1u; mclu.pad2[1] = 0u;
          mclu.box.m_min.x = fminf(a.box.m_min.x, b.box.m_min.x); mclu.box.m_min.y = fminf(a.box.m_min.y, b.box.m_min.y); mclu.box.m_min.z = fminf(a.box.m_min.z, b.box.m_min.z);
          mclu.box.m_max.x = fmaxf(a.box.m_max.x, b.box.m_max.x); mclu.box.m_max.y = fmaxf(a.box.m_max.y, b.box.m_max.y); mclu.box.m_max.z = fmaxf(a.box.m_max.z, b.box.m_max.z);
          b2bvh_top_node t;
          t.index = nid; t.left = a.node; t.right = b.node; t.pad = 0u; t.box = mclu.box; t.pad2[0] = t.pad2[1] = 0u;
          top[nTop++] = t;
          if (isRoot) root = nid;
          nxt[out++] = mclu;
          i += 2u;
          merged = true;
          continue;
        }
      }
      nxt[out++] = cur[i];
      i++;
    }
    if (!merged) { result[0] = nTop; result[1] = B2_INVALID; result[2] = 1u; return; } /* inconsistent boundary depths: report, never spin */
    b2bvh_cluster* t = cur; cur = nxt; nxt = t;
    len = out;
  }
  result[0] = nTop; result[1] = root; result[2] = 0u;
}

extern "C" {

int b2bvh_global_partition(b2bvh_ctx* ctx, const uint32_t* d_codes, const b2bvh_aabb* d_boxes, uint32_t n, uint32_t first_gid, const uint32_t* d_splitters,
                           uint32_t world, uint32_t* d_sendCodes, uint32_t* d_sendGids, b2bvh_aabb* d_sendBoxes, uint32_t* d_sendCounts) {
  if (!ctx || !d_sendCounts || world == 0 || world > 256u) return b2_fail(B2BVH_ERR_INVALID, "global_partition: bad argument (1 <= world <= 256)");
  if (n && (!d_codes || !d_boxes || !d_sendCodes || !d_sendGids || !d_sendBoxes || (world > 1 && !d_splitters)))
    return b2_fail(B2BVH_ERR_INVALID, "global_partition: null argument");
  if (n > 0x3FFFFFFFu) return b2_fail(B2BVH_ERR_INVALID, "global_partition: n=%u exceeds 2^30-1", n);
  B2_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) { B2_CUDA(cudaMemsetAsync(d_sendCounts, 0, world * sizeof(u32), ctx->stream)); return 0; }
  void *dDest, *dSorted, *dOrder, *tk, *tv, *sc;
  B2_TRY(b2_reserve(ctx, SLOT_KEYS, (size_t)n * 4, &dDest));
  B2_TRY(b2_reserve(ctx, SLOT_SKEYS, (size_t)n * 4, &dSorted));
  B2_TRY(b2_reserve(ctx, SLOT_SVALS, (size_t)n * 4, &dOrder));
  B2_TRY(b2_reserve(ctx, SLOT_TKEYS, (size_t)n * 4, &tk));
  B2_TRY(b2_reserve(ctx, SLOT_TVALS, (size_t)n * 4, &tv));
  B2_TRY(b2_reserve(ctx, SLOT_SORT, b2_sort_scratch_bytes(n), &sc));
  const u32 grid = (n + GB_THREADS - 1) / GB_THREADS;
  B2_KERNEL(ctx, "global_dest");
  gb_dest_kernel<<<grid, GB_THREADS, 0, ctx->stream>>>(d_codes, n, d_splitters, world, (u32*)dDest);
  B2_LAUNCH_CHECK(ctx);
  /* stable partition by destination = one pass of the stable radix sort over the destination byte, values = local index */
  B2_TRY(b2_launch_sort(ctx, (const u32*)dDest, nullptr, (u32*)dSorted, (u32*)dOrder, (u32*)tk, (u32*)tv, sc, n, 0, 8));
  B2_KERNEL(ctx, "global_gather");
  gb_gather_kernel<<<grid, GB_THREADS, 0, ctx->stream>>>((const u32*)dOrder, d_codes, d_boxes, n, first_gid, d_sendCodes, d_sendGids, d_sendBoxes);
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "global_counts");
  gb_counts_kernel<<<1, 256, 0, ctx->stream>>>((const u32*)dSorted, n, world, d_sendCounts);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2bvh_global_sort(b2bvh_ctx* ctx, const uint32_t* d_codes, uint32_t cnt, uint32_t* d_sortedCodes, uint32_t* d_perm, uint32_t* d_edge2) {
  if (!ctx || !d_edge2 || (cnt && (!d_codes || !d_sortedCodes || !d_perm))) return b2_fail(B2BVH_ERR_INVALID, "global_sort: null argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  if (cnt) {
    if (((uintptr_t)d_codes | (uintptr_t)d_sortedCodes | (uintptr_t)d_perm) & 15) return b2_fail(B2BVH_ERR_INVALID, "global_sort: arrays must be 16-byte aligned");
    void *tk, *tv, *sc;
    B2_TRY(b2_reserve(ctx, SLOT_TKEYS, (size_t)cnt * 4, &tk));
    B2_TRY(b2_reserve(ctx, SLOT_TVALS, (size_t)cnt * 4, &tv));
    B2_TRY(b2_reserve(ctx, SLOT_SORT, b2_sort_scratch_bytes(cnt), &sc));
    B2_TRY(b2_launch_sort(ctx, d_codes, nullptr, d_sortedCodes, d_perm, (u32*)tk, (u32*)tv, sc, cnt, 0, 32));
  }
  B2_KERNEL(ctx, "global_edges");
  gb_edges_kernel<<<1, 32, 0, ctx->stream>>>(d_sortedCodes, cnt, d_edge2);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2bvh_global_tree(b2bvh_ctx* ctx, const uint32_t* d_sortedCodes, const uint32_t* d_perm, const uint32_t* d_gids, const b2bvh_aabb* d_boxes, uint32_t cnt,
                      const uint32_t* d_allEdges, int prev_rank, int next_rank, uint32_t first_pos, uint32_t n_total, int karras, b2bvh_bvh2_node* d_nodesOut,
                      b2bvh_cluster* d_clusters, uint32_t* d_clusterCount) {
  if (!ctx || !d_clusters || !d_clusterCount) return b2_fail(B2BVH_ERR_INVALID, "global_tree: null argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const u32 ghostL = prev_rank >= 0 ? 1u : 0u, ghostR = next_rank >= 0 ? 1u : 0u;
  const u32 m = cnt ? cnt + ghostL + ghostR : 0u;
  if (cnt == 0) { /* an empty rank contributes nothing */
    B2_CUDA(cudaMemsetAsync(d_clusterCount, 0, sizeof(u32), s));
    B2_CUDA(cudaMemsetAsync(d_clusters, 0, GB_MAX_CLUSTERS * sizeof(b2bvh_cluster), s));
    return 0;
  }
  if (!d_sortedCodes || !d_perm || !d_gids || !d_boxes || !d_nodesOut || ((ghostL || ghostR) && !d_allEdges)) return b2_fail(B2BVH_ERR_INVALID, "global_tree: null argument");
  if (m < 2) return b2_fail(B2BVH_ERR_INVALID, "global_tree: one primitive in total (a rank with a single leaf needs a neighbour)");
  if ((uint64_t)first_pos + cnt > n_total) return b2_fail(B2BVH_ERR_INVALID, "global_tree: range [%u, %u + %u) does not fit %u positions", first_pos, first_pos, cnt, n_total);
  const u32 a2 = first_pos - ghostL; /* global position of local leaf 0 (the left ghost, when there is one) */
  void *dK64, *dVals, *dLocal, *dParents = nullptr, *dLbvh, *dFlags, *dLeafPrim;
  B2_TRY(b2_reserve(ctx, SLOT_KEYS64, (size_t)m * 8, &dK64));
  B2_TRY(b2_reserve(ctx, SLOT_M60_VALS, (size_t)m * 4, &dVals));
  B2_TRY(b2_reserve(ctx, SLOT_NODES, (2 * (size_t)m - 1) * sizeof(b2bvh_bvh2_node), &dLocal));
  if (karras) B2_TRY(b2_reserve(ctx, SLOT_PARENTS, (2 * (size_t)m - 1) * 4, &dParents));
  B2_TRY(b2_reserve(ctx, SLOT_LBVH, b2_lbvh_scratch_bytes(m), &dLbvh));
  B2_TRY(b2_reserve(ctx, SLOT_MISC, (size_t)m + 64, &dFlags));
  B2_TRY(b2_reserve(ctx, SLOT_SPLIT_LEAFPRIM, (size_t)m * 4, &dLeafPrim));
  u32* dRoot = (u32*)((unsigned char*)ctx->bufs[SLOT_CTL].p + 96);
  B2_KERNEL(ctx, "global_keys");
  gb_keys_kernel<<<(m + GB_THREADS - 1) / GB_THREADS, GB_THREADS, 0, s>>>(d_sortedCodes, d_perm, cnt, ghostL, ghostR, d_allEdges, (u32)(prev_rank < 0 ? 0 : prev_rank),
                                                                       (u32)(next_rank < 0 ? 0 : next_rank), a2, (u64*)dK64, (u32*)dVals);
  B2_LAUNCH_CHECK(ctx);
  /* the leaves name the GLOBAL primitive: the hierarchy kernels translate the received position through d_gids (the same hook the
   * early-split references use, InitBvhNodesPrimRef) */
  ctx->ref_prim = d_gids;
  ctx->ref_leaf_prim = (u32*)dLeafPrim;
  ctx->lbvh_second_level = 0;
  const int st = b2_launch_lbvh_fused64(ctx, (const u64*)dK64, (const u32*)dVals, d_boxes, m, (b2bvh_bvh2_node*)dLocal, karras ? (u32*)dParents : nullptr, (u32*)dLbvh,
                                        dRoot, karras ? 1 : 0);
  ctx->ref_prim = nullptr;
  ctx->ref_leaf_prim = nullptr;
  B2_TRY(st);
  if (karras) B2_CUDA(cudaMemsetAsync(dRoot, 0, 4, s));
  B2_TRY(b2_launch_range_extract(ctx, (const b2bvh_bvh2_node*)dLocal, m, dRoot, karras, ghostL, ghostR, a2, n_total, (unsigned char*)dFlags, d_nodesOut, d_clusters,
                                 d_clusterCount));
  B2_KERNEL(ctx, "global_cluster_depth");
  gb_cluster_depth_kernel<<<1, GB_MAX_CLUSTERS, 0, s>>>(d_clusters, d_clusterCount, (const u64*)dK64, m, a2);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

int b2bvh_global_top(b2bvh_ctx* ctx, const b2bvh_cluster* d_allClusters, const uint32_t* d_allCounts, uint32_t world, uint32_t n_total, int karras,
                     b2bvh_top_node* d_topNodes, uint32_t* d_result3) {
  if (!ctx || !d_allClusters || !d_allCounts || !d_topNodes || !d_result3 || world == 0 || world > 256u) return b2_fail(B2BVH_ERR_INVALID, "global_top: bad argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  void* work;
  B2_TRY(b2_reserve(ctx, SLOT_COLLAPSE, 2 * (size_t)world * GB_MAX_CLUSTERS * sizeof(b2bvh_cluster), &work));
  B2_KERNEL(ctx, "global_top");
  if (karras) gb_top_kernel<true><<<1, 32, 0, ctx->stream>>>(d_allClusters, d_allCounts, world, n_total, (b2bvh_cluster*)work, d_topNodes, d_result3);
  else gb_top_kernel<false><<<1, 32, 0, ctx->stream>>>(d_allClusters, d_allCounts, world, n_total, (b2bvh_cluster*)work, d_topNodes, d_result3);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

} /* extern "C" */

/* ---- 5. the whole tree on every rank (for the 4-wide collapse, whose breadth-first numbering is a property of the WHOLE tree): the ranks
 * all-gather their pieces (padded to the longest) and every rank assembles the one-GPU node array, then runs the ordinary collapse over it ---- */
__global__ void __launch_bounds__(GB_THREADS) gb_assemble_kernel(const b2bvh_bvh2_node* __restrict__ pieceNodes, const b2bvh_bvh2_node* __restrict__ pieceLeaves,
                                                                 u32 world, u32 maxNodes, u32 maxLeaves, const u32* __restrict__ layout /* world x 4: nodeFirst,
                                                                 nodeCount, leafFirst, leafCount */, u32 nTotal, b2bvh_bvh2_node* __restrict__ full,
                                                                 u32* __restrict__ leafPrim) {
  const u32 r = blockIdx.y;
  const u32 nodeFirst = layout[4 * r], nodeCount = layout[4 * r + 1], leafFirst = layout[4 * r + 2], leafCount = layout[4 * r + 3];
  const u32 nInt = nTotal - 1u;
  for (u32 i = blockIdx.x * GB_THREADS + threadIdx.x; i < nodeCount + leafCount; i += gridDim.x * GB_THREADS) {
    if (i < nodeCount) {
      const Node2 nd = load_node2_ro(pieceNodes + (size_t)r * maxNodes + i);
      if (nd.left != B2_INVALID) store_node2(full + nodeFirst + i, nd.left, nd.right, nd.box); /* artefact (ghost spine) nodes are skipped */
    } else {
      const u32 j = i - nodeCount;
      const Node2 nd = load_node2_ro(pieceLeaves + (size_t)r * maxLeaves + j);
      store_node2(full + nInt + leafFirst + j, nd.left, nd.right, nd.box);
      leafPrim[leafFirst + j] = nd.left; /* leaf slot -> (global) primitive: what the collapse's PrimNode records name */
    }
  }
}
__global__ void gb_assemble_top_kernel(const b2bvh_top_node* __restrict__ top, const u32* __restrict__ result3, b2bvh_bvh2_node* __restrict__ full) {
  const u32 nTop = result3[0];
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < nTop; i += gridDim.x * blockDim.x) {
    const b2bvh_top_node t = top[i];
    const Box b = Box{t.box.m_min.x, t.box.m_min.y, t.box.m_min.z, t.box.m_max.x, t.box.m_max.y, t.box.m_max.z};
    store_node2(full + t.index, t.left, t.right, b);
  }
}

extern "C" {

int b2bvh_global_assemble(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_pieceNodes, const b2bvh_bvh2_node* d_pieceLeaves, uint32_t world, uint32_t max_nodes,
                          uint32_t max_leaves, const uint32_t* h_layout4, const b2bvh_top_node* d_top, const uint32_t* d_result3, uint32_t n_total,
                          b2bvh_bvh2_node* d_fullNodes, uint32_t* d_leafPrim) {
  if (!ctx || !d_pieceNodes || !d_pieceLeaves || !h_layout4 || !d_top || !d_result3 || !d_fullNodes || !d_leafPrim || world == 0 || world > 256u || n_total < 2)
    return b2_fail(B2BVH_ERR_INVALID, "global_assemble: bad argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  void* dLayout;
  B2_TRY(b2_reserve(ctx, SLOT_MISC, 4096 + 64, &dLayout));
  B2_CUDA(cudaMemcpyAsync(dLayout, h_layout4, (size_t)world * 16, cudaMemcpyHostToDevice, ctx->stream)); /* pageable source: staged before the call returns */
  u32 gx = (max_nodes + max_leaves + GB_THREADS - 1) / GB_THREADS;
  const u32 cap = (u32)ctx->sm_count * 8u;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  B2_KERNEL(ctx, "global_assemble");
  gb_assemble_kernel<<<dim3(gx, world), GB_THREADS, 0, ctx->stream>>>(d_pieceNodes, d_pieceLeaves, world, max_nodes, max_leaves, (const u32*)dLayout, n_total, d_fullNodes,
                                                                      d_leafPrim);
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "global_assemble_top");
  gb_assemble_top_kernel<<<8, 256, 0, ctx->stream>>>(d_top, d_result3, d_fullNodes);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}

/* The collapse stage alone over a Bvh2 in the LBVH layout (2n-1 nodes, leaves at n-1 + slot): CollapseToWide4Bvh (TwoPassLbvhKernel.h:237-337) as an
 * individually callable stage.  d_root: DEVICE pointer of the root index (the globally sorted build leaves it in d_result3[1]); d_leafPrim[slot] =
 * primitive of leaf slot.  Synchronises once, for *n_wide. */
int b2bvh_collapse_bvh2(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const uint32_t* d_leafPrim, const uint32_t* d_root, uint32_t n, b2bvh_bvh4_node* d_wide,
                        b2bvh_prim_node* d_wideLeaves, uint32_t* n_wide) {
  if (!ctx || !d_nodes || !d_leafPrim || !d_root || !d_wide || !d_wideLeaves || !n_wide || n < 2 || n > 0x3FFFFFFFu) return b2_fail(B2BVH_ERR_INVALID, "collapse_bvh2: bad argument");
  B2_CUDA(cudaSetDevice(ctx->device));
  void* scratch;
  B2_TRY(b2_reserve(ctx, SLOT_COLLAPSE, b2_collapse_scratch_bytes(n), &scratch));
  B2_TRY(b2_launch_collapse(ctx, d_nodes, nullptr, d_leafPrim, d_root, n, d_wide, d_wideLeaves, scratch, nullptr));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_wide = b2_mailbox(ctx, B2_MB_COLLAPSE)[1];
  if (*n_wide == B2_INVALID) { *n_wide = 0; return b2_fail(B2BVH_ERR_INTERNAL, "collapse_bvh2: the nodes do not form a tree"); }
  return 0;
}

} /* extern "C" */
