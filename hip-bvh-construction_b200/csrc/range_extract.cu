/*
 * range_extract.cu — second building block of the globally sorted multi-GPU build (no reference counterpart).
 *
 * A rank has run the ordinary hierarchy stage (b2bvh_lbvh_from_sorted64) over its range of the GLOBAL sorted order, extended by one
 * ghost leaf per inner edge, with keys widened to (code << 32 | global position).  Of that local tree
 *   - every node that contains no ghost IS a node of the one-GPU tree: it keeps its box and only needs its indices moved from the
 *     local to the global numbering (internal i -> first_pos + i, leaf slot g -> (n_global - 1) + first_pos + g);
 *   - the nodes that contain a ghost — the spine from the root down to the first leaf and / or to the last leaf — are artefacts;
 *   - the ghost-free children hanging off those spines are the rank's LEFT-OVER clusters: their parents straddle a rank boundary and
 *     are formed after the ranks have exchanged these few records.
 * range_spine_kernel walks the (at most two) spines with one thread — a node's range follows from its index in O(1) in both
 * numberings (Apetrei: the index is the split position; Karras: a left child carries the last leaf of its range, a right child the
 * first) — flags the artefacts and emits the left-over clusters in position order.  range_renumber_kernel streams over the nodes.
 */
#include "common.cuh"

#define RX_MAX_CLUSTERS 256u

/* first position of the right child's range = split of internal node `idx` whose left child is `l` (local indices, m leaves) */
template <bool KARRAS>
__device__ __forceinline__ u32 rx_split(u32 idx, u32 l, u32 m) {
  if (!KARRAS) return idx + 1u;                 /* Apetrei: the node sits at its split position */
  return (l >= m - 1u ? l - (m - 1u) : l) + 1u; /* Karras: a left child is named after the LAST leaf of its range (a leaf: its slot) */
}

template <bool KARRAS>
__global__ void range_spine_kernel(const b2bvh_bvh2_node* __restrict__ loc, u32 m, const u32* __restrict__ rootPtr, u32 ghostL, u32 ghostR, u32 firstPos,
                                   u32 nGlobal, unsigned char* __restrict__ artefact, b2bvh_cluster* __restrict__ out, u32* __restrict__ count) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const u32 root = *rootPtr; /* left on the device by the hierarchy stage: no host round trip between the two */
  const u32 nIntG = nGlobal - 1u;
  u32 n = 0;
  /* iterative in-order walk over the ghost-containing nodes only; everything else is emitted whole */
  u32 stIdx[256], stLo[256], stHi[256];
  int top = 0;
  stIdx[0] = root; stLo[0] = 0; stHi[0] = m; top = 1;
  while (top > 0) {
    --top;
    const u32 idx = stIdx[top], lo = stLo[top], hi = stHi[top];
    const bool hasGhost = (ghostL && lo == 0u) || (ghostR && hi == m);
    const bool leaf = idx >= m - 1u;
    if (!hasGhost) {
      if (n < RX_MAX_CLUSTERS) {
        b2bvh_cluster c;
        c.lo = firstPos + lo; c.hi = firstPos + hi;
        c.node = leaf ? nIntG + firstPos + (idx - (m - 1u)) : firstPos + idx;
        c.pad = 0; c.pad2[0] = c.pad2[1] = 0;
        c.box = loc[idx].m_aabb;
        out[n] = c;
      }
      n++;
      continue;
    }
    if (leaf) continue; /* a ghost leaf */
    artefact[idx] = 1;
    const u32 l = loc[idx].m_leftChildIdx, r = loc[idx].m_rightChildIdx;
    const u32 split = rx_split<KARRAS>(idx, l, m);
    if (top + 2 > 256) { n = 0xFFFFFFFFu; break; } /* deeper than any radix tree over 96-bit keys: report, never overrun */
    /* right first so that the left subtree is handled next: clusters come out in position order */
    stIdx[top] = r; stLo[top] = split; stHi[top] = hi; top++;
    stIdx[top] = l; stLo[top] = lo; stHi[top] = split; top++;
  }
  *count = n;
}

__global__ void __launch_bounds__(256) range_renumber_kernel(const b2bvh_bvh2_node* __restrict__ loc, u32 m, u32 firstPos, u32 nGlobal,
                                                             const unsigned char* __restrict__ artefact, b2bvh_bvh2_node* __restrict__ out) {
  const u32 i = blockIdx.x * 256u + threadIdx.x;
  if (i >= 2u * m - 1u) return;
  const u32 nIntG = nGlobal - 1u;
  Node2 nd = load_node2_ro(loc + i);
  if (i < m - 1u) {
    if (artefact[i]) { nd.left = B2_INVALID; nd.right = B2_INVALID; nd.box = box_empty(); }
    else {
      nd.left = nd.left >= m - 1u ? nIntG + firstPos + (nd.left - (m - 1u)) : firstPos + nd.left;
      nd.right = nd.right >= m - 1u ? nIntG + firstPos + (nd.right - (m - 1u)) : firstPos + nd.right;
    }
  }
  store_node2(out + i, nd.left, nd.right, nd.box); /* leaves: unchanged (m_leftChildIdx is the primitive) */
}

int b2_launch_range_extract(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_local, u32 m, const u32* d_root, int karras, u32 ghostL, u32 ghostR, u32 firstPos, u32 nGlobal,
                            unsigned char* d_flags, b2bvh_bvh2_node* d_out, b2bvh_cluster* d_clusters, u32* d_count) {
  B2_CUDA(cudaMemsetAsync(d_flags, 0, m, ctx->stream));
  B2_KERNEL(ctx, "range_spine");
  if (karras) range_spine_kernel<true><<<1, 32, 0, ctx->stream>>>(d_local, m, d_root, ghostL, ghostR, firstPos, nGlobal, d_flags, d_clusters, d_count);
  else range_spine_kernel<false><<<1, 32, 0, ctx->stream>>>(d_local, m, d_root, ghostL, ghostR, firstPos, nGlobal, d_flags, d_clusters, d_count);
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "range_renumber");
  range_renumber_kernel<<<(2 * m - 1 + 255) / 256, 256, 0, ctx->stream>>>(d_local, m, firstPos, nGlobal, d_flags, d_out);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
