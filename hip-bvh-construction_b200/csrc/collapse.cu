/*
 * collapse.cu — stage S5: Bvh2 -> 4-wide Bvh (Bvh4Node[] + PrimNode[]).
 *
 * Replaces CollapseToWide4Bvh (TwoPassLbvhKernel.h:237-337 for the LBVH layout, Ploc++Kernel.h:364-465 for the
 * separate-leaf layout; host setup TwoPassLbvh.cpp:154-183).  The reference runs ONE persistent launch whose threads
 * spin on a task queue and allocate wide nodes with a global atomicAdd: node numbering is timing dependent.  Here the
 * wide tree is numbered breadth-first — what a sequential execution of the reference's task loop produces — so the output
 * is deterministic and comparable with memcmp (canonical numbering of the oracle).  Three launches:
 *
 *   1. collapse_expand_kernel      for EVERY internal Bvh2 node, the (up to 4) children it would have as a wide node: twice,
 *                                  the internal child with the largest area is replaced by its two children (strict '>' so
 *                                  the first of equals wins, areas without FMA).  One thread per node, no synchronisation,
 *                                  16 B out.  Keeps the dependent node reads out of the level-synchronous part.
 *   2. collapse_number_kernel      one cooperative launch, every CTA resident.  One task = one wide node = one thread; a task is
 *                                  (Bvh2 node, wide parent).  The tasks of a level are a contiguous index range.  After a grid barrier
 *                                  CTA c takes chunk c of it; from then on a CTA keeps the children it numbered itself (its part of
 *                                  the next level is contiguous and in chunk order: same numbering) until the parts drift apart.
 *                                  Pass A copies the expansion of every task's node into the task's record (independent gathers, four
 *                                  tiles in flight) and counts the internal children; the CTA posts the count, counts itself in on an
 *                                  arrival counter (one polling thread per CTA) and sums the counts of all parts — all due at the
 *                                  same moment: no chain of dependent look-backs, ONE grid-wide step per level.  Pass C numbers the
 *                                  children consecutively in (task, slot) order, four tiles per barrier pair, appends their tasks
 *                                  and notes each task's first child index.  Every CTA sees the same counts, so all decide alike
 *                                  when to re-cut the level into equal chunks (one grid barrier).  Runs of levels of at most 1024
 *                                  tasks (the top of the tree, the tail of a deep one) are processed by CTA 0 alone between two
 *                                  barriers; the very first levels hand their nodes from level to level through shared memory.
 *   3. collapse_emit_kernel        one thread per wide node, no synchronisation: boxes of the internal children gathered
 *                                  with L2::64B loads, the 128-byte node written through a swizzled shared-memory transpose
 *                                  as full contiguous lines, PrimNode records for the leaf children.
 *
 * Traffic per primitive (10 M uniform): expansion 32 r + 16 w; ~0.47 wide nodes x (numbering: 8 task r + 16 expansion gathered
 * + 16 record w + 16 record r + 12 task/first-child w; emit: 24 record r + 32 child box gathered + 128 node w) + 8 PrimNode
 * + 4 sorted value ~ 180 B, of which ~133 B are compulsory; a gather moves 64 B of DRAM whatever it asks for.
 */
#include "common.cuh"
#ifdef COL_TRACE
#include <stdio.h>
#endif

#define COL_THREADS 256
#ifndef NUM_SOLO
#define NUM_SOLO 1024    /* levels of at most this many tasks are numbered by CTA 0 alone (a grid-wide level costs ~7 us of synchronisation) */
#endif
/* numbering kernel: CTA size by input size (measured, 10 M / bunny / sponza in us: 128 threads x 12 CTAs per SM 325 / 52 / 78,
 * 256 x 6 247 / 50 / 89, 512 x 2 210 / 60 / 121, 1024 x 1 234 / 85 / 129): large levels want few arrivals per grid-wide step,
 * the deep thin trees of small scenes want a short tile loop */
#ifndef NUM_OWN_SLACK_DIV
#define NUM_OWN_SLACK_DIV 4   /* a CTA keeps its own children while the largest part is below mean x (1 + 1/4) + 2 tiles */
#endif
#ifndef NUM_OWN_SLACK
#define NUM_OWN_SLACK 2
#endif
#define NUM_THREADS_LARGE 512
#define NUM_THREADS_SMALL 128

/* scratch layout: CollapseCtrl (256 B) | uint4 expansion[n] | uint4 taskCh[n] | u32 taskParent[n] | u32 firstChild[n] | u32 taskNode[n] | u64 counts[2][G] */
struct CollapseCtrl {
  u32 bar;        /* grid barrier: arrivals so far */
  u32 nWide;      /* result: number of wide nodes */
  u32 arrive;     /* chunks that have posted their count, all levels so far */
  u32 error;      /* set when a level would hold more tasks than there are internal nodes: the input is not a tree (cannot happen after a correct
                     hierarchy stage; the numbering stops and the build fails instead of running away over a cyclic input) */
  uint4 next[2];  /* {level, start, end, -}: the level to process after barrier b is published in next[b & 1] (a CTA that
                     leaves barrier b early may publish the level after it before a late CTA has read this one) */
};

size_t b2_collapse_scratch_bytes(u32 n) {
  /* two buffers (consecutive levels) of one count word per CTA, G <= 8 CTAs x 1024 SMs */
  return 256 + (size_t)n * (2 * sizeof(uint4) + 3 * sizeof(u32)) + 16 + 16384 * sizeof(u64);
}

template <int NUM_THREADS>
struct ColSmem {
  u32 warpSum[3][NUM_THREADS / 32];
  u32 warpMax[NUM_THREADS / 32];
  u32 tileSum[4][NUM_THREADS / 32];
  u32 topNode[2][NUM_THREADS];   /* number_top: Bvh2 node of every task of the level, and of the next one */
};

__device__ __forceinline__ void publish_level(CollapseCtrl* ctrl, u32 barrier, u32 level, u32 start, u32 end) {
  u32* nx = reinterpret_cast<u32*>(&ctrl->next[barrier & 1u]);
  st_relaxed(nx, level); st_relaxed(nx + 1, start); st_relaxed(nx + 2, end);
}

/* ---- 1. expansion of every internal node: CollapseToWide4Bvh, TwoPassLbvhKernel.h:262-296 — two rounds of "replace the
 * internal child with the largest area by its two children" (the new right child is appended) ---- */
#ifndef COL_EXPAND_MINB
#define COL_EXPAND_MINB 1
#endif
__global__ void __launch_bounds__(COL_THREADS, COL_EXPAND_MINB) collapse_expand_kernel(const b2bvh_bvh2_node* __restrict__ nodes, u32 nInt, uint4* __restrict__ expansion,
                                                                                         uint4* __restrict__ ctrlWords) {
  const u32 i = blockIdx.x * COL_THREADS + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x < 256 / 16) ctrlWords[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u); /* CollapseCtrl of the numbering kernel that follows (a memset node costs ~2 us of the stream) */
  if (i >= nInt) return;
  const uint2 top = __ldg(reinterpret_cast<const uint2*>(nodes + i));
  u32 ch[4] = {top.x, top.y, B2_INVALID, B2_INVALID};
  /* first expansion: the internal child with the larger area (left wins ties, a leaf never wins) */
  Node2 c0, c1;
  float a0 = 0.0f, a1 = 0.0f;
  if (ch[0] < nInt) { c0 = load_node2_ro(nodes + ch[0]); a0 = box_area(c0.box); }
  if (ch[1] < nInt) { c1 = load_node2_ro(nodes + ch[1]); a1 = box_area(c1.box); }
  float best = 0.0f;
  int pos = -1;
  if (ch[0] < nInt && a0 > best) { best = a0; pos = 0; }
  if (ch[1] < nInt && a1 > best) { best = a1; pos = 1; }
  if (pos >= 0) {
    const Node2 e = pos == 0 ? c0 : c1;
    /* slots now: [pos] = e.left, [2] = e.right, [1 - pos] unchanged */
    float ar[3];
    u32 id3[3];
    id3[pos] = e.left; id3[1 - pos] = ch[1 - pos]; id3[2] = e.right;
    ar[1 - pos] = pos == 0 ? a1 : a0;
    Node2 g0, g1; /* the two new children: their boxes are needed for the second choice, their children for the second expansion */
    ar[pos] = 0.0f; ar[2] = 0.0f;
    if (e.left < nInt) { g0 = load_node2_ro(nodes + e.left); ar[pos] = box_area(g0.box); }
    if (e.right < nInt) { g1 = load_node2_ro(nodes + e.right); ar[2] = box_area(g1.box); }
    ch[0] = id3[0]; ch[1] = id3[1]; ch[2] = id3[2];
    /* second expansion */
    best = 0.0f;
    int pos2 = -1;
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (id3[k] < nInt && ar[k] > best) { best = ar[k]; pos2 = k; }
    if (pos2 >= 0) {
      u32 l2, r2;
      if (pos2 == pos) { l2 = g0.left; r2 = g0.right; }
      else if (pos2 == 2) { l2 = g1.left; r2 = g1.right; }
      else { const Node2 o = pos == 0 ? c1 : c0; l2 = o.left; r2 = o.right; }
#pragma unroll
      for (int k = 0; k < 3; k++) if (k == pos2) ch[k] = l2;
      ch[3] = r2;
    }
  }
  expansion[i] = make_uint4(ch[0], ch[1], ch[2], ch[3]);
}

/* ---- 2. numbering ---- */
__device__ __forceinline__ u32 count_internal(const uint4& t, u32 nInt) {
  return (t.x < nInt ? 1u : 0u) + (t.y < nInt ? 1u : 0u) + (t.z < nInt ? 1u : 0u) + (t.w < nInt ? 1u : 0u);
}

/* A. tasks [a, b): fetch the expansion of each task's Bvh2 node into its record and count the internal children.  The
 * iterations are independent, so the gathers of several tiles are in flight together (four per thread). */
template <int NUM_THREADS>
__device__ __forceinline__ u32 number_fetch(const uint4* __restrict__ expansion, u32 nInt, const u32* taskNode, uint4* taskCh, u32 a, u32 b) {
  u32 cnt = 0;
  u32 g = a + threadIdx.x;
  for (; g + 3 * NUM_THREADS < b; g += 4 * NUM_THREADS) {
    u32 nd[4];
    uint4 ex[4];
#pragma unroll
    for (int k = 0; k < 4; k++) nd[k] = __ldcg(taskNode + g + k * NUM_THREADS); /* written by another CTA one level earlier */
#pragma unroll
    for (int k = 0; k < 4; k++) ex[k] = ldg_gather_u4(expansion + nd[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) { taskCh[g + k * NUM_THREADS] = ex[k]; cnt += count_internal(ex[k], nInt); }
  }
  for (; g < b; g += NUM_THREADS) {
    const uint4 ex = ldg_gather_u4(expansion + __ldcg(taskNode + g));
    taskCh[g] = ex;
    cnt += count_internal(ex, nInt);
  }
  return cnt;
}

/* C. B consecutive tiles of NUM_THREADS tasks, [tileStart, min(tileStart + B * NUM_THREADS, end)): number the internal children from
 * childBase + (exclusive count before the task) in (task, slot) order and append their tasks.  The records were written by the same CTA in
 * number_fetch.  B tiles share one pair of barriers and have their record loads in flight together: tile by tile, a CTA's part of a large
 * level was a chain of ~1.5 us steps (25 of the 58 us of the largest level at 10 M primitives).  Returns the number of internal children
 * of the B tiles (same value in every thread). */
template <int NUM_THREADS, int B>
__device__ __forceinline__ u32 number_tiles(u32 nInt, u32* taskNode, const uint4* taskCh, u32* taskParent, u32* firstChild, ColSmem<NUM_THREADS>& S, u32 tileStart,
                                            u32 end, u32 childBase) {
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  uint4 t[B];
  u32 nInternal[B], incl[B];
#pragma unroll
  for (int k = 0; k < B; k++) {
    const u32 g = tileStart + (u32)k * NUM_THREADS + tid;
    t[k] = make_uint4(B2_INVALID, B2_INVALID, B2_INVALID, B2_INVALID);
    if (g < end) t[k] = __ldcg(taskCh + g); /* own part: written by this CTA, possibly by another thread (a part does not start on a tile boundary) */
  }
#pragma unroll
  for (int k = 0; k < B; k++) {
    nInternal[k] = count_internal(t[k], nInt);
    incl[k] = nInternal[k];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(B2_FULL, incl[k], o);
      if ((int)l >= o) incl[k] += v;
    }
    if (l == 31) S.tileSum[k][w] = incl[k];
  }
  __syncthreads();
  u32 running = childBase;
#pragma unroll
  for (int k = 0; k < B; k++) {
    u32 warpBase = 0, tileTotal = 0;
#pragma unroll
    for (int q = 0; q < NUM_THREADS / 32; q++) { const u32 v = S.tileSum[k][q]; if (q < (int)w) warpBase += v; tileTotal += v; }
    const u32 g = tileStart + (u32)k * NUM_THREADS + tid;
    u32 nextId = running + warpBase + incl[k] - nInternal[k];
    if (g < end) firstChild[g] = nextId;
    const u32 ch[4] = {t[k].x, t[k].y, t[k].z, t[k].w};
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (ch[j] < nInt) {
        if (nextId < nInt) { taskNode[nextId] = ch[j]; taskParent[nextId] = g; } /* never past the arrays, whatever the input */
        nextId++;
      }
    running += tileTotal;
  }
  __syncthreads(); /* tileSum is reused by the next call */
  return running - childBase;
}

/* tasks [a, b) of one CTA, numbered from childBase: batches of four tiles, then single tiles */
template <int NUM_THREADS>
__device__ __forceinline__ u32 number_range(u32 nInt, u32* taskNode, const uint4* taskCh, u32* taskParent, u32* firstChild, ColSmem<NUM_THREADS>& S, u32 a, u32 b,
                                            u32 childBase) {
  u32 running = childBase, tileStart = a;
  for (; tileStart + 2 * NUM_THREADS < b; tileStart += 4 * NUM_THREADS)
    running += number_tiles<NUM_THREADS, 4>(nInt, taskNode, taskCh, taskParent, firstChild, S, tileStart, b, running);
  for (; tileStart < b; tileStart += NUM_THREADS)
    running += number_tiles<NUM_THREADS, 1>(nInt, taskNode, taskCh, taskParent, firstChild, S, tileStart, b, running);
  return running - childBase;
}

/* The first levels, while a level fits one tile (1, 4, 16, 64 ... tasks): CTA 0 alone, one task per thread, and the Bvh2 node of every task
 * handed from level to level through shared memory, so that a level is ONE dependent memory access (the gather of the expansion) instead
 * of three (task node, expansion, record): ~1.3 us per level instead of ~2.8 — the top of the tree is a fixed cost of every build,
 * a tenth of the collapse of a 150 K-triangle scene.  Writes the same arrays as the general path (which takes over at the first level
 * that does not fit).  Returns true when the input is not a tree. */
template <int NUM_THREADS>
__device__ __forceinline__ bool number_top(const uint4* __restrict__ expansion, u32 nInt, u32* taskNode, uint4* taskCh, u32* taskParent, u32* firstChild,
                                           ColSmem<NUM_THREADS>& S, u32& level, u32& start, u32& end) {
  const u32 tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  if (tid == 0) S.topNode[0][0] = taskNode[0];
  __syncthreads();
  u32 cur = 0;
  while (end - start <= NUM_THREADS && end != start) {
    const u32 size = end - start, g = start + tid;
    uint4 t = make_uint4(B2_INVALID, B2_INVALID, B2_INVALID, B2_INVALID);
    if (tid < size) { t = ldg_gather_u4(expansion + S.topNode[cur][tid]); taskCh[g] = t; }
    const u32 nInternal = count_internal(t, nInt);
    u32 incl = nInternal;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const u32 v = __shfl_up_sync(B2_FULL, incl, o);
      if ((int)l >= o) incl += v;
    }
    if (l == 31) S.tileSum[0][w] = incl;
    __syncthreads();
    u32 warpBase = 0, total = 0;
#pragma unroll
    for (int q = 0; q < NUM_THREADS / 32; q++) { const u32 v = S.tileSum[0][q]; if (q < (int)w) warpBase += v; total += v; }
    if (total > nInt - end) return true; /* more tasks than internal nodes (same value in every thread) */
    u32 nextId = end + warpBase + incl - nInternal;
    if (tid < size) firstChild[g] = nextId;
    const u32 ch[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (ch[k] < nInt) {
        taskNode[nextId] = ch[k]; taskParent[nextId] = g;
        if (nextId - end < (u32)NUM_THREADS) S.topNode[cur ^ 1u][nextId - end] = ch[k];
        nextId++;
      }
    __syncthreads();
    start = end; end += total; level++; cur ^= 1u;
  }
  return false;
}

template <int NUM_THREADS>
__global__ void __launch_bounds__(NUM_THREADS, NUM_THREADS == NUM_THREADS_LARGE ? 2 : 4) collapse_number_kernel(const uint4* __restrict__ expansion, u32 nInt, const u32* __restrict__ rootIdx,
                                                                        u32* taskNode, uint4* taskCh, u32* taskParent, u32* firstChild,
                                                                        CollapseCtrl* ctrl, u64* counts) {
  __shared__ ColSmem<NUM_THREADS> S;
  const u32 G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, w = tid >> 5, l = tid & 31u;
  u32 level = 0, start = 0, end = 1, barriers = 0, arriveTarget = 0;
  if (c == 0 && tid == 0) { taskNode[0] = *rootIdx; taskParent[0] = B2_INVALID; }
  __syncthreads();

  bool own = false;   /* this CTA's part of the level is [oa, ob): the children it numbered itself one level earlier */
  u32 oa = 0, ob = 0;
#ifdef COL_TRACE
  __shared__ unsigned long long trT[256], trP[256][3]; __shared__ u32 trS[256], trM[256]; u32 trN = 0;
#define TR_STAMP(k) if (c == 0 && tid == 0 && trN - 1 < 256) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trP[trN - 1][k] = t; }
#else
#define TR_STAMP(k)
#endif
  while (true) {
    const u32 size = end - start;
#ifdef COL_TRACE
    if (c == 0 && tid == 0 && trN < 256) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); trT[trN] = t; trS[trN] = size; trM[trN] = (own ? 1u : 0u) | (level << 8); trN++; }
#endif
    if (size == 0) break;
    if (size <= NUM_SOLO) {
      /* ---- a run of small levels (the top of the tree, the tail of a deep one): CTA 0 alone, no grid-wide step in between ---- */
      if (c == 0) {
        if (level == 0 && number_top<NUM_THREADS>(expansion, nInt, taskNode, taskCh, taskParent, firstChild, S, level, start, end) && tid == 0) st_relaxed(&ctrl->error, 1u);
        while (end - start <= NUM_SOLO && end != start) {
          number_fetch<NUM_THREADS>(expansion, nInt, taskNode, taskCh, start, end);
          const u32 running = end + number_range<NUM_THREADS>(nInt, taskNode, taskCh, taskParent, firstChild, S, start, end, end);
          start = end; end = running; level++;
          if (end > nInt) { if (tid == 0) st_relaxed(&ctrl->error, 1u); end = start; } /* not a tree: stop */
        }
        if (tid == 0) publish_level(ctrl, barriers + 1u, level, start, end);
      }
      barriers++;
      grid_barrier(&ctrl->bar, barriers * G);
      const u32* nx = reinterpret_cast<const u32*>(&ctrl->next[barriers & 1u]);
      level = ld_relaxed(nx); start = ld_relaxed(nx + 1); end = ld_relaxed(nx + 2);
      own = false;
      continue;
    }
    /* ---- a level of many tiles.  After a grid barrier CTA c takes the contiguous chunk c of the level; from then on it keeps the
     * children it numbered itself (its part of the next level is contiguous and in chunk order, so the numbering is the same) and the
     * levels follow each other with ONE grid-wide step — the exchange of counts — instead of two, until the parts drift apart ---- */
    if (!own) {
      const u32 chunk = ((size + G - 1) / G + NUM_THREADS - 1) / NUM_THREADS * NUM_THREADS;
      oa = min(end, start + c * chunk);
      ob = min(end, oa + chunk);
    }
    arriveTarget += G;
    /* A. records + internal children of the whole part */
    u32 cnt = number_fetch<NUM_THREADS>(expansion, nInt, taskNode, taskCh, oa, ob);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(B2_FULL, cnt, o);
    if (l == 0) S.warpSum[1][w] = cnt;
    __syncthreads();
    u32 partTotal = 0;
#pragma unroll
    for (int q = 0; q < NUM_THREADS / 32; q++) partTotal += S.warpSum[1][q];
    /* B. children of the parts before this one: every CTA posts after the same pass and counts itself in on an arrival
     * counter that ONE thread per CTA watches (every thread spinning on the posted words — 227 K pollers — starves the
     * CTAs still at work, as measured for the PLOC++ merge kernel, profiles/r01m).  The words of two consecutive levels live in two
     * buffers: nobody posts level L+2 before everybody has posted L+1, that is, after everybody has read the words of level L. */
    u64* lvlCounts = counts + (size_t)(level & 1u) * G;
    TR_STAMP(0)
    if (tid == 0) {
      st_relaxed64(lvlCounts + c, (u64)partTotal);
      __threadfence();
      atomicAdd(&ctrl->arrive, 1u);
      SpinGuard guard;
      while (ld_acquire(&ctrl->arrive) < arriveTarget) {
        __nanosleep(32);
        guard.tick();
      }
    }
    __syncthreads();
    TR_STAMP(1)
    u32 before = 0, total = 0, most = 0;
    for (u32 i = tid; i < G; i += NUM_THREADS) {
      const u32 v = (u32)ld_relaxed64(lvlCounts + i);
      if (i < c) before += v;
      total += v;
      most = max(most, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      before += __shfl_xor_sync(B2_FULL, before, o);
      total += __shfl_xor_sync(B2_FULL, total, o);
      most = max(most, __shfl_xor_sync(B2_FULL, most, o));
    }
    if (l == 0) { S.warpSum[2][w] = before; S.warpSum[0][w] = total; S.warpMax[w] = most; }
    __syncthreads();
    before = 0; total = 0; most = 0;
#pragma unroll
    for (int q = 0; q < NUM_THREADS / 32; q++) { before += S.warpSum[2][q]; total += S.warpSum[0][q]; most = max(most, S.warpMax[q]); }
    __syncthreads(); /* warpSum is rewritten in the next level */
    if (total > nInt - end) { /* more tasks than internal nodes: not a tree — every CTA sees the same counts and stops here */
      if (c == 0 && tid == 0) st_relaxed(&ctrl->error, 1u);
      break;
    }
    /* C. the part tile by tile: no CTA waits for another one here */
    const u32 childBase = end + before;
    number_range<NUM_THREADS>(nInt, taskNode, taskCh, taskParent, firstChild, S, oa, ob, childBase);
    TR_STAMP(2)
    oa = childBase; ob = childBase + partTotal;
    level++; start = end; end += total;
    if (total == 0) break;
    /* keep the parts while the largest is within reach of the mean (all CTAs decide alike, from the same words); otherwise — and before a
     * run of small levels, which CTA 0 must see complete — one grid barrier and equal chunks again */
    const u32 mean = total / G;
    if (total <= NUM_SOLO || most > mean + mean / NUM_OWN_SLACK_DIV + NUM_OWN_SLACK * NUM_THREADS) {
      barriers++;
      grid_barrier(&ctrl->bar, barriers * G);
      own = false;
    } else own = true;
  }
#ifdef COL_TRACE
  if (c == 0 && tid == 0) for (u32 i = 0; i + 1 < trN; i++) printf("TRACE lvl %u own %u size %u us %.2f  A %.2f wait %.2f C %.2f rest %.2f\n", trM[i] >> 8, trM[i] & 1u, trS[i], (double)(trT[i + 1] - trT[i]) * 1e-3, (double)(trP[i][0] - trT[i]) * 1e-3, (double)(trP[i][1] - trP[i][0]) * 1e-3, (double)(trP[i][2] - trP[i][1]) * 1e-3, (double)(trT[i + 1] - trP[i][2]) * 1e-3);
#endif
  if (c == 0 && tid == 0) ctrl->nWide = ld_relaxed(&ctrl->error) ? B2_INVALID : end;
}

/* ---- 3. wide nodes + leaf records: one thread per wide node, no synchronisation between CTAs.  Boxes of the internal
 * children are gathered (leaf slots keep the empty box, TwoPassLbvhKernel.h:320-325); the 128-byte node goes out through a
 * swizzled shared-memory transpose as full, contiguous lines (eight 16-byte stores per thread at a 128-byte stride run at a
 * third of the speed, tools/micro/mem_micro.cu). ---- */
#ifndef COL_EMIT_THREADS
#define COL_EMIT_THREADS 256
#endif
struct EmitSmem {
  uint4 stage[COL_EMIT_THREADS * 8]; /* 256 wide nodes, 16-byte pieces, piece p of node t at t*8 + (p ^ (t & 7)) */
};

#ifndef COL_EMIT_MINB
#define COL_EMIT_MINB 5
#endif
__global__ void __launch_bounds__(COL_EMIT_THREADS, COL_EMIT_MINB) collapse_emit_kernel(const b2bvh_bvh2_node* __restrict__ nodes, const u32* __restrict__ sortedVals, u32 nInt,
                                                                      const uint4* __restrict__ taskCh, const u32* __restrict__ taskParent,
                                                                      const u32* __restrict__ firstChild, const CollapseCtrl* __restrict__ ctrl,
                                                                      b2bvh_bvh4_node* __restrict__ wide, b2bvh_prim_node* __restrict__ wideLeaves) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  EmitSmem& S = *reinterpret_cast<EmitSmem*>(smemRaw);
  pdl_wait(); /* launched programmatically behind the numbering kernel */
  const u32 nWide = ctrl->nWide, tid = threadIdx.x;
  if (nWide > nInt) return; /* the numbering gave up (CollapseCtrl::error) */
  for (u32 tileStart = blockIdx.x * COL_EMIT_THREADS; tileStart < nWide; tileStart += gridDim.x * COL_EMIT_THREADS) {
    const u32 g = tileStart + tid;
    u32 ch[4] = {B2_INVALID, B2_INVALID, B2_INVALID, B2_INVALID};
    u32 parent = B2_INVALID, nextId = 0;
    if (g < nWide) {
      const uint4 t = __ldg(taskCh + g);
      parent = __ldg(taskParent + g);
      nextId = __ldg(firstChild + g);
      ch[0] = t.x; ch[1] = t.y; ch[2] = t.z; ch[3] = t.w;
    }
    Box box[4];
    u32 prim[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      box[k] = box_empty();
      prim[k] = 0;
      if (ch[k] < nInt) box[k] = load_node2_gather(nodes + ch[k]).box;
      /* leaf slot s holds primitive sortedVals[s] in both layouts (Bvh2 leaf m_leftChildIdx / PrimRef m_primIdx) */
      else if (ch[k] != B2_INVALID) prim[k] = ldg_gather_u32(sortedVals + (ch[k] - nInt));
    }
    u32 outChild[4], cc = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      outChild[k] = ch[k];
      if (ch[k] != B2_INVALID) {
        cc++;
        if (ch[k] < nInt) outChild[k] = nextId++;
        else reinterpret_cast<uint2*>(wideLeaves)[ch[k] - nInt] = make_uint2(prim[k], g);
      }
    }
    {
      const u32 sw = tid & 7u;
      uint4* row = S.stage + tid * 8;
      const Box &b0 = box[0], &b1 = box[1], &b2 = box[2], &b3 = box[3];
#define FU(x) __float_as_uint(x)
      row[0 ^ sw] = make_uint4(FU(b0.lx), FU(b0.ly), FU(b0.lz), FU(b0.hx));
      row[1 ^ sw] = make_uint4(FU(b0.hy), FU(b0.hz), FU(b1.lx), FU(b1.ly));
      row[2 ^ sw] = make_uint4(FU(b1.lz), FU(b1.hx), FU(b1.hy), FU(b1.hz));
      row[3 ^ sw] = make_uint4(FU(b2.lx), FU(b2.ly), FU(b2.lz), FU(b2.hx));
      row[4 ^ sw] = make_uint4(FU(b2.hy), FU(b2.hz), FU(b3.lx), FU(b3.ly));
      row[5 ^ sw] = make_uint4(FU(b3.lz), FU(b3.hx), FU(b3.hy), FU(b3.hz));
      row[6 ^ sw] = make_uint4(outChild[0], outChild[1], outChild[2], outChild[3]);
      row[7 ^ sw] = make_uint4(parent, cc, 0u, 0u);
#undef FU
    }
    /* (one transpose per CTA; a per-warp variant without the two barriers — 20 % of the stall samples sit at the first — was measured
     * 1-2 % slower: 0.610 vs 0.602 ms for the stage at 10 M, 5.40 vs 5.30 ms at 100 M, gpurun r2m) */
    __syncthreads();
    {
      const u32 valid = min((u32)COL_EMIT_THREADS, nWide - tileStart) * 8u;
      uint4* out = reinterpret_cast<uint4*>(wide + tileStart);
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const u32 q = (u32)i * COL_EMIT_THREADS + tid;
        const u32 node = q >> 3, part = q & 7u;
        if (q < valid) out[q] = S.stage[node * 8 + (part ^ (node & 7u))];
      }
    }
    __syncthreads(); /* the transpose buffer is reused by the next tile */
  }
}

int b2_launch_collapse(b2bvh_ctx* ctx, const b2bvh_bvh2_node* d_nodes, const b2bvh_prim_ref* d_leaves, const u32* d_sortedVals, const u32* d_rootIdx, u32 n,
                       b2bvh_bvh4_node* d_wide, b2bvh_prim_node* d_wideLeaves, void* d_scratch, u32* h_nWide) {
  (void)d_leaves; /* both layouts name leaves by slot; the primitive index comes from the sorted value array */
  if (n < 2) return b2_fail(B2BVH_ERR_INVALID, "collapse needs at least 2 primitives");
  int &occLarge = ctx->occ[B2_OCC_COLLAPSE_LARGE], &occSmall = ctx->occ[B2_OCC_COLLAPSE_SMALL];
  const size_t emitSmem = sizeof(EmitSmem);
  if (!(ctx->once_mask & B2_ONCE_COLLAPSE)) {
    B2_CUDA(cudaFuncSetAttribute(collapse_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitSmem));
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occLarge, collapse_number_kernel<NUM_THREADS_LARGE>, NUM_THREADS_LARGE, 0));
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occSmall, collapse_number_kernel<NUM_THREADS_SMALL>, NUM_THREADS_SMALL, 0));
    if (occLarge < 1 || occSmall < 1) return b2_fail(B2BVH_ERR_INTERNAL, "collapse: kernel does not fit on an SM");
    if (occLarge > 2) occLarge = 2; /* fewer arrivals per grid-wide step beat more warps (see NUM_THREADS_LARGE) */
    ctx->once_mask |= B2_ONCE_COLLAPSE;
  }
  const bool large = n >= (1u << 20);
  /* every CTA must be resident (grid barrier): at most SMs x occupancy; small inputs use fewer CTAs (cheaper barriers) */
  u32 grid = (u32)ctx->sm_count * (u32)(large ? occLarge : occSmall);
  const u32 want = (n + 2047u) >> 11; /* one CTA per 2048 primitives (us for 144 K / 262 K / 900 K with one per 1024: 52 / 77 / 99, 2048: 55 / 72 / 70, 4096: 60 / 76 / 70) */
  if (grid > want) grid = want;
  if (grid < 1) grid = 1;
  unsigned char* base = reinterpret_cast<unsigned char*>(d_scratch);
  CollapseCtrl* ctrl = reinterpret_cast<CollapseCtrl*>(base);
  uint4* expansion = reinterpret_cast<uint4*>(base + 256);
  uint4* taskCh = reinterpret_cast<uint4*>(base + 256 + (size_t)n * sizeof(uint4));
  u32* taskParent = reinterpret_cast<u32*>(base + 256 + (size_t)n * 2 * sizeof(uint4));
  u32* firstChild = taskParent + n;
  u32* taskNode = firstChild + n;
  u64* counts = reinterpret_cast<u64*>(taskNode + n + (n & 1u)); /* counts[2][G]: one word per CTA and level parity; 3n (+1) words after a 16-byte aligned start */
  u32 nInt = n - 1;
  /* no memsets: the expansion kernel clears CollapseCtrl, and every count word is posted before it is read */
  B2_KERNEL(ctx, "collapse_expand");
  collapse_expand_kernel<<<(nInt + COL_THREADS - 1) / COL_THREADS, COL_THREADS, 0, ctx->stream>>>(d_nodes, nInt, expansion, reinterpret_cast<uint4*>(ctrl));
  B2_LAUNCH_CHECK(ctx);
  B2_KERNEL(ctx, "collapse_number");
  void* args[] = {(void*)&expansion, (void*)&nInt, (void*)&d_rootIdx, (void*)&taskNode, (void*)&taskCh, (void*)&taskParent, (void*)&firstChild, (void*)&ctrl, (void*)&counts};
  if (large) B2_CUDA(cudaLaunchCooperativeKernel((const void*)collapse_number_kernel<NUM_THREADS_LARGE>, dim3(grid), dim3(NUM_THREADS_LARGE), args, 0, ctx->stream));
  else B2_CUDA(cudaLaunchCooperativeKernel((const void*)collapse_number_kernel<NUM_THREADS_SMALL>, dim3(grid), dim3(NUM_THREADS_SMALL), args, 0, ctx->stream));
  B2_LAUNCH_CHECK(ctx);
  /* the number of wide nodes stays on the device: the emit grid is sized for the worst case and strides over ctrl->nWide */
  u32 egrid = (nInt + COL_EMIT_THREADS - 1) / COL_EMIT_THREADS;
  const u32 ecap = (u32)ctx->sm_count * 2u * COL_EMIT_MINB * (256 / COL_EMIT_THREADS);
  if (egrid > ecap) egrid = ecap;
  B2_KERNEL(ctx, "collapse_emit");
  B2_LAUNCH_PDL(collapse_emit_kernel, egrid, COL_EMIT_THREADS, emitSmem, ctx->stream, d_nodes, d_sortedVals, nInt, taskCh, taskParent, firstChild, ctrl, d_wide, d_wideLeaves);
  B2_LAUNCH_CHECK(ctx);
  /* CollapseCtrl::nWide comes back through the mailbox: b2_mailbox(ctx, B2_MB_COLLAPSE)[1] after the build's final synchronisation
   * (no host round trip in the middle of a build) */
  (void)h_nWide;
  B2_TRY(b2_fetch_words(ctx, ctrl, 4, B2_MB_COLLAPSE));
  return 0;
}
