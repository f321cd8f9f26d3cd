/* TEST INFRASTRUCTURE ONLY — the reference's Ploc and SinglePassPloc kernels (Ploc++Kernel.h:98-362), UNMODIFIED, run block
 * by block with one cooperative fiber per GPU thread (cuda_emul_mt.h), driven by the host loop of PLOCNew::build
 * (PLOC++Bvh.cpp:131-152).  Node numbering inside an iteration follows the order in which warps reach an atomicAdd
 * (binaryWarpPrefixSum, :57-68) — timing dependent on a GPU, highest warp first here: callers compare trees up to numbering. */
#include <functional>
#include <vector>

#include <src/Common.h>
namespace {
#include <src/Ploc++Kernel.h>
}

thread_local dim3e threadIdx, blockIdx, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
namespace b2emul {
Block g_blk;
void yield_to_scheduler() { swapcontext(&g_blk.fibers[g_blk.current].ctx, &g_blk.main); }
}
static std::function<void()>* g_kernel;
static bool g_lowestFirst = false; /* order of the runnable fibers between two synchronisation points (see ref_hploc_mt) */
static void fiber_entry() {
  (*g_kernel)();
  b2emul::g_blk.fibers[b2emul::g_blk.current].state = b2emul::EXITED;
  b2emul::yield_to_scheduler();
}

static void run_block(uint32_t block, uint32_t nThreads, uint32_t nBlocks, std::function<void()> kernel) {
  using namespace b2emul;
  const size_t stackBytes = 128 * 1024;
  std::vector<Fiber> fibers(nThreads);
  std::vector<char> stacks((size_t)nThreads * stackBytes);
  g_blk.nThreads = (int)nThreads; g_blk.fibers = fibers.data();
  g_kernel = &kernel;
  for (uint32_t t = 0; t < nThreads; t++) {
    Fiber& f = fibers[t];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = stacks.data() + (size_t)t * stackBytes;
    f.ctx.uc_stack.ss_size = stackBytes;
    f.ctx.uc_link = &g_blk.main;
    makecontext(&f.ctx, fiber_entry, 0);
    f.state = READY;
  }
  blockIdx = {block, 0, 0}; blockDim = {nThreads, 1, 1}; gridDim = {nBlocks, 1, 1};
  const int nW = ((int)nThreads + kWarp - 1) / kWarp;
  while (true) {
    /* run every READY fiber once, highest index first */
    bool ran = false;
    for (int k = 0; k < (int)nThreads; k++) {
      const int t = g_lowestFirst ? k : (int)nThreads - 1 - k;
      if (fibers[t].state != READY) continue;
      g_blk.current = t;
      threadIdx = {(uint32_t)t, 0, 0};
      swapcontext(&g_blk.main, &fibers[t].ctx);
      ran = true;
    }
    /* warp collectives: complete where every lane that is neither at the barrier nor gone has arrived */
    bool progressed = false;
    for (int w = 0; w < nW; w++) {
      const int lo = w * kWarp, hi = std::min((int)nThreads, lo + kWarp);
      int arrived = 0, ready = 0;
      for (int t = lo; t < hi; t++) { arrived += fibers[t].state == AT_COLLECTIVE; ready += fibers[t].state == READY; }
      if (arrived == 0 || ready != 0) continue;
      uint64_t ballot = 0;
      for (int t = lo; t < hi; t++) if (fibers[t].state == AT_COLLECTIVE && fibers[t].pred) ballot |= 1ull << (t - lo);
      for (int t = lo; t < hi; t++)
        if (fibers[t].state == AT_COLLECTIVE) {
          fibers[t].resBallot = ballot;
          fibers[t].resVal = fibers[lo + fibers[t].src < hi ? lo + fibers[t].src : lo].val;
        }
      for (int t = lo; t < hi; t++) if (fibers[t].state == AT_COLLECTIVE) fibers[t].state = READY;
      progressed = true;
    }
    if (progressed) continue;
    /* block barrier: everybody who has not returned waits */
    int atBarrier = 0, exited = 0;
    for (uint32_t t = 0; t < nThreads; t++) { atBarrier += fibers[t].state == AT_BARRIER; exited += fibers[t].state == EXITED; }
    if (exited == (int)nThreads) break;
    if (atBarrier + exited == (int)nThreads) {
      for (uint32_t t = 0; t < nThreads; t++) if (fibers[t].state == AT_BARRIER) fibers[t].state = READY;
      continue;
    }
    if (!ran) { fprintf(stderr, "ref_emul_ploc_mt: block %u cannot make progress (divergent barrier?)\n", block); abort(); }
  }
}

extern "C" {
/* nodeIdx0 = SetupClusters output (n entries), nodeIdx1 = scratch (n entries).  Returns the number of kernel launches.
 * bvhNodes[n-1] internal nodes are written. */
uint32_t ref_ploc_build_mt(Bvh2Node* bvhNodes, PrimRef* primRefs, int* nodeIdx0, int* nodeIdx1, uint32_t n) {
  uint32_t nClusters = n, launches = 0;
  const uint32_t nInternalNodes = n - 1;
  bool swapBuffer = false;
  while (nClusters > 1) {
    int nMerged = 0, blockOffsetSum = 0, atomicBlockCounter = 0;
    int* in = !swapBuffer ? nodeIdx0 : nodeIdx1;
    int* out = !swapBuffer ? nodeIdx1 : nodeIdx0;
    launches++;
    if (nClusters < (uint32_t)PlocBlockSize) { /* PLOC++Bvh.cpp:138-143 */
      run_block(0, PlocBlockSize, 1, [&] { SinglePassPloc(in, bvhNodes, primRefs, nClusters, nInternalNodes); });
      break;
    }
    const uint32_t nBlocks = (nClusters + PlocBlockSize - 1) / PlocBlockSize;
    for (uint32_t b = 0; b < nBlocks; b++)
      run_block(b, PlocBlockSize, nBlocks, [&] { Ploc(in, out, bvhNodes, primRefs, &nMerged, &blockOffsetSum, &atomicBlockCounter, nClusters, nInternalNodes); });
    nClusters -= (uint32_t)nMerged;
    swapBuffer = !swapBuffer;
  }
  return launches;
}
}

/* ---- the reference's batched build kernel (BatchedBuildKernel.h:218-312) under the same block emulator.  The header is compiled from
 * a temporary copy in which ONE token is removed — `__shared__` in the parameter list of blockReduce (:76), where this prelude's
 * `static` cannot stand (oracle/Makefile makes and deletes the copy) — and with ExtentCacheSize, which the reference never defines,
 * set on the command line.  One 32-thread block per item; all items must have the same size (the kernel's offsets, :234-235). ---- */
#ifdef B2_BATCHED_KERNEL
namespace refbatched {
#include B2_BATCHED_KERNEL
}
extern "C" void ref_batched_build_mt(const Triangle* tris, uint32_t nItems, uint32_t primCount, Bvh2Node* nodes, PrimRef* leaves, uint32_t* roots, Aabb* scenes) {
  std::vector<D_BatchedBuildInputs> in(nItems);
  for (uint32_t i = 0; i < nItems; i++) { in[i].m_prims = const_cast<Triangle*>(tris) + (size_t)i * primCount; in[i].m_nPrimtives = primCount; }
  std::vector<uint2> spans((size_t)nItems * primCount);
  for (uint32_t b = 0; b < nItems; b++)
    run_block(b, MaxBatchedBlockSize, nItems, [&] { refbatched::BatchedBuildKernelLbvh(in.data(), nodes, leaves, spans.data(), roots, nItems, scenes); });
}
#endif


/* ---- the reference's H-PLOC kernel (HplocKernel.h:257-315 with findParent, plocMerge, loadIndices, findNearestNeighbours, mergeClusters,
 * storeIndices) under the block emulator: one 32-thread block = one wavefront (HPlocBlockSize = 32, Common.h:594), blocks one after the
 * other (siblings meet through atomicExch on parentIdx and the first arriver leaves: nobody ever waits for another block, so any block
 * order is a schedule the GPU could have produced).  The kernel relies on LOCK-STEP execution of the wavefront in two places, and the
 * emulation reproduces exactly that:
 *   1. mergeClusters reads nearestNeighbours[] of OTHER lanes right after findNearestNeighbours' atomicMin, with no barrier in between
 *      (HplocKernel.h:117 -> :137-142): a lock-step wavefront has finished every lane's atomicMin before any lane reads.  The header is
 *      compiled from a temporary copy with ONE statement added — `__syncthreads();` as the first statement of mergeClusters (oracle/Makefile
 *      makes and deletes the copy) — a no-op on lock-step hardware, the reconvergence point for the fibers here.
 *   2. the compaction at the end of mergeClusters (:183-185) lets the lanes WITHOUT a surviving cluster write to the slot of the next
 *      surviving lane; the surviving lane is the highest lane writing that slot and must win.  Between two synchronisation points the
 *      fibers run one at a time in ASCENDING lane order: the highest lane writes last (and a merging lane reads its partner's slot, which
 *      belongs to a higher lane, before the partner invalidates it, :139-141).
 * Launch as Hploc.cpp:120: nInternalNodes threads rounded up to whole blocks — leaf n-1 gets no thread when (n-1) % 32 == 0 (a defect of the
 * reference; callers use other sizes).  Node numbering follows the order of the atomicAdd on nMergedClusters (timing dependent on a GPU):
 * callers compare trees up to numbering. ---- */
#ifdef B2_HPLOC_KERNEL
namespace refhploc {
#include B2_HPLOC_KERNEL
}
extern "C" uint32_t ref_hploc_mt(Bvh2Node* bvhNodes, PrimRef* primRefs, uint32_t* sortedKeys, uint32_t* nodeIdx0, uint32_t* parentIdx, uint32_t n) {
  uint32_t nMerged = 0;
  const uint32_t nInternalNodes = n - 1;
  const uint32_t nBlocks = (nInternalNodes + HPlocBlockSize - 1) / HPlocBlockSize;
  g_lowestFirst = true;
  for (uint32_t b = 0; b < nBlocks; b++)
    run_block(b, HPlocBlockSize, nBlocks, [&] { refhploc::HPloc(bvhNodes, primRefs, sortedKeys, nodeIdx0, parentIdx, &nMerged, n, nInternalNodes); });
  g_lowestFirst = false;
  return nMerged;
}
#endif
