/*
 * bvh_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain, single-threaded C++ restatement of the reference algorithms on the
 * BVH-build hot path of Niravaana/HIP-BVH-Construction.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it (as the checker or as the timed CPU arm).
 * The product library (libb2bvh.so) never links, loads or calls anything here.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared  (oracle/Makefile).
 * -ffp-contract=off fixes the floating-point contract: every +,-,*,/ is rounded
 * individually (SURVEY.md §7 "Floating-point contract").
 *
 * Pinning (see oracle/README.md, tests/test_oracle_vs_ref.py, tests/golden/):
 *   - Morton coding, Karras range/split, Apetrei build+fit, refit, SetupClusters,
 *     Bvh4 collapse: checked bit-for-bit against the UNMODIFIED reference kernels
 *     executed by the sequential CUDA-thread emulator in oracle/ref_shim
 *     (oracle/_ref/libref_emul.so, built from /root/reference in place).
 *   - cost functions, validators, early-split (no-split) PrimRefs, CPU traversal,
 *     OBJ loader: checked against the reference's own Utility.cpp compiled as is.
 *   - sort: std::stable_sort, the oracle Orochi's own test uses
 *     (dependencies/Orochi/Test/RadixSort/main.cpp:130,239).
 *   - PLOC++ / H-PLOC: pinned by the README SAH costs (README.md:145,165,187,207)
 *     reproduced on bunny/sponza; node numbering is canonical (see below) because
 *     the reference's is atomic-order dependent.
 *
 * Each function cites the reference file:line it follows.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "../include/b2bvh_types.h"

typedef uint32_t u32;
typedef uint64_t u64;
typedef b2bvh_float3 F3;
typedef b2bvh_aabb Box;
static const u32 INVALID = B2BVH_INVALID;
static const float FLTMAX = B2BVH_FLT_MAX;

/* ------------------------------------------------------------------ math */

/* fminf/fmaxf as the GPU evaluates them (Common.h:224-270 use fminf/fmaxf):
 * a NaN operand is ignored, and -0 orders below +0 (PTX min/max semantics). */
static inline u32 fbits(float f) { u32 u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(u32 u) { float f; memcpy(&f, &u, 4); return f; }
static inline float fminr(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return (fbits(a) >> 31) ? a : b;
  return a < b ? a : b;
}
static inline float fmaxr(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  if (a == b) return (fbits(a) >> 31) ? b : a;
  return a > b ? a : b;
}
static inline F3 f3(float x, float y, float z) { F3 r = {x, y, z}; return r; }
static inline F3 vmin(F3 a, F3 b) { return f3(fminr(a.x, b.x), fminr(a.y, b.y), fminr(a.z, b.z)); }
static inline F3 vmax(F3 a, F3 b) { return f3(fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z)); }
static inline F3 vadd(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline F3 vsub(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline F3 vmul(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline F3 vdiv(F3 a, F3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline F3 vscale(F3 a, float c) { return f3(c * a.x, c * a.y, c * a.z); }   /* Common.h:197,217 */
static inline F3 vdivs(F3 a, float c) { return f3(a.x / c, a.y / c, a.z / c); }     /* Common.h:174 */
static inline float vdot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  /* Common.h:280 */
static inline F3 vcross(F3 a, F3 b) {                                                /* Common.h:284 */
  return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline F3 vnormalize(F3 a) { return vdivs(a, sqrtf(vdot(a, a))); }           /* Common.h:282 */

/* Aabb, Common.h:310-416 */
static inline Box box_empty() { Box b; b.m_min = f3(FLTMAX, FLTMAX, FLTMAX); b.m_max = f3(-FLTMAX, -FLTMAX, -FLTMAX); return b; }
static inline void box_grow_p(Box& b, F3 p) { b.m_min = vmin(b.m_min, p); b.m_max = vmax(b.m_max, p); }
static inline void box_grow(Box& b, const Box& o) { b.m_min = vmin(b.m_min, o.m_min); b.m_max = vmax(b.m_max, o.m_max); }
static inline Box box_merge(const Box& l, const Box& r) { Box b; b.m_min = vmin(l.m_min, r.m_min); b.m_max = vmax(l.m_max, r.m_max); return b; }
static inline F3 box_center(const Box& b) { return vscale(vadd(b.m_max, b.m_min), 0.5f); }   /* Common.h:347 */
static inline F3 box_extent(const Box& b) { return vsub(b.m_max, b.m_min); }
static inline float box_area(const Box& b) {                                                /* Common.h:361-365 */
  F3 e = box_extent(b);
  float xy = e.x * e.y, xz = e.x * e.z, yz = e.y * e.z;
  float s = xy + xz;
  s = s + yz;
  return 2 * s;
}
static inline int box_max_dim(const Box& b) {                                               /* Common.h:351-359 */
  F3 d = box_extent(b);
  if (d.x > d.y && d.x > d.z) return 0;
  return d.y > d.z ? 1 : 2;
}
static inline F3 box_offset(const Box& b, F3 p) {                                           /* Common.h:367-374 */
  F3 o = vsub(p, b.m_min);
  if (b.m_max.x > b.m_min.x) o.x /= b.m_max.x - b.m_min.x;
  if (b.m_max.y > b.m_min.y) o.y /= b.m_max.y - b.m_min.y;
  if (b.m_max.z > b.m_min.z) o.z /= b.m_max.z - b.m_min.z;
  return o;
}
static inline Box tri_box(const b2bvh_triangle& t) {
  Box b = box_empty();
  box_grow_p(b, t.v1); box_grow_p(b, t.v2); box_grow_p(b, t.v3);
  return b;
}

extern "C" {

float orc_area(const Box* b) { return box_area(*b); }

/* FNV-1a-32 over u32 words (SURVEY.md Appendix D hash definition). */
u32 orc_fnv1a_words(const u32* a, const u32* b, u64 n) {
  u32 h = 2166136261u;
  for (u64 i = 0; i < n; i++) {
    h = (h ^ a[i]) * 16777619u;
    if (b) h = (h ^ b[i]) * 16777619u;
  }
  return h;
}

/* ------------------------------------------------- S1: PrimRefs + scene box
 * Utility::doEarlySplitClipping with saMax = FltMax (Utility.cpp:456-476: no
 * splitting, output order = input order) followed by CalculatePrimRefExtents
 * (CommonBlocksKernel.h:116-137): scene box = union of primitive boxes.
 * triAabb (optional) is CalculateSceneExtents' per-primitive Aabb output
 * (CommonBlocksKernel.h:92-114).                                             */
void orc_primrefs(const b2bvh_triangle* tris, u32 n, b2bvh_prim_ref* refs, Box* triAabb, Box* scene) {
  Box s = box_empty();
  for (u32 i = 0; i < n; i++) {
    Box b = tri_box(tris[i]);
    if (refs) { refs[i].m_primIdx = i; refs[i].m_aabb = b; }
    if (triAabb) triAabb[i] = b;
    box_grow(s, b);
  }
  *scene = s;
}

/* ------------------------------------------------- S0: early split clipping
 * Utility::doEarlySplitClipping (Utility.cpp:456-538), the host step in front of the path when the builders are
 * compiled with USE_PRIM_SPLITTING (TwoPassLbvh.cpp:23-28, saMax = 10): a FIFO of PrimRefs, seeded with one per
 * triangle in input order; a reference whose box area is <= saMax is emitted, any other is cut at the MIDDLE of its
 * box along the largest extent (no triangle clipping: the two halves are plain box halves carrying the same m_primIdx)
 * and both halves go to the back of the queue.  The emission order is therefore level by level, and inside a level
 * the order of the queue.  Returns the number of references; writes at most `cap` of them.  `maxLevels` bounds the
 * generations (the reference loops forever when a box never gets small enough, e.g. an infinite or NaN area — those
 * are reported at once): 0 is returned then. */
u32 orc_early_split(const b2bvh_triangle* tris, u32 n, float saMax, b2bvh_prim_ref* out, u32 cap, u32 maxLevels) {
  std::queue<std::pair<b2bvh_prim_ref, u32>> q; /* reference + generation */
  for (u32 i = 0; i < n; i++) {
    b2bvh_prim_ref r; r.m_primIdx = i; r.m_aabb = tri_box(tris[i]);
    q.push(std::make_pair(r, 0u));
  }
  u64 count = 0;
  while (!q.empty()) {
    const b2bvh_prim_ref ref = q.front().first; const u32 gen = q.front().second; q.pop();
    if (box_area(ref.m_aabb) <= saMax) {            /* :477-480 */
      if (count < cap && out) out[count] = ref;
      count++;
      continue;
    }
    if (gen + 1 >= maxLevels) return 0;
    if (!(box_area(ref.m_aabb) < INFINITY)) return 0; /* inf / NaN area: halving never gets below saMax */
    const int dim = box_max_dim(ref.m_aabb);        /* :484 */
    const F3 c = box_center(ref.m_aabb);            /* :485 */
    b2bvh_prim_ref L = ref, R = ref;                /* :488-532: L = [min, max with centre on dim], R = [min with centre on dim, max] */
    if (dim == 0) { L.m_aabb.m_max.x = c.x; R.m_aabb.m_min.x = c.x; }
    if (dim == 1) { L.m_aabb.m_max.y = c.y; R.m_aabb.m_min.y = c.y; }
    if (dim == 2) { L.m_aabb.m_max.z = c.z; R.m_aabb.m_min.z = c.z; }
    q.push(std::make_pair(L, gen + 1)); q.push(std::make_pair(R, gen + 1));
    if (count + q.size() > 0x3FFFFFFFull) return 0;
  }
  return (u32)count;
}

/* the per-triangle box table of the cost report when references were split (TwoPassLbvh.cpp:188-193):
 * triangleAabb[ref.m_primIdx] = ref.m_aabb in reference order, so the LAST fragment of a triangle wins;
 * the table has one entry per REFERENCE (entries past the triangle count stay empty). */
void orc_split_cost_boxes(const b2bvh_prim_ref* refs, u32 m, Box* table) {
  for (u32 i = 0; i < m; i++) table[i] = box_empty();
  for (u32 i = 0; i < m; i++) table[refs[i].m_primIdx] = refs[i].m_aabb;
}

/* ------------------------------------------------- S2: extended Morton code
 * computeExtendedMortonCode, CommonBlocksKernel.h:159-359.  The bit allocation
 * depends only on the scene extent, so it is computed once (orc_morton_config)
 * and applied per primitive (orc_morton_code_cfg).                            */
struct MortonCfg {
  int axis[3];     /* startAxis: axis[0] has the largest extent            */
  int pre[2];      /* numPrebits.x, numPrebits.y after clamping            */
  int swap;        /* > 0: one extra leading bit for axis[0]               */
  int sum;         /* numPrebitsSum after the swap adjustment              */
  int nb[3];       /* numBits.x/y/z                                        */
};

/* (int)log2(a/b) as the GPU evaluates it: float log2, then a float->int conversion that SATURATES
 * (NaN -> 0, +inf/huge -> INT_MAX, -inf -> INT_MIN; cvt.rzi.s32.f32 on NVIDIA, v_cvt_i32_f32 on AMD),
 * followed by two's-complement wrapping integer arithmetic.  This matters only for scenes with a zero
 * extent (e2 == 0 -> ratio inf); there the x86 conversion of the emulated reference (INT_MIN) is not
 * what the reference's GPU path computes, so the GPU semantic is the one pinned here.                */
static int ilog2_ratio(float a, float b) {
  float r = a / b;
  float l = log2f(r);
  if (l != l) return 0;
  if (l >= 2147483648.0f) return 2147483647;
  if (l <= -2147483648.0f) return (int)0x80000000u;
  return (int)l;
}
static inline int wadd(int a, int b) { return (int)((u32)a + (u32)b); }
static inline int wsub(int a, int b) { return (int)((u32)a - (u32)b); }
static inline int wmul2(int a) { return (int)((u32)a * 2u); }
static inline u32 shl32(u32 v, int n) { return (n < 0 || n > 31) ? 0u : (v << n); }
static inline u32 shr32(u32 v, int n) { return (n < 0 || n > 31) ? 0u : (v >> n); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

void orc_morton_config(const float ext[3], MortonCfg* c) {
  /* comparison tree of CommonBlocksKernel.h:167-250 as a decision table on
   * (x<y, x<z, y<z); strict '<' so ties fall to the else branches.           */
  static const int order[8][3] = {
      {0, 1, 2}, /* 000 */ {0, 2, 1}, /* 001 */ {0, 1, 2}, /* 010 */ {2, 0, 1}, /* 011 */
      {1, 0, 2}, /* 100 */ {1, 0, 2}, /* 101 */ {1, 2, 0}, /* 110 */ {2, 1, 0}  /* 111 */
  };
  int k = (ext[0] < ext[1] ? 4 : 0) | (ext[0] < ext[2] ? 2 : 0) | (ext[1] < ext[2] ? 1 : 0);
  for (int i = 0; i < 3; i++) c->axis[i] = order[k][i];
  float e0 = ext[c->axis[0]], e1 = ext[c->axis[1]], e2 = ext[c->axis[2]];
  int px = ilog2_ratio(e0, e1), py = ilog2_ratio(e1, e2), pz = ilog2_ratio(e0, e2);
  int swap = wsub(pz, wadd(px, py));                           /* :252 */
  px = imin(px, 30);                                           /* :254 */
  py = imin(wmul2(py), wsub(30, px)) / 2;                      /* :255 */
  int sum = wadd(px, wmul2(py));                               /* :257 */
  if (sum != 30) sum = wadd(sum, swap); else swap = 0;         /* :259-262 */
  int nbz = (e2 != 0.0f) ? imax(0, wsub(30, sum) / 3) : 0;     /* :264 */
  int nbx, nby;
  if (swap > 0) { nbx = imax(0, wadd(wadd(wadd(wsub(wsub(30, nbz), sum) / 2, py), px), 1)); nby = wsub(wsub(30, nbx), nbz); }  /* :266-270 */
  else { nby = imax(0, wadd(wsub(wsub(30, nbz), sum) / 2, py)); nbx = wsub(wsub(30, nby), nbz); }                             /* :271-275 */
  c->pre[0] = px; c->pre[1] = py; c->swap = swap; c->sum = sum;
  c->nb[0] = nbx; c->nb[1] = nby; c->nb[2] = nbz;
}

static inline u32 spread2(u32 v) {   /* morton2D, :139-147 */
  v &= 0x0000ffffu;
  v = (v ^ (v << 8)) & 0x00ff00ffu;
  v = (v ^ (v << 4)) & 0x0f0f0f0fu;
  v = (v ^ (v << 2)) & 0x33333333u;
  v = (v ^ (v << 1)) & 0x55555555u;
  return v;
}
static inline u32 spread3(u32 x) {   /* morton3D, :149-156 */
  x = (x * 0x00010001u) & 0xFF0000FFu;
  x = (x * 0x00000101u) & 0x0F00F00Fu;
  x = (x * 0x00000011u) & 0xC30C30C3u;
  x = (x * 0x00000005u) & 0x49249249u;
  return x;
}
static inline u32 quantize(float p, int nb) {   /* :281-283 */
  float scale = (float)shl32(1u, nb);
  float v = fmaxr(p * scale, 0.0f);
  u32 q = (v >= 4294967296.0f) ? 0xFFFFFFFFu : (u32)v;
  u32 lim = shl32(1u, nb) - 1u;
  return q < lim ? q : lim;
}

u32 orc_morton_code_cfg(const float p[3], const MortonCfg* c) {
  int nbx = c->nb[0], nby = c->nb[1], nbz = c->nb[2];
  u32 ax = quantize(p[c->axis[0]], nbx);
  u32 ay = quantize(p[c->axis[1]], nby);
  u32 az = quantize(p[c->axis[2]], nbz);
  u32 code = 0;
  int d0 = 0, d1 = 0;
  if (c->sum > 0) {                                            /* :289-338 */
    int px = c->pre[0], py = c->pre[1];
    nbx -= px;
    code = shr32(ax & shl32(shl32(1u, px) - 1u, nbx), nbx);
    code = shl32(code, py * 2);
    nbx -= py; nby -= py;
    u32 t0 = spread2(shr32(ax & shl32(shl32(1u, py) - 1u, nbx), nbx));
    u32 t1 = spread2(shr32(ay & shl32(shl32(1u, py) - 1u, nby), nby));
    code |= t0 * 2 + t1;
    if (c->swap > 0) {
      code = shl32(code, 1);
      nbx -= 1;
      code |= shr32(ax & shl32(1u, nbx), nbx);
    }
    code = shl32(code, nbx + nby + nbz);
    ax &= shl32(1u, nbx) - 1u;
    ay &= shl32(1u, nby) - 1u;
    if (c->swap > 0) { d0 = nby - nbx; ax = shl32(ax, d0); d1 = nby - nbz; az = shl32(az, d1); }
    else { d0 = nbx - nby; ay = shl32(ay, d0); d1 = nbx - nbz; az = shl32(az, d1); }
  }
  if (nbz == 0) {                                              /* :340-345 */
    code |= spread2(ax) * 2 + spread2(ay);
  } else {                                                     /* :346-356 */
    u32 mx = ax ? spread3(ax) : 0, my = ay ? spread3(ay) : 0, mz = az ? spread3(az) : 0;
    if (c->swap > 0) code |= shr32(my * 4 + mx * 2 + mz, d0 + d1);
    else code |= shr32(mx * 4 + my * 2 + mz, d0 + d1);
  }
  return code;
}

/* CalculateMortonCodes / CalculateMortonCodesPrimRef, :374-398.
 * boxes: n Aabbs at `strideBytes` apart (24 for Aabb[], 28 for PrimRef[] + 4). */
void orc_morton_codes(const void* boxes, u32 strideBytes, const Box* scene, u32 n, u32* keys, u32* vals) {
  F3 ext = box_extent(*scene);
  float e[3] = {ext.x, ext.y, ext.z};
  MortonCfg cfg;
  orc_morton_config(e, &cfg);
  for (u32 i = 0; i < n; i++) {
    const Box* b = (const Box*)((const char*)boxes + (size_t)i * strideBytes);
    F3 c = box_center(*b);
    F3 q = vdiv(vsub(c, scene->m_min), ext);
    float p[3] = {q.x, q.y, q.z};
    keys[i] = orc_morton_code_cfg(p, &cfg);
    vals[i] = i;
  }
}

/* ------------------------------------------------- S3: sort
 * Oro::RadixSort::sort KV, bits [0,32) (RadixSort.cpp:291-318) == stable sort
 * by key (Orochi Test/RadixSort/main.cpp:130,239).                           */
void orc_sort_kv(const u32* keys, const u32* vals, u32 n, u32* keysOut, u32* valsOut) {
  std::vector<u32> perm(n);
  for (u32 i = 0; i < n; i++) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });
  for (u32 i = 0; i < n; i++) { keysOut[i] = keys[perm[i]]; valsOut[i] = vals[perm[i]]; }
}

/* ------------------------------------------------- S4a: Karras two-pass LBVH
 * InitBvhNodesPrimRef (TwoPassLbvhKernel.h:164-194), determineRange (:42-100),
 * findSplit (:102-130), BvhBuild (:196-216), FitBvhNodes (:217-235).          */
}  /* extern "C" */
static inline int clz32(u32 v) { return v ? __builtin_clz(v) : 32; }
static inline int karras_delta(const u32* k, u32 n, int i, int j) {
  if (j < 0 || j >= (int)n) return -1;
  if (k[i] != k[j]) return clz32(k[i] ^ k[j]);     /* :27-30 */
  return 32 + clz32((u32)i ^ (u32)j);             /* :32-40 (clzll of (0<<32 | i^j)) */
}
/* 64-bit keys (the 60-bit Morton variant, no reference counterpart): the same rule one word wider —
 * common prefix of the keys, ties broken by the common prefix of the indices. */
static inline int karras_delta(const u64* k, u32 n, int i, int j) {
  if (j < 0 || j >= (int)n) return -1;
  if (k[i] != k[j]) return __builtin_clzll(k[i] ^ k[j]);
  return 64 + clz32((u32)i ^ (u32)j);
}

template <typename K>
static void lbvh_karras_t(const b2bvh_prim_ref* refs, const K* keys, const u32* vals, u32 n,
                          b2bvh_bvh2_node* nodes, u32* parents) {
  const u32 nInt = n - 1;
  std::vector<u32> par(2 * (size_t)n - 1, INVALID);
  for (u32 g = 0; g < n; g++) {                       /* leaves */
    b2bvh_bvh2_node& nd = nodes[nInt + g];
    nd.m_aabb = refs[vals[g]].m_aabb;
    nd.m_leftChildIdx = refs[vals[g]].m_primIdx;
    nd.m_rightChildIdx = INVALID;
  }
  for (u32 i = 0; i < nInt; i++) {
    int first, last;
    if (i == 0) { first = 0; last = (int)n - 1; }
    else {
      int dl = karras_delta(keys, n, i, (int)i - 1), dr = karras_delta(keys, n, i, (int)i + 1);
      int d = dr > dl ? 1 : -1;
      int dmin = dl < dr ? dl : dr;
      int lmax = 2;
      while (karras_delta(keys, n, i, (int)i + d * lmax) > dmin) lmax <<= 1;
      int l = 0;
      for (int t = lmax >> 1; t > 0; t >>= 1)
        if (karras_delta(keys, n, i, (int)i + (l + t) * d) > dmin) l += t;
      int j = (int)i + l * d;
      first = d < 0 ? j : (int)i; last = d < 0 ? (int)i : j;
    }
    int dnode = karras_delta(keys, n, first, last);
    int split = first, stride = last - first;
    do {
      stride = (stride + 1) >> 1;
      int mid = split + stride;
      if (mid < last && karras_delta(keys, n, first, mid) > dnode) split = mid;
    } while (stride > 1);
    u32 l = (split == first) ? (u32)split + nInt : (u32)split;
    u32 r = (split + 1 == last) ? (u32)split + 1 + nInt : (u32)split + 1;
    nodes[i].m_leftChildIdx = l; nodes[i].m_rightChildIdx = r;
    nodes[i].m_aabb = box_empty();
    par[l] = i; par[r] = i;
  }
  /* refit: second arrival at a node merges its children and continues upward */
  std::vector<u32> flags(2 * (size_t)n - 1, 0);
  for (u32 g = 0; g < n; g++) {
    u32 p = par[nInt + g];
    while (p != INVALID && flags[p]++ > 0) {
      nodes[p].m_aabb = box_merge(nodes[nodes[p].m_leftChildIdx].m_aabb, nodes[nodes[p].m_rightChildIdx].m_aabb);
      p = par[p];
    }
  }
  if (parents) memcpy(parents, par.data(), par.size() * sizeof(u32));
}
extern "C" {
void orc_lbvh_karras(const b2bvh_prim_ref* refs, const u32* keys, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, u32* parents) {
  lbvh_karras_t<u32>(refs, keys, vals, n, nodes, parents);
}
void orc_lbvh_karras64(const b2bvh_prim_ref* refs, const u64* keys, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, u32* parents) {
  lbvh_karras_t<u64>(refs, keys, vals, n, nodes, parents);
}

/* ------------------------------------------------- S4b: Apetrei single-pass LBVH
 * InitBvhNodes (SinglePassLbvhKernel.h:27-54), findHighestDiffBit (:56-62),
 * findParent (:64-86), BvhBuildAndFit (:88-126).  Returns the root index that
 * the kernel leaves in bvhNodeCounter[nLeafNodes-1] (:118).                   */
}  /* extern "C" */
static inline u64 aug_xor(const u32* k, int n, int i, int j) {
  if (j < 0 || j >= n) return ~0ull;
  return (((u64)k[i] << 32) | (u32)i) ^ (((u64)k[j] << 32) | (u32)j);
}
/* 64-bit keys: the augmented XOR is 96 bits wide (key XOR, then index XOR); out of range = all ones */
typedef unsigned __int128 u128;
static inline u128 aug_xor(const u64* k, int n, int i, int j) {
  if (j < 0 || j >= n) return ~(u128)0;
  return ((u128)(k[i] ^ k[j]) << 32) | (u32)((u32)i ^ (u32)j);
}

template <typename K>
static u32 lbvh_apetrei_t(const b2bvh_triangle* tris, const K* keys, const u32* vals, u32 n, b2bvh_bvh2_node* nodes) {
  const u32 nInt = n - 1;
  const int N = (int)n;
  for (u32 g = 0; g < n; g++) {
    b2bvh_bvh2_node& nd = nodes[nInt + g];
    nd.m_aabb = tri_box(tris[vals[g]]);
    nd.m_leftChildIdx = vals[g];
    nd.m_rightChildIdx = INVALID;
  }
  for (u32 i = 0; i < nInt; i++) { nodes[i].m_aabb = box_empty(); nodes[i].m_leftChildIdx = nodes[i].m_rightChildIdx = INVALID; }
  std::vector<int> lo(n, 0), hi(n, 0), cnt(n, 0);
  u32 root = INVALID;
  auto choose_parent = [&](u32 self, int i, int j) -> u32 {
    if (i == 0 && j == N) return INVALID;
    if (i == 0 || (j != N && aug_xor(keys, N, j - 1, j) < aug_xor(keys, N, i - 1, i))) {
      nodes[j - 1].m_leftChildIdx = self; lo[j - 1] = i; return (u32)(j - 1);
    }
    nodes[i - 1].m_rightChildIdx = self; hi[i - 1] = j; return (u32)(i - 1);
  };
  for (u32 g = 0; g < n; g++) {
    u32 cur = choose_parent(nInt + g, (int)g, (int)g + 1);
    while (cur != INVALID && cnt[cur]++ > 0) {
      b2bvh_bvh2_node& nd = nodes[cur];
      nd.m_aabb = box_merge(nodes[nd.m_leftChildIdx].m_aabb, nodes[nd.m_rightChildIdx].m_aabb);
      u32 p = choose_parent(cur, lo[cur], hi[cur]);
      if (p == INVALID) { root = cur; break; }
      cur = p;
    }
  }
  return root;
}
extern "C" {
u32 orc_lbvh_apetrei(const b2bvh_triangle* tris, const u32* keys, const u32* vals, u32 n, b2bvh_bvh2_node* nodes) {
  return lbvh_apetrei_t<u32>(tris, keys, vals, n, nodes);
}
u32 orc_lbvh_apetrei64(const b2bvh_triangle* tris, const u64* keys, const u32* vals, u32 n, b2bvh_bvh2_node* nodes) {
  return lbvh_apetrei_t<u64>(tris, keys, vals, n, nodes);
}

/* ------------------------------------------------- 60-bit Morton codes (north star "30/60-bit Morton coding")
 * The reference has 30-bit codes only; at 100 M primitives neighbouring primitives share codes and the order below the
 * code resolution falls back to the input index.  Defined here: the PLAIN interleave of computeMortonCode
 * (CommonBlocksKernel.h:361-372) with 20 bits per axis instead of 10 — x = min(max(p.x * 2^20, 0), 2^20 - 1) per axis
 * (float, then truncation), bits interleaved x,y,z from the top.  Equivalently: (30-bit interleave of the upper 10 bits
 * of each axis) << 30 | (30-bit interleave of the lower 10 bits).  Parity unpinned by the reference (there is nothing
 * to pin it to); the 10-bit interleave it is made of is pinned (orc_morton_plain). */
static inline u32 morton3d_10(u32 x);
u64 orc_morton60_point(const float p[3]) {
  u32 q[3];
  for (int a = 0; a < 3; a++) q[a] = (u32)fminr(fmaxr(p[a] * 1048576.0f, 0.0f), 1048575.0f);
  const u32 hi = morton3d_10(q[0] >> 10) * 4 + morton3d_10(q[1] >> 10) * 2 + morton3d_10(q[2] >> 10);
  const u32 lo = morton3d_10(q[0] & 1023u) * 4 + morton3d_10(q[1] & 1023u) * 2 + morton3d_10(q[2] & 1023u);
  return ((u64)hi << 30) | lo;
}
void orc_morton60_codes(const void* boxes, u32 strideBytes, const Box* scene, u32 n, u64* keys, u32* vals) {
  F3 ext = box_extent(*scene);
  for (u32 i = 0; i < n; i++) {
    const Box* b = (const Box*)((const char*)boxes + (size_t)i * strideBytes);
    F3 q = vdiv(vsub(box_center(*b), scene->m_min), ext);
    float p[3] = {q.x, q.y, q.z};
    keys[i] = orc_morton60_point(p);
    vals[i] = i;
  }
}
void orc_sort_kv64(const u64* keys, const u32* vals, u32 n, u64* keysOut, u32* valsOut) {
  std::vector<u32> perm(n);
  for (u32 i = 0; i < n; i++) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](u32 a, u32 b) { return keys[a] < keys[b]; });
  for (u32 i = 0; i < n; i++) { keysOut[i] = keys[perm[i]]; valsOut[i] = vals[perm[i]]; }
}

/* ------------------------------------------------- batched builder
 * BatchedBuildKernelLbvh (BatchedBuildKernel.h:218-312), one small BVH per batch item (<= MaxBatchedBlockSize = 32
 * triangles, Common.h:597), built by one block: per-item scene box (:237-259), PLAIN 10/10/10 Morton code of the
 * normalised centroid (computeMortonCode, :98-110 = CommonBlocksKernel.h:361-372), stable sort of (code, index)
 * (32 one-bit split passes, :285-297, == stable sort), Apetrei build + fit on the sorted codes (:136-216, the code of
 * SinglePassLbvhKernel.h:64-126), root in rootNodes[item] (:311).  Node / leaf indices are LOCAL to the item
 * (ptrBvhNodes / ptrLeafNodes, :300-303): internal k in [0, n-1), leaf g = (n-1) + g.
 * The reference kernel is work in progress (it does not compile: ExtentCacheSize is undefined; main.cpp:38-52 keeps
 * it behind USE_BATCHED_BUILDER).  Three repairs, each stated in DESIGN.md: (1) leaf slot g holds the primitive with the
 * g-th smallest code — the reference fills the leaf records before sorting and never permutes them (:241-242), so its
 * tree joins sorted codes to unsorted boxes; (2) item offsets are prefix sums of the item sizes (the reference uses
 * item * itemSize, :234-235, correct only for equal sizes); (3) a one-triangle item has no internal node and root 0.
 * normalisation: 0/0 -> NaN -> 0 through fmaxf (GPU min/max semantics), as in orc_morton_codes.
 * PINNED: the reference kernel itself, run by the block emulator (ref_shim/ref_emul_ploc_mt.cpp, ref_batched_build_mt), gives the
 * same topology / roots / scene boxes on any input and byte-identical nodes and leaves on items pre-sorted by code, where repair (1)
 * changes nothing (tests/test_oracle_batched.py). */
static inline u32 morton3d_10(u32 x) {                   /* BatchedBuildKernel.h:89-96 */
  x = (x * 0x00010001u) & 0xFF0000FFu;
  x = (x * 0x00000101u) & 0x0F00F00Fu;
  x = (x * 0x00000011u) & 0xC30C30C3u;
  x = (x * 0x00000005u) & 0x49249249u;
  return x;
}
u32 orc_morton_plain(const float p[3]) {                 /* :98-110 */
  float x = fminr(fmaxr(p[0] * 1024.0f, 0.0f), 1023.0f);
  float y = fminr(fmaxr(p[1] * 1024.0f, 0.0f), 1023.0f);
  float z = fminr(fmaxr(p[2] * 1024.0f, 0.0f), 1023.0f);
  return morton3d_10((u32)x) * 4 + morton3d_10((u32)y) * 2 + morton3d_10((u32)z);
}
/* tris: all items back to back; counts[item]; nodes: sum(n-1); leaves: sum(n); roots, scenes: one per item. */
void orc_batched_build(const b2bvh_triangle* tris, const u32* counts, u32 nItems, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves, u32* roots, Box* scenes) {
  u64 triOff = 0, nodeOff = 0;
  for (u32 it = 0; it < nItems; it++) {
    const u32 n = counts[it];
    const b2bvh_triangle* T = tris + triOff;
    std::vector<Box> box(n);
    Box scene = box_empty();
    for (u32 i = 0; i < n; i++) { box[i] = tri_box(T[i]); box_grow(scene, box[i]); }
    scenes[it] = scene;
    const F3 ext = box_extent(scene);
    std::vector<std::pair<u32, u32>> kv(n);
    for (u32 i = 0; i < n; i++) {
      const F3 q = vdiv(vsub(box_center(box[i]), scene.m_min), ext);   /* :276-278 */
      const float qq[3] = {q.x, q.y, q.z};
      kv[i] = std::make_pair(orc_morton_plain(qq), i);
    }
    std::stable_sort(kv.begin(), kv.end(), [](const std::pair<u32, u32>& a, const std::pair<u32, u32>& b) { return a.first < b.first; });
    std::vector<u32> sk(n), sv(n);
    for (u32 i = 0; i < n; i++) { sk[i] = kv[i].first; sv[i] = kv[i].second; }
    for (u32 g = 0; g < n; g++) { leaves[triOff + g].m_primIdx = sv[g]; leaves[triOff + g].m_aabb = box[sv[g]]; }
    if (n == 1) roots[it] = 0;
    else {
      std::vector<b2bvh_bvh2_node> tmp(2 * (size_t)n - 1);
      roots[it] = orc_lbvh_apetrei(T, sk.data(), sv.data(), n, tmp.data());
      for (u32 k = 0; k + 1 < n; k++) nodes[nodeOff + k] = tmp[k];
    }
    triOff += n; nodeOff += n - 1;
  }
}

/* ------------------------------------------------- SetupClusters
 * Ploc++Kernel.h:39-55 / HplocKernel.h:39-56.                               */
static void setup_clusters(const Box* triAabb, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves) {
  for (u32 g = 0; g < n; g++) { leaves[g].m_primIdx = vals[g]; leaves[g].m_aabb = triAabb[vals[g]]; }
  for (u32 g = 0; g + 1 < n; g++) { nodes[g].m_leftChildIdx = nodes[g].m_rightChildIdx = INVALID; nodes[g].m_aabb = box_empty(); }
}
static inline u64 nn_key(const Box& a, const Box& b, u32 idx) {  /* Ploc++Kernel.h:259-265 */
  Box u = b; box_grow(u, a);
  return ((u64)fbits(box_area(u)) << 32) | idx;
}

/* ------------------------------------------------- S6: PLOC++
 * Ploc (Ploc++Kernel.h:211-362), SinglePassPloc (:98-209), host loop
 * PLOC++Bvh.cpp:132-152.  One iteration = windowed (radius 8) nearest
 * neighbour by packed (area bits, index) min, mutual pairs merge, the lower
 * index keeps the slot, order-preserving compaction.  Merged node index =
 * C-2-rank (:311) with rank = number of merging clusters with a smaller index
 * (CANONICAL numbering; the kernel's rank is atomicAdd order, :57-68).
 * stats[0] = iterations, stats[1] = iterations with C >= 1024, stats[2..3] = sum of C (lo,hi). */
void orc_ploc(const Box* triAabb, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves, u32* stats) {
  const u32 nInt = n - 1;
  setup_clusters(triAabb, vals, n, nodes, leaves);
  std::vector<u32> id(n), id2(n);
  std::vector<Box> bx(n), bx2(n);
  std::vector<u64> nn(n);
  for (u32 g = 0; g < n; g++) { id[g] = g + nInt; bx[g] = leaves[g].m_aabb; }
  u32 C = n, iters = 0, bigIters = 0; u64 sumC = 0;
  const int R = B2BVH_PLOC_RADIUS;
  while (C > 1) {
    iters++; if (C >= B2BVH_PLOC_BLOCK) bigIters++; sumC += C;
    for (u32 c = 0; c < C; c++) nn[c] = ~0ull;
    for (u32 c = 0; c < C; c++)
      for (u32 j = c + 1; j < C && j <= c + (u32)R; j++) {
        Box u = bx[j]; box_grow(u, bx[c]);
        u64 a = (u64)fbits(box_area(u)) << 32;
        nn[c] = std::min(nn[c], a | j);
        nn[j] = std::min(nn[j], a | c);
      }
    u32 rank = 0, out = 0;
    for (u32 c = 0; c < C; c++) {
      u32 p = (u32)nn[c];
      bool mutual = ((u32)nn[p] == c);
      if (mutual && c > p) continue;                     /* absorbed by partner */
      if (mutual) {
        u32 m = C - 2 - rank++;
        Box b = bx[c]; box_grow(b, bx[p]);
        nodes[m].m_leftChildIdx = id[c]; nodes[m].m_rightChildIdx = id[p]; nodes[m].m_aabb = b;
        id2[out] = m; bx2[out] = b;
      } else { id2[out] = id[c]; bx2[out] = bx[c]; }
      out++;
    }
    id.swap(id2); bx.swap(bx2); C = out;
  }
  if (stats) { stats[0] = iters; stats[1] = bigIters; stats[2] = (u32)sumC; stats[3] = (u32)(sumC >> 32); }
}

/* ------------------------------------------------- S7: H-PLOC
 * HPloc (HplocKernel.h:257-315), plocMerge (:220-255), loadIndices (:192-206),
 * findNearestNeighbours (:83-117), mergeClusters (:126-190), storeIndices
 * (:208-218).  The LBVH hierarchy (findParent, :66-81) is the radix tree over
 * (key<<32 | index); a hierarchy node whose range is larger than 16 leaves, or
 * the root, PLOC-merges the <=16 leading cluster ids of each child range inside
 * a 32-slot list.  The memory contents of nodeIndices are simulated literally.
 *
 * CANONICAL numbering.  The kernel hands out node indices from a global
 * atomicAdd (:162-168), i.e. in warp-arrival order; topology and boxes do not
 * depend on it.  The canonical order used here needs no global counter, so the
 * GPU builder reproduces it without serialising: every cluster carries one
 * FREE index; leaf g (g >= 1) starts with g-1, leaf 0 with none.  When lanes
 * l < p merge, the new node takes the free index carried by p and the merged
 * cluster keeps the one carried by l (the cluster without an index contains
 * leaf 0, is always first in its list and therefore never a partner).  All
 * N-1 indices are used exactly once.  Finally the root is exchanged with the
 * node that received index 0, so that the root index is 0 as in the reference
 * (the last allocation of :167 is index 0).
 * stats[0] = number of plocMerge calls, stats[1] = nodes created.             */
struct HplocState {
  const u32* keys; const u64* keys64; /* one of the two: the reference's 32-bit codes, or the 60-bit variant */
  u32 n; b2bvh_bvh2_node* nodes; b2bvh_prim_ref* leaves;
  std::vector<u32> nodeIdx, freeIdx; u32 allocated; u32 calls;
};
static void hploc_merge(HplocState& S, u32 L, u32 R, u32 split, bool fin) {
  const u32 nInt = S.n - 1;
  u32 cl[32], fr[32]; Box bx[32]; u64 nn[32];
  for (int i = 0; i < 32; i++) { cl[i] = INVALID; fr[i] = INVALID; bx[i] = box_empty(); }
  auto load = [&](u32 start, u32 end, u32 offset) -> u32 {
    u32 cnt = std::min(end - start, 16u);
    for (u32 l = 0; l < cnt; l++) { cl[l + offset] = S.nodeIdx[start + l]; fr[l + offset] = S.freeIdx[start + l]; }
    u32 valid = 0; for (int i = 0; i < 32; i++) valid += cl[i] != INVALID;
    return std::min(cnt, valid - offset);
  };
  u32 nLeft = load(L, split, 0);
  u32 nRight = load(split, R + 1, nLeft);
  u32 np = nLeft + nRight;
  const u32 stored = np;
  u32 threshold = fin ? 1 : 16;
  for (int i = 0; i < 32; i++)
    if (cl[i] != INVALID) bx[i] = cl[i] >= nInt ? S.leaves[cl[i] - nInt].m_aabb : S.nodes[cl[i]].m_aabb;
  while (np > threshold) {
    for (int i = 0; i < 32; i++) nn[i] = ~0ull;
    for (u32 l = 0; l < np; l++)
      for (u32 r = 1; r <= (u32)B2BVH_PLOC_RADIUS; r++) {
        u32 j = l + r;
        if (j < 32 && j < np) {
          Box u = bx[j]; box_grow(u, bx[l]);
          u64 a = (u64)fbits(box_area(u)) << 32;
          nn[l] = std::min(nn[l], a | j);
          nn[j] = std::min(nn[j], a | l);
        }
      }
    u32 ncl[32], nfr[32]; Box nbx[32]; u32 out = 0;
    for (u32 l = 0; l < np; l++) {
      u32 p = (u32)nn[l];
      bool mutual = ((u32)nn[p] == l);
      if (mutual && l > p) continue;
      if (mutual) {
        u32 m = fr[p];                       /* canonical: the partner's free index */
        S.allocated++;
        Box b = bx[l]; box_grow(b, bx[p]);
        S.nodes[m].m_leftChildIdx = cl[l]; S.nodes[m].m_rightChildIdx = cl[p]; S.nodes[m].m_aabb = b;
        ncl[out] = m; nbx[out] = b;
      } else { ncl[out] = cl[l]; nbx[out] = bx[l]; }
      nfr[out] = fr[l];
      out++;
    }
    for (u32 l = 0; l < np; l++) { cl[l] = l < out ? ncl[l] : INVALID; fr[l] = l < out ? nfr[l] : INVALID; if (l < out) bx[l] = nbx[l]; }
    np = out;
  }
  for (u32 l = 0; l < stored; l++) { S.nodeIdx[L + l] = cl[l]; S.freeIdx[L + l] = fr[l]; }
  S.calls++;
}
static void hploc_visit(HplocState& S, u32 L, u32 R) {
  if (L == R) return;
  /* radix-tree split of [L,R] on the augmented key (key<<32 | index) */
  typedef unsigned __int128 u128;
  auto aug = [&](u32 i) -> u128 { return ((u128)(S.keys64 ? S.keys64[i] : (u64)S.keys[i]) << 32) | i; };
  const u128 x = aug(L) ^ aug(R);
  const u64 xh = (u64)(x >> 64), xl = (u64)x;
  int top = xh ? 127 - __builtin_clzll(xh) : 63 - __builtin_clzll(xl);
  u32 lo = L, hi = R;                 /* first index whose bit `top` is set */
  while (lo < hi) {
    u32 mid = lo + (hi - lo) / 2;
    if ((aug(mid) >> top) & 1) hi = mid; else lo = mid + 1;
  }
  u32 split = lo;
  hploc_visit(S, L, split - 1);
  hploc_visit(S, split, R);
  u32 size = R - L + 1;
  bool fin = size == S.n;
  if (size > 16 || fin) hploc_merge(S, L, R, split, fin);
}
static void hploc_run(const Box* triAabb, const u32* keys, const u64* keys64, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves, u32* stats);
void orc_hploc(const Box* triAabb, const u32* keys, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves, u32* stats) {
  hploc_run(triAabb, keys, nullptr, vals, n, nodes, leaves, stats);
}
/* the same walk over 64-bit sorted keys (60-bit Morton variant; no reference counterpart) */
void orc_hploc64(const Box* triAabb, const u64* keys64, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves, u32* stats) {
  hploc_run(triAabb, nullptr, keys64, vals, n, nodes, leaves, stats);
}
static void hploc_run(const Box* triAabb, const u32* keys, const u64* keys64, const u32* vals, u32 n, b2bvh_bvh2_node* nodes, b2bvh_prim_ref* leaves, u32* stats) {
  setup_clusters(triAabb, vals, n, nodes, leaves);
  HplocState S; S.keys = keys; S.keys64 = keys64; S.n = n; S.nodes = nodes; S.leaves = leaves; S.allocated = 0; S.calls = 0;
  S.nodeIdx.resize(n); S.freeIdx.resize(n);
  for (u32 g = 0; g < n; g++) { S.nodeIdx[g] = g + (n - 1); S.freeIdx[g] = g ? g - 1 : INVALID; }
  hploc_visit(S, 0, n - 1);
  /* root -> index 0 */
  u32 root = S.nodeIdx[0];
  if (root != 0) {
    std::swap(nodes[0], nodes[root]);       /* nodes[0] = root node, nodes[root] = the node that had index 0 */
    for (u32 i = 0; i + 1 < n; i++) {       /* its parent (possibly the root itself) must point at the new place */
      if (nodes[i].m_leftChildIdx == 0) nodes[i].m_leftChildIdx = root;
      else if (nodes[i].m_rightChildIdx == 0) nodes[i].m_rightChildIdx = root;
      else continue;
      break;
    }
  }
  if (stats) { stats[0] = S.calls; stats[1] = S.allocated; }
}

/* ------------------------------------------------- S5: collapse Bvh2 -> Bvh4
 * CollapseToWide4Bvh (TwoPassLbvhKernel.h:237-337; PLOC layout twin
 * Ploc++Kernel.h:364-465).  leaves == NULL: LBVH layout (leaf i is
 * nodes[n-1+i], prim index in m_leftChildIdx); otherwise PLOC layout (nodes has
 * n-1 entries, leaf prim index from leaves[]).  CANONICAL numbering = tasks
 * processed in index order, children allocated from a running counter that
 * starts at 1 (BFS).  wide must hold n-1 nodes (upper bound); returns the
 * wide-node count (= internalNodeOffset read back at TwoPassLbvh.cpp:187).     */
u32 orc_collapse4(const b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* leaves, u32 root, u32 n,
                  b2bvh_bvh4_node* wide, b2bvh_prim_node* wideLeaves) {
  const u32 nInt = n - 1;
  std::vector<u32> taskNode(nInt ? nInt : 1, INVALID), taskParent(nInt ? nInt : 1, INVALID);
  taskNode[0] = root; taskParent[0] = INVALID;
  u32 count = 1;
  for (u32 g = 0; g < count; g++) {
    const b2bvh_bvh2_node& n2 = nodes[taskNode[g]];
    u32 ch[4] = {n2.m_leftChildIdx, n2.m_rightChildIdx, INVALID, INVALID};
    u32 cc = 2;
    for (int pass = 0; pass < 2; pass++) {                 /* :270-296 */
      float best = 0.0f; u32 pos = INVALID;
      for (u32 k = 0; k < cc; k++)
        if (ch[k] < nInt) { float a = box_area(nodes[ch[k]].m_aabb); if (a > best) { best = a; pos = k; } }
      if (pos == INVALID) break;
      const b2bvh_bvh2_node& mc = nodes[ch[pos]];
      ch[pos] = mc.m_leftChildIdx; ch[cc++] = mc.m_rightChildIdx;
    }
    b2bvh_bvh4_node w;
    memset(&w, 0, sizeof(w));
    for (int k = 0; k < 4; k++) { w.m_aabb[k] = box_empty(); w.m_child[k] = INVALID; }
    w.m_parent = taskParent[g]; w.m_childCount = cc;
    for (u32 k = 0; k < cc; k++) {
      if (ch[k] < nInt) {                                  /* :312-319 */
        u32 id = count++;
        w.m_child[k] = id; w.m_aabb[k] = nodes[ch[k]].m_aabb;
        taskNode[id] = ch[k]; taskParent[id] = g;
      } else {                                             /* :320-325 */
        w.m_child[k] = ch[k];
        u32 slot = ch[k] - nInt;
        wideLeaves[slot].m_parent = g;
        wideLeaves[slot].m_primIdx = leaves ? leaves[slot].m_primIdx : nodes[ch[k]].m_leftChildIdx;
      }
    }
    wide[g] = w;
  }
  return count;
}

/* ------------------------------------------------- costs (host, float accumulate in index order)
 * calculatebvh4Cost Utility.cpp:351-396; calculateLbvhCost :317-349;
 * calculateBinnedSahBvhCost :398-422 (quirks kept).            */
float orc_cost_bvh4(const b2bvh_bvh4_node* w, const b2bvh_prim_node* wl, const Box* primAabbs, u32 root, u32 total, u32 nInt) {
  Box rb = box_empty();
  for (int k = 0; k < 4; k++) if (w[root].m_child[k] != INVALID) box_grow(rb, w[root].m_aabb[k]);
  const float inv = 1.0f / box_area(rb);
  float cost = 0.0f;
  cost += 1.0f;
  for (u32 i = 0; i < total; i++)
    for (int k = 0; k < 4; k++)
      if (w[i].m_child[k] != INVALID && w[i].m_child[k] < nInt) cost += 1.0f * box_area(w[i].m_aabb[k]) * inv;
  for (u32 i = 0; i < nInt + 1; i++) cost += box_area(primAabbs[wl[i].m_primIdx]) * inv;
  return cost;
}
float orc_cost_lbvh(const b2bvh_bvh2_node* nodes, u32 root, u32 nLeaf, u32 nInt) {
  const float inv = 1.0f / box_area(nodes[root].m_aabb);
  float cost = 0.0f;
  cost += 1.0f;
  for (u32 i = 0; i < nInt; i++) {
    if (nodes[i].m_leftChildIdx != INVALID) cost += 1.0f * box_area(nodes[nodes[i].m_leftChildIdx].m_aabb) * inv;
    if (nodes[i].m_rightChildIdx != INVALID) cost += 1.0f * box_area(nodes[nodes[i].m_rightChildIdx].m_aabb) * inv;
  }
  for (u32 i = nInt; i < nLeaf + nInt; i++)
    if (nodes[i].m_leftChildIdx != INVALID) cost += 1.0f * box_area(nodes[i].m_aabb) * inv;
  return cost;
}
/* Bvh2 cost for the separate-leaf (PLOC/HPLOC) layout, same formula. */
float orc_cost_bvh2_ploc(const b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* leaves, u32 root, u32 n) {
  const u32 nInt = n - 1;
  const float inv = 1.0f / box_area(nodes[root].m_aabb);
  auto area_of = [&](u32 id) { return id >= nInt ? box_area(leaves[id - nInt].m_aabb) : box_area(nodes[id].m_aabb); };
  float cost = 0.0f;
  cost += 1.0f;
  for (u32 i = 0; i < nInt; i++) { cost += area_of(nodes[i].m_leftChildIdx) * inv; cost += area_of(nodes[i].m_rightChildIdx) * inv; }
  for (u32 i = 0; i < n; i++) cost += box_area(leaves[i].m_aabb) * inv;
  return cost;
}
float orc_cost_binned_sah_quirk(const b2bvh_sah_node* nodes, u32 root, u32 total) {
  const float inv = 1.0f / box_area(nodes[root].m_aabb);
  float cost = 0.0f;
  cost += 1.0f;
  for (u32 i = 0; i < total; i++) {
    if (nodes[i].m_firstChildIdx != INVALID) { u32 l = nodes[i].m_firstChildIdx; cost += 1.0f * box_area(nodes[l].m_aabb) * inv; }
    if (nodes[i].m_firstChildIdx + 1 != INVALID) { u32 r = nodes[i].m_firstChildIdx + 1; cost += 1.0f * box_area(nodes[r].m_aabb) * inv; }
  }
  return cost;
}
/* proper SAH of the binned tree: 1 + sum over non-root nodes area/rootArea (Ct = Ci = 1) */
float orc_cost_binned_sah_proper(const b2bvh_sah_node* nodes, u32 total) {
  const float inv = 1.0f / box_area(nodes[0].m_aabb);
  float cost = 0.0f;
  cost += 1.0f;
  for (u32 i = 0; i < total; i++)
    if (nodes[i].m_primCount == 0) {
      cost += box_area(nodes[nodes[i].m_firstChildIdx].m_aabb) * inv;
      cost += box_area(nodes[nodes[i].m_firstChildIdx + 1].m_aabb) * inv;
    } else cost += box_area(nodes[i].m_aabb) * inv;
  return cost;
}

/* ------------------------------------------------- validators (Utility.cpp:15-159), with stacks deep enough */
int orc_check_root_aabb(const b2bvh_bvh2_node* nodes, u32 root, u32 nLeaf, u32 nInt) {
  Box b = box_empty();
  for (u32 i = 0; i < nLeaf; i++) box_grow(b, nodes[nInt + i].m_aabb);
  return memcmp(&b, &nodes[root].m_aabb, sizeof(Box)) == 0;
}
static int perm_ok(std::vector<u32>& p, u32 n) {
  if (p.size() != n) return 0;
  std::sort(p.begin(), p.end());
  for (u32 i = 0; i < n; i++) if (p[i] != i) return 0;
  return 1;
}
int orc_check_bvh2(const b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* leaves, u32 root, u32 n) {
  const u32 nInt = n - 1;
  std::vector<u32> prims, stack; stack.push_back(root);
  u64 guard = 0;
  while (!stack.empty()) {
    u32 id = stack.back(); stack.pop_back();
    if (++guard > 4ull * n) return 0;
    if (id >= nInt) { if (id - nInt >= n) return 0; prims.push_back(leaves ? leaves[id - nInt].m_primIdx : nodes[id].m_leftChildIdx); }
    else { stack.push_back(nodes[id].m_leftChildIdx); stack.push_back(nodes[id].m_rightChildIdx); }
  }
  return perm_ok(prims, n);
}
int orc_check_bvh4(const b2bvh_bvh4_node* w, const b2bvh_prim_node* wl, u32 root, u32 nInt) {
  std::vector<u32> prims, stack; stack.push_back(root);
  u64 guard = 0;
  while (!stack.empty()) {
    u32 id = stack.back(); stack.pop_back();
    if (++guard > 4ull * (nInt + 1)) return 0;
    if (id >= nInt) prims.push_back(wl[id - nInt].m_primIdx);
    else for (int k = 0; k < 4; k++) if (w[id].m_child[k] != INVALID) stack.push_back(w[id].m_child[k]);
  }
  return perm_ok(prims, nInt + 1);
}
/* max leaf depth of a Bvh2 (root depth 0) */
u32 orc_bvh2_depth(const b2bvh_bvh2_node* nodes, u32 root, u32 n) {
  const u32 nInt = n - 1;
  std::vector<std::pair<u32, u32>> st; st.push_back({root, 0});
  u32 best = 0;
  while (!st.empty()) {
    auto [id, d] = st.back(); st.pop_back();
    if (id >= nInt) { best = std::max(best, d); continue; }
    st.push_back({nodes[id].m_leftChildIdx, d + 1}); st.push_back({nodes[id].m_rightChildIdx, d + 1});
  }
  return best;
}

/* ------------------------------------------------- rays + traversal
 * GenerateRays CommonBlocksKernel.h:432-463 (quaternion helpers Common.h:461-514). */
static inline b2bvh_float4 q4(float x, float y, float z, float w) { b2bvh_float4 q = {x, y, z, w}; return q; }
static b2bvh_float4 qt_mul(b2bvh_float4 a, b2bvh_float4 b) {   /* Common.h:483-492 */
  F3 c = vcross(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z));
  b2bvh_float4 r;
  r.x = (c.x + a.w * b.x) + b.w * a.x;
  r.y = (c.y + a.w * b.y) + b.w * a.y;
  r.z = (c.z + a.w * b.z) + b.w * a.z;
  r.w = a.w * b.w - vdot(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z));
  return r;
}
static inline b2bvh_float4 qt_inv(b2bvh_float4 q) { return q4(-q.x, -q.y, -q.z, q.w); }
static F3 qt_rotate(b2bvh_float4 q, F3 p) {                    /* Common.h:502-508 */
  b2bvh_float4 o = qt_mul(qt_mul(q, q4(p.x, p.y, p.z, 0.0f)), qt_inv(q));
  return f3(o.x, o.y, o.z);
}
static inline F3 qt_inv_rotate(b2bvh_float4 q, F3 v) { return qt_rotate(qt_inv(q), v); }
static inline F3 inv_transform(F3 p, F3 s, b2bvh_float4 r, F3 t) { return vdiv(qt_inv_rotate(r, vsub(p, t)), s); }  /* :512 */
static inline F3 fwd_transform(F3 p, F3 s, b2bvh_float4 r, F3 t) { return vadd(qt_rotate(r, vmul(s, p)), t); }      /* :514 */

void orc_qt_rotation(const float axisAngle[4], float out[4]) {  /* qtRotation, Common.h:461-472 */
  F3 ax = vnormalize(f3(axisAngle[0], axisAngle[1], axisAngle[2]));
  float ang = axisAngle[3];
  out[0] = ax.x * sinf(ang / 2.0f); out[1] = ax.y * sinf(ang / 2.0f); out[2] = ax.z * sinf(ang / 2.0f); out[3] = cosf(ang / 2.0f);
}

/* zDir = sensorSize.y / (2 tan(fov/2)) is uniform; it is evaluated once on the
 * host (tanf) and handed to the GPU kernel so both sides use the same bits.   */
float orc_ray_zdir(float fov) { return 0.024f / (2.f * tanf(fov / 2.f)); }

void orc_generate_rays(const b2bvh_camera* cam, u32 width, u32 height, b2bvh_ray* rays) {
  const float sx = 0.024f * (width / (float)height), sy = 0.024f;
  const float zd = orc_ray_zdir(cam->m_fov);
  const F3 hol = qt_rotate(cam->m_quat, f3(1.0f, 0.0f, 0.0f));
  const F3 up = qt_rotate(cam->m_quat, f3(0.0f, -1.0f, 0.0f));
  const F3 view = qt_rotate(cam->m_quat, f3(0.0f, 0.0f, -1.0f));
  for (u32 gx = 0; gx < width; gx++)
    for (u32 gy = 0; gy < height; gy++) {
      float px = ((float)gx + 0.5f) / width - 0.5f, py = ((float)gy + 0.5f) / height - 0.5f;
      F3 d = f3(px * sx, py * sy, zd);
      F3 dir = vnormalize(vadd(vadd(vscale(hol, d.x), vscale(up, d.y)), vscale(view, d.z)));
      b2bvh_ray& r = rays[gx * height + gy];
      r.m_origin = f3(cam->m_eye.x, cam->m_eye.y, cam->m_eye.z);
      F3 far = f3(cam->m_eye.x + dir.x * cam->m_far, cam->m_eye.y + dir.y * cam->m_far, cam->m_eye.z + dir.z * cam->m_far);
      r.m_direction = vnormalize(far);
      r.m_tMin = 0.0f; r.m_tMax = FLTMAX;
    }
}

static inline void slab(const Box& b, F3 from, F3 inv, float maxt, float& tn, float& tf) {   /* Aabb::intersect, Common.h:384-397 */
  F3 dF = vmul(vsub(b.m_max, from), inv), dN = vmul(vsub(b.m_min, from), inv);
  F3 tF = vmax(dF, dN), tN = vmin(dF, dN);
  float minFar = fminr(tF.x, fminr(tF.y, tF.z));
  float maxNear = fmaxr(tN.x, fmaxr(tN.y, tN.z));
  tf = fminr(maxt, minFar); tn = fmaxr(0.0f, maxNear);
}
static inline void tri_hit(F3 v0, F3 v1, F3 v2, F3 o, F3 d, float out[4]) {   /* intersectTriangle, Common.h:516-531 */
  F3 p0 = vsub(v0, o), p1 = vsub(v1, o), p2 = vsub(v2, o);
  F3 e0 = vsub(v2, v0), e1 = vsub(v0, v1), e2 = vsub(v1, v2);
  F3 nrm = vcross(e1, e0);
  float u = vdot(vcross(vadd(p0, p2), e0), d);
  float v = vdot(vcross(vadd(p1, p0), e1), d);
  float w = vdot(vcross(vadd(p2, p1), e2), d);
  float t = vdot(p0, nrm) * 2.0f;
  float den = vdot(nrm, d) * 2.0f;
  out[0] = u / den; out[1] = v / den; out[2] = w / den; out[3] = t / den;
}

/* Closest-hit traversal, Utility::TraversalLbvhCPU (Utility.cpp:161-237) ==
 * BvhTraversalWhile (TraversalKernel.h:238-335): 64-entry stack, far child is
 * dropped when the stack is full.  leaves == NULL: LBVH layout.  Ray index is
 * gIdx*width+gIdy (:246).  Writes one HitInfo per ray; returns the hit count. */
u32 orc_traverse(const b2bvh_ray* rays, const b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* leaves, const b2bvh_triangle* tris,
                 const b2bvh_transform* tr, u32 root, u32 nInt, u32 nRays, b2bvh_hit* hits) {
  u32 nHit = 0;
  for (u32 idx = 0; idx < nRays; idx++) {
    const b2bvh_ray& ray = rays[idx];
    F3 o = inv_transform(ray.m_origin, tr->m_scale, tr->m_quat, tr->m_translation);
    F3 d = inv_transform(ray.m_direction, tr->m_scale, tr->m_quat, f3(0, 0, 0));
    F3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    b2bvh_hit hit; hit.m_primIdx = INVALID; hit.m_t = FLTMAX; hit.m_u = 0; hit.m_v = 0;
    u32 stack[64]; u32 top = 0; stack[top++] = INVALID;
    u32 node = root;
    while (node != INVALID) {
      if (node >= nInt) {
        u32 prim = leaves ? leaves[node - nInt].m_primIdx : nodes[node].m_leftChildIdx;
        const b2bvh_triangle& t = tris[prim];
        F3 a = fwd_transform(t.v1, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 b = fwd_transform(t.v2, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 c = fwd_transform(t.v3, tr->m_scale, tr->m_quat, tr->m_translation);
        float r[4]; tri_hit(a, b, c, ray.m_origin, ray.m_direction, r);
        if (r[0] > 0.0f && r[1] > 0.0f && r[2] > 0.0f && r[3] > 0.0f && r[3] < hit.m_t) { hit.m_primIdx = prim; hit.m_t = r[3]; hit.m_u = r[0]; hit.m_v = r[1]; }
      } else {
        u32 l = nodes[node].m_leftChildIdx, r = nodes[node].m_rightChildIdx;
        const Box& lb = l >= nInt && leaves ? leaves[l - nInt].m_aabb : nodes[l].m_aabb;
        const Box& rb = r >= nInt && leaves ? leaves[r - nInt].m_aabb : nodes[r].m_aabb;
        float n0, f0, n1, f1; slab(lb, o, inv, hit.m_t, n0, f0); slab(rb, o, inv, hit.m_t, n1, f1);
        bool hl = n0 <= f0, hr = n1 <= f1;
        if (hl || hr) {
          if (hl && hr) { node = n0 < n1 ? l : r; if (top < 64) stack[top++] = n0 < n1 ? r : l; }
          else node = hl ? l : r;
          continue;
        }
      }
      node = stack[--top];
    }
    hits[idx] = hit;
    nHit += hit.m_primIdx != INVALID;
  }
  return nHit;
}

/* The same closest-hit query with the reference's other two Bvh2 kernels, plus the per-ray leaf-test counter they keep
 * (rayCounter, TraversalKernel.h:88,192) and the 4-wide traversal this project adds.
 *   kind 0: BvhTraversalifif (TraversalKernel.h:148-236) — one node per step; same visiting order as orc_traverse.
 *   kind 1: BvhTraversalRestartTrail (:49-146) with pop() (:32-47) — stackless, restarts from the root (the reference
 *           restarts from node 0, :44, which is the root only for TwoPassLbvh; rootIdx is used here), near child on a tie
 *           is the LEFT one (:112-116; the stack kernels take the right one, :219).  Trees deeper than 63 are not supported
 *           by a 64-bit trail.
 * counter may be NULL.  Returns the hit count. */
u32 orc_traverse_kind(u32 kind, const b2bvh_ray* rays, const b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* leaves, const b2bvh_triangle* tris,
                      const b2bvh_transform* tr, u32 root, u32 nInt, u32 nRays, b2bvh_hit* hits, u32* counter) {
  if (kind == 0) {
    u32 nHit = orc_traverse(rays, nodes, leaves, tris, tr, root, nInt, nRays, hits);
    if (counter) { /* leaf tests per ray: replay the walk, counting */
      for (u32 idx = 0; idx < nRays; idx++) {
        const b2bvh_ray& ray = rays[idx];
        F3 o = inv_transform(ray.m_origin, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 d = inv_transform(ray.m_direction, tr->m_scale, tr->m_quat, f3(0, 0, 0));
        F3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        float ht = FLTMAX;
        u32 stack[64]; u32 top = 0; stack[top++] = INVALID;
        u32 node = root, cnt = 0;
        while (node != INVALID) {
          if (node >= nInt) {
            u32 prim = leaves ? leaves[node - nInt].m_primIdx : nodes[node].m_leftChildIdx;
            const b2bvh_triangle& t = tris[prim];
            F3 a = fwd_transform(t.v1, tr->m_scale, tr->m_quat, tr->m_translation);
            F3 b = fwd_transform(t.v2, tr->m_scale, tr->m_quat, tr->m_translation);
            F3 c = fwd_transform(t.v3, tr->m_scale, tr->m_quat, tr->m_translation);
            float r[4]; tri_hit(a, b, c, ray.m_origin, ray.m_direction, r);
            cnt++;
            if (r[0] > 0.0f && r[1] > 0.0f && r[2] > 0.0f && r[3] > 0.0f && r[3] < ht) ht = r[3];
          } else {
            u32 l = nodes[node].m_leftChildIdx, r = nodes[node].m_rightChildIdx;
            const Box& lb = l >= nInt && leaves ? leaves[l - nInt].m_aabb : nodes[l].m_aabb;
            const Box& rb = r >= nInt && leaves ? leaves[r - nInt].m_aabb : nodes[r].m_aabb;
            float n0, f0, n1, f1; slab(lb, o, inv, ht, n0, f0); slab(rb, o, inv, ht, n1, f1);
            bool hl = n0 <= f0, hr = n1 <= f1;
            if (hl || hr) {
              if (hl && hr) { node = n0 < n1 ? l : r; if (top < 64) stack[top++] = n0 < n1 ? r : l; }
              else node = hl ? l : r;
              continue;
            }
          }
          node = stack[--top];
        }
        counter[idx] = cnt;
      }
    }
    return nHit;
  }
  u32 nHit = 0;
  const unsigned long long TOP = 0x8000000000000000ull;
  for (u32 idx = 0; idx < nRays; idx++) {
    const b2bvh_ray& ray = rays[idx];
    F3 o = inv_transform(ray.m_origin, tr->m_scale, tr->m_quat, tr->m_translation);
    F3 d = inv_transform(ray.m_direction, tr->m_scale, tr->m_quat, f3(0, 0, 0));
    F3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    b2bvh_hit hit; hit.m_primIdx = INVALID; hit.m_t = FLTMAX; hit.m_u = 0; hit.m_v = 0;
    unsigned long long trail = TOP, level = TOP, popLevel = 0;
    u32 node = root, cnt = 0;
    bool done = false;
    auto pop = [&]() -> bool {   /* pop(), TraversalKernel.h:32-47 */
      trail &= (0ull - level);
      trail += level;
      unsigned long long temp = trail >> 1;
      level = (((temp - 1) ^ temp) + 1);
      if (!(trail & TOP)) return true;
      popLevel = level;
      node = root;
      level = TOP;
      return false;
    };
    while (!done) {
      if (node >= nInt) {
        u32 prim = leaves ? leaves[node - nInt].m_primIdx : nodes[node].m_leftChildIdx;
        const b2bvh_triangle& t = tris[prim];
        F3 a = fwd_transform(t.v1, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 b = fwd_transform(t.v2, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 c = fwd_transform(t.v3, tr->m_scale, tr->m_quat, tr->m_translation);
        float r[4]; tri_hit(a, b, c, ray.m_origin, ray.m_direction, r);
        cnt++;
        if (r[0] > 0.0f && r[1] > 0.0f && r[2] > 0.0f && r[3] > 0.0f && r[3] < hit.m_t) { hit.m_primIdx = prim; hit.m_t = r[3]; hit.m_u = r[0]; hit.m_v = r[1]; }
        done = pop();
      } else {
        u32 l = nodes[node].m_leftChildIdx, r = nodes[node].m_rightChildIdx;
        const Box& lb = l >= nInt && leaves ? leaves[l - nInt].m_aabb : nodes[l].m_aabb;
        const Box& rb = r >= nInt && leaves ? leaves[r - nInt].m_aabb : nodes[r].m_aabb;
        float n0, f0, n1, f1; slab(lb, o, inv, hit.m_t, n0, f0); slab(rb, o, inv, hit.m_t, n1, f1);
        bool hl = n0 <= f0, hr = n1 <= f1;
        if (hl || hr) {
          if (hl && hr) {
            u32 nearC = l, farC = r;
            if (n0 > n1) { nearC = r; farC = l; }
            level >>= 1;
            node = (trail & level) ? farC : nearC;
          } else {
            level >>= 1;
            if (level != popLevel) { trail |= level; node = hr ? r : l; }
            else done = pop();
          }
        } else done = pop();
      }
    }
    hits[idx] = hit;
    if (counter) counter[idx] = cnt;
    nHit += hit.m_primIdx != INVALID;
  }
  return nHit;
}

/* Closest hit through the 4-wide tree (the reference builds it, TwoPassLbvh.cpp:154-183, but has no kernel that walks it;
 * defined here and mirrored by traverse_wide4_kernel).  Visiting a wide node: (1) its leaf children are handled in slot
 * order — a wide node stores no box for them (TwoPassLbvhKernel.h:320-325), so box and primitive come from the Bvh2 leaf
 * record (LBVH layout: nodes[c]; separate-leaf layout: leaves[c - nInt]); the triangle is intersected when the slab test
 * with the current hit distance passes; (2) the boxes of its internal children are slab-tested with the hit distance
 * after (1); (3) the children hit are visited nearest first (entry distance, slot order among equals): the nearest is
 * next, the others are pushed far to near on a 128-entry stack (dropped when full).  Same triangle test, transform and
 * acceptance rule as the Bvh2 kernels.  counter = triangle tests per ray. */
u32 orc_traverse_wide4(const b2bvh_ray* rays, const b2bvh_bvh4_node* wide, const b2bvh_bvh2_node* nodes, const b2bvh_prim_ref* leaves,
                       const b2bvh_triangle* tris, const b2bvh_transform* tr, u32 nInt, u32 nRays, b2bvh_hit* hits, u32* counter) {
  u32 nHit = 0;
  for (u32 idx = 0; idx < nRays; idx++) {
    const b2bvh_ray& ray = rays[idx];
    F3 o = inv_transform(ray.m_origin, tr->m_scale, tr->m_quat, tr->m_translation);
    F3 d = inv_transform(ray.m_direction, tr->m_scale, tr->m_quat, f3(0, 0, 0));
    F3 inv = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    b2bvh_hit hit; hit.m_primIdx = INVALID; hit.m_t = FLTMAX; hit.m_u = 0; hit.m_v = 0;
    u32 stack[128]; u32 top = 0; stack[top++] = INVALID;
    u32 node = 0, cnt = 0;
    while (node != INVALID) {
      const b2bvh_bvh4_node& w = wide[node];
      for (int k = 0; k < 4; k++) {
        const u32 c = w.m_child[k];
        if (c == INVALID || c < nInt) continue;
        const Box& lb = leaves ? leaves[c - nInt].m_aabb : nodes[c].m_aabb;
        float n0, f0; slab(lb, o, inv, hit.m_t, n0, f0);
        if (!(n0 <= f0)) continue;
        const u32 prim = leaves ? leaves[c - nInt].m_primIdx : nodes[c].m_leftChildIdx;
        const b2bvh_triangle& t = tris[prim];
        F3 a = fwd_transform(t.v1, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 b = fwd_transform(t.v2, tr->m_scale, tr->m_quat, tr->m_translation);
        F3 cc = fwd_transform(t.v3, tr->m_scale, tr->m_quat, tr->m_translation);
        float r[4]; tri_hit(a, b, cc, ray.m_origin, ray.m_direction, r);
        cnt++;
        if (r[0] > 0.0f && r[1] > 0.0f && r[2] > 0.0f && r[3] > 0.0f && r[3] < hit.m_t) { hit.m_primIdx = prim; hit.m_t = r[3]; hit.m_u = r[0]; hit.m_v = r[1]; }
      }
      u32 ids[4]; float tn[4]; int m = 0;
      for (int k = 0; k < 4; k++) {
        const u32 c = w.m_child[k];
        if (c == INVALID || c >= nInt) continue;
        float n0, f0; slab(w.m_aabb[k], o, inv, hit.m_t, n0, f0);
        if (!(n0 <= f0)) continue;
        int pos = m++;   /* insertion by entry distance, stable in slot order */
        while (pos > 0 && tn[pos - 1] > n0) { tn[pos] = tn[pos - 1]; ids[pos] = ids[pos - 1]; pos--; }
        tn[pos] = n0; ids[pos] = c;
      }
      if (m == 0) { node = stack[--top]; continue; }
      for (int k = m - 1; k >= 1; k--) if (top < 128) stack[top++] = ids[k];
      node = ids[0];
    }
    hits[idx] = hit;
    if (counter) counter[idx] = cnt;
    nHit += hit.m_primIdx != INVALID;
  }
  return nHit;
}

/* Utility::generateTraversalHeatMap (Utility.cpp:424-454) without the PNG write: rgba[index] = (c/max*150, c/max*255, 255, 255). */
void orc_heat_map(const u32* counter, u32 count, unsigned char* rgba) {
  u32 mx = 0;
  for (u32 i = 0; i < count; i++) if (counter[i] > mx) mx = counter[i];
  for (u32 i = 0; i < count; i++) {
    rgba[i * 4 + 0] = (unsigned char)((counter[i] / (float)mx) * 150);
    rgba[i * 4 + 1] = (unsigned char)((counter[i] / (float)mx) * 255);
    rgba[i * 4 + 2] = 255;
    rgba[i * 4 + 3] = 255;
  }
}

/* ------------------------------------------------- CPU baseline: binned SAH
 * SahBvh::build, BinnedSahBvh.cpp:13-204.  Faithful: BFS queue, 32 buckets,
 * O(32^2) sweep, min over buckets 0..30, std::partition / std::nth_element over
 * [start, end-1) (the last element is never moved), the two fallbacks.  One
 * deviation, stated: the bucket index is clamped to 31 when counting (:101-104
 * writes out of bounds when a centroid sits on the max face; :151 clamps).
 * nodes must hold 3n-1 entries (:40).  Returns the number of nodes used.       */
struct SahRef { Box box; size_t prim; };
struct SahBucket { int count; Box box; };
u32 orc_binned_sah_build(const b2bvh_triangle* tris, u32 n, b2bvh_sah_node* nodes) {
  std::vector<SahRef> refs; refs.reserve(n);
  for (u32 i = 0; i < n; i++) refs.push_back({tri_box(tris[i]), i});
  const u32 NB = 32;
  struct Task { u32 node, start, end; };
  std::queue<Task> q; q.push({0, 0, n});
  u32 next = 0;
  size_t cap = (2 * (size_t)n - 1) + n;
  for (size_t i = 0; i < cap; i++) { nodes[i].m_aabb = box_empty(); nodes[i].m_firstChildIdx = 0; nodes[i].m_primCount = 0; }
  next++;
  auto centroid_dim = [](const Box& b, int dim) { F3 c = box_center(b); return dim == 0 ? c.x : dim == 1 ? c.y : c.z; };
  auto comp = [](F3 v, int dim) { return dim == 0 ? v.x : dim == 1 ? v.y : v.z; };
  while (!q.empty()) {
    Task t = q.front(); q.pop();
    b2bvh_sah_node& node = nodes[t.node];
    if (t.end - t.start == 1) {
      node.m_aabb = refs[t.start].box; node.m_firstChildIdx = (u32)refs[t.start].prim; node.m_primCount = 1;
      continue;
    }
    Box nb = box_empty();
    for (u32 i = t.start; i < t.end; i++) box_grow(nb, refs[i].box);
    node.m_aabb = nb;
    int dim = box_max_dim(nb);
    node.m_firstChildIdx = next++; node.m_primCount = 0; next++;
    u32 split = 0;
    auto by_centroid = [&](const SahRef& a, const SahRef& b) { return centroid_dim(a.box, dim) < centroid_dim(b.box, dim); };
    if (t.end - t.start <= 2) {
      split = (t.start + t.end) / 2;
      std::nth_element(&refs[t.start], &refs[split], &refs[t.end - 1], by_centroid);
    } else {
      SahBucket buckets[NB];
      for (u32 b = 0; b < NB; b++) { buckets[b].count = 0; buckets[b].box = box_empty(); }
      auto bucket_of = [&](const SahRef& r) { u32 b = (u32)(NB * comp(box_offset(nb, box_center(r.box)), dim)); return b >= NB ? NB - 1 : b; };
      for (u32 i = t.start; i < t.end; i++) { u32 b = bucket_of(refs[i]); buckets[b].count++; box_grow(buckets[b].box, refs[i].box); }
      float cost[NB];
      for (u32 b = 0; b < NB; b++) {
        Box lh = box_empty(), rh = box_empty(); int lc = 0, rc = 0;
        for (u32 j = 0; j <= b; j++) if (buckets[j].count) { box_grow(lh, buckets[j].box); lc += buckets[j].count; }
        for (u32 j = b + 1; j < NB; j++) if (buckets[j].count) { box_grow(rh, buckets[j].box); rc += buckets[j].count; }
        float ls = lc == 0 ? 0.0f : lc * box_area(lh), rs = rc == 0 ? 0.0f : rc * box_area(rh);
        float tot = (lc + rc) == 0 ? 0.0f : ((ls + rs) / box_area(nb));
        cost[b] = tot == 0.0f ? FLTMAX : 0.125f + tot;
      }
      float best = cost[0]; int sb = 0;
      for (u32 i = 0; i < NB - 1; i++) if (cost[i] < best) { best = cost[i]; sb = (int)i; }
      split = (u32)(std::partition(&refs[t.start], &refs[t.end - 1], [&](const SahRef& r) { return (int)bucket_of(r) <= sb; }) - &refs[0]);
      if (split <= t.start || split >= t.end) {
        float mid = comp(box_offset(nb, box_center(nb)), dim) / 2.0f;
        split = (u32)(std::partition(&refs[t.start], &refs[t.end - 1], [&](const SahRef& r) { return centroid_dim(r.box, dim) < mid; }) - &refs[0]);
      }
      if (split <= t.start || split >= t.end) {
        split = (t.start + t.end) / 2;
        std::nth_element(&refs[t.start], &refs[split], &refs[t.end - 1], by_centroid);
      }
    }
    q.push({node.m_firstChildIdx, t.start, split});
    q.push({node.m_firstChildIdx + 1, split, t.end});
  }
  return next;
}
int orc_check_sah(const b2bvh_sah_node* nodes, u32 n) {   /* checkSahCorrectness, Utility.cpp:132-159 */
  std::vector<u32> prims, st; st.push_back(0);
  while (!st.empty()) {
    u32 id = st.back(); st.pop_back();
    if (nodes[id].m_primCount != 0) prims.push_back(nodes[id].m_firstChildIdx);
    else { st.push_back(nodes[id].m_firstChildIdx); st.push_back(nodes[id].m_firstChildIdx + 1); }
    if (prims.size() > n) return 0;
  }
  return perm_ok(prims, n);
}

/* ------------------------------------------------- synthetic input synth_uniform_v1
 * RNG = the reference's tea<16>/lcg/randf (CommonBlocksKernel.h:401-430).      */
static inline u32 lcg_next(u32& s) { s = 1103515245u * s + 12345u; return s & 0x00FFFFFFu; }
static inline float rand01(u32& s) { return (float)lcg_next(s) / (float)0x01000000; }
static inline u32 tea16(u32 v0, u32 v1) {
  u32 s0 = 0;
  for (int r = 0; r < 16; r++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
void orc_synth_uniform(u64 first, u32 count, u32 seed, float half, b2bvh_triangle* out) {
  const float two_h = 2 * half;
  for (u32 k = 0; k < count; k++) {
    u32 i = (u32)(first + k);
    u32 s = tea16(i, seed);
    float r[12];
    for (int j = 0; j < 12; j++) r[j] = rand01(s);
    float c[3];
    for (int j = 0; j < 3; j++) { float m = 2000.0f * r[j]; c[j] = -1000.0f + m; }
    float v[9];
    for (int j = 0; j < 9; j++) { float a = r[3 + j] - 0.5f; float m = a * two_h; v[j] = c[j % 3] + m; }
    memset(&out[k], 0, sizeof(b2bvh_triangle));
    out[k].v1 = f3(v[0], v[1], v[2]); out[k].v2 = f3(v[3], v[4], v[5]); out[k].v3 = f3(v[6], v[7], v[8]);
  }
}

/* synth_clustered_v1: see b2bvh_synth_clustered (include/b2bvh.h) — cluster j = floor(r0 * 4096), centre from tea16(j, seed ^ 0xC1),
 * triangle centre = that + 20 * (r - 0.5), vertices as above. */
void orc_synth_clustered(u64 first, u32 count, u32 seed, float half, b2bvh_triangle* out) {
  const float two_h = 2 * half;
  for (u32 k = 0; k < count; k++) {
    u32 s = tea16((u32)(first + k), seed);
    float r0 = rand01(s);
    u32 j = (u32)(r0 * 4096.0f);
    u32 sj = tea16(j, seed ^ 0xC1u);
    float c[3];
    for (int a = 0; a < 3; a++) { float m = 2000.0f * rand01(sj); c[a] = -1000.0f + m; }
    for (int a = 0; a < 3; a++) { float d = rand01(s) - 0.5f; float m = 20.0f * d; c[a] = c[a] + m; }
    float v[9];
    for (int a = 0; a < 9; a++) { float d = rand01(s) - 0.5f; float m = d * two_h; v[a] = c[a % 3] + m; }
    memset(&out[k], 0, sizeof(b2bvh_triangle));
    out[k].v1 = f3(v[0], v[1], v[2]); out[k].v2 = f3(v[3], v[4], v[5]); out[k].v3 = f3(v[6], v[7], v[8]);
  }
}

/* ------------------------------------------------- sharded build: top-level tree over G sub-tree roots
 * (new work; no reference counterpart).  Spec: Morton-code the
 * root boxes with the same extended code in the frame of their union, stable
 * sort, Karras LBVH + refit over the G leaves.  Leaf g's m_leftChildIdx = rank.
 * nodes must hold 2G-1 entries.  G == 1: nodes[0] is the single leaf.          */
void orc_top_level(const Box* roots, u32 G, b2bvh_bvh2_node* nodes) {
  if (G == 1) { nodes[0].m_aabb = roots[0]; nodes[0].m_leftChildIdx = 0; nodes[0].m_rightChildIdx = INVALID; return; }
  std::vector<b2bvh_prim_ref> refs(G);
  Box scene = box_empty();
  for (u32 g = 0; g < G; g++) { refs[g].m_primIdx = g; refs[g].m_aabb = roots[g]; box_grow(scene, roots[g]); }
  std::vector<u32> k(G), v(G), ks(G), vs(G);
  orc_morton_codes(&refs[0].m_aabb, sizeof(b2bvh_prim_ref), &scene, G, k.data(), v.data());
  orc_sort_kv(k.data(), v.data(), G, ks.data(), vs.data());
  orc_lbvh_karras(refs.data(), ks.data(), vs.data(), G, nodes, nullptr);
}

} /* extern "C" */
