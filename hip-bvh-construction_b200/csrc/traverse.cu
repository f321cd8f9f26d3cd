/*
 * traverse.cu — primary-ray generation, closest-hit traversal of the Bvh2 (while-while, speculative while-while, if-if,
 * restart trail) and of the Bvh4, and the top-level tree of a sharded build.
 *
 * Replaces GenerateRays (CommonBlocksKernel.h:432-463), BvhTraversalWhile (TraversalKernel.h:238-335),
 * BvhTraversalSpeculativeWhile (:337-451), BvhTraversalifif (:148-236), BvhTraversalRestartTrail (:49-146) and their host
 * side TwoPassLbvh::traverseBvh (TwoPassLbvh.cpp:199-311).
 * Kept: pinhole camera with a 0.024 m sensor, ray index gIdx*height+gIdy at generation and gIdx*width+gIdy at traversal,
 * boxes tested in object space (inverse transform of the ray), triangles tested in world space, the slab test with
 * maxt = current hit distance, near child first / far child pushed, hit accepted when u,v,w,t > 0 and t < hit.t, the
 * per-ray count of triangle tests of the if-if and restart-trail kernels (rayCounter).
 * Changed: the stack has 64 entries (the reference's 32-entry stack overflows on sponza, depth 35); the tree may be in the
 * LBVH layout or in the separate-leaf layout of PLOC++/H-PLOC (the reference never traverses those); a HitInfo buffer can
 * be returned (parity is checked on hits, not pixels); the Bvh4 of the build can be traced (traverse_wide4_kernel; the
 * reference builds it and only reports its cost).  Arithmetic is evaluated without FMA, in the reference's order.
 * Round-2 experiment, taken out again: carrying the child indices of the node the ray descends into (they sit in the 32-byte record that
 * was just fetched for its box) instead of re-reading them at the visit — same Mray/s for while-while (sponza 981 vs 1003), 12 % less for
 * if-if: the re-read hits L1, there was no second memory round trip to save (gpurun r2j).  512 x 512 rays are one partial wave of a B200
 * (4096 CTAs of 64 threads, all resident): the time is the longest ray's path; at 4096 x 4096 the same kernels reach 4.0-4.3 Gray/s.
 */
#include <math.h>

#include "common.cuh"
#include "morton.cuh"

struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };
__device__ __forceinline__ F3 f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 operator/(F3 a, F3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
__device__ __forceinline__ F3 scale(F3 a, float c) { return f3(c * a.x, c * a.y, c * a.z); }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross3(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ F3 normalize3(F3 a) { const float l = sqrtf(dot3(a, a)); return f3(a.x / l, a.y / l, a.z / l); }
/* quaternion product, rotation and the two transforms (Common.h:483-514) */
__device__ __forceinline__ F4 qmul(F4 a, F4 b) {
  const F3 c = cross3(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z));
  F4 r;
  r.x = (c.x + a.w * b.x) + b.w * a.x;
  r.y = (c.y + a.w * b.y) + b.w * a.y;
  r.z = (c.z + a.w * b.z) + b.w * a.z;
  r.w = a.w * b.w - dot3(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z));
  return r;
}
__device__ __forceinline__ F4 qinv(F4 q) { return F4{-q.x, -q.y, -q.z, q.w}; }
__device__ __forceinline__ F3 qrotate(F4 q, F3 p) { const F4 o = qmul(qmul(q, F4{p.x, p.y, p.z, 0.0f}), qinv(q)); return f3(o.x, o.y, o.z); }
__device__ __forceinline__ F3 to_object(F3 p, F3 s, F4 r, F3 t) { return qrotate(qinv(r), p - t) / s; }
__device__ __forceinline__ F3 to_world(F3 p, F3 s, F4 r, F3 t) { return qrotate(r, s * p) + t; }

__global__ void __launch_bounds__(64) generate_rays_kernel(b2bvh_camera cam, float zDir, u32 width, u32 height, b2bvh_ray* __restrict__ rays) {
  const u32 gx = blockIdx.x * 8 + (threadIdx.x & 7u), gy = blockIdx.y * 8 + (threadIdx.x >> 3);
  if (gx >= width || gy >= height) return;
  const float sx = 0.024f * ((float)width / (float)height), sy = 0.024f;
  const F4 q = F4{cam.m_quat.x, cam.m_quat.y, cam.m_quat.z, cam.m_quat.w};
  const F3 hol = qrotate(q, f3(1.0f, 0.0f, 0.0f)), up = qrotate(q, f3(0.0f, -1.0f, 0.0f)), view = qrotate(q, f3(0.0f, 0.0f, -1.0f));
  const float px = ((float)gx + 0.5f) / (float)width - 0.5f, py = ((float)gy + 0.5f) / (float)height - 0.5f;
  const F3 d = f3(px * sx, py * sy, zDir);
  const F3 dir = normalize3((scale(hol, d.x) + scale(up, d.y)) + scale(view, d.z));
  const F3 eye = f3(cam.m_eye.x, cam.m_eye.y, cam.m_eye.z);
  const F3 dst = normalize3(f3(eye.x + dir.x * cam.m_far, eye.y + dir.y * cam.m_far, eye.z + dir.z * cam.m_far));
  float4* out = reinterpret_cast<float4*>(rays + (size_t)gx * height + gy);
  out[0] = make_float4(eye.x, eye.y, eye.z, dst.x);
  out[1] = make_float4(dst.y, dst.z, 0.0f, B2_FLT_MAX);
}

struct TravArgs {
  const b2bvh_ray* rays;
  const b2bvh_bvh2_node* nodes;
  const b2bvh_prim_ref* leaves; /* null: LBVH layout */
  const b2bvh_triangle* tris;
  b2bvh_transform tr;
  b2bvh_hit* hits;
  uint8_t* rgba;
  u32* counter;                 /* leaf (triangle) tests per ray, or null */
  const b2bvh_bvh4_node* wide;  /* 4-wide traversal only */
  u32 root, nInt, width, height;
  u32* overflow;                /* set to 1 by a ray whose stack (or 64-bit trail) is too short for the tree: the call then fails, a subtree is never dropped silently */
};

__device__ __forceinline__ void slab(const Box& b, F3 o, F3 inv, float maxt, float& tn, float& tf) {
  const F3 dF = (f3(b.hx, b.hy, b.hz) - o) * inv, dN = (f3(b.lx, b.ly, b.lz) - o) * inv;
  const F3 tF = f3(fmaxf(dF.x, dN.x), fmaxf(dF.y, dN.y), fmaxf(dF.z, dN.z)), tN = f3(fminf(dF.x, dN.x), fminf(dF.y, dN.y), fminf(dF.z, dN.z));
  const float minFar = fminf(tF.x, fminf(tF.y, tF.z)), maxNear = fmaxf(tN.x, fmaxf(tN.y, tN.z));
  tf = fminf(maxt, minFar);
  tn = fmaxf(0.0f, maxNear);
}

template <bool SEPARATE>
__device__ __forceinline__ Box child_box(const TravArgs& A, u32 id) {
  if (SEPARATE && id >= A.nInt) {
    const float* f = reinterpret_cast<const float*>(A.leaves + (id - A.nInt)) + 1;
    return Box{__ldg(f), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5)};
  }
  return load_node2_ro(A.nodes + id).box;
}

struct Hit { u32 prim; float t, u, v; };

template <bool SEPARATE>
__device__ __forceinline__ void test_leaf(const TravArgs& A, u32 leaf, F3 ro, F3 rd, F3 ts, F4 tq, F3 tt, Hit& hit) {
  const u32 prim = SEPARATE ? __ldg(&A.leaves[leaf - A.nInt].m_primIdx) : __ldg(&A.nodes[leaf].m_leftChildIdx);
  const float4* p = reinterpret_cast<const float4*>(A.tris + prim);
  const float4 a = __ldg(p), c = __ldg(p + 1);
  const float z3 = __ldg(reinterpret_cast<const float*>(p + 2));
  const F3 v0 = to_world(f3(a.x, a.y, a.z), ts, tq, tt), v1 = to_world(f3(a.w, c.x, c.y), ts, tq, tt), v2 = to_world(f3(c.z, c.w, z3), ts, tq, tt);
  /* intersectTriangle, Common.h:516-531 */
  const F3 p0 = v0 - ro, p1 = v1 - ro, p2 = v2 - ro;
  const F3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
  const F3 nrm = cross3(e1, e0);
  const float den = dot3(nrm, rd) * 2.0f;
  const float u = dot3(cross3(p0 + p2, e0), rd) / den;
  const float v = dot3(cross3(p1 + p0, e1), rd) / den;
  const float w = dot3(cross3(p2 + p1, e2), rd) / den;
  const float t = (dot3(p0, nrm) * 2.0f) / den;
  if (u > 0.0f && v > 0.0f && w > 0.0f && t > 0.0f && t < hit.t) { hit.prim = prim; hit.t = t; hit.u = u; hit.v = v; }
}

__device__ __forceinline__ void write_hit(const TravArgs& A, u32 index, const Hit& hit, u32 tests) {
  if (A.hits) {
    float4* h = reinterpret_cast<float4*>(A.hits + index);
    h[0] = make_float4(__uint_as_float(hit.prim), hit.t, hit.u, hit.v);
    h[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (A.counter) A.counter[index] = tests;
  if (A.rgba && hit.prim != B2_INVALID) {
    uchar4 c;
    c.x = (unsigned char)(hit.u * 255); c.y = (unsigned char)(hit.v * 255); c.z = (unsigned char)((1 - hit.u - hit.v) * 255); c.w = 255;
    reinterpret_cast<uchar4*>(A.rgba)[index] = c;
  }
}

template <bool SEPARATE, bool SPECULATIVE>
__global__ void __launch_bounds__(64) traverse_kernel(TravArgs A) {
  const u32 gx = blockIdx.x * 8 + (threadIdx.x & 7u), gy = blockIdx.y * 8 + (threadIdx.x >> 3);
  const bool inside = gx < A.width && gy < A.height;
  const u32 index = inside ? gx * A.width + gy : 0;
  const float4* rp = reinterpret_cast<const float4*>(A.rays + index);
  const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
  const F3 ro = f3(r0.x, r0.y, r0.z), rd = f3(r0.w, r1.x, r1.y);
  const F3 ts = f3(A.tr.m_scale.x, A.tr.m_scale.y, A.tr.m_scale.z), tt = f3(A.tr.m_translation.x, A.tr.m_translation.y, A.tr.m_translation.z);
  const F4 tq = F4{A.tr.m_quat.x, A.tr.m_quat.y, A.tr.m_quat.z, A.tr.m_quat.w};
  const F3 oo = to_object(ro, ts, tq, tt), od = to_object(rd, ts, tq, f3(0.0f, 0.0f, 0.0f));
  const F3 inv = f3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
  Hit hit = Hit{B2_INVALID, B2_FLT_MAX, 0.0f, 0.0f};
  u32 stack[64];
  u32 top = 0;
  stack[top++] = B2_INVALID;
  u32 node = inside ? A.root : B2_INVALID;
  u32 postponed = B2_INVALID;
  while (node != B2_INVALID) {
    /* ---- inner loop: descend through internal nodes ---- */
    bool searching = true;
    while (node != B2_INVALID && node < A.nInt) {
      const uint2 ch = __ldg(reinterpret_cast<const uint2*>(A.nodes + node));
      const Box lb = child_box<SEPARATE>(A, ch.x), rb = child_box<SEPARATE>(A, ch.y);
      float n0, f0, n1, f1;
      slab(lb, oo, inv, hit.t, n0, f0);
      slab(rb, oo, inv, hit.t, n1, f1);
      const bool hl = n0 <= f0, hr = n1 <= f1;
      if (hl && hr) {
        const bool leftFirst = n0 < n1;
        node = leftFirst ? ch.x : ch.y;
        if (top < 64) stack[top++] = leftFirst ? ch.y : ch.x; else *A.overflow = 1u;
      } else if (hl || hr) {
        node = hl ? ch.x : ch.y;
      } else {
        node = stack[--top];
      }
      if (SPECULATIVE) {
        /* keep the first leaf found aside and go on looking for the next one while other lanes are still descending */
        if (node != B2_INVALID && node >= A.nInt && postponed == B2_INVALID) {
          searching = false;
          postponed = node;
          node = stack[--top];
        }
        if (!__any_sync(__activemask(), searching)) break;
      }
    }
    /* ---- leaf loop ---- */
    if (SPECULATIVE) {
      while (postponed != B2_INVALID) {
        test_leaf<SEPARATE>(A, postponed, ro, rd, ts, tq, tt, hit);
        postponed = B2_INVALID;
        if (node != B2_INVALID && node >= A.nInt) { postponed = node; node = stack[--top]; }
      }
    } else if (node != B2_INVALID) {
      test_leaf<SEPARATE>(A, node, ro, rd, ts, tq, tt, hit);
      node = stack[--top];
    }
  }
  if (!inside) return;
  write_hit(A, index, hit, 0u);
}

/* ---- the reference's two other Bvh2 kernels: BvhTraversalifif (TraversalKernel.h:148-236; one node per step, 64-entry
 * stack — the kernel TwoPassLbvh::traverseBvh launches under IFIF, TwoPassLbvh.cpp:250-269) and BvhTraversalRestartTrail
 * (:49-146 with pop() :32-47; stackless, a 64-bit trail, restarts from the root).  Both keep the per-ray count of triangle
 * tests (rayCounter) that feeds the heat map (Utility.cpp:424-454).  Restart trail: the reference restarts from node 0
 * (:44), which is the root only for the Karras numbering; the root index is used here.  On a tie of the entry distances the
 * stack kernels visit the right child first (:219), the trail kernel the left one (:112-116). ---- */
template <bool SEPARATE, bool RESTART>
__global__ void __launch_bounds__(64) traverse_step_kernel(TravArgs A) {
  const u32 gx = blockIdx.x * 8 + (threadIdx.x & 7u), gy = blockIdx.y * 8 + (threadIdx.x >> 3);
  if (gx >= A.width || gy >= A.height) return;
  const u32 index = gx * A.width + gy;
  const float4* rp = reinterpret_cast<const float4*>(A.rays + index);
  const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
  const F3 ro = f3(r0.x, r0.y, r0.z), rd = f3(r0.w, r1.x, r1.y);
  const F3 ts = f3(A.tr.m_scale.x, A.tr.m_scale.y, A.tr.m_scale.z), tt = f3(A.tr.m_translation.x, A.tr.m_translation.y, A.tr.m_translation.z);
  const F4 tq = F4{A.tr.m_quat.x, A.tr.m_quat.y, A.tr.m_quat.z, A.tr.m_quat.w};
  const F3 oo = to_object(ro, ts, tq, tt), od = to_object(rd, ts, tq, f3(0.0f, 0.0f, 0.0f));
  const F3 inv = f3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
  Hit hit = Hit{B2_INVALID, B2_FLT_MAX, 0.0f, 0.0f};
  u32 tests = 0;
  u32 node = A.root;
  if (!RESTART) {
    u32 stack[64];
    u32 top = 0;
    stack[top++] = B2_INVALID;
    while (node != B2_INVALID) {
      if (node >= A.nInt) {
        test_leaf<SEPARATE>(A, node, ro, rd, ts, tq, tt, hit);
        tests++;
      } else {
        const uint2 ch = __ldg(reinterpret_cast<const uint2*>(A.nodes + node));
        const Box lb = child_box<SEPARATE>(A, ch.x), rb = child_box<SEPARATE>(A, ch.y);
        float n0, f0, n1, f1;
        slab(lb, oo, inv, hit.t, n0, f0);
        slab(rb, oo, inv, hit.t, n1, f1);
        const bool hl = n0 <= f0, hr = n1 <= f1;
        if (hl || hr) {
          if (hl && hr) {
            const bool leftFirst = n0 < n1;
            node = leftFirst ? ch.x : ch.y;
            if (top < 64) stack[top++] = leftFirst ? ch.y : ch.x; else *A.overflow = 1u;
          } else {
            node = hl ? ch.x : ch.y;
          }
          continue;
        }
      }
      node = stack[--top];
    }
  } else {
    const u64 TOP = 0x8000000000000000ull;
    u64 trail = TOP, level = TOP, popLevel = 0;
    bool done = false;
    auto pop = [&]() -> bool {
      trail &= (0ull - level);
      trail += level;
      const u64 temp = trail >> 1;
      level = ((temp - 1) ^ temp) + 1;
      if (!(trail & TOP)) return true;
      popLevel = level;
      node = A.root;
      level = TOP;
      return false;
    };
    while (!done) {
      if (node >= A.nInt) {
        test_leaf<SEPARATE>(A, node, ro, rd, ts, tq, tt, hit);
        tests++;
        done = pop();
      } else {
        const uint2 ch = __ldg(reinterpret_cast<const uint2*>(A.nodes + node));
        const Box lb = child_box<SEPARATE>(A, ch.x), rb = child_box<SEPARATE>(A, ch.y);
        float n0, f0, n1, f1;
        slab(lb, oo, inv, hit.t, n0, f0);
        slab(rb, oo, inv, hit.t, n1, f1);
        const bool hl = n0 <= f0, hr = n1 <= f1;
        if (hl && hr) {
          const bool swap = n0 > n1;
          const u32 nearC = swap ? ch.y : ch.x, farC = swap ? ch.x : ch.y;
          level >>= 1;
          if (level == 0) { *A.overflow = 1u; break; } /* deeper than the 64-bit trail can record */
          node = (trail & level) ? farC : nearC;
        } else if (hl || hr) {
          level >>= 1;
          if (level == 0) { *A.overflow = 1u; break; }
          if (level != popLevel) { trail |= level; node = hr ? ch.y : ch.x; }
          else done = pop();
        } else {
          done = pop();
        }
      }
    }
  }
  write_hit(A, index, hit, tests);
}

/* ---- closest hit through the 4-wide tree (new: the reference builds the Bvh4 but never walks it).  Visiting a wide node:
 * leaf children first, in slot order — box and primitive from the Bvh2 leaf record (a wide node stores no box for them,
 * TwoPassLbvhKernel.h:320-325), triangle tested when the slab test passes; then the internal children are slab-tested
 * with the updated hit distance and visited nearest first (entry distance, slot order among equals), the others pushed
 * far to near on a 128-entry stack.  Mirrors orc_traverse_wide4 of the oracle. ---- */
template <bool SEPARATE>
__global__ void __launch_bounds__(64) traverse_wide4_kernel(TravArgs A) {
  const u32 gx = blockIdx.x * 8 + (threadIdx.x & 7u), gy = blockIdx.y * 8 + (threadIdx.x >> 3);
  if (gx >= A.width || gy >= A.height) return;
  const u32 index = gx * A.width + gy;
  const float4* rp = reinterpret_cast<const float4*>(A.rays + index);
  const float4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
  const F3 ro = f3(r0.x, r0.y, r0.z), rd = f3(r0.w, r1.x, r1.y);
  const F3 ts = f3(A.tr.m_scale.x, A.tr.m_scale.y, A.tr.m_scale.z), tt = f3(A.tr.m_translation.x, A.tr.m_translation.y, A.tr.m_translation.z);
  const F4 tq = F4{A.tr.m_quat.x, A.tr.m_quat.y, A.tr.m_quat.z, A.tr.m_quat.w};
  const F3 oo = to_object(ro, ts, tq, tt), od = to_object(rd, ts, tq, f3(0.0f, 0.0f, 0.0f));
  const F3 inv = f3(1.0f / od.x, 1.0f / od.y, 1.0f / od.z);
  Hit hit = Hit{B2_INVALID, B2_FLT_MAX, 0.0f, 0.0f};
  u32 tests = 0;
  u32 stack[128];
  u32 top = 0;
  stack[top++] = B2_INVALID;
  u32 node = 0;
  while (node != B2_INVALID) {
    const float2* wb = reinterpret_cast<const float2*>(A.wide + node); /* 4 boxes of 24 B, then m_child[4] at byte 96 */
    const uint4 chv = __ldg(reinterpret_cast<const uint4*>(A.wide + node) + 6);
    const u32 ch[4] = {chv.x, chv.y, chv.z, chv.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const u32 c = ch[k];
      if (c == B2_INVALID || c < A.nInt) continue;
      float n0, f0;
      slab(child_box<SEPARATE>(A, c), oo, inv, hit.t, n0, f0);
      if (!(n0 <= f0)) continue;
      test_leaf<SEPARATE>(A, c, ro, rd, ts, tq, tt, hit);
      tests++;
    }
    float tn[4];
    bool hitK[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      hitK[k] = false; tn[k] = 0.0f;
      const u32 c = ch[k];
      if (c == B2_INVALID || c >= A.nInt) continue;
      const float2 q0 = __ldg(wb + 3 * k), q1 = __ldg(wb + 3 * k + 1), q2 = __ldg(wb + 3 * k + 2);
      float f0;
      slab(Box{q0.x, q0.y, q1.x, q1.y, q2.x, q2.y}, oo, inv, hit.t, tn[k], f0);
      hitK[k] = tn[k] <= f0;
    }
    /* rank among the children hit: by entry distance, slot order among equals */
    u32 rank[4], m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      rank[k] = 0;
      if (hitK[k]) m++;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j != k && hitK[j] && (tn[j] < tn[k] || (tn[j] == tn[k] && j < k))) rank[k]++;
    }
    if (m == 0) { node = stack[--top]; continue; }
#pragma unroll
    for (int r = 3; r >= 1; r--)
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (hitK[k] && rank[k] == (u32)r) { if (top < 128) stack[top++] = ch[k]; else *A.overflow = 1u; }
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (hitK[k] && rank[k] == 0) node = ch[k];
  }
  write_hit(A, index, hit, tests);
}

/* ---- top-level tree over the G sub-tree root boxes of a sharded build (G <= 256): Morton codes of the box centres in the
 * frame of their union, stable sort, Karras hierarchy + refit — the same three steps as the per-shard build, done by one
 * thread because G is the GPU count.  Leaf g of the result holds the rank in m_leftChildIdx. ---- */
__global__ void top_level_kernel(const b2bvh_aabb* __restrict__ roots, u32 G, b2bvh_bvh2_node* nodes) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (G == 1) { store_node2(nodes, 0u, B2_INVALID, load_aabb(roots)); return; }
  __shared__ u32 key[256], val[256];
  __shared__ unsigned char done[512];
  Box scene = box_empty();
  for (u32 g = 0; g < G; g++) scene = box_union(scene, load_aabb(roots + g));
  const float ex = scene.hx - scene.lx, ey = scene.hy - scene.ly, ez = scene.hz - scene.lz;
  MortonCfg cfg;
  morton_make_cfg(ex, ey, ez, cfg);
  for (u32 g = 0; g < G; g++) {
    const Box b = load_aabb(roots + g);
    float p[3];
    p[0] = (0.5f * (b.hx + b.lx) - scene.lx) / ex;
    p[1] = (0.5f * (b.hy + b.ly) - scene.ly) / ey;
    p[2] = (0.5f * (b.hz + b.lz) - scene.lz) / ez;
    const u32 k = morton_code(p, cfg);
    u32 pos = g; /* stable insertion sort */
    while (pos > 0 && key[pos - 1] > k) { key[pos] = key[pos - 1]; val[pos] = val[pos - 1]; pos--; }
    key[pos] = k; val[pos] = g;
  }
  const u32 nInt = G - 1;
  for (u32 g = 0; g < G; g++) { store_node2(nodes + nInt + g, val[g], B2_INVALID, load_aabb(roots + val[g])); done[nInt + g] = 1; }
  auto delta = [&](u32 i, int j) -> int {
    if (j < 0 || j >= (int)G) return -1;
    const u64 a = ((u64)key[i] << 32) | i, b = ((u64)key[j] << 32) | (u32)j;
    return __clzll((long long)(a ^ b));
  };
  u32 lc[256], rc[256];
  for (u32 i = 0; i < nInt; i++) {
    u32 first, last;
    if (i == 0) { first = 0; last = G - 1; }
    else {
      const int dR = delta(i, (int)i + 1), dL = delta(i, (int)i - 1);
      const int dir = dR > dL ? 1 : -1, dMin = min(dL, dR);
      int len = 0;
      while (delta(i, (int)i + (len + 1) * dir) > dMin) len++;
      const u32 other = (u32)((int)i + len * dir);
      first = dir > 0 ? i : other; last = dir > 0 ? other : i;
    }
    const int dNode = delta(first, (int)last);
    u32 split = first;
    while (split + 1 < last && delta(first, (int)split + 1) > dNode) split++;
    lc[i] = (split == first) ? split + nInt : split;
    rc[i] = (split + 1 == last) ? split + 1 + nInt : split + 1;
    done[i] = 0;
  }
  for (u32 round = 0; round < nInt; round++)
    for (u32 i = 0; i < nInt; i++)
      if (!done[i] && done[lc[i]] && done[rc[i]]) {
        const b2bvh_bvh2_node* l = nodes + lc[i];
        const b2bvh_bvh2_node* r = nodes + rc[i];
        const Box lb = Box{l->m_aabb.m_min.x, l->m_aabb.m_min.y, l->m_aabb.m_min.z, l->m_aabb.m_max.x, l->m_aabb.m_max.y, l->m_aabb.m_max.z};
        const Box rb = Box{r->m_aabb.m_min.x, r->m_aabb.m_min.y, r->m_aabb.m_min.z, r->m_aabb.m_max.x, r->m_aabb.m_max.y, r->m_aabb.m_max.z};
        store_node2(nodes + i, lc[i], rc[i], box_union(lb, rb));
        done[i] = 1;
      }
}

extern "C" {

int b2bvh_generate_rays(b2bvh_ctx* ctx, const b2bvh_camera* cam, uint32_t width, uint32_t height, b2bvh_ray* d_rays, float* ms) {
  if (!ctx || !cam || !d_rays || width == 0 || height == 0) return b2_fail(B2BVH_ERR_INVALID, "generate_rays: bad argument");
  const float zDir = 0.024f / (2.f * tanf(cam->m_fov / 2.f)); /* uniform: evaluated once on the host */
  B2_CUDA(cudaEventRecord(ctx->ev[10], ctx->stream));
  B2_KERNEL(ctx, "generate_rays");
  generate_rays_kernel<<<dim3((width + 7) / 8, (height + 7) / 8), 64, 0, ctx->stream>>>(*cam, zDir, width, height, d_rays);
  B2_LAUNCH_CHECK(ctx);
  B2_CUDA(cudaEventRecord(ctx->ev[11], ctx->stream));
  B2_CUDA(cudaEventSynchronize(ctx->ev[11]));
  if (ms) B2_CUDA(cudaEventElapsedTime(ms, ctx->ev[10], ctx->ev[11]));
  return 0;
}

int b2bvh_traverse_ex(b2bvh_ctx* ctx, const b2bvh_tree* tree, const b2bvh_ray* d_rays, uint32_t n_rays, const b2bvh_transform* xform, int kernel,
                      b2bvh_hit* d_hits, uint8_t* d_rgba, uint32_t* d_rayCounter, float* ms) {
  if (!ctx || !tree || !d_rays || !xform || n_rays == 0) return b2_fail(B2BVH_ERR_INVALID, "traverse: bad argument");
  if (kernel < B2BVH_TRAVERSE_WHILE || kernel > B2BVH_TRAVERSE_WIDE4) return b2_fail(B2BVH_ERR_INVALID, "traverse: unknown kernel %d", kernel);
  if (kernel == B2BVH_TRAVERSE_WIDE4 && (tree->n_wide == 0 || !tree->d_wideBvhNodes))
    return b2_fail(B2BVH_ERR_INVALID, "traverse: the tree has no 4-wide nodes (build with collapse = 1)");
  if (d_rayCounter && (kernel == B2BVH_TRAVERSE_WHILE || kernel == B2BVH_TRAVERSE_SPECULATIVE_WHILE))
    return b2_fail(B2BVH_ERR_INVALID, "traverse: the while-while kernels keep no ray counter (as in the reference); use IFIF, RESTART_TRAIL or WIDE4");
  /* the reference traces square images (width == height == 512, TwoPassLbvh.cpp:221-222); n_rays must be a square */
  const u32 side = (u32)(sqrt((double)n_rays) + 0.5);
  if ((uint64_t)side * side != n_rays) return b2_fail(B2BVH_ERR_INVALID, "traverse: n_rays=%u is not a square image", n_rays);
  TravArgs A;
  A.rays = d_rays; A.nodes = tree->d_bvhNodes; A.leaves = tree->leaves_separate ? tree->d_leafNodes : nullptr; A.tris = tree->d_triangleBuff;
  A.tr = *xform; A.hits = d_hits; A.rgba = d_rgba; A.counter = d_rayCounter; A.wide = tree->d_wideBvhNodes;
  A.root = tree->root; A.nInt = tree->n_internal; A.width = side; A.height = side;
  A.overflow = reinterpret_cast<u32*>(reinterpret_cast<unsigned char*>(ctx->bufs[SLOT_CTL].p) + 160);
  B2_CUDA(cudaMemsetAsync(A.overflow, 0, 4, ctx->stream));
  if (d_rgba) B2_CUDA(cudaMemsetAsync(d_rgba, 0, (size_t)n_rays * 4, ctx->stream));
  const dim3 grid((side + 7) / 8, (side + 7) / 8);
  static const char* const names[] = {"traverse_while", "traverse_speculative_while", "traverse_ifif", "traverse_restart_trail", "traverse_wide4"};
  B2_CUDA(cudaEventRecord(ctx->ev[10], ctx->stream));
  B2_KERNEL(ctx, names[kernel]);
  const bool sep = tree->leaves_separate != 0;
  switch (kernel) {
    case B2BVH_TRAVERSE_WHILE:
      if (sep) traverse_kernel<true, false><<<grid, 64, 0, ctx->stream>>>(A); else traverse_kernel<false, false><<<grid, 64, 0, ctx->stream>>>(A);
      break;
    case B2BVH_TRAVERSE_SPECULATIVE_WHILE:
      if (sep) traverse_kernel<true, true><<<grid, 64, 0, ctx->stream>>>(A); else traverse_kernel<false, true><<<grid, 64, 0, ctx->stream>>>(A);
      break;
    case B2BVH_TRAVERSE_IFIF:
      if (sep) traverse_step_kernel<true, false><<<grid, 64, 0, ctx->stream>>>(A); else traverse_step_kernel<false, false><<<grid, 64, 0, ctx->stream>>>(A);
      break;
    case B2BVH_TRAVERSE_RESTART_TRAIL:
      if (sep) traverse_step_kernel<true, true><<<grid, 64, 0, ctx->stream>>>(A); else traverse_step_kernel<false, true><<<grid, 64, 0, ctx->stream>>>(A);
      break;
    default:
      if (sep) traverse_wide4_kernel<true><<<grid, 64, 0, ctx->stream>>>(A); else traverse_wide4_kernel<false><<<grid, 64, 0, ctx->stream>>>(A);
      break;
  }
  B2_LAUNCH_CHECK(ctx);
  B2_CUDA(cudaEventRecord(ctx->ev[11], ctx->stream));
  B2_TRY(b2_fetch_words(ctx, A.overflow, 1, B2_MB_TRAVERSE));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ms) B2_CUDA(cudaEventElapsedTime(ms, ctx->ev[10], ctx->ev[11]));
  if (b2_mailbox(ctx, B2_MB_TRAVERSE)[0] != 0u)
    return b2_fail(B2BVH_ERR_INTERNAL, "traverse: the tree is deeper than the %s of kernel %d holds; hits are incomplete",
                   kernel == B2BVH_TRAVERSE_RESTART_TRAIL ? "64-bit trail" : (kernel == B2BVH_TRAVERSE_WIDE4 ? "128-entry stack" : "64-entry stack"), kernel);
  return 0;
}

int b2bvh_traverse(b2bvh_ctx* ctx, const b2bvh_tree* tree, const b2bvh_ray* d_rays, uint32_t n_rays, const b2bvh_transform* xform, int kernel,
                   b2bvh_hit* d_hits, uint8_t* d_rgba, float* ms) {
  return b2bvh_traverse_ex(ctx, tree, d_rays, n_rays, xform, kernel, d_hits, d_rgba, nullptr, ms);
}

/* Utility::generateTraversalHeatMap (Utility.cpp:424-454) without the PNG write; host function on host buffers. */
int b2bvh_heat_map(const uint32_t* rayCounter, uint32_t count, uint8_t* rgba) {
  if (!rayCounter || !rgba || count == 0) return b2_fail(B2BVH_ERR_INVALID, "heat_map: bad argument");
  uint32_t mx = 0;
  for (uint32_t i = 0; i < count; i++) if (rayCounter[i] > mx) mx = rayCounter[i];
  for (uint32_t i = 0; i < count; i++) {
    rgba[i * 4 + 0] = (uint8_t)((rayCounter[i] / (float)mx) * 150);
    rgba[i * 4 + 1] = (uint8_t)((rayCounter[i] / (float)mx) * 255);
    rgba[i * 4 + 2] = 255;
    rgba[i * 4 + 3] = 255;
  }
  return 0;
}

int b2bvh_top_level(b2bvh_ctx* ctx, const b2bvh_aabb* d_rootBoxes, uint32_t g, b2bvh_bvh2_node* d_topNodes) {
  if (!ctx || !d_rootBoxes || !d_topNodes || g == 0 || g > 256) return b2_fail(B2BVH_ERR_INVALID, "top_level: bad argument (1 <= g <= 256)");
  B2_KERNEL(ctx, "top_level");
  top_level_kernel<<<1, 32, 0, ctx->stream>>>(d_rootBoxes, g, d_topNodes);
  B2_LAUNCH_CHECK(ctx);
  return 0;
}
}
