// aob_bvh.cuh — 8-wide compressed BVH: node layout, LBVH construction bodies, SAH-guided
// collapse, and any-hit traversal.  Replaces the closed OptiX Prime builder + traversal the
// reference called through rtpModelUpdate / rtpQueryExecute (bake_ao_optix_prime.cpp;
// SURVEY §8 a10/a11).
//
// Layout (HBM):
//   nodes : 80-byte Node8 records = 5 x 16-byte loads (Ylitie, Karras, Laine 2017 style
//           compressed wide BVH): parent-relative 8-bit child boxes, a shared power-of-two
//           scale per axis, an internal-child mask, and a meta byte per slot.
//   tris  : 48-byte records = 3 x float4 (v0, v1, v2; w lanes carry the source primitive id),
//           in leaf order, so a leaf is (prim_base + offset, count <= 3).
//   insts : 80-byte records = 3 x float4 rows of the world->object 3x4 + uint4{blas_root, id, -, -}
//           + float4 world-space bounding sphere (centre, radius^2) for the pre-test.
//
// Kernel bodies are written as `*_body(tid, ...)` functions that also compile under plain
// g++ (AOB_HOST_EMU) — tests/emu runs the identical code serially on the CPU.
#pragma once
#include "aob_math.cuh"
#include <string.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

namespace aob {

// ---- host/device shims ---------------------------------------------------------------
struct alignas(16) U4 { uint32_t x, y, z, w; };
struct alignas(16) F4 { float x, y, z, w; };
struct alignas(8) U2 { uint32_t x, y; };

#if defined(__CUDA_ARCH__)
AOB_D U4 ld_u4(const U4* p) { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); U4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
AOB_D F4 ld_f4(const F4* p) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
AOB_D F4 ld_f4_cg(const F4* p) { float4 v = __ldcg(reinterpret_cast<const float4*>(p)); F4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w; return r; }
AOB_D float as_float(uint32_t u) { return __uint_as_float(u); }
AOB_D uint32_t as_uint(float f) { return __float_as_uint(f); }
AOB_D uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
AOB_D int clz32(uint32_t v) { return __clz((int)v); }
AOB_D int clz64(uint64_t v) { return __clzll((long long)v); }
AOB_D int popc32(uint32_t v) { return __popc(v); }
AOB_D int ffs32(uint32_t v) { return __ffs((int)v); }
AOB_D uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
AOB_D void thread_fence() { __threadfence(); }
// PRMT with a register selector (no 0x7777 masking as in __byte_perm: selectors here never set bit 3)
AOB_D uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r; }
// packed fp16 pairs for the node test: F2FP.SATFINITE / HFMA2 / HMNMX2 / HADD2 with lane swizzles
AOB_D uint32_t h2_pack_sat(float hi, float lo) { uint32_t r; asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
AOB_D uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
AOB_D uint32_t h2_min(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
AOB_D uint32_t h2_lane_sum(uint32_t a) {  // lane0 + lane1, in both lanes (HADD2 R, R.H0_H0, R.H1_H1)
  const __half2 h = *reinterpret_cast<const __half2*>(&a);
  const __half2 r = __hadd2(__low2half2(h), __high2half2(h));
  return *reinterpret_cast<const uint32_t*>(&r);
}
#else
inline U4 ld_u4(const U4* p) { return *p; }
inline F4 ld_f4(const F4* p) { return *p; }
inline F4 ld_f4_cg(const F4* p) { return *p; }
inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) {
  uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
  return r;
}
inline int clz32(uint32_t v) { return v ? __builtin_clz(v) : 32; }
inline int clz64(uint64_t v) { return v ? __builtin_clzll(v) : 64; }
inline int popc32(uint32_t v) { return __builtin_popcount(v); }
inline int ffs32(uint32_t v) { return __builtin_ffs((int)v); }
inline uint32_t atomic_add_u32(uint32_t* p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
inline void thread_fence() {}
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return byte_perm(a, b, s); }
// fp16 emulation for tests/emu: exact decode, round-to-nearest-even encode from a double (the
// products of two halves are exact in a double; the one rounding of an fp16 FMA is reproduced up to
// double rounding in cases that need more than 53 bits, which the padding of the node test covers).
inline double h_to_d(uint32_t h) {
  const uint32_t sgn = (h >> 15) & 1u, e = (h >> 10) & 31u, m = h & 1023u;
  double v;
  if (e == 0) v = ldexp((double)m, -24);
  else if (e == 31) v = m ? NAN : INFINITY;
  else v = ldexp((double)(m | 1024u), (int)e - 25);
  return sgn ? -v : v;
}
inline uint32_t d_to_h(double x, bool satfinite) {
  if (x != x) return 0x7fffu;
  const uint32_t sgn = signbit(x) ? 0x8000u : 0u;
  double a = fabs(x);
  if (a >= 65520.0) return sgn | (satfinite ? 0x7bffu : 0x7c00u);   // 65520 = halfway to 2^16: rounds to inf
  if (a < ldexp(1.0, -14)) {   // subnormal half: multiples of 2^-24
    const double r = nearbyint(ldexp(a, 24));   // ties to even (default rounding mode)
    return sgn | (uint32_t)r;                   // r == 1024 encodes the smallest normal: still correct
  }
  int e;
  const double f = frexp(a, &e);                // a = f * 2^e, f in [0.5, 1)
  double m = nearbyint(ldexp(f, 11));           // 11 significant bits: [1024, 2048]
  if (m == 2048.0) { m = 1024.0; e += 1; }
  const int be = e - 1 + 15;                    // biased exponent of 1.xxx * 2^(e-1)
  if (be >= 31) return sgn | (satfinite ? 0x7bffu : 0x7c00u);
  return sgn | ((uint32_t)be << 10) | ((uint32_t)m - 1024u);
}
inline uint32_t h2_pack_sat(float hi, float lo) { return (d_to_h(hi, true) << 16) | d_to_h(lo, true); }
inline uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (int l = 0; l < 2; l++) {
    const int sh = 16 * l;
    r |= d_to_h(h_to_d((a >> sh) & 0xffffu) * h_to_d((b >> sh) & 0xffffu) + h_to_d((c >> sh) & 0xffffu), false) << sh;
  }
  return r;
}
inline uint32_t h2_min(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int l = 0; l < 2; l++) {
    const int sh = 16 * l;
    const uint32_t x = (a >> sh) & 0xffffu, y = (b >> sh) & 0xffffu;
    const double dx = h_to_d(x), dy = h_to_d(y);
    r |= ((dx != dx) ? y : (dy != dy) ? x : (dy < dx ? y : x)) << sh;
  }
  return r;
}
inline uint32_t h2_lane_sum(uint32_t a) {
  const uint32_t h = d_to_h(h_to_d(a & 0xffffu) + h_to_d(a >> 16), false);
  return (h << 16) | h;
}
#endif

// ---- node layout ---------------------------------------------------------------------
struct alignas(16) Node8 {
  float px, py, pz;            // quantisation origin = node box min
  uint8_t ex, ey, ez;          // biased fp32 exponents of the per-axis grid step
  uint8_t imask;               // bit s: slot s holds an internal child
  uint32_t child_base;         // index of the first internal child (children contiguous, slot order)
  uint32_t prim_base;          // index of the first leaf primitive referenced by this node
  uint8_t meta[8];             // 0 empty | internal: 0x20|(24+slot) | leaf: unary(count)<<5 | offset
  // child boxes on the parent's 8-bit grid, [axis][slot][0 = lo, 1 = hi]: the lo and hi bytes of a
  // slot sit side by side so that ONE byte permute yields the (near, far) pair of a child as two
  // fp16 lanes (intersect_node8_h2); words 2j, 2j+1 of an axis hold slots 2j and 2j+1.
  uint8_t q[3][8][2];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

constexpr uint32_t kLeafBit = 0x80000000u;
constexpr uint32_t kSentinel = 0xFFFFFFFFu;
constexpr int kStackSize = 48;
constexpr int kInstF4 = 5;  // float4s per instance record
constexpr int kExpSpread = 14;  // max log2 ratio between the grid steps of one node (see collapse_body)

struct BvhView {
  const U4* nodes;   // 5 x U4 per node
  const F4* tris;    // 3 x F4 per triangle
  const F4* insts;   // kInstF4 x F4 per instance (two-level only)
  uint32_t root;     // node index where traversal starts (TLAS root when two_level)
  uint32_t two_level;
};

// =======================================================================================
// Construction bodies
// =======================================================================================
AOB_HD uint64_t expand21(uint32_t v) {  // spread the low 21 bits to every third bit
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}
// 63-bit Morton code of the box centroid inside the centroid bounds [cmin, cmin + 1/cinv].
AOB_HD uint64_t morton63(F4 lo, F4 hi, V3 cmin, V3 cinv) {
  float cx = 0.5f * (lo.x + hi.x), cy = 0.5f * (lo.y + hi.y), cz = 0.5f * (lo.z + hi.z);
  float fx = fminf(fmaxf((cx - cmin.x) * cinv.x, 0.0f), 1.0f);
  float fy = fminf(fmaxf((cy - cmin.y) * cinv.y, 0.0f), 1.0f);
  float fz = fminf(fmaxf((cz - cmin.z) * cinv.z, 0.0f), 1.0f);
  uint32_t ix = (uint32_t)fminf(fx * 2097152.0f, 2097151.0f);
  uint32_t iy = (uint32_t)fminf(fy * 2097152.0f, 2097151.0f);
  uint32_t iz = (uint32_t)fminf(fz * 2097152.0f, 2097151.0f);
  return (expand21(ix) << 2) | (expand21(iy) << 1) | expand21(iz);
}

// Binary LBVH arrays (Karras 2012).  n leaves, n-1 internal nodes; node refs carry kLeafBit.
struct Lbvh {
  const uint64_t* keys;   // sorted Morton keys [n]
  const uint32_t* prim;   // sorted primitive ids [n]
  const F4* plo;          // primitive boxes (indexed by primitive id)
  const F4* phi;
  uint32_t* left;         // [n-1]
  uint32_t* right;        // [n-1]
  uint32_t* first;        // [n-1] first leaf of the node's range
  uint32_t* last;         // [n-1]
  uint32_t* parent_int;   // [n-1]
  uint32_t* parent_leaf;  // [n]
  F4* ilo;                // [n-1] internal boxes
  F4* ihi;
  uint32_t* flags;        // [n-1] zeroed arrival counters
  uint32_t* count;        // [n-1] primitives under each internal node
  uint32_t n;
};

AOB_HD int lbvh_delta(const Lbvh& L, int i, int j) {
  if (j < 0 || j >= (int)L.n) return -1;
  uint64_t a = L.keys[i], b = L.keys[j];
  if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
  return clz64(a ^ b);
}
// one thread per internal node i in [0, n-1)
AOB_HD void lbvh_hierarchy_body(uint32_t tid, const Lbvh& L) {
  if (tid + 1 >= L.n) return;
  const int i = (int)tid;
  const int d = (lbvh_delta(L, i, i + 1) - lbvh_delta(L, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(L, i, i - d);
  int lmax = 2;
  while (lbvh_delta(L, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (lbvh_delta(L, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvh_delta(L, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (lbvh_delta(L, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + (d < 0 ? d : 0);
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  uint32_t lref, rref;
  if (lo == gamma) { lref = (uint32_t)gamma | kLeafBit; L.parent_leaf[gamma] = (uint32_t)i; }
  else { lref = (uint32_t)gamma; L.parent_int[gamma] = (uint32_t)i; }
  if (hi == gamma + 1) { rref = (uint32_t)(gamma + 1) | kLeafBit; L.parent_leaf[gamma + 1] = (uint32_t)i; }
  else { rref = (uint32_t)(gamma + 1); L.parent_int[gamma + 1] = (uint32_t)i; }
  L.left[i] = lref;
  L.right[i] = rref;
  L.first[i] = (uint32_t)lo;
  L.last[i] = (uint32_t)hi;
  L.count[i] = (uint32_t)(hi - lo + 1);
  if (i == 0) L.parent_int[0] = kSentinel;
}

AOB_HD void lbvh_ref_box(const Lbvh& L, uint32_t ref, F4* lo, F4* hi) {
  if (ref & kLeafBit) {
    uint32_t p = L.prim[ref & ~kLeafBit];
    *lo = L.plo[p]; *hi = L.phi[p];
  } else {
    *lo = L.ilo[ref]; *hi = L.ihi[ref];
  }
}
// one thread per leaf: walk up, the second arrival at a node computes its box.
AOB_HD void lbvh_refit_body(uint32_t tid, const Lbvh& L) {
  if (tid >= L.n || L.n < 2) return;
  uint32_t cur = L.parent_leaf[tid];
  while (cur != kSentinel) {
    thread_fence();
    if (atomic_add_u32(&L.flags[cur], 1u) == 0u) return;
    thread_fence();
    // internal child boxes were written by other threads: read them through L2 (ld.cg)
    F4 alo, ahi, blo, bhi;
    const uint32_t lr = L.left[cur], rr = L.right[cur];
    if (lr & kLeafBit) { const uint32_t p = L.prim[lr & ~kLeafBit]; alo = L.plo[p]; ahi = L.phi[p]; }
    else { alo = ld_f4_cg(&L.ilo[lr]); ahi = ld_f4_cg(&L.ihi[lr]); }
    if (rr & kLeafBit) { const uint32_t p = L.prim[rr & ~kLeafBit]; blo = L.plo[p]; bhi = L.phi[p]; }
    else { blo = ld_f4_cg(&L.ilo[rr]); bhi = ld_f4_cg(&L.ihi[rr]); }
    F4 lo, hi;
    lo.x = fminf(alo.x, blo.x); lo.y = fminf(alo.y, blo.y); lo.z = fminf(alo.z, blo.z); lo.w = 0.f;
    hi.x = fmaxf(ahi.x, bhi.x); hi.y = fmaxf(ahi.y, bhi.y); hi.z = fmaxf(ahi.z, bhi.z); hi.w = 0.f;
    L.ilo[cur] = lo;
    L.ihi[cur] = hi;
    cur = L.parent_int[cur];
  }
}

AOB_HD float box_half_area(F4 lo, F4 hi) {
  float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
  return dx * dy + dy * dz + dz * dx;
}
AOB_HD uint32_t lbvh_ref_count(const Lbvh& L, uint32_t ref) {
  return (ref & kLeafBit) ? 1u : L.count[ref];
}

// smallest biased exponent e with 255 * 2^(e-127) >= extent (0 when extent == 0)
AOB_HD uint32_t quant_exponent(float extent) {
  if (!(extent > 0.0f)) return 0u;
  float s = (extent / 255.0f) * 1.000001f;
  uint32_t b = as_uint(s);
  uint32_t e = (b >> 23) & 0xffu;
  if (b & 0x7fffffu) e += 1u;
  if (e < 1u) e = 1u;
  if (e > 254u) e = 254u;
  return e;
}
AOB_HD void quantize_axis(float p, uint32_t e, float clo, float chi, uint8_t* qlo, uint8_t* qhi) {
  const float scale = as_float(e << 23);
  if (!(scale > 0.0f)) { *qlo = 0; *qhi = 0; return; }
  int lo = (int)floorf((clo - p) / scale);
  int hi = (int)ceilf((chi - p) / scale);
  lo = lo < 0 ? 0 : (lo > 255 ? 255 : lo);
  hi = hi < 0 ? 0 : (hi > 255 ? 255 : hi);
  while (lo > 0 && p + (float)lo * scale > clo) lo--;
  while (hi < 255 && p + (float)hi * scale < chi) hi++;
  *qlo = (uint8_t)lo;
  *qhi = (uint8_t)hi;
}

struct CollapseArgs {
  Lbvh L;
  Node8* nodes;          // absolute array
  uint32_t* wide2bin;    // wide node index (relative to node_offset) -> binary node ref
  uint32_t* leaf_prims;  // leaf order (relative to prim_offset) -> primitive id
  uint32_t* node_count;  // relative counters (start at 1 / 0)
  uint32_t* prim_count;
  uint32_t node_offset;  // added to child_base
  uint32_t prim_offset;  // added to prim_base
  uint32_t max_leaf;     // primitives per leaf slot: 3 for triangles, 1 for instances
};

// one thread per wide node w (relative index) of the current level.
AOB_HD void collapse_body(uint32_t w, const CollapseArgs& A) {
  const Lbvh& L = A.L;
  const uint32_t bref = A.wide2bin[w];
  uint32_t slot[8];
  float area[8];
  int ns = 0;
  F4 nlo, nhi;
  if ((bref & kLeafBit) || lbvh_ref_count(L, bref) <= A.max_leaf) {
    // degenerate root: the whole tree fits one leaf slot
    lbvh_ref_box(L, bref, &nlo, &nhi);
    slot[0] = bref; area[0] = -1.0f; ns = 1;
  } else {
    nlo = L.ilo[bref]; nhi = L.ihi[bref];
    slot[0] = L.left[bref]; slot[1] = L.right[bref]; ns = 2;
    for (int k = 0; k < 2; k++) {
      F4 lo, hi;
      lbvh_ref_box(L, slot[k], &lo, &hi);
      area[k] = (lbvh_ref_count(L, slot[k]) <= A.max_leaf) ? -1.0f : box_half_area(lo, hi);
    }
    // greedy surface-area-ordered expansion (the SAH heuristic of the wide collapse): open the
    // largest openable child until 8 slots are used.
    while (ns < 8) {
      int best = -1;
      float ba = -1.0f;
      for (int k = 0; k < ns; k++)
        if (area[k] > ba) { ba = area[k]; best = k; }
      if (best < 0) break;
      const uint32_t b = slot[best];
      const uint32_t c0 = L.left[b], c1 = L.right[b];
      slot[best] = c0;
      slot[ns] = c1;
      const int idx[2] = {best, ns};
      ns++;
      for (int k = 0; k < 2; k++) {
        F4 lo, hi;
        lbvh_ref_box(L, slot[idx[k]], &lo, &hi);
        area[idx[k]] = (lbvh_ref_count(L, slot[idx[k]]) <= A.max_leaf) ? -1.0f : box_half_area(lo, hi);
      }
    }
  }
  // classify and allocate
  uint32_t n_int = 0, n_prims = 0;
  for (int k = 0; k < ns; k++) {
    if (area[k] >= 0.0f) n_int++;
    else n_prims += lbvh_ref_count(L, slot[k]);
  }
  uint32_t cbase = n_int ? atomic_add_u32(A.node_count, n_int) : 0u;
  uint32_t pbase = n_prims ? atomic_add_u32(A.prim_count, n_prims) : 0u;
  Node8 nd;
  nd.px = nlo.x; nd.py = nlo.y; nd.pz = nlo.z;
  uint32_t ex = quant_exponent(nhi.x - nlo.x), ey = quant_exponent(nhi.y - nlo.y), ez = quant_exponent(nhi.z - nlo.z);
  {
    // Keep the three grid steps within 2^kExpSpread of the widest one (and the widest away from
    // zero).  The fp16 node test works in a frame scaled by the widest axis and needs every
    // per-axis step to be a normal fp16 number there; a step 2^-14 of the widest one is far below
    // anything that changes which rays meet the box.
    uint32_t em = ex > ey ? ex : ey;
    em = em > ez ? em : ez;
    if (em < 24u) em = 24u;
    const uint32_t fl = em - (uint32_t)kExpSpread;
    ex = ex > fl ? ex : fl; ey = ey > fl ? ey : fl; ez = ez > fl ? ez : fl;
  }
  nd.ex = (uint8_t)ex; nd.ey = (uint8_t)ey; nd.ez = (uint8_t)ez;
  nd.child_base = A.node_offset + cbase;
  nd.prim_base = A.prim_offset + pbase;
  uint32_t imask = 0, ci = 0, po = 0;
  for (int k = 0; k < 8; k++) {
    if (k >= ns) {
      nd.meta[k] = 0;
      for (int a = 0; a < 3; a++) { nd.q[a][k][0] = 255; nd.q[a][k][1] = 0; }
      continue;
    }
    F4 lo, hi;
    lbvh_ref_box(L, slot[k], &lo, &hi);
    quantize_axis(nlo.x, ex, lo.x, hi.x, &nd.q[0][k][0], &nd.q[0][k][1]);
    quantize_axis(nlo.y, ey, lo.y, hi.y, &nd.q[1][k][0], &nd.q[1][k][1]);
    quantize_axis(nlo.z, ez, lo.z, hi.z, &nd.q[2][k][0], &nd.q[2][k][1]);
    if (area[k] >= 0.0f) {
      imask |= 1u << k;
      nd.meta[k] = (uint8_t)(0x20u | (24u + (uint32_t)k));
      A.wide2bin[cbase + ci] = slot[k];
      ci++;
    } else {
      // enumerate the (<= max_leaf) primitives of the subtree, depth first
      uint32_t cnt = 0, st[8];
      int sp = 0;
      st[sp++] = slot[k];
      while (sp) {
        const uint32_t rr = st[--sp];
        if (rr & kLeafBit) A.leaf_prims[pbase + po + cnt++] = L.prim[rr & ~kLeafBit];
        else { st[sp++] = L.right[rr]; st[sp++] = L.left[rr]; }
      }
      nd.meta[k] = (uint8_t)((((1u << cnt) - 1u) << 5) | po);
      po += cnt;
    }
  }
  nd.imask = (uint8_t)imask;
  A.nodes[A.node_offset + w] = nd;
}

// ---- oversized primitives (see k_flag_big in aob_kernels.cuh) ---------------------------
// A primitive is oversized when its box is longer than a quarter of the scene's largest extent.
AOB_HD bool box_is_oversized(F4 lo, F4 hi, float scene_extent) {
  const float pe = fmaxf(hi.x - lo.x, fmaxf(hi.y - lo.y, hi.z - lo.z));
  return pe > 0.25f * scene_extent;
}
// The extra root: slot 0 = the tree over the ordinary primitives (node `main_root`, box [small_lo, small_hi]), the
// following slots = the oversized primitives (sorted positions [n_small, n_small + n_big), `per_slot` per slot), which
// also complete the leaf order.  Shared by k_super_root and the CPU emulation.
AOB_HD void super_root_body(Node8* out, uint32_t main_root, F4 nlo, F4 nhi, F4 small_lo, F4 small_hi, const F4* plo, const F4* phi,
                            const uint32_t* sorted_prims, uint32_t n_small, uint32_t n_big, uint32_t per_slot, uint32_t prim_offset,
                            uint32_t* leaf_prims) {
  Node8 nd;
  nd.px = nlo.x; nd.py = nlo.y; nd.pz = nlo.z;
  uint32_t ex = quant_exponent(nhi.x - nlo.x), ey = quant_exponent(nhi.y - nlo.y), ez = quant_exponent(nhi.z - nlo.z);
  {
    uint32_t em = ex > ey ? ex : ey;   // same rule as collapse_body: steps within 2^kExpSpread of the widest
    em = em > ez ? em : ez;
    if (em < 24u) em = 24u;
    const uint32_t fl = em - (uint32_t)kExpSpread;
    ex = ex > fl ? ex : fl; ey = ey > fl ? ey : fl; ez = ez > fl ? ez : fl;
  }
  nd.ex = (uint8_t)ex; nd.ey = (uint8_t)ey; nd.ez = (uint8_t)ez;
  nd.child_base = main_root;
  nd.prim_base = prim_offset + n_small;
  nd.imask = 1u;
  for (int k = 0; k < 8; k++) {
    nd.meta[k] = 0;
    for (int a = 0; a < 3; a++) { nd.q[a][k][0] = 255; nd.q[a][k][1] = 0; }
  }
  nd.meta[0] = (uint8_t)(0x20u | 24u);
  quantize_axis(nlo.x, ex, small_lo.x, small_hi.x, &nd.q[0][0][0], &nd.q[0][0][1]);
  quantize_axis(nlo.y, ey, small_lo.y, small_hi.y, &nd.q[1][0][0], &nd.q[1][0][1]);
  quantize_axis(nlo.z, ez, small_lo.z, small_hi.z, &nd.q[2][0][0], &nd.q[2][0][1]);
  uint32_t po = 0;
  for (uint32_t k = 1; k < 8 && po < n_big; k++) {
    const uint32_t cnt = per_slot < n_big - po ? per_slot : n_big - po;
    F4 lo, hi;
    lo.x = lo.y = lo.z = 3.0e38f; hi.x = hi.y = hi.z = -3.0e38f;
    lo.w = hi.w = 0.f;
    for (uint32_t j = 0; j < cnt; j++) {
      const uint32_t p = sorted_prims[n_small + po + j];
      leaf_prims[n_small + po + j] = p;
      lo.x = fminf(lo.x, plo[p].x); lo.y = fminf(lo.y, plo[p].y); lo.z = fminf(lo.z, plo[p].z);
      hi.x = fmaxf(hi.x, phi[p].x); hi.y = fmaxf(hi.y, phi[p].y); hi.z = fmaxf(hi.z, phi[p].z);
    }
    quantize_axis(nlo.x, ex, lo.x, hi.x, &nd.q[0][k][0], &nd.q[0][k][1]);
    quantize_axis(nlo.y, ey, lo.y, hi.y, &nd.q[1][k][0], &nd.q[1][k][1]);
    quantize_axis(nlo.z, ez, lo.z, hi.z, &nd.q[2][k][0], &nd.q[2][k][1]);
    nd.meta[k] = (uint8_t)((((1u << cnt) - 1u) << 5) | po);
    po += cnt;
  }
  *out = nd;
}

// =======================================================================================
// Traversal
// =======================================================================================
struct RayState {
  V3 org, dir;
  float tmin, tmax;
  V3 idir;     // clamped reciprocal for the slab tests
  bool wide;   // fp16 node test not applicable (a direction component below 2^-12, or a non-rigid
               // instance transform): use the fp32 test for this ray
};
AOB_HD float safe_rcp(float d) {  // for the slab tests only; their padding absorbs 2 ulp here
  const float tiny = 1e-18f;
  float a = fabsf(d) > tiny ? d : (d < 0.0f ? -tiny : tiny);
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));  // one MUFU.RCP, <= 1 ulp; no slow path
  return r;
#else
  return 1.0f / a;
#endif
}
// |idir| above this (a direction component below 2^-12) sends the ray to the fp32 node test: the
// fp16 test keeps |A| = 8 |idir| inside the fp16 range only up to here.  About 5 rays in 10^4.
constexpr float kH2MaxIdir = 4096.0f;
AOB_HD bool ray_is_wide(V3 dir, V3 idir) {
  const float m = fmaxf(fmaxf(fabsf(idir.x), fabsf(idir.y)), fabsf(idir.z));
  const float d2 = (dir.x * dir.x + dir.y * dir.y) + dir.z * dir.z;   // object-space rays: rigid instances only
  return !(m <= kH2MaxIdir) || !(fabsf(d2 - 1.0f) <= 1.0e-3f);
}
AOB_HD void ray_setup(RayState& r, V3 org, V3 dir, float tmin, float tmax) {
  r.org = org; r.dir = dir; r.tmin = tmin; r.tmax = tmax;
  r.idir = v3(safe_rcp(dir.x), safe_rcp(dir.y), safe_rcp(dir.z));
  r.wide = ray_is_wide(dir, r.idir);
}

// Constants of the node test, pinned in registers once per thread: ptxas otherwise re-materialises
// the PRMT selectors from uniform registers with a MOV in front of every PRMT (22 MOVs per node on
// the binding ALU pipe).
struct NodeConsts {
  uint32_t k47;     // 0x47000000 = 32768.0f
};
#if defined(__CUDACC__)
// 32768.0f lives in the constant bank on purpose.  PRMT takes one immediate; when both the
// selector and 0x47000000 are compile-time constants ptxas keeps one of them in a scratch
// register that the PRMT itself overwrites and re-materialises it with a MOV before every one of
// the 48 PRMTs of a node test — on the binding ALU pipe.  Read from c[], the value is opaque,
// stays in one register (or is used as a constant-bank operand), and the selectors are immediates.
__constant__ uint32_t c_k47 = 0x47000000u;
#endif
AOB_D NodeConsts make_node_consts() {
  NodeConsts c;
#if defined(__CUDA_ARCH__)
  c.k47 = c_k47;
#else
  c.k47 = 0x47000000u;
#endif
  return c;
}
// (hi << 1) | (lo >> 31): shifts the sign bit of `lo` into an accumulator with one SHF
AOB_D uint32_t shift_in_sign(uint32_t acc, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(lo, acc, 1);
#else
  return (acc << 1) | (lo >> 31);
#endif
}
// Expands the 8 slot hit bits into the traversal mask [31:24] internal slots | [23:0] leaf primitive bits.
AOB_D uint32_t expand_hit_bits(uint32_t hb, uint32_t im, uint32_t meta_lo, uint32_t meta_hi) {
  uint32_t hitmask = (hb & im) << 24;
  uint32_t lh = hb & ~im;
  while (lh) {
    const int s = ffs32(lh) - 1;
    lh &= lh - 1u;
    const uint32_t m = (((s & 4) ? meta_hi : meta_lo) >> (8 * (s & 3))) & 0xffu;
    hitmask |= (m >> 5) << (m & 31u);
  }
  return hitmask;
}

// Slab-tests the 8 quantised child boxes of node `idx` in fp32; returns the hit mask in the layout
// [31:24] internal slots | [23:0] leaf primitive bits.  Used for the few rays the fp16 test below
// cannot take (RayState::wide) and as its cross-check in the emulation tests.
//
// Conservative slabs: t(q) = (32768 + q) * ad + o with o = b - 32768 * ad, b = (p - org) * idir,
// ad = 2^e * idir; the byte q is dropped into the mantissa of 32768.0f with one PRMT (no
// int->float conversion).  Rounding of b (incl. an approximate reciprocal), of the folded
// constant o and of the final FMA is bounded by 5e-7*|b| + 0.008*|ad|; near/far are pushed
// apart by pad = 1e-6*|b| + 0.0234*|ad| (2.3 % of one quantisation step), so a box the exact
// ray touches is never culled and no per-child padding multiply is needed.
// CLAMP_TMAX = false drops the far clamp of the slab interval (8 FMNMX per node on the binding ALU
// pipe): legal whenever tmax exceeds the scene diagonal, which is the AO default (10 x the scene
// extent) — culling by tmax can then never reject a box inside the scene, and the triangle test
// still enforces t < tmax exactly.
// intersect_node8_raw returns the 8 slot hit bits; intersect_node8 expands them into the traversal mask.
template <bool CLAMP_TMAX = true>
AOB_D uint32_t intersect_node8_raw(const U4* nodes, uint32_t idx, const RayState& r, const NodeConsts& nc, uint32_t* child_base,
                                   uint32_t* prim_base, uint32_t* imask, uint32_t* meta_lo, uint32_t* meta_hi) {
  const U4* p = nodes + 5ull * idx;
  const U4 n0 = ld_u4(p), n1 = ld_u4(p + 1), n2 = ld_u4(p + 2), n3 = ld_u4(p + 3), n4 = ld_u4(p + 4);
  *child_base = n1.x;
  *prim_base = n1.y;
  *meta_lo = n1.z;
  *meta_hi = n1.w;
  const uint32_t im = n0.w >> 24;
  *imask = im;
  const float adx = as_float((n0.w & 0xffu) << 23) * r.idir.x;
  const float ady = as_float(((n0.w >> 8) & 0xffu) << 23) * r.idir.y;
  const float adz = as_float(((n0.w >> 16) & 0xffu) << 23) * r.idir.z;
  const float bx = (as_float(n0.x) - r.org.x) * r.idir.x;
  const float by = (as_float(n0.y) - r.org.y) * r.idir.y;
  const float bz = (as_float(n0.z) - r.org.z) * r.idir.z;
  const float ox = fmaf(-32768.0f, adx, bx), oy = fmaf(-32768.0f, ady, by), oz = fmaf(-32768.0f, adz, bz);
  const float padx = fmaf(0.0234375f, fabsf(adx), 1.0e-6f * fabsf(bx));
  const float pady = fmaf(0.0234375f, fabsf(ady), 1.0e-6f * fabsf(by));
  const float padz = fmaf(0.0234375f, fabsf(adz), 1.0e-6f * fabsf(bz));
  const float onx = ox - padx, ofx = ox + padx, ony = oy - pady, ofy = oy + pady, onz = oz - padz, ofz = oz + padz;
  // Slot 2j + s of an axis is (lo, hi) = bytes (2s, 2s + 1) of word j.  For a negative direction
  // component the near plane is the hi byte: swap the two bytes of every pair once per word (one
  // PRMT with a per-ray selector), so that the 48 decoding PRMTs below keep immediate selectors.
  const uint32_t swx = r.dir.x < 0.0f ? 0x2301u : 0x3210u, swy = r.dir.y < 0.0f ? 0x2301u : 0x3210u,
                 swz = r.dir.z < 0.0f ? 0x2301u : 0x3210u;
  const uint32_t wx[4] = {prmt(n2.x, 0u, swx), prmt(n2.y, 0u, swx), prmt(n2.z, 0u, swx), prmt(n2.w, 0u, swx)};
  const uint32_t wy[4] = {prmt(n3.x, 0u, swy), prmt(n3.y, 0u, swy), prmt(n3.z, 0u, swy), prmt(n3.w, 0u, swy)};
  const uint32_t wz[4] = {prmt(n4.x, 0u, swz), prmt(n4.y, 0u, swz), prmt(n4.z, 0u, swz), prmt(n4.w, 0u, swz)};
  // Children are tested from slot 7 down to slot 0; each pushes the sign bit of (tf - tn) into
  // `miss` (one FADD on the FMA pipe + one SHF), so slot s ends up in bit s and a set bit means
  // "missed" (tf < tn).  tf - tn is never NaN (all operands finite) and x - x = +0.
  // 0x74n4 = (k47.b0, byte n of the word, k47.b2, k47.b3) = 32768 + q as a float.
  uint32_t miss = 0;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    const uint32_t sn = (k & 1) ? 0x7424u : 0x7404u, sf = (k & 1) ? 0x7434u : 0x7414u;
    const float tnx = fmaf(as_float(byte_perm(wx[k >> 1], nc.k47, sn)), adx, onx), tfx = fmaf(as_float(byte_perm(wx[k >> 1], nc.k47, sf)), adx, ofx);
    const float tny = fmaf(as_float(byte_perm(wy[k >> 1], nc.k47, sn)), ady, ony), tfy = fmaf(as_float(byte_perm(wy[k >> 1], nc.k47, sf)), ady, ofy);
    const float tnz = fmaf(as_float(byte_perm(wz[k >> 1], nc.k47, sn)), adz, onz), tfz = fmaf(as_float(byte_perm(wz[k >> 1], nc.k47, sf)), adz, ofz);
    const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, r.tmin));
    const float tf = CLAMP_TMAX ? fminf(fminf(tfx, tfy), fminf(tfz, r.tmax)) : fminf(fminf(tfx, tfy), tfz);
    miss = shift_in_sign(miss, as_uint(tf - tn));
  }
  return ~miss & 0xffu;
}
template <bool CLAMP_TMAX = true>
AOB_D uint32_t intersect_node8(const U4* nodes, uint32_t idx, const RayState& r, const NodeConsts& nc, uint32_t* child_base,
                               uint32_t* prim_base, uint32_t* imask) {
  uint32_t mlo, mhi;
  const uint32_t hb = intersect_node8_raw<CLAMP_TMAX>(nodes, idx, r, nc, child_base, prim_base, imask, &mlo, &mhi);
  return expand_hit_bits(hb, *imask, mlo, mhi);
}

// The same test in packed fp16, two planes per instruction (HFMA2 / HMNMX2): the (near, far) bytes of
// a slot are adjacent in the node, so one PRMT against RZ yields them as two fp16 *subnormals*
// (q * 2^-24, exact, no int->float conversion and no bias to remove); one HFMA2 per axis gives
// (-tn_a, tf_a), two 3-input minima reduce over the axes and the ray interval, and one HADD2 with
// lane swizzles gives tf - tn, whose sign bit is the miss bit.  Per slot: 3 PRMT + 3 HFMA2 +
// 2 HMNMX + 1 HADD2 + 1 SHF, against 6 PRMT + 6 FFMA + 3 FMNMX + 1 FADD + 1 SHF in fp32.
//
// fp16 has 11 significant bits, so the planes are evaluated in a frame local to the node:
//   T = (t - t_c) / W,   t_c = (c - org) . dir  (the ray's closest approach to the grid centre c),
//   W = 2^21 * (widest grid step) = 2^13 grid widths.
// If the ray meets the node's box at all, it does so within 0.87 grid widths of t_c, i.e. at
// |T| <= 1.1e-4, where fp16 resolves 6e-8..1.2e-7 = 1/8..1/4 of the widest axis' quantisation step;
// far from t_c the resolution degrades but nothing there can turn a hit into a miss.  Plane q of
// axis a is T = q_h * A + B with q_h = q * 2^-24, A = 2^24 * step_a * idir_a / W (= 8 idir_a on the
// widest axis, so |idir| <= 4096 keeps A finite; other rays take the fp32 test), B = ((p_a - org_a)
// * idir_a - t_c) / W.  Conservative by construction: every plane is pushed outwards by
//   P = 1.6e-8 |A| + 1.0e-3 |B| + 5e-7 |t_c / W| + 6.5e-8
// which bounds the rounding of A to fp16 (255 * 2^-11 steps), of B and of the FMA result (2^-11
// relative, or 2^-25 absolute below the fp16 normal range), the fp32 set-up arithmetic including the
// approximate reciprocal, and the cancellation in B; saturating conversions only ever move a plane
// that is > 65504 away from a hit (any hit has |B| < 1).  NaN can only arise as inf - inf in the
// last add, i.e. for a box that is a miss anyway, and reads as a hit.  tests/emu checks this test
// against brute force and against the fp32 test.
#ifndef AOB_H2_CORNER_FRAME
#define AOB_H2_CORNER_FRAME 0
#endif
// CLAMP_TMAX = false drops the far clamp of the ray interval (legal whenever tmax exceeds the diagonal of the tree being
// walked — the AO default: boxes beyond tmax cannot exist inside it, and the triangle test enforces t < tmax exactly).
template <bool CLAMP_TMAX = true>
AOB_D uint32_t intersect_node8_h2_raw(const U4* nodes, uint32_t idx, const RayState& r, uint32_t* child_base, uint32_t* prim_base,
                                      uint32_t* imask, uint32_t* meta_lo, uint32_t* meta_hi) {
  const U4* p = nodes + 5ull * idx;
  const U4 n0 = ld_u4(p), n1 = ld_u4(p + 1), n2 = ld_u4(p + 2), n3 = ld_u4(p + 3), n4 = ld_u4(p + 4);
  *child_base = n1.x;
  *prim_base = n1.y;
  *meta_lo = n1.z;
  *meta_hi = n1.w;
  const uint32_t im = n0.w >> 24;
  *imask = im;
  const float sx = as_float((n0.w & 0xffu) << 23), sy = as_float(((n0.w >> 8) & 0xffu) << 23), sz = as_float(((n0.w >> 16) & 0xffu) << 23);
  const float smax = fmaxf(fmaxf(sx, sy), sz);
  const float inv = as_float(0x7f000000u - (21u << 23) - as_uint(smax));   // 1 / W, exact (smax is a power of two >= 2^-103)
  const float dx = as_float(n0.x) - r.org.x, dy = as_float(n0.y) - r.org.y, dz = as_float(n0.z) - r.org.z;
#if AOB_H2_CORNER_FRAME
  // round-2 candidate (not measured on the GPU yet): frame origin at the ray's closest approach to
  // the grid *corner* — 3 FFMA less and a shorter dependent chain; hits then sit at |T| <= 2.1e-4.
  const float tc = fmaf(dz, r.dir.z, fmaf(dy, r.dir.y, dx * r.dir.x));
#else
  const float tc = fmaf(fmaf(128.0f, sz, dz), r.dir.z, fmaf(fmaf(128.0f, sy, dy), r.dir.y, fmaf(128.0f, sx, dx) * r.dir.x));
#endif
  const float tcs = tc * inv;
  const float ux = r.idir.x * inv, uy = r.idir.y * inv, uz = r.idir.z * inv;
  const float ax = (sx * 16777216.0f) * ux, ay = (sy * 16777216.0f) * uy, az = (sz * 16777216.0f) * uz;
  const float bx = fmaf(dx, ux, -tcs), by = fmaf(dy, uy, -tcs), bz = fmaf(dz, uz, -tcs);
  const float g = fmaf(5.0e-7f, fabsf(tcs), 6.5e-8f);
  const float px = fmaf(1.6e-8f, fabsf(ax), fmaf(1.0e-3f, fabsf(bx), g));
  const float py = fmaf(1.6e-8f, fabsf(ay), fmaf(1.0e-3f, fabsf(by), g));
  const float pz = fmaf(1.6e-8f, fabsf(az), fmaf(1.0e-3f, fabsf(bz), g));
  // lanes: low = -(near plane), high = far plane
  const uint32_t A2x = h2_pack_sat(ax, -ax), A2y = h2_pack_sat(ay, -ay), A2z = h2_pack_sat(az, -az);
  const uint32_t B2x = h2_pack_sat(bx + px, px - bx), B2y = h2_pack_sat(by + py, py - by), B2z = h2_pack_sat(bz + pz, pz - bz);
  const float q0 = (r.tmin - tc) * inv;
  uint32_t Q2;
  if (CLAMP_TMAX) {
    const float q1 = (r.tmax - tc) * inv;
    Q2 = h2_pack_sat(fmaf(5.2e-4f, fabsf(q1), q1 + 6.5e-8f), fmaf(5.2e-4f, fabsf(q0), 6.5e-8f - q0));
  } else {
    Q2 = h2_pack_sat(65504.0f, fmaf(5.2e-4f, fabsf(q0), 6.5e-8f - q0));   // far lane: the largest finite half
  }
  // byte selectors: (near byte, 0, far byte, 0) with RZ as the second PRMT source; slot 2j + 1 is +0x0202
  const uint32_t s0x = r.dir.x < 0.0f ? 0x4041u : 0x4140u, s0y = r.dir.y < 0.0f ? 0x4041u : 0x4140u,
                 s0z = r.dir.z < 0.0f ? 0x4041u : 0x4140u;
  const uint32_t wx[4] = {n2.x, n2.y, n2.z, n2.w}, wy[4] = {n3.x, n3.y, n3.z, n3.w}, wz[4] = {n4.x, n4.y, n4.z, n4.w};
  uint32_t miss = 0;
#pragma unroll
  for (int k = 7; k >= 0; k--) {
    const uint32_t up = (k & 1) ? 0x0202u : 0u;
    const uint32_t tx = h2_fma(prmt(wx[k >> 1], 0u, s0x + up), A2x, B2x);
    const uint32_t ty = h2_fma(prmt(wy[k >> 1], 0u, s0y + up), A2y, B2y);
    const uint32_t tz = h2_fma(prmt(wz[k >> 1], 0u, s0z + up), A2z, B2z);
    const uint32_t m = h2_min(h2_min(tx, ty), h2_min(tz, Q2));   // (-tn, tf)
    miss = shift_in_sign(miss, h2_lane_sum(m));                  // sign(tf - tn)
  }
  return ~miss & 0xffu;
}
AOB_D uint32_t intersect_node8_h2(const U4* nodes, uint32_t idx, const RayState& r, uint32_t* child_base, uint32_t* prim_base,
                                  uint32_t* imask) {
  uint32_t mlo, mhi;
  const uint32_t hb = intersect_node8_h2_raw<true>(nodes, idx, r, child_base, prim_base, imask, &mlo, &mhi);
  return expand_hit_bits(hb, *imask, mlo, mhi);
}
// The leaf slots `lh` (bits 0..7) of node `idx`: primitive base and the 24-bit primitive mask relative to it.  The fused
// kernel keeps (node, hit leaf slots) while a lane is paused and expands them here, inside the batched block, instead of
// after every node test (the expansion loop runs for two lanes at a time there).
AOB_D uint32_t leaf_slots_to_prims(const U4* nodes, uint32_t idx, uint32_t lh, uint32_t* prim_base) {
  const U4 n1 = ld_u4(nodes + 5ull * idx + 1);
  *prim_base = n1.y;
  uint32_t mask = 0;
  while (lh) {
    const int s = ffs32(lh) - 1;
    lh &= lh - 1u;
    const uint32_t m = (((s & 4) ? n1.w : n1.z) >> (8 * (s & 3))) & 0xffu;
    mask |= (m >> 5) << (m & 31u);
  }
  return mask;
}

struct TraceCounters {
  uint32_t nodes, tris, insts;
};

// Any-hit over one leaf primitive group (bits of `mask` index triangles from `base`).
template <int KZ>
AOB_D bool test_tri_group_k(const F4* tris, uint32_t base, uint32_t mask, V3 org, const Shear& sh, float tmin, float tmax,
                            uint32_t* tested) {
  do {
    const int b = ffs32(mask) - 1;
    mask &= mask - 1u;
    const uint64_t prim = (uint64_t)base + (uint32_t)b;
    const F4 a = ld_f4(tris + 3 * prim), bb = ld_f4(tris + 3 * prim + 1), c = ld_f4(tris + 3 * prim + 2);
    (*tested)++;
    if (woop_hit_k<KZ>(org, sh, tmin, tmax, v3(a.x, a.y, a.z), v3(bb.x, bb.y, bb.z), v3(c.x, c.y, c.z))) return true;
  } while (mask);
  return false;
}
AOB_D bool test_tri_group_sel(const F4* tris, uint32_t base, uint32_t mask, V3 org, const Shear& sh, float tmin, float tmax,
                              uint32_t* tested) {
  do {
    const int b = ffs32(mask) - 1;
    mask &= mask - 1u;
    const uint64_t prim = (uint64_t)base + (uint32_t)b;
    const F4 a = ld_f4(tris + 3 * prim), bb = ld_f4(tris + 3 * prim + 1), c = ld_f4(tris + 3 * prim + 2);
    (*tested)++;
    if (woop_hit_sel(org, sh, tmin, tmax, v3(a.x, a.y, a.z), v3(bb.x, bb.y, bb.z), v3(c.x, c.y, c.z))) return true;
  } while (mask);
  return false;
}
// The Woop shear constants are only needed by rays that reach a triangle: computed here, lazily.
// When every lane of the warp that is in a triangle test right now has the same dominant axis
// the select-free copy for that axis runs (a warp-uniform branch); otherwise one select-based
// copy serves all of them instead of up to three serialised specialised ones.
AOB_D bool test_tri_group(const F4* tris, uint32_t base, uint32_t mask, V3 org, V3 dir, float tmin, float tmax, uint32_t* tested) {
  const Shear sh = make_shear(dir);
#if defined(__CUDA_ARCH__)
  const unsigned am = __activemask();
  const int kz0 = __shfl_sync(am, sh.kz, __ffs((int)am) - 1);
  if (!__all_sync(am, sh.kz == kz0)) return test_tri_group_sel(tris, base, mask, org, sh, tmin, tmax, tested);
#endif
  switch (sh.kz) {
    case 0: return test_tri_group_k<0>(tris, base, mask, org, sh, tmin, tmax, tested);
    case 1: return test_tri_group_k<1>(tris, base, mask, org, sh, tmin, tmax, tested);
    default: return test_tri_group_k<2>(tris, base, mask, org, sh, tmin, tmax, tested);
  }
}

// Conservative world-space pre-test of an instance: can the ray touch its bounding sphere
// (centre s.xyz, padded radius^2 s.w) at all?  ~20 instructions against the ~100 of the exact ray
// transform plus the BLAS root test it avoids; for compact meshes the sphere is far tighter than
// the world AABB in the TLAS (a ball fills 52 % of its bounding cube).
AOB_D bool sphere_may_hit(V3 o, V3 d, F4 s) {
  const float ocx = s.x - o.x, ocy = s.y - o.y, ocz = s.z - o.z;
  const float c = ocx * ocx + ocy * ocy + ocz * ocz - s.w;
  if (c <= 0.0f) return true;                 // origin inside the sphere
  const float b = ocx * d.x + ocy * d.y + ocz * d.z;
  if (b <= 0.0f) return false;                // sphere entirely behind the origin
  const float dd = d.x * d.x + d.y * d.y + d.z * d.z;
  return b * b * 1.00001f >= dd * c;          // discriminant >= 0, with slack for rounding
}

// Node test selection: AOB_H2 = 1 (default) uses the packed-fp16 test for every ray it applies to
// and the fp32 test for the rest; 0 builds the fp32 test only.  trace_any_hit (one ray per thread:
// k_ao_simple, k_trace_rays, the emulation) carries both tests inline and picks per ray; the fused
// persistent kernel cannot afford the registers of both (96 instead of 72) and instead hands the few
// rays the fp16 test cannot take to a second, tiny launch (k_ao_deferred).
#ifndef AOB_H2
#define AOB_H2 1
#endif
template <bool CLAMP_TMAX, bool H2>
AOB_D uint32_t node_test(const U4* nodes, uint32_t idx, const RayState& r, const NodeConsts& nc, uint32_t* cb, uint32_t* pb, uint32_t* im) {
  if (H2 && !r.wide) return intersect_node8_h2(nodes, idx, r, cb, pb, im);
  return intersect_node8<CLAMP_TMAX>(nodes, idx, r, nc, cb, pb, im);
}

// Any-hit traversal of one ray with a caller-provided stack of kStackSize entries.
template <bool STATS, bool H2 = (AOB_H2 != 0)>
AOB_D bool trace_any_hit(const BvhView& bvh, V3 org, V3 dir, float tmin, float tmax, U2* stack, TraceCounters* cnt) {
  RayState r;
  ray_setup(r, org, dir, tmin, tmax);
  const NodeConsts nc = make_node_consts();
  int sp = 0;
  bool in_blas = !bvh.two_level;
  U2 G;
  G.x = bvh.root;
  G.y = (1u << 24) | 1u;
  while (true) {
    U2 T;
    if (G.y & 0xff000000u) {
      const int bit = 31 - clz32(G.y);
      G.y &= ~(1u << bit);
      const uint32_t slot = (uint32_t)bit - 24u;
      const uint32_t node = G.x + (uint32_t)popc32(G.y & 0xffu & ((1u << slot) - 1u));
      if (G.y & 0xff000000u) stack[sp++] = G;
      uint32_t cb, pb, im;
      const uint32_t hm = node_test<true, H2>(bvh.nodes, node, r, nc, &cb, &pb, &im);
      if (STATS) cnt->nodes++;
      G.x = cb; G.y = (hm & 0xff000000u) | im;
      T.x = pb; T.y = hm & 0x00ffffffu;
    } else {
      T = G;
      G.x = 0; G.y = 0;
    }
    if (T.y && in_blas) {
      uint32_t tested = 0;
      const bool hit = test_tri_group(bvh.tris, T.x, T.y, r.org, r.dir, r.tmin, r.tmax, &tested);
      if (STATS) cnt->tris += tested;
      if (hit) return true;
      T.y = 0;
    }
    while (T.y) {
      const int b = ffs32(T.y) - 1;
      T.y &= T.y - 1u;
      const uint32_t prim = T.x + (uint32_t)b;
      if (!sphere_may_hit(org, dir, ld_f4(bvh.insts + (uint64_t)kInstF4 * prim + 4))) continue;
      {
        // instance leaf: save the TLAS continuation, switch to object space
        if (T.y) stack[sp++] = T;
        if (G.y & 0xff000000u) stack[sp++] = G;
        U2 s; s.x = kSentinel; s.y = 0;
        stack[sp++] = s;
        const F4* rec = bvh.insts + (uint64_t)kInstF4 * prim;
        const F4 r0 = ld_f4(rec), r1 = ld_f4(rec + 1), r2 = ld_f4(rec + 2), r3 = ld_f4(rec + 3);
        const float m[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
        if (STATS) cnt->insts++;
        ray_setup(r, xf_point(m, org), xf_vector(m, dir), tmin, tmax);
        in_blas = true;
        G.x = as_uint(r3.x);
        G.y = (1u << 24) | 1u;
        T.y = 0;
      }
    }
    if ((G.y & 0xff000000u) == 0u) {
      while (true) {
        if (sp == 0) return false;
        G = stack[--sp];
        if (G.x == kSentinel && G.y == 0u) {
          ray_setup(r, org, dir, tmin, tmax);
          in_blas = false;
          continue;
        }
        break;
      }
    }
  }
}

}  // namespace aob
