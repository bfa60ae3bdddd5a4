// aob_math.cuh — exactly-rounded fp32/fp64 building blocks shared by every kernel.
//
// Every operation that feeds a bit-exact parity check (sample placement, ray generation,
// the watertight triangle test, instance transforms) goes through ex::mul/add/sub/div/sqrt,
// which map to __fmul_rn/__fadd_rn/... on the device (never contracted into FMAs by nvcc)
// and to plain operators on the host (compiled with -ffp-contract=off).  One IEEE rounding
// per operation, evaluated in the order written — the arithmetic contract of BASELINE.md §4.
//
// The header also compiles under plain g++ (AOB_HOST_EMU) so the host-emulation tests can
// run the same code on CPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define AOB_HD __host__ __device__ __forceinline__
#define AOB_D __device__ __forceinline__
#else
#define AOB_HD inline
#define AOB_D inline
#endif

namespace aob {

struct V3 {
  float x, y, z;
};
AOB_HD V3 v3(float x, float y, float z) {
  V3 r;
  r.x = x; r.y = y; r.z = z;
  return r;
}

namespace ex {
#if defined(__CUDA_ARCH__)
AOB_HD float mul(float a, float b) { return __fmul_rn(a, b); }
AOB_HD float add(float a, float b) { return __fadd_rn(a, b); }
AOB_HD float sub(float a, float b) { return __fsub_rn(a, b); }
AOB_HD float div(float a, float b) { return __fdiv_rn(a, b); }
AOB_HD float sqrt(float a) { return __fsqrt_rn(a); }
AOB_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
AOB_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
AOB_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
AOB_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
AOB_HD double dsqrt(double a) { return __dsqrt_rn(a); }
#else
AOB_HD float mul(float a, float b) { return a * b; }
AOB_HD float add(float a, float b) { return a + b; }
AOB_HD float sub(float a, float b) { return a - b; }
AOB_HD float div(float a, float b) { return a / b; }
AOB_HD float sqrt(float a) { return ::sqrtf(a); }
AOB_HD double dmul(double a, double b) { return a * b; }
AOB_HD double dadd(double a, double b) { return a + b; }
AOB_HD double dsub(double a, double b) { return a - b; }
AOB_HD double ddiv(double a, double b) { return a / b; }
AOB_HD double dsqrt(double a) { return ::sqrt(a); }
#endif
// (a*b + c*d) + e*f
AOB_HD float dot3(float a, float b, float c, float d, float e, float f) {
  return add(add(mul(a, b), mul(c, d)), mul(e, f));
}
}  // namespace ex

AOB_HD V3 sub(V3 a, V3 b) { return v3(ex::sub(a.x, b.x), ex::sub(a.y, b.y), ex::sub(a.z, b.z)); }
AOB_HD V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
AOB_HD float dot(V3 a, V3 b) { return ex::dot3(a.x, b.x, a.y, b.y, a.z, b.z); }
AOB_HD V3 cross(V3 a, V3 b) {
  return v3(ex::sub(ex::mul(a.y, b.z), ex::mul(a.z, b.y)), ex::sub(ex::mul(a.z, b.x), ex::mul(a.x, b.z)),
            ex::sub(ex::mul(a.x, b.y), ex::mul(a.y, b.x)));
}
AOB_HD V3 normalize(V3 a) {
  float len = ex::sqrt(dot(a, a));
  if (!(len > 0.0f)) return a;
  return v3(ex::div(a.x, len), ex::div(a.y, len), ex::div(a.z, len));
}
AOB_HD float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

// world = M * (v,1) for a row-major matrix whose first 12 floats are the 3x4 affine part.
AOB_HD V3 xf_point(const float* m, V3 v) {
  return v3(ex::add(ex::add(ex::add(ex::mul(m[0], v.x), ex::mul(m[1], v.y)), ex::mul(m[2], v.z)), m[3]),
            ex::add(ex::add(ex::add(ex::mul(m[4], v.x), ex::mul(m[5], v.y)), ex::mul(m[6], v.z)), m[7]),
            ex::add(ex::add(ex::add(ex::mul(m[8], v.x), ex::mul(m[9], v.y)), ex::mul(m[10], v.z)), m[11]));
}
AOB_HD V3 xf_vector(const float* m, V3 d) {
  return v3(ex::dot3(m[0], d.x, m[1], d.y, m[2], d.z), ex::dot3(m[4], d.x, m[5], d.y, m[6], d.z),
            ex::dot3(m[8], d.x, m[9], d.y, m[10], d.z));
}
// n_world = (inv 3x3)^T * n
AOB_HD V3 xf_normal(const float* inv, V3 n) {
  return v3(ex::dot3(inv[0], n.x, inv[4], n.y, inv[8], n.z), ex::dot3(inv[1], n.x, inv[5], n.y, inv[9], n.z),
            ex::dot3(inv[2], n.x, inv[6], n.y, inv[10], n.z));
}

// ---- RNG: tea<N>, lcg, rnd of the reference's random.h -----------------
template <unsigned N>
AOB_HD uint32_t tea(uint32_t v0, uint32_t v1) {
  uint32_t s0 = 0;
#pragma unroll
  for (unsigned n = 0; n < N; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
AOB_HD uint32_t lcg(uint32_t& s) {
  s = 1664525u * s + 1013904223u;
  return s & 0x00FFFFFFu;
}
// lcg / 2^24: the division by a power of two is exact, so the multiply is bit-identical
AOB_HD float rnd(uint32_t& s) { return ex::mul((float)lcg(s), 5.9604644775390625e-8f); }

// radical inverse, fp32 accumulate (bake_sample.cpp; decision #4)
AOB_HD float halton(uint32_t i, uint32_t base) {
  const float inv_base = ex::div(1.0f, (float)base);
  float f = inv_base, r = 0.0f;
  while (i) {
    r = ex::add(r, ex::mul(f, (float)(i % base)));
    i /= base;
    f = ex::mul(f, inv_base);
  }
  return r;
}

// cos/sin(2*pi*u), u in [0,1): quadrant reduction + fixed polynomials (decision #11).
AOB_HD void sincos2pi(float u, float* c, float* s) {
  int qi = (int)floorf(ex::add(ex::mul(4.0f, u), 0.5f));
  float r = ex::sub(u, ex::mul(0.25f, (float)qi));
  float th = ex::mul(6.28318548202514648f, r);
  float t2 = ex::mul(th, th);
  float sp = ex::add(ex::mul(-1.9515295891e-4f, t2), 8.3321608736e-3f);
  sp = ex::add(ex::mul(sp, t2), -1.6666654611e-1f);
  float sn = ex::add(ex::mul(ex::mul(sp, t2), th), th);
  float cp = ex::add(ex::mul(2.443315711809948e-5f, t2), -1.388731625493765e-3f);
  cp = ex::add(ex::mul(cp, t2), 4.166664568298827e-2f);
  float cs = ex::add(ex::mul(ex::mul(cp, t2), t2), ex::sub(1.0f, ex::mul(0.5f, t2)));
  switch (qi & 3) {
    case 0: *c = cs; *s = sn; break;
    case 1: *c = -sn; *s = cs; break;
    case 2: *c = -cs; *s = -sn; break;
    default: *c = sn; *s = -cs; break;
  }
}

// 0.5*|e0 x e1|: cross in fp32, norm in fp64
AOB_HD double tri_area(V3 w0, V3 w1, V3 w2) {
  V3 c = cross(sub(w1, w0), sub(w2, w0));
  double cx = c.x, cy = c.y, cz = c.z;
  return ex::dmul(0.5, ex::dsqrt(ex::dadd(ex::dadd(ex::dmul(cx, cx), ex::dmul(cy, cy)), ex::dmul(cz, cz))));
}

// ---- ray generation (bake_kernels.cu generateRaysKernel) ------------------
AOB_HD int sqrt_rays(int rays_per_sample) { return (int)(ex::add(ex::sqrt((float)rays_per_sample), 0.5f)); }

struct Onb {
  V3 t, b;
};
AOB_HD Onb make_onb(V3 n) {
  Onb o;
  if (fabsf(n.x) > fabsf(n.z)) o.b = v3(-n.y, n.x, 0.0f);
  else o.b = v3(0.0f, -n.z, n.y);
  o.b = normalize(o.b);
  o.t = cross(o.b, n);
  return o;
}
AOB_HD V3 cosine_dir(float u0, float u1, V3 n, const Onb& o) {
  float r = ex::sqrt(u0), c, s;
  sincos2pi(u1, &c, &s);
  float x = ex::mul(r, c), y = ex::mul(r, s);
  float z = ex::sqrt(fmaxf(0.0f, ex::sub(ex::sub(1.0f, ex::mul(x, x)), ex::mul(y, y))));
  return v3(ex::add(ex::add(ex::mul(x, o.t.x), ex::mul(y, o.b.x)), ex::mul(z, n.x)),
            ex::add(ex::add(ex::mul(x, o.t.y), ex::mul(y, o.b.y)), ex::mul(z, n.y)),
            ex::add(ex::add(ex::mul(x, o.t.z), ex::mul(y, o.b.z)), ex::mul(z, n.z)));
}
// direction of stratum `pass` (= px*q+py) of global sample g.
AOB_HD V3 ao_ray_dir(uint32_t g, uint32_t pass, int q, V3 n, V3 fn, const Onb& onb) {
  // (p + rnd) / q: for power-of-two q (all BASELINE configs) the reciprocal is exact and the
  // multiply is bit-identical to the IEEE division; other q take the division.
  const bool pow2 = (q & (q - 1)) == 0;
#if defined(__CUDA_ARCH__)
  const uint32_t px = pow2 ? pass >> (__ffs(q) - 1) : pass / (uint32_t)q;  // no runtime integer division on the common path
#else
  const uint32_t px = pass / (uint32_t)q;
#endif
  const uint32_t py = pass - px * (uint32_t)q;
  uint32_t seed = tea<2>((pass << 16) | pass, g);
  const float fq = (float)q, rq = ex::div(1.0f, fq);
  const float s0 = ex::add((float)px, rnd(seed)), s1 = ex::add((float)py, rnd(seed));
  float u0 = pow2 ? ex::mul(s0, rq) : ex::div(s0, fq);
  float u1 = pow2 ? ex::mul(s1, rq) : ex::div(s1, fq);
  V3 d = v3(0.f, 0.f, 0.f);
  for (int attempt = 0; attempt < 5; attempt++) {
    d = cosine_dir(u0, u1, n, onb);
    if (dot(d, fn) > 0.0f) break;
    u0 = rnd(seed);
    u1 = rnd(seed);
  }
  return d;
}
AOB_HD V3 ao_ray_origin(V3 p, V3 n, float offset) {
  return v3(ex::add(p.x, ex::mul(offset, n.x)), ex::add(p.y, ex::mul(offset, n.y)), ex::add(p.z, ex::mul(offset, n.z)));
}

// ---- watertight ray/triangle (Woop, Benthin, Wald 2013), any-hit with (tmin, tmax) --------
// kz = dominant axis of the direction, (kx, ky) the next two cyclically; Sz = 1/d[kz] (IEEE),
// Sx = d[kx]*Sz, Sy = d[ky]*Sz.  The paper's kx<->ky swap for d[kz] < 0 only flips the sign of
// U, V, W, det and T *exactly* (fl(a-b) = -fl(b-a)), which neither the same-sign test nor the
// range test can see, so it is omitted here (the oracle keeps it; decisions are identical).
// The test is specialised on kz at compile time: no runtime component selects, everything on
// the FMA pipe, and the three copies cost nothing extra because only a lane or two of a warp
// are in a triangle test at any time.
struct Shear {
  int kz;
  float Sx, Sy, Sz;
};
AOB_HD Shear make_shear(V3 d) {
  Shear s;
  float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
  s.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  const float dz = s.kz == 0 ? d.x : (s.kz == 1 ? d.y : d.z);
  const float dx = s.kz == 0 ? d.y : (s.kz == 1 ? d.z : d.x);
  const float dy = s.kz == 0 ? d.z : (s.kz == 1 ? d.x : d.y);
  s.Sz = ex::div(1.0f, dz);
  s.Sx = ex::mul(dx, s.Sz);
  s.Sy = ex::mul(dy, s.Sz);
  return s;
}
template <int KZ>
AOB_HD bool woop_hit_k(V3 org, const Shear& s, float tmin, float tmax, V3 p0, V3 p1, V3 p2) {
  const V3 A = sub(p0, org), B = sub(p1, org), C = sub(p2, org);
  // (kx, ky, kz) = (KZ+1, KZ+2, KZ) mod 3
  const float Akz = KZ == 0 ? A.x : (KZ == 1 ? A.y : A.z), Akx = KZ == 0 ? A.y : (KZ == 1 ? A.z : A.x), Aky = KZ == 0 ? A.z : (KZ == 1 ? A.x : A.y);
  const float Bkz = KZ == 0 ? B.x : (KZ == 1 ? B.y : B.z), Bkx = KZ == 0 ? B.y : (KZ == 1 ? B.z : B.x), Bky = KZ == 0 ? B.z : (KZ == 1 ? B.x : B.y);
  const float Ckz = KZ == 0 ? C.x : (KZ == 1 ? C.y : C.z), Ckx = KZ == 0 ? C.y : (KZ == 1 ? C.z : C.x), Cky = KZ == 0 ? C.z : (KZ == 1 ? C.x : C.y);
  const float Ax = ex::sub(Akx, ex::mul(s.Sx, Akz)), Ay = ex::sub(Aky, ex::mul(s.Sy, Akz));
  const float Bx = ex::sub(Bkx, ex::mul(s.Sx, Bkz)), By = ex::sub(Bky, ex::mul(s.Sy, Bkz));
  const float Cx = ex::sub(Ckx, ex::mul(s.Sx, Ckz)), Cy = ex::sub(Cky, ex::mul(s.Sy, Ckz));
  float U = ex::sub(ex::mul(Cx, By), ex::mul(Cy, Bx));
  float V = ex::sub(ex::mul(Ax, Cy), ex::mul(Ay, Cx));
  float W = ex::sub(ex::mul(Bx, Ay), ex::mul(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)ex::dsub(ex::dmul((double)Cx, (double)By), ex::dmul((double)Cy, (double)Bx));
    V = (float)ex::dsub(ex::dmul((double)Ax, (double)Cy), ex::dmul((double)Ay, (double)Cx));
    W = (float)ex::dsub(ex::dmul((double)Bx, (double)Ay), ex::dmul((double)By, (double)Ax));
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = ex::add(ex::add(U, V), W);
  if (det == 0.0f) return false;
  const float Az = ex::mul(s.Sz, Akz), Bz = ex::mul(s.Sz, Bkz), Cz = ex::mul(s.Sz, Ckz);
  const float T = ex::add(ex::add(ex::mul(U, Az), ex::mul(V, Bz)), ex::mul(W, Cz));
  // division-free range test: T*sign(det) against t*|det|
  const float ad = fabsf(det);
  const float Ts = det < 0.0f ? -T : T;
  return (Ts > ex::mul(tmin, ad)) && (Ts < ex::mul(tmax, ad));
}
// Same test with the axis permutation done by selects (branch-free): identical arithmetic after
// the selection, so identical decisions.  Used when the lanes of a warp that are in a triangle
// test at the same time disagree on kz — one select path then beats up to three serialised
// specialised copies.
AOB_HD bool woop_hit_sel(V3 org, const Shear& s, float tmin, float tmax, V3 p0, V3 p1, V3 p2) {
  const V3 A = sub(p0, org), B = sub(p1, org), C = sub(p2, org);
  const bool z0 = s.kz == 0, z1 = s.kz == 1;
  const float Akz = z0 ? A.x : (z1 ? A.y : A.z), Akx = z0 ? A.y : (z1 ? A.z : A.x), Aky = z0 ? A.z : (z1 ? A.x : A.y);
  const float Bkz = z0 ? B.x : (z1 ? B.y : B.z), Bkx = z0 ? B.y : (z1 ? B.z : B.x), Bky = z0 ? B.z : (z1 ? B.x : B.y);
  const float Ckz = z0 ? C.x : (z1 ? C.y : C.z), Ckx = z0 ? C.y : (z1 ? C.z : C.x), Cky = z0 ? C.z : (z1 ? C.x : C.y);
  const float Ax = ex::sub(Akx, ex::mul(s.Sx, Akz)), Ay = ex::sub(Aky, ex::mul(s.Sy, Akz));
  const float Bx = ex::sub(Bkx, ex::mul(s.Sx, Bkz)), By = ex::sub(Bky, ex::mul(s.Sy, Bkz));
  const float Cx = ex::sub(Ckx, ex::mul(s.Sx, Ckz)), Cy = ex::sub(Cky, ex::mul(s.Sy, Ckz));
  float U = ex::sub(ex::mul(Cx, By), ex::mul(Cy, Bx));
  float V = ex::sub(ex::mul(Ax, Cy), ex::mul(Ay, Cx));
  float W = ex::sub(ex::mul(Bx, Ay), ex::mul(By, Ax));
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)ex::dsub(ex::dmul((double)Cx, (double)By), ex::dmul((double)Cy, (double)Bx));
    V = (float)ex::dsub(ex::dmul((double)Ax, (double)Cy), ex::dmul((double)Ay, (double)Cx));
    W = (float)ex::dsub(ex::dmul((double)Bx, (double)Ay), ex::dmul((double)By, (double)Ax));
  }
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  const float det = ex::add(ex::add(U, V), W);
  if (det == 0.0f) return false;
  const float Az = ex::mul(s.Sz, Akz), Bz = ex::mul(s.Sz, Bkz), Cz = ex::mul(s.Sz, Ckz);
  const float T = ex::add(ex::add(ex::mul(U, Az), ex::mul(V, Bz)), ex::mul(W, Cz));
  const float ad = fabsf(det);
  const float Ts = det < 0.0f ? -T : T;
  return (Ts > ex::mul(tmin, ad)) && (Ts < ex::mul(tmax, ad));
}
AOB_HD bool woop_hit(V3 org, const Shear& s, float tmin, float tmax, V3 p0, V3 p1, V3 p2) {
  switch (s.kz) {
    case 0: return woop_hit_k<0>(org, s, tmin, tmax, p0, p1, p2);
    case 1: return woop_hit_k<1>(org, s, tmin, tmax, p0, p1, p2);
    default: return woop_hit_k<2>(org, s, tmin, tmax, p0, p1, p2);
  }
}

}  // namespace aob
