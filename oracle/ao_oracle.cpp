// ao_oracle.cpp — single-source C++/OpenMP CPU restatement of the optix_prime_baking
// ambient-occlusion bake path.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product (libaobake.so, the
// optix_prime_baking_b200 package) may include, link, import or execute this file.
// It is the checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs, and nothing else.
//
// *** PARITY UNPINNED. ***  /root/reference holds only a deprecation README and a licence
// (SURVEY.md §0): the pre-deprecation sources (bake_api.h, bake_sample.cpp,
// bake_kernels.cu, bake_ao_optix_prime.cpp, bake_filter*.cpp, random.h) are absent, the
// dominant step lived in the closed OptiX Prime library, and the sample shipped no tests
// or golden vectors.  No file:line into the reference can be checked.  Each function below
// therefore cites the *recalled* reference file (no line numbers — any would be
// fabricated) and the BASELINE.md §4 / SURVEY.md §8(a) row it restates.  Where recall
// could not settle a detail the choice made here is normative (DESIGN.md "Oracle
// decisions").
//
// Two walks of the same SAH tree answer the any-hit query — binary with a scalar slab test, and (where the CPU has AVX2,
// chosen at run time) the tree collapsed 8-wide with an 8-lane slab test of the same arithmetic — and return the same
// answer for every ray; the wide one exists so that bench.py's CPU arm is a fair yardstick, not a naive one.
//
// Build: g++ -O2 -fopenmp -ffp-contract=off -shared -fPIC  (see oracle/Makefile).
// -ffp-contract=off matters: every fp32/fp64 expression below is evaluated exactly as
// written (one IEEE rounding per operation, left to right), which is what the CUDA path
// reproduces with __fmul_rn/__fadd_rn/... so that sample points and rays are bit-exact.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__) && (defined(__GNUC__) || defined(__clang__))
#include <immintrin.h>
#define AO_ORACLE_HAVE_AVX2_PATH 1
#else
#define AO_ORACLE_HAVE_AVX2_PATH 0
#endif

// ----------------------------------------------------------------------------------
// POD types — layout-identical to include/aobake.h (restated, not included, so that the
// oracle stays single-source).  Mirror bake::Mesh/Instance/Scene/SampleInfo/AOSamples
// (bake_api.h, SURVEY §8 a1–a4).
// ----------------------------------------------------------------------------------
extern "C" {
typedef struct {
  uint64_t num_vertices;
  const float* vertices;
  uint32_t vertex_stride_bytes;  // 0 => 12
  const float* normals;          // nullable
  uint32_t normal_stride_bytes;  // 0 => 12
  uint64_t num_triangles;
  const uint32_t* tri_vertex_indices;
  float bbox_min[3];
  float bbox_max[3];
} OrMesh;
typedef struct {
  float xform[16];  // row-major 4x4, affine
  uint64_t storage_identifier;
  uint32_t mesh_index;
  float bbox_min[3];
  float bbox_max[3];
} OrInstance;
typedef struct {
  const OrMesh* meshes;
  uint64_t num_meshes;
  const OrInstance* instances;
  uint64_t num_instances;
} OrScene;
typedef struct {
  uint32_t tri_idx;
  float bary[3];
  float dA;
} OrSampleInfo;
typedef struct {
  uint64_t num_samples;
  float* sample_positions;
  float* sample_normals;
  float* sample_face_normals;
  OrSampleInfo* sample_infos;
} OrSamples;
}

namespace {

struct V3 {
  float x, y, z;
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// normalize: v / sqrt(dot(v,v)) with IEEE sqrt and IEEE divide; zero vectors stay zero.
inline V3 normalize(V3 a) {
  float len = std::sqrt(dot(a, a));
  if (!(len > 0.0f)) return a;
  return v3(a.x / len, a.y / len, a.z / len);
}
inline float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

inline const float* vertex_ptr(const OrMesh& m, uint64_t i) {
  uint32_t s = m.vertex_stride_bytes ? m.vertex_stride_bytes : 12u;
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(m.vertices) + i * s);
}
inline const float* normal_ptr(const OrMesh& m, uint64_t i) {
  uint32_t s = m.normal_stride_bytes ? m.normal_stride_bytes : 12u;
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(m.normals) + i * s);
}
inline V3 load3(const float* p) { return v3(p[0], p[1], p[2]); }

// world = M * (v,1), fp32, ((m0*x + m1*y) + m2*z) + m3   [bake_sample.cpp: xform*v]
inline V3 xf_point(const float* m, V3 v) {
  return v3(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3],
            ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7],
            ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11]);
}
// inverse of the affine part (3x4, row-major) in fp64 by cofactors, rounded to fp32 once.
// [bake_sample.cpp: xform.inverse(); bake_ao_optix_prime.cpp passes xforms to Prime]
void affine_inverse(const float* m, float* inv /*12*/) {
  double a00 = m[0], a01 = m[1], a02 = m[2], t0 = m[3];
  double a10 = m[4], a11 = m[5], a12 = m[6], t1 = m[7];
  double a20 = m[8], a21 = m[9], a22 = m[10], t2 = m[11];
  double c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  double c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
  double c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
  double det = (a00 * c00 + a01 * c10) + a02 * c20;
  double i00 = c00 / det, i01 = c01 / det, i02 = c02 / det;
  double i10 = c10 / det, i11 = c11 / det, i12 = c12 / det;
  double i20 = c20 / det, i21 = c21 / det, i22 = c22 / det;
  inv[0] = (float)i00; inv[1] = (float)i01; inv[2] = (float)i02;
  inv[3] = (float)(-((i00 * t0 + i01 * t1) + i02 * t2));
  inv[4] = (float)i10; inv[5] = (float)i11; inv[6] = (float)i12;
  inv[7] = (float)(-((i10 * t0 + i11 * t1) + i12 * t2));
  inv[8] = (float)i20; inv[9] = (float)i21; inv[10] = (float)i22;
  inv[11] = (float)(-((i20 * t0 + i21 * t1) + i22 * t2));
}
// n_world = (M^-1)^T * n  using the fp32 3x4 inverse (column access = transpose).
inline V3 xf_normal(const float* inv, V3 n) {
  return v3((inv[0] * n.x + inv[4] * n.y) + inv[8] * n.z,
            (inv[1] * n.x + inv[5] * n.y) + inv[9] * n.z,
            (inv[2] * n.x + inv[6] * n.y) + inv[10] * n.z);
}
inline V3 xf_vector(const float* m12, V3 d) {  // 3x3 part of a 3x4
  return v3((m12[0] * d.x + m12[1] * d.y) + m12[2] * d.z,
            (m12[4] * d.x + m12[5] * d.y) + m12[6] * d.z,
            (m12[8] * d.x + m12[9] * d.y) + m12[10] * d.z);
}

// ----------------------------------------------------------------------------------
// RNG — random.h (OptiX SDK sample header): tea<N>, lcg, rnd.  SURVEY §8 a8.
// ----------------------------------------------------------------------------------
inline uint32_t tea(uint32_t N, uint32_t v0, uint32_t v1) {
  uint32_t s0 = 0;
  for (uint32_t n = 0; n < N; n++) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
inline uint32_t lcg(uint32_t& s) {
  s = 1664525u * s + 1013904223u;
  return s & 0x00FFFFFFu;
}
inline float rnd(uint32_t& s) { return (float)lcg(s) / 16777216.0f; }

// halton radical inverse in fp32 (bake_sample.cpp; decision #4): index i >= 1.
inline float halton(uint32_t i, uint32_t base) {
  const float inv_base = 1.0f / (float)base;
  float f = inv_base, r = 0.0f;
  while (i) {
    r = r + f * (float)(i % base);
    i /= base;
    f = f * inv_base;
  }
  return r;
}

// cos/sin of 2*pi*u, u in [0,1): quadrant reduction + fixed polynomials in fp32.
// The reference called cosf/sinf (optixu cosine_sample_hemisphere), whose last-ulp results
// differ between glibc and CUDA libdevice; the oracle pins one formula so that rays are
// bit-exact across CPU and GPU (decision #11).  Max abs error ~1.2e-7.
inline void sincos2pi(float u, float* c, float* s) {
  int qi = (int)std::floor(4.0f * u + 0.5f);
  float r = u - 0.25f * (float)qi;          // exact
  float th = 6.28318548202514648f * r;      // |th| <= pi/4
  float t2 = th * th;
  float sp = -1.9515295891e-4f * t2 + 8.3321608736e-3f;
  sp = sp * t2 + -1.6666654611e-1f;
  float sn = (sp * t2) * th + th;
  float cp = 2.443315711809948e-5f * t2 + -1.388731625493765e-3f;
  cp = cp * t2 + 4.166664568298827e-2f;
  float cs = (cp * t2) * t2 + (1.0f - 0.5f * t2);
  switch (qi & 3) {
    case 0: *c = cs; *s = sn; break;
    case 1: *c = -sn; *s = cs; break;
    case 2: *c = -cs; *s = -sn; break;
    default: *c = sn; *s = -cs; break;
  }
}

// ----------------------------------------------------------------------------------
// Areas and the sample budget — bake_sample.cpp / bake_sample_internal.h, SURVEY a5–a6.
// ----------------------------------------------------------------------------------
inline void tri_world(const OrMesh& m, const float* xf, uint64_t t, V3* w) {
  const uint32_t* idx = m.tri_vertex_indices + 3 * t;
  for (int k = 0; k < 3; k++) w[k] = xf_point(xf, load3(vertex_ptr(m, idx[k])));
}
// 0.5*|e0 x e1|: cross in fp32, norm in fp64.
inline double tri_area(const V3* w) {
  V3 c = cross(sub(w[1], w[0]), sub(w[2], w[0]));
  double cx = c.x, cy = c.y, cz = c.z;
  return 0.5 * std::sqrt((cx * cx + cy * cy) + cz * cz);
}
// Fixed-shape fp64 sum (decision #3): sequential inside 1024-element blocks, then
// sequential over the block sums.
double blocked_sum(const double* a, uint64_t n) {
  double total = 0.0;
  for (uint64_t b = 0; b < n; b += 1024) {
    uint64_t e = std::min<uint64_t>(n, b + 1024);
    double s = 0.0;
    for (uint64_t i = b; i < e; i++) s = s + a[i];
    total = total + s;
  }
  return total;
}
void instance_tri_areas(const OrScene& sc, uint64_t inst, std::vector<double>& areas) {
  const OrInstance& I = sc.instances[inst];
  const OrMesh& m = sc.meshes[I.mesh_index];
  areas.resize(m.num_triangles);
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < (int64_t)m.num_triangles; t++) {
    V3 w[3];
    tri_world(m, I.xform, (uint64_t)t, w);
    areas[t] = tri_area(w);
  }
}
// distribute_samples_generic (bake_sample_internal.h): minimum per element, then
// floor(Na*area/total) each, then +1 to elements 0,1,2,... until the total is met.
// Returns 0, or -1 if fp rounding made the floors exceed the budget (never seen; loud).
int distribute_generic(uint64_t n, const uint64_t* mins, const double* areas, double total_area,
                       uint64_t N, uint64_t* counts) {
  uint64_t summin = 0;
  for (uint64_t i = 0; i < n; i++) summin += mins[i];
  if (N < summin) N = summin;
  const uint64_t Na = N - summin;
  uint64_t assigned = 0;
  for (uint64_t i = 0; i < n; i++) {
    uint64_t c = mins[i];
    if (Na > 0 && total_area > 0.0) c += (uint64_t)(((double)Na * areas[i]) / total_area);
    counts[i] = c;
    assigned += c;
  }
  if (assigned > N) return -1;
  uint64_t left = N - assigned;
  if (n == 0) return left ? -1 : 0;
  // left < n whenever total_area > 0; if every area is zero the sweep wraps round.
  for (uint64_t i = 0; left > 0; i = (i + 1) % n, left--) counts[i] += 1;
  return 0;
}

// ----------------------------------------------------------------------------------
// Ray / triangle / box primitives.  The any-hit predicate (decision #1): a ray is occluded
// iff some triangle is hit with tmin < t < tmax under the watertight test of Woop, Benthin,
// Wald, "Watertight Ray/Triangle Intersection" (JCGT 2013), no back-face culling.
// ----------------------------------------------------------------------------------
struct Ray {
  V3 o, d;
  float tmin, tmax;
};
struct RayShear {
  int kx, ky, kz;
  float Sx, Sy, Sz;
};
inline RayShear make_shear(V3 d) {
  RayShear s;
  float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
  s.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
  s.kx = s.kz + 1; if (s.kx == 3) s.kx = 0;
  s.ky = s.kx + 1; if (s.ky == 3) s.ky = 0;
  if (comp(d, s.kz) < 0.0f) std::swap(s.kx, s.ky);
  s.Sz = 1.0f / comp(d, s.kz);
  s.Sx = comp(d, s.kx) * s.Sz;
  s.Sy = comp(d, s.ky) * s.Sz;
  return s;
}
// Returns true on hit; *margin (optional) = min(|U|,|V|,|W|)/|det| (distance to the nearest
// edge in barycentric units) for diagnostics.
inline bool woop_hit(const Ray& r, const RayShear& s, V3 p0, V3 p1, V3 p2, float* t_out,
                     float* margin) {
  V3 A = sub(p0, r.o), B = sub(p1, r.o), C = sub(p2, r.o);
  float Ax = comp(A, s.kx) - s.Sx * comp(A, s.kz), Ay = comp(A, s.ky) - s.Sy * comp(A, s.kz);
  float Bx = comp(B, s.kx) - s.Sx * comp(B, s.kz), By = comp(B, s.ky) - s.Sy * comp(B, s.kz);
  float Cx = comp(C, s.kx) - s.Sx * comp(C, s.kz), Cy = comp(C, s.ky) - s.Sy * comp(C, s.kz);
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if (margin) *margin = std::numeric_limits<float>::infinity();
  if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
  float det = (U + V) + W;
  if (det == 0.0f) return false;
  float Az = s.Sz * comp(A, s.kz), Bz = s.Sz * comp(B, s.kz), Cz = s.Sz * comp(C, s.kz);
  float T = (U * Az + V * Bz) + W * Cz;
  // range test without a division (Woop et al. §3.2): compare T*sign(det) with t*|det|
  const float ad = std::fabs(det);
  const float Ts = det < 0.0f ? -T : T;
  if (t_out) *t_out = T / det;
  if (margin) *margin = std::min(std::fabs(U), std::min(std::fabs(V), std::fabs(W))) / ad;
  return (Ts > r.tmin * ad) && (Ts < r.tmax * ad);
}

struct Box {
  float lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; k++) { lo[k] = std::numeric_limits<float>::max(); hi[k] = -std::numeric_limits<float>::max(); }
  }
  void grow(V3 p) {
    lo[0] = std::min(lo[0], p.x); lo[1] = std::min(lo[1], p.y); lo[2] = std::min(lo[2], p.z);
    hi[0] = std::max(hi[0], p.x); hi[1] = std::max(hi[1], p.y); hi[2] = std::max(hi[2], p.z);
  }
  void grow(const Box& b) {
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); }
  }
  float half_area() const {
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
  }
};
// Conservative slab test: subtract-then-multiply (no cancellation against a far origin),
// far distance padded by 4 ulp.
inline bool box_hit(const Box& b, const Ray& r, V3 id) {
  float tn = r.tmin, tf = r.tmax;
  const float o[3] = {r.o.x, r.o.y, r.o.z}, iv[3] = {id.x, id.y, id.z};
  for (int k = 0; k < 3; k++) {
    const float t0 = (b.lo[k] - o[k]) * iv[k], t1 = (b.hi[k] - o[k]) * iv[k];
    // 0 * inf: the origin sits exactly on a slab plane of an exactly-zero direction component —
    // it is inside this slab (conservative), no constraint
    if (t0 != t0 || t1 != t1) continue;
    const float a = t0 < t1 ? t0 : t1, c = t0 < t1 ? t1 : t0;
    if (a > tn) tn = a;
    if (c < tf) tf = c;
  }
  return tn <= tf * 1.0000005f;
}

// Binary BVH, binned SAH (16 bins), leaves <= 4 primitives.  Generic over primitive boxes.
struct BNode {
  Box box;
  uint32_t left;   // internal: index of left child (right = left+1); leaf: first prim
  uint32_t count;  // 0 => internal
};
// 8-wide collapse of the binary tree for the AVX2 traversal (below): child boxes in SoA form, one 8-lane slab test per
// node.  A slot is empty (cnt == 0), a leaf (cnt = 1..4 primitives from prim[ref]) or an internal child (cnt == kWideInternal,
// ref = index of the wide node).  Every box is the box of a node of the binary tree, unchanged.
constexpr uint32_t kWideInternal = 0xffffffffu;
struct alignas(32) WNode {
  float lo[3][8], hi[3][8];
  uint32_t ref[8], cnt[8];
};
struct Bvh {
  std::vector<BNode> nodes;
  std::vector<uint32_t> prim;  // permutation
  std::vector<WNode> wide;     // empty: binary traversal only
};
// Traversal used by TriSoup::any_hit: 0 = auto (8-wide AVX2 where the CPU has it, else binary scalar), 1 = binary scalar,
// 2 = 8-wide AVX2 (binary if the CPU lacks AVX2).  Both walk boxes of the same SAH tree with the same slab arithmetic and
// test triangles with the same scalar watertight test, so they return the same answer for every ray (any-hit does not
// depend on the order of the walk); the wide one is what a production CPU tracer does and is the fairer yardstick for
// bench.py's CPU arm.
int g_traversal_mode = 0;
bool cpu_has_avx2() {
#if AO_ORACLE_HAVE_AVX2_PATH
  static const bool ok = __builtin_cpu_supports("avx2");
  return ok;
#else
  return false;
#endif
}
bool use_wide_traversal() { return g_traversal_mode != 1 && cpu_has_avx2(); }
void bvh_build(const std::vector<Box>& pb, Bvh& out) {
  const uint32_t n = (uint32_t)pb.size();
  out.prim.resize(n);
  std::iota(out.prim.begin(), out.prim.end(), 0u);
  out.nodes.clear();
  out.nodes.reserve(n ? 2 * n : 1);
  std::vector<float> cx(n), cy(n), cz(n);
  for (uint32_t i = 0; i < n; i++) {
    cx[i] = 0.5f * (pb[i].lo[0] + pb[i].hi[0]);
    cy[i] = 0.5f * (pb[i].lo[1] + pb[i].hi[1]);
    cz[i] = 0.5f * (pb[i].lo[2] + pb[i].hi[2]);
  }
  const std::vector<float>* cc[3] = {&cx, &cy, &cz};
  struct Item { uint32_t node, first, count; };
  std::vector<Item> todo;
  out.nodes.push_back(BNode());
  todo.push_back({0, 0, n});
  while (!todo.empty()) {
    Item it = todo.back();
    todo.pop_back();
    Box nb, cb;
    nb.reset(); cb.reset();
    for (uint32_t i = it.first; i < it.first + it.count; i++) {
      uint32_t p = out.prim[i];
      nb.grow(pb[p]);
      cb.grow(v3(cx[p], cy[p], cz[p]));
    }
    BNode& nd = out.nodes[it.node];
    nd.box = nb;
    if (it.count <= 4) { nd.left = it.first; nd.count = it.count; continue; }
    // pick split
    const int NB = 16;
    float best = std::numeric_limits<float>::infinity();
    int baxis = -1, bsplit = -1;
    for (int ax = 0; ax < 3; ax++) {
      float lo = cb.lo[ax], ext = cb.hi[ax] - cb.lo[ax];
      if (!(ext > 0.0f)) continue;
      Box bb[NB]; uint32_t bc[NB];
      for (int b = 0; b < NB; b++) { bb[b].reset(); bc[b] = 0; }
      float scale = (float)NB / ext;
      for (uint32_t i = it.first; i < it.first + it.count; i++) {
        uint32_t p = out.prim[i];
        int b = std::min(NB - 1, std::max(0, (int)(((*cc[ax])[p] - lo) * scale)));
        bb[b].grow(pb[p]); bc[b]++;
      }
      float ra[NB]; Box acc; acc.reset(); uint32_t cnt = 0;
      for (int b = NB - 1; b > 0; b--) { acc.grow(bb[b]); cnt += bc[b]; ra[b] = cnt ? acc.half_area() * (float)cnt : 0.0f; }
      acc.reset(); cnt = 0;
      for (int b = 0; b < NB - 1; b++) {
        acc.grow(bb[b]); cnt += bc[b];
        if (cnt == 0 || cnt == it.count) continue;
        float cost = acc.half_area() * (float)cnt + ra[b + 1];
        if (cost < best) { best = cost; baxis = ax; bsplit = b; }
      }
    }
    uint32_t mid;
    if (baxis < 0) {
      mid = it.first + it.count / 2;  // all centroids coincide: median split
    } else {
      float lo = cb.lo[baxis], scale = (float)NB / (cb.hi[baxis] - cb.lo[baxis]);
      const std::vector<float>& c = *cc[baxis];
      auto midit = std::partition(out.prim.begin() + it.first, out.prim.begin() + it.first + it.count,
                                  [&](uint32_t p) {
                                    int b = std::min(NB - 1, std::max(0, (int)((c[p] - lo) * scale)));
                                    return b <= bsplit;
                                  });
      mid = (uint32_t)(midit - out.prim.begin());
      if (mid == it.first || mid == it.first + it.count) mid = it.first + it.count / 2;
    }
    uint32_t l = (uint32_t)out.nodes.size();
    out.nodes.push_back(BNode());
    out.nodes.push_back(BNode());
    out.nodes[it.node].left = l;
    out.nodes[it.node].count = 0;
    todo.push_back({l, it.first, mid - it.first});
    todo.push_back({l + 1, mid, it.first + it.count - mid});
  }
}

// Collapses the binary tree into 8-wide nodes: starting from a node's two children, the internal slot with the largest
// surface area is opened (replaced by its two children) until eight slots are filled or only leaves remain.
void bvh_build_wide(Bvh& bvh) {
  bvh.wide.clear();
  if (bvh.nodes.empty()) return;
  struct Item { uint32_t bnode, wnode; };
  std::vector<Item> todo;
  bvh.wide.push_back(WNode());
  todo.push_back({0u, 0u});
  while (!todo.empty()) {
    const Item it = todo.back();
    todo.pop_back();
    uint32_t slots[8];
    int ns = 0;
    const BNode& root = bvh.nodes[it.bnode];
    if (root.count) slots[ns++] = it.bnode;   // a tree that is a single leaf
    else { slots[ns++] = root.left; slots[ns++] = root.left + 1; }
    while (ns < 8) {
      int best = -1;
      float best_area = -1.0f;
      for (int k = 0; k < ns; k++) {
        const BNode& c = bvh.nodes[slots[k]];
        if (c.count == 0 && c.box.half_area() > best_area) { best_area = c.box.half_area(); best = k; }
      }
      if (best < 0) break;
      const uint32_t l = bvh.nodes[slots[best]].left;
      slots[best] = l;
      slots[ns++] = l + 1;
    }
    WNode w;
    for (int k = 0; k < 8; k++) {
      for (int a = 0; a < 3; a++) { w.lo[a][k] = std::numeric_limits<float>::max(); w.hi[a][k] = -std::numeric_limits<float>::max(); }
      w.ref[k] = 0; w.cnt[k] = 0;
    }
    for (int k = 0; k < ns; k++) {
      const BNode& c = bvh.nodes[slots[k]];
      for (int a = 0; a < 3; a++) { w.lo[a][k] = c.box.lo[a]; w.hi[a][k] = c.box.hi[a]; }
      if (c.count) { w.ref[k] = c.left; w.cnt[k] = c.count; }
      else {
        w.cnt[k] = kWideInternal;
        w.ref[k] = (uint32_t)bvh.wide.size();
        bvh.wide.push_back(WNode());
        todo.push_back({slots[k], w.ref[k]});
      }
    }
    bvh.wide[it.wnode] = w;
  }
}

struct TriSoup {
  std::vector<V3> v;  // 3 per triangle
  Bvh bvh;
  void build() {
    std::vector<Box> pb(v.size() / 3);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)pb.size(); t++) {
      pb[t].reset();
      pb[t].grow(v[3 * t]); pb[t].grow(v[3 * t + 1]); pb[t].grow(v[3 * t + 2]);
    }
    bvh_build(pb, bvh);
    if (cpu_has_avx2()) {
      bvh_build_wide(bvh);
      // the wide walk reads a leaf's triangles from one contiguous run (leaf order) instead of through prim[]
      vleaf.resize(v.size());
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < (int64_t)bvh.prim.size(); i++) {
        const uint32_t t = bvh.prim[i];
        vleaf[3 * i] = v[3 * t]; vleaf[3 * i + 1] = v[3 * t + 1]; vleaf[3 * i + 2] = v[3 * t + 2];
      }
    }
  }
  std::vector<V3> vleaf;   // the triangles in leaf order (8-wide walk only)
#if AO_ORACLE_HAVE_AVX2_PATH
  // The 8-lane form of box_hit: the same operations in the same order per lane ((plane - origin) * reciprocal, NaN planes
  // skipped, far distance padded by the same factor), so a lane decides exactly as box_hit does for that box.
  __attribute__((target("avx2"))) bool any_hit_wide(const Ray& r) const {
    const RayShear sh = make_shear(r.d);
    const __m256 o[3] = {_mm256_set1_ps(r.o.x), _mm256_set1_ps(r.o.y), _mm256_set1_ps(r.o.z)};
    const __m256 iv[3] = {_mm256_set1_ps(1.0f / r.d.x), _mm256_set1_ps(1.0f / r.d.y), _mm256_set1_ps(1.0f / r.d.z)};
    const __m256 tmin = _mm256_set1_ps(r.tmin), tmax = _mm256_set1_ps(r.tmax), pad = _mm256_set1_ps(1.0000005f);
    uint32_t stack[256];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const WNode& n = bvh.wide[stack[--sp]];
      __m256 tn = tmin, tf = tmax;
      for (int k = 0; k < 3; k++) {
        const __m256 t0 = _mm256_mul_ps(_mm256_sub_ps(_mm256_load_ps(n.lo[k]), o[k]), iv[k]);
        const __m256 t1 = _mm256_mul_ps(_mm256_sub_ps(_mm256_load_ps(n.hi[k]), o[k]), iv[k]);
        const __m256 nan = _mm256_or_ps(_mm256_cmp_ps(t0, t0, _CMP_UNORD_Q), _mm256_cmp_ps(t1, t1, _CMP_UNORD_Q));
        const __m256 lt = _mm256_cmp_ps(t0, t1, _CMP_LT_OQ);
        const __m256 a = _mm256_blendv_ps(t1, t0, lt), c = _mm256_blendv_ps(t0, t1, lt);
        const __m256 tn2 = _mm256_blendv_ps(tn, a, _mm256_cmp_ps(a, tn, _CMP_GT_OQ));
        const __m256 tf2 = _mm256_blendv_ps(tf, c, _mm256_cmp_ps(c, tf, _CMP_LT_OQ));
        tn = _mm256_blendv_ps(tn2, tn, nan);
        tf = _mm256_blendv_ps(tf2, tf, nan);
      }
      unsigned mask = (unsigned)_mm256_movemask_ps(_mm256_cmp_ps(tn, _mm256_mul_ps(tf, pad), _CMP_LE_OQ));
      if (!mask) continue;
      alignas(32) float tnear[8];
      _mm256_store_ps(tnear, tn);
      // leaves are tested at once (a hit ends the ray); internal children go on the stack with the nearest one on top —
      // an occluded ray usually finds its blocker in the first subtree it enters (the order never changes the answer)
      uint32_t nearest = 0xffffffffu;
      float nearest_t = std::numeric_limits<float>::infinity();
      while (mask) {
        const int k = __builtin_ctz(mask);
        mask &= mask - 1u;
        const uint32_t cnt = n.cnt[k];
        if (cnt == 0u) continue;
        if (cnt == kWideInternal) {
          if (sp >= 255) return any_hit_binary(r);   // unreachable for SAH trees; the binary walk has its own guard
          const char* nx = reinterpret_cast<const char*>(&bvh.wide[n.ref[k]]);   // 4 cache lines; the walk is latency bound on big scenes
          _mm_prefetch(nx, _MM_HINT_T0); _mm_prefetch(nx + 64, _MM_HINT_T0); _mm_prefetch(nx + 128, _MM_HINT_T0); _mm_prefetch(nx + 192, _MM_HINT_T0);
          if (tnear[k] < nearest_t) {
            if (nearest != 0xffffffffu) stack[sp++] = nearest;
            nearest = n.ref[k];
            nearest_t = tnear[k];
          } else {
            stack[sp++] = n.ref[k];
          }
        } else {
          const V3* tv = &vleaf[3ull * n.ref[k]];
          for (uint32_t i = 0; i < cnt; i++)
            if (woop_hit(r, sh, tv[3 * i], tv[3 * i + 1], tv[3 * i + 2], nullptr, nullptr)) return true;
        }
      }
      if (nearest != 0xffffffffu) stack[sp++] = nearest;
    }
    return false;
  }
#endif
  bool any_hit(const Ray& r) const {
    if (bvh.nodes.empty() || v.empty()) return false;
#if AO_ORACLE_HAVE_AVX2_PATH
    if (!bvh.wide.empty() && use_wide_traversal()) return any_hit_wide(r);
#endif
    return any_hit_binary(r);
  }
  bool any_hit_binary(const Ray& r) const {
    if (bvh.nodes.empty() || v.empty()) return false;
    RayShear sh = make_shear(r.d);
    V3 id = v3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const BNode& nd = bvh.nodes[stack[--sp]];
      if (!box_hit(nd.box, r, id)) continue;
      if (nd.count) {
        for (uint32_t i = nd.left; i < nd.left + nd.count; i++) {
          uint32_t t = bvh.prim[i];
          if (woop_hit(r, sh, v[3 * t], v[3 * t + 1], v[3 * t + 2], nullptr, nullptr)) return true;
        }
      } else {
        if (sp + 2 > 128) return false;  // unreachable for SAH trees
        stack[sp++] = nd.left;
        stack[sp++] = nd.left + 1;
      }
    }
    return false;
  }
  bool any_hit_brute(const Ray& r) const {
    RayShear sh = make_shear(r.d);
    for (size_t t = 0; t < v.size() / 3; t++)
      if (woop_hit(r, sh, v[3 * t], v[3 * t + 1], v[3 * t + 2], nullptr, nullptr)) return true;
    return false;
  }
  // smallest "distance to a decision boundary" over all triangles the ray nearly hits.
  float min_margin(const Ray& r) const {
    RayShear sh = make_shear(r.d);
    float best = std::numeric_limits<float>::infinity();
    for (size_t t = 0; t < v.size() / 3; t++) {
      float tt = 0.0f, m = 0.0f;
      V3 p0 = v[3 * t], p1 = v[3 * t + 1], p2 = v[3 * t + 2];
      // edge margin irrespective of the sign test: recompute U,V,W magnitudes
      V3 A = sub(p0, r.o), B = sub(p1, r.o), C = sub(p2, r.o);
      float Ax = comp(A, sh.kx) - sh.Sx * comp(A, sh.kz), Ay = comp(A, sh.ky) - sh.Sy * comp(A, sh.kz);
      float Bx = comp(B, sh.kx) - sh.Sx * comp(B, sh.kz), By = comp(B, sh.ky) - sh.Sy * comp(B, sh.kz);
      float Cx = comp(C, sh.kx) - sh.Sx * comp(C, sh.kz), Cy = comp(C, sh.ky) - sh.Sy * comp(C, sh.kz);
      double U = (double)Cx * By - (double)Cy * Bx, V = (double)Ax * Cy - (double)Ay * Cx,
             W = (double)Bx * Ay - (double)By * Ax;
      double det = U + V + W;
      if (det == 0.0) { best = 0.0f; continue; }
      double u = U / det, vv = V / det, w = W / det;
      double Az = sh.Sz * comp(A, sh.kz), Bz = sh.Sz * comp(B, sh.kz), Cz = sh.Sz * comp(C, sh.kz);
      double tval = (U * Az + V * Bz + W * Cz) / det;
      double edge = std::min(std::fabs(u), std::min(std::fabs(vv), std::fabs(w)));
      double inside = std::min(u, std::min(vv, w));
      double trel = std::min(std::fabs(tval - r.tmin), std::fabs(tval - r.tmax)) /
                    std::max(1e-30, (double)std::fabs(r.tmax));
      (void)tt; (void)m;
      if (inside > -1e-4 && tval > r.tmin - 1e-4 * std::fabs(r.tmax) && tval < r.tmax * 1.0001) {
        best = std::min(best, (float)std::min(edge, trel));
      }
    }
    return best;
  }
};

struct Tracer {
  bool two_level = false;
  // flatten mode
  TriSoup world;
  // two-level mode
  std::vector<TriSoup> blas;         // scene meshes then blocker meshes, object space
  struct Inst { float inv[12]; uint32_t blas; };
  std::vector<Inst> insts;           // scene instances then blocker instances
  Bvh tlas;
  bool any_hit(const Ray& r) const {
    if (!two_level) return world.any_hit(r);
    if (tlas.nodes.empty() || insts.empty()) return false;
    V3 id = v3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    uint32_t stack[128];
    int sp = 0;
    stack[sp++] = 0;
    while (sp) {
      const BNode& nd = tlas.nodes[stack[--sp]];
      if (!box_hit(nd.box, r, id)) continue;
      if (nd.count) {
        for (uint32_t i = nd.left; i < nd.left + nd.count; i++) {
          const Inst& I = insts[tlas.prim[i]];
          Ray lr;
          lr.o = xf_point(I.inv, r.o);
          lr.d = xf_vector(I.inv, r.d);
          lr.tmin = r.tmin; lr.tmax = r.tmax;
          if (blas[I.blas].any_hit(lr)) return true;
        }
      } else {
        stack[sp++] = nd.left;
        stack[sp++] = nd.left + 1;
      }
    }
    return false;
  }
  bool any_hit_brute(const Ray& r) const {
    if (!two_level) return world.any_hit_brute(r);
    for (const Inst& I : insts) {
      Ray lr;
      lr.o = xf_point(I.inv, r.o); lr.d = xf_vector(I.inv, r.d); lr.tmin = r.tmin; lr.tmax = r.tmax;
      if (blas[I.blas].any_hit_brute(lr)) return true;
    }
    return false;
  }
  float min_margin(const Ray& r) const {
    if (!two_level) return world.min_margin(r);
    float best = std::numeric_limits<float>::infinity();
    for (const Inst& I : insts) {
      Ray lr;
      lr.o = xf_point(I.inv, r.o); lr.d = xf_vector(I.inv, r.d); lr.tmin = r.tmin; lr.tmax = r.tmax;
      best = std::min(best, blas[I.blas].min_margin(lr));
    }
    return best;
  }
};

// AUTO instancing rule (decision #12): flatten to one world-space triangle soup when no
// mesh is referenced by more than one instance; otherwise two-level TLAS/BLAS.
bool scene_needs_two_level(const OrScene* scenes[2]) {
  for (int s = 0; s < 2; s++) {
    if (!scenes[s]) continue;
    std::vector<uint32_t> refs(scenes[s]->num_meshes, 0);
    for (uint64_t i = 0; i < scenes[s]->num_instances; i++)
      if (++refs[scenes[s]->instances[i].mesh_index] > 1) return true;
  }
  return false;
}

// ----------------------------------------------------------------------------------
// Sample placement — bake_sample.cpp sample_instance / sample_triangle, SURVEY a7.
// ----------------------------------------------------------------------------------
void sample_instance(const OrScene& sc, uint64_t inst, uint64_t n_samples, uint64_t min_per_tri,
                     OrSamples& out, uint64_t base, int* status) {
  const OrInstance& I = sc.instances[inst];
  const OrMesh& m = sc.meshes[I.mesh_index];
  const uint64_t nT = m.num_triangles;
  std::vector<double> areas;
  instance_tri_areas(sc, inst, areas);
  const double total = blocked_sum(areas.data(), nT);
  std::vector<uint64_t> mins(nT, min_per_tri), counts(nT), offs(nT + 1);
  if (distribute_generic(nT, mins.data(), areas.data(), total, n_samples, counts.data()) != 0) {
    *status = -1;
    return;
  }
  offs[0] = 0;
  for (uint64_t t = 0; t < nT; t++) offs[t + 1] = offs[t] + counts[t];
  if (offs[nT] != n_samples) { *status = -2; return; }
  float inv[12];
  affine_inverse(I.xform, inv);
  const uint32_t seed_inst = (uint32_t)inst;  // decision #10
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t tt = 0; tt < (int64_t)nT; tt++) {
    const uint64_t t = (uint64_t)tt;
    const uint64_t c = counts[t];
    if (!c) continue;
    const uint32_t* idx = m.tri_vertex_indices + 3 * t;
    V3 p0 = load3(vertex_ptr(m, idx[0])), p1 = load3(vertex_ptr(m, idx[1])), p2 = load3(vertex_ptr(m, idx[2]));
    V3 fn = normalize(cross(sub(p1, p0), sub(p2, p0)));
    V3 n0 = fn, n1 = fn, n2 = fn;
    if (m.normals) {
      n0 = load3(normal_ptr(m, idx[0])); n1 = load3(normal_ptr(m, idx[1])); n2 = load3(normal_ptr(m, idx[2]));
      if (dot(n0, fn) < 0.0f) n0 = neg(n0);
      if (dot(n1, fn) < 0.0f) n1 = neg(n1);
      if (dot(n2, fn) < 0.0f) n2 = neg(n2);
    }
    V3 fnw = normalize(xf_normal(inv, fn));
    uint32_t seed = tea(4, seed_inst, (uint32_t)t);
    const float ox = rnd(seed), oy = rnd(seed);
    const float dA = (float)(areas[t] / (double)c);
    for (uint64_t k = 0; k < c; k++) {
      float r1 = ox + halton((uint32_t)(k + 1), 2); r1 = r1 - std::floor(r1);
      float r2 = oy + halton((uint32_t)(k + 1), 3); r2 = r2 - std::floor(r2);
      float s = std::sqrt(r1);
      float b0 = 1.0f - s, b1 = r2 * s, b2 = (1.0f - b0) - b1;
      V3 po = v3((b0 * p0.x + b1 * p1.x) + b2 * p2.x, (b0 * p0.y + b1 * p1.y) + b2 * p2.y,
                 (b0 * p0.z + b1 * p1.z) + b2 * p2.z);
      V3 pw = xf_point(I.xform, po);
      V3 no = v3((b0 * n0.x + b1 * n1.x) + b2 * n2.x, (b0 * n0.y + b1 * n1.y) + b2 * n2.y,
                 (b0 * n0.z + b1 * n1.z) + b2 * n2.z);
      V3 nw = normalize(xf_normal(inv, no));
      uint64_t g = base + offs[t] + k;
      out.sample_positions[3 * g] = pw.x; out.sample_positions[3 * g + 1] = pw.y; out.sample_positions[3 * g + 2] = pw.z;
      out.sample_normals[3 * g] = nw.x; out.sample_normals[3 * g + 1] = nw.y; out.sample_normals[3 * g + 2] = nw.z;
      out.sample_face_normals[3 * g] = fnw.x; out.sample_face_normals[3 * g + 1] = fnw.y; out.sample_face_normals[3 * g + 2] = fnw.z;
      OrSampleInfo& si = out.sample_infos[g];
      si.tri_idx = (uint32_t)t; si.bary[0] = b0; si.bary[1] = b1; si.bary[2] = b2; si.dA = dA;
    }
  }
}

// ----------------------------------------------------------------------------------
// Ray generation — bake_kernels.cu generateRaysKernel / generateRaysHost, SURVEY a9.
// ----------------------------------------------------------------------------------
inline int sqrt_rays(int rays_per_sample) { return (int)(std::sqrt((float)rays_per_sample) + 0.5f); }

// `g` indexes the sample arrays; `gid` is the global sample index that keys the RNG (they differ
// only when a caller passes a subset of a larger sample set).
inline Ray make_ray(const OrSamples& S, uint64_t g, int px, int py, int q, float offset, float maxdist, uint64_t gid = ~0ull) {
  if (gid == ~0ull) gid = g;
  V3 p = load3(S.sample_positions + 3 * g), n = load3(S.sample_normals + 3 * g),
     fn = load3(S.sample_face_normals + 3 * g);
  const uint32_t pass = (uint32_t)(px * q + py);
  uint32_t seed = tea(2, (pass << 16) | pass, (uint32_t)gid);  // decision #10
  // optix::Onb about the shading normal
  V3 b;
  if (std::fabs(n.x) > std::fabs(n.z)) b = v3(-n.y, n.x, 0.0f); else b = v3(0.0f, -n.z, n.y);
  b = normalize(b);
  V3 t = cross(b, n);
  float u0 = ((float)px + rnd(seed)) / (float)q;
  float u1 = ((float)py + rnd(seed)) / (float)q;
  V3 d = v3(0, 0, 0);
  for (int attempt = 0; attempt < 5; attempt++) {  // decision #5
    float r = std::sqrt(u0), c, s;
    sincos2pi(u1, &c, &s);
    float x = r * c, y = r * s;
    float z = std::sqrt(std::fmax(0.0f, (1.0f - x * x) - y * y));
    d = v3((x * t.x + y * b.x) + z * n.x, (x * t.y + y * b.y) + z * n.y, (x * t.z + y * b.z) + z * n.z);
    if (dot(d, fn) > 0.0f) break;
    u0 = rnd(seed);
    u1 = rnd(seed);
  }
  Ray r;
  r.o = v3(p.x + offset * n.x, p.y + offset * n.y, p.z + offset * n.z);
  r.d = d;
  r.tmin = 0.0f;
  r.tmax = maxdist;
  return r;
}

}  // namespace

// ==================================================================================
// C entry points (ctypes)
// ==================================================================================
extern "C" {

int ao_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is told to use all host cores
void ao_oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
uint32_t ao_oracle_tea(uint32_t rounds, uint32_t v0, uint32_t v1) { return tea(rounds, v0, v1); }
uint32_t ao_oracle_lcg(uint32_t* state) { return lcg(*state); }
float ao_oracle_rnd(uint32_t* state) { return rnd(*state); }
float ao_oracle_halton(uint32_t i, uint32_t base) { return halton(i, base); }
void ao_oracle_sincos2pi(float u, float* c, float* s) { sincos2pi(u, c, s); }
void ao_oracle_affine_inverse(const float* m16, float* inv12) { affine_inverse(m16, inv12); }

// per-instance world-space surface area (fixed-shape sum).
int ao_oracle_instance_areas(const OrScene* sc, double* out) {
  std::vector<double> a;
  for (uint64_t i = 0; i < sc->num_instances; i++) {
    instance_tri_areas(*sc, i, a);
    out[i] = blocked_sum(a.data(), a.size());
  }
  return 0;
}

// bake::distributeSamples (bake_sample.cpp) — SURVEY a5.  Returns the total, or 0 on error.
uint64_t ao_oracle_distribute_samples(const OrScene* sc, uint64_t min_per_tri, uint64_t requested,
                                      uint64_t* per_instance) {
  const uint64_t n = sc->num_instances;
  std::vector<double> areas(n);
  std::vector<uint64_t> mins(n);
  ao_oracle_instance_areas(sc, areas.data());
  uint64_t summin = 0;
  for (uint64_t i = 0; i < n; i++) {
    mins[i] = min_per_tri * sc->meshes[sc->instances[i].mesh_index].num_triangles;
    summin += mins[i];
  }
  const uint64_t N = std::max(requested, summin);
  const double total = blocked_sum(areas.data(), n);
  if (distribute_generic(n, mins.data(), areas.data(), total, N, per_instance) != 0) return 0;
  return N;
}

// per-triangle counts of one instance (diagnostic for the bit-exact index tests).
int ao_oracle_triangle_counts(const OrScene* sc, uint64_t inst, uint64_t n_samples, uint64_t min_per_tri,
                              uint64_t* counts) {
  std::vector<double> areas;
  instance_tri_areas(*sc, inst, areas);
  const uint64_t nT = areas.size();
  std::vector<uint64_t> mins(nT, min_per_tri);
  return distribute_generic(nT, mins.data(), areas.data(), blocked_sum(areas.data(), nT), n_samples, counts);
}

// bake::sampleInstances (bake_sample.cpp sample_instances) — SURVEY a7.
int ao_oracle_sample_instances(const OrScene* sc, const uint64_t* per_instance, uint64_t min_per_tri,
                               OrSamples* out) {
  uint64_t base = 0;
  int status = 0;
  for (uint64_t i = 0; i < sc->num_instances; i++) {
    sample_instance(*sc, i, per_instance[i], min_per_tri, *out, base, &status);
    if (status) return status;
    base += per_instance[i];
  }
  return (base == out->num_samples) ? 0 : -3;
}

// make_ground_plane (main.cpp) — SURVEY a14.  upaxis 0..5 = +X,+Y,+Z,-X,-Y,-Z.
// Writes 4 vertices (12 floats) and 2 triangles (6 indices).
void ao_oracle_make_ground_plane(const float* bbox_min, const float* bbox_max, int upaxis,
                                 float scale_factor, float offset_factor, float* verts, uint32_t* tris) {
  const int axis = upaxis % 3;
  const bool flip = upaxis >= 3;
  const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
  float ext = 0.0f;
  for (int k = 0; k < 3; k++) ext = std::max(ext, bbox_max[k] - bbox_min[k]);
  const float h = flip ? bbox_max[axis] + offset_factor * ext : bbox_min[axis] - offset_factor * ext;
  const float c1 = 0.5f * (bbox_min[a1] + bbox_max[a1]), c2 = 0.5f * (bbox_min[a2] + bbox_max[a2]);
  const float h1 = 0.5f * scale_factor * (bbox_max[a1] - bbox_min[a1]);
  const float h2 = 0.5f * scale_factor * (bbox_max[a2] - bbox_min[a2]);
  const float s1[4] = {-1, 1, 1, -1}, s2[4] = {-1, -1, 1, 1};
  for (int i = 0; i < 4; i++) {
    verts[3 * i + axis] = h;
    verts[3 * i + a1] = c1 + s1[i] * h1;
    verts[3 * i + a2] = c2 + s2[i] * h2;
  }
  const uint32_t up[6] = {0, 1, 2, 0, 2, 3}, dn[6] = {0, 2, 1, 0, 3, 2};
  for (int i = 0; i < 6; i++) tris[i] = flip ? dn[i] : up[i];
}

// ---- tracer ----------------------------------------------------------------------
// mode: 0 = auto (decision #12), 1 = flatten, 2 = two-level.
void* ao_oracle_tracer_create(const OrScene* scene, const OrScene* blockers, int mode) {
  Tracer* T = new Tracer();
  const OrScene* scenes[2] = {scene, (blockers && blockers->num_instances) ? blockers : nullptr};
  T->two_level = (mode == 2) || (mode == 0 && scene_needs_two_level(scenes));
  if (!T->two_level) {
    for (int s = 0; s < 2; s++) {
      if (!scenes[s]) continue;
      for (uint64_t i = 0; i < scenes[s]->num_instances; i++) {
        const OrInstance& I = scenes[s]->instances[i];
        const OrMesh& m = scenes[s]->meshes[I.mesh_index];
        size_t base = T->world.v.size();
        T->world.v.resize(base + 3 * m.num_triangles);
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < (int64_t)m.num_triangles; t++) tri_world(m, I.xform, (uint64_t)t, &T->world.v[base + 3 * t]);
      }
    }
    T->world.build();
  } else {
    uint32_t mesh_base = 0;
    std::vector<Box> ib;
    for (int s = 0; s < 2; s++) {
      if (!scenes[s]) continue;
      for (uint64_t mi = 0; mi < scenes[s]->num_meshes; mi++) {
        const OrMesh& m = scenes[s]->meshes[mi];
        T->blas.emplace_back();
        TriSoup& ts = T->blas.back();
        ts.v.resize(3 * m.num_triangles);
        for (uint64_t t = 0; t < m.num_triangles; t++)
          for (int k = 0; k < 3; k++) ts.v[3 * t + k] = load3(vertex_ptr(m, m.tri_vertex_indices[3 * t + k]));
        ts.build();
      }
      for (uint64_t i = 0; i < scenes[s]->num_instances; i++) {
        const OrInstance& I = scenes[s]->instances[i];
        Tracer::Inst in;
        affine_inverse(I.xform, in.inv);
        in.blas = mesh_base + I.mesh_index;
        T->insts.push_back(in);
        // world box of the instance: transform the 8 corners of the BLAS root box, pad.
        Box wb; wb.reset();
        const TriSoup& ts = T->blas[in.blas];
        if (!ts.bvh.nodes.empty() && !ts.v.empty()) {
          const Box& rb = ts.bvh.nodes[0].box;
          for (int c = 0; c < 8; c++) {
            V3 p = v3((c & 1) ? rb.hi[0] : rb.lo[0], (c & 2) ? rb.hi[1] : rb.lo[1], (c & 4) ? rb.hi[2] : rb.lo[2]);
            wb.grow(xf_point(I.xform, p));
          }
          for (int k = 0; k < 3; k++) {
            float pad = 3.8e-6f * std::max(std::fabs(wb.lo[k]), std::fabs(wb.hi[k]));
            wb.lo[k] -= pad; wb.hi[k] += pad;
          }
        } else {
          for (int k = 0; k < 3; k++) { wb.lo[k] = 0; wb.hi[k] = 0; }
        }
        ib.push_back(wb);
      }
      mesh_base += (uint32_t)scenes[s]->num_meshes;
    }
    bvh_build(ib, T->tlas);
  }
  return T;
}
void ao_oracle_tracer_destroy(void* h) { delete static_cast<Tracer*>(h); }
// Traversal of the triangle BVHs: 0 = auto, 1 = binary scalar, 2 = 8-wide AVX2 (see g_traversal_mode).  Returns the kind in
// effect for tracers built on this CPU: 1 = binary scalar, 2 = 8-wide AVX2.
int ao_oracle_set_traversal(int mode) {
  g_traversal_mode = (mode == 1 || mode == 2) ? mode : 0;
  return use_wide_traversal() ? 2 : 1;
}
int ao_oracle_tracer_is_two_level(void* h) { return static_cast<Tracer*>(h)->two_level ? 1 : 0; }

// rays: n x 8 floats (o.xyz, tmin, d.xyz, tmax).  hit: n bytes (1 = occluded).
// which: 0 = BVH, 1 = brute force.
int ao_oracle_trace_rays(void* h, const float* rays, uint64_t n, uint8_t* hit, int which) {
  const Tracer* T = static_cast<Tracer*>(h);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    const float* p = rays + 8 * i;
    Ray r;
    r.o = v3(p[0], p[1], p[2]); r.tmin = p[3]; r.d = v3(p[4], p[5], p[6]); r.tmax = p[7];
    hit[i] = (which == 1 ? T->any_hit_brute(r) : T->any_hit(r)) ? 1 : 0;
  }
  return 0;
}
// margin[i] = distance (barycentric units / relative t) from ray i to the nearest hit/miss
// decision boundary, by brute force.  Used to check that CUDA/oracle disagreements are edge cases.
int ao_oracle_ray_margin(void* h, const float* rays, uint64_t n, float* margin) {
  const Tracer* T = static_cast<Tracer*>(h);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = 0; i < (int64_t)n; i++) {
    const float* p = rays + 8 * i;
    Ray r;
    r.o = v3(p[0], p[1], p[2]); r.tmin = p[3]; r.d = v3(p[4], p[5], p[6]); r.tmax = p[7];
    margin[i] = T->min_margin(r);
  }
  return 0;
}

int ao_oracle_sqrt_rays(int rays_per_sample) { return sqrt_rays(rays_per_sample); }

// rays_out: (end-begin) x q*q x 8 floats, stratum-major inside a sample (px*q+py).
int ao_oracle_generate_rays(const OrSamples* S, uint64_t begin, uint64_t end, int rays_per_sample,
                            float offset, float maxdist, float* rays_out) {
  const int q = sqrt_rays(rays_per_sample);
#pragma omp parallel for schedule(static)
  for (int64_t g = (int64_t)begin; g < (int64_t)end; g++) {
    for (int px = 0; px < q; px++)
      for (int py = 0; py < q; py++) {
        Ray r = make_ray(*S, (uint64_t)g, px, py, q, offset, maxdist);
        float* o = rays_out + (((uint64_t)g - begin) * (uint64_t)(q * q) + (uint64_t)(px * q + py)) * 8;
        o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.tmin;
        o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.tmax;
      }
  }
  return 0;
}

// rays of local sample `k` generated with the RNG streams of global sample index `gid`
int ao_oracle_generate_rays_for(const OrSamples* S, uint64_t k, uint64_t gid, int rays_per_sample, float offset,
                                float maxdist, float* rays_out) {
  const int q = sqrt_rays(rays_per_sample);
  for (int px = 0; px < q; px++)
    for (int py = 0; py < q; py++) {
      Ray r = make_ray(*S, k, px, py, q, offset, maxdist, gid);
      float* o = rays_out + (uint64_t)(px * q + py) * 8;
      o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.tmin;
      o[4] = r.d.x; o[5] = r.d.y; o[6] = r.d.z; o[7] = r.tmax;
    }
  return 0;
}

// bake::computeAO (bake_ao_optix_prime.cpp ao_optix_prime + bake_kernels.cu), SURVEY a9–a13:
// ao[g-begin] = 1 - hits/q^2 (decision #2).  hit_counts optional.
int ao_oracle_compute_ao(void* h, const OrSamples* S, uint64_t begin, uint64_t end, int rays_per_sample,
                         float offset, float maxdist, float* ao, uint32_t* hit_counts) {
  const Tracer* T = static_cast<Tracer*>(h);
  const int q = sqrt_rays(rays_per_sample);
  const float denom = (float)(q * q);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t g = (int64_t)begin; g < (int64_t)end; g++) {
    uint32_t hits = 0;
    for (int px = 0; px < q; px++)
      for (int py = 0; py < q; py++) {
        Ray r = make_ray(*S, (uint64_t)g, px, py, q, offset, maxdist);
        hits += T->any_hit(r) ? 1u : 0u;
      }
    if (hit_counts) hit_counts[g - begin] = hits;
    ao[g - begin] = 1.0f - (float)hits / denom;
  }
  return 0;
}

// ---- vertex maps -------------------------------------------------------------------
// bake_filter.cpp filter/filter_mesh — SURVEY a15.  vertex_ao[i] has mesh.num_vertices floats.
int ao_oracle_filter_area(const OrScene* sc, const uint64_t* per_instance, const OrSamples* S,
                          const float* ao, float** vertex_ao) {
  uint64_t base = 0;
  for (uint64_t i = 0; i < sc->num_instances; i++) {
    const OrMesh& m = sc->meshes[sc->instances[i].mesh_index];
    std::vector<double> num(m.num_vertices, 0.0), wgt(m.num_vertices, 0.0);
    for (uint64_t k = 0; k < per_instance[i]; k++) {
      const OrSampleInfo& si = S->sample_infos[base + k];
      const uint32_t* idx = m.tri_vertex_indices + 3 * (uint64_t)si.tri_idx;
      const double val = (double)ao[base + k] * (double)si.dA;
      for (int c = 0; c < 3; c++) {
        num[idx[c]] += (double)si.bary[c] * val;
        wgt[idx[c]] += (double)si.bary[c] * (double)si.dA;
      }
    }
    for (uint64_t v = 0; v < m.num_vertices; v++) vertex_ao[i][v] = wgt[v] > 0.0 ? (float)(num[v] / wgt[v]) : 0.0f;
    base += per_instance[i];
  }
  return 0;
}

// bake_filter_least_squares.cpp — SURVEY a16, decisions #6/#7: solve (M + w R) x = b in
// fp64 per instance.  M = sum_samples dA b b^T, b = sum dA ao b.  R = sum over interior edges of the
// energy of the gradient jump of the piecewise-linear interpolant across the edge (Kavan, Bargteil,
// Sloan 2011, "Least Squares Vertex Baking"):
//   E_edge = (A1 + A2) * |grad(T1) - grad(T2)|^2      (all three components of the 3-D gradients)
// With (i,j) the edge ends, p/q the opposite vertices, h the altitude of the opposite vertex, s its foot
// parameter along the edge and m1, m2 the unit in-plane normals of the edge pointing at p and q:
//   grad(T1) - grad(T2) = a1 m1 - a2 m2,   a1 = (x_p - s1 x_j - (1-s1) x_i) / h1,   a2 likewise with q
// (the components along the edge cancel), so  |.|^2 = a1^2 + a2^2 - 2 c a1 a2  with c = m1 . m2
// (c = -1 for a flat pair: the energy is then (a1 + a2)^2, the squared jump of the co-normal derivative).
// energy = 1 selects the round-1 form instead: (A1 + A2)^2 (a1 + a2)^2 — the unfolded jump with the area
// squared, which makes w scale-free; kept as an option (AoBakeParams::ls_energy = 1 on the GPU side).
// Geometry uses world-space vertices.  Vertices with zero lumped mass get M_vv = 1, rhs 0 (decision #7).
// Solver: Jacobi-preconditioned CG from x = 0 to |r|/|b| <= tol.  Returns iterations used (>= 0) or < 0.
struct LsEdge { uint32_t i, j, p, q; double s1, h1, s2, h2, c, W; };
// y += w * dE/dx for one edge; returns nothing.  a1 = al . x on (i,j,p), a2 = be . x on (i,j,q).
static inline void ls_edge_coeffs(const LsEdge& E, double al[3], double be[3]) {
  al[0] = -(1.0 - E.s1) / E.h1; al[1] = -E.s1 / E.h1; al[2] = 1.0 / E.h1;
  be[0] = -(1.0 - E.s2) / E.h2; be[1] = -E.s2 / E.h2; be[2] = 1.0 / E.h2;
}

static void ls_build_edges(const OrMesh& m, const float* xf, int energy, std::vector<LsEdge>& edges) {
  struct Half { uint64_t key; uint32_t tri; uint32_t opp; };
  std::vector<Half> hs;
  hs.reserve(3 * m.num_triangles);
  for (uint64_t t = 0; t < m.num_triangles; t++) {
    const uint32_t* idx = m.tri_vertex_indices + 3 * t;
    for (int e = 0; e < 3; e++) {
      uint32_t a = idx[e], b = idx[(e + 1) % 3], o = idx[(e + 2) % 3];
      if (a == b) continue;
      uint64_t key = ((uint64_t)std::min(a, b) << 32) | std::max(a, b);
      hs.push_back({key, (uint32_t)t, o});
    }
  }
  std::sort(hs.begin(), hs.end(), [](const Half& x, const Half& y) { return x.key != y.key ? x.key < y.key : x.tri < y.tri; });
  for (size_t k = 0; k < hs.size();) {
    size_t e = k;
    while (e < hs.size() && hs[e].key == hs[k].key) e++;
    if (e - k >= 2) {  // interior (non-manifold: first two by triangle index)
      LsEdge E;
      E.i = (uint32_t)(hs[k].key >> 32); E.j = (uint32_t)(hs[k].key & 0xffffffffu);
      E.p = hs[k].opp; E.q = hs[k + 1].opp;
      V3 pi = xf_point(xf, load3(vertex_ptr(m, E.i))), pj = xf_point(xf, load3(vertex_ptr(m, E.j)));
      V3 pp = xf_point(xf, load3(vertex_ptr(m, E.p))), pq = xf_point(xf, load3(vertex_ptr(m, E.q)));
      double ex = (double)pj.x - pi.x, ey = (double)pj.y - pi.y, ez = (double)pj.z - pi.z;
      double L2 = ex * ex + ey * ey + ez * ez;
      double r1[3], r2[3];
      auto foot = [&](V3 o, double* s, double* h, double* area, double* r) {
        double ox = (double)o.x - pi.x, oy = (double)o.y - pi.y, oz = (double)o.z - pi.z;
        *s = (ox * ex + oy * ey + oz * ez) / L2;
        r[0] = ox - *s * ex; r[1] = oy - *s * ey; r[2] = oz - *s * ez;
        *h = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
        *area = 0.5 * std::sqrt(L2) * *h;
      };
      double A1 = 0, A2 = 0;
      bool ok = L2 > 0.0;
      if (ok) { foot(pp, &E.s1, &E.h1, &A1, r1); foot(pq, &E.s2, &E.h2, &A2, r2); ok = E.h1 > 0.0 && E.h2 > 0.0; }
      if (ok) {
        if (energy == 1) { E.c = -1.0; E.W = (A1 + A2) * (A1 + A2); }
        else { E.c = (r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2]) / (E.h1 * E.h2); E.W = A1 + A2; }
        edges.push_back(E);
      }
    }
    k = e;
  }
}

// check_only: vertex_ao holds a candidate solution; nothing is solved, *relres receives
// max over instances of |b - (M + wR) x| / |b| under THIS file's operator (full-size parity checks, where
// a CPU solve of the same system would take minutes).
static int ls_impl(const OrScene* sc, const uint64_t* per_instance, const OrSamples* S, const float* ao, float weight, double tol,
                   int max_iter, float** vertex_ao, int energy, bool check_only, double* relres) {
  uint64_t base = 0;
  int total_iters = 0;
  if (relres) *relres = 0.0;
  for (uint64_t inst = 0; inst < sc->num_instances; inst++) {
    const OrInstance& I = sc->instances[inst];
    const OrMesh& m = sc->meshes[I.mesh_index];
    const uint64_t nV = m.num_vertices, nT = m.num_triangles;
    // per-triangle sampled mass blocks (6 unique entries) and rhs
    std::vector<double> Mt(6 * nT, 0.0), b(nV, 0.0), diag(nV, 0.0);
    for (uint64_t k = 0; k < per_instance[inst]; k++) {
      const OrSampleInfo& si = S->sample_infos[base + k];
      const uint32_t* idx = m.tri_vertex_indices + 3 * (uint64_t)si.tri_idx;
      const double dA = si.dA, a = ao[base + k];
      const double b0 = si.bary[0], b1 = si.bary[1], b2 = si.bary[2];
      double* M = &Mt[6 * (uint64_t)si.tri_idx];
      M[0] += dA * b0 * b0; M[1] += dA * b0 * b1; M[2] += dA * b0 * b2;
      M[3] += dA * b1 * b1; M[4] += dA * b1 * b2; M[5] += dA * b2 * b2;
      b[idx[0]] += dA * a * b0; b[idx[1]] += dA * a * b1; b[idx[2]] += dA * a * b2;
    }
    std::vector<LsEdge> edges;
    if (weight != 0.0f) ls_build_edges(m, I.xform, energy, edges);
    const double w = weight;
    auto apply = [&](const std::vector<double>& x, std::vector<double>& y) {
      std::fill(y.begin(), y.end(), 0.0);
      for (uint64_t t = 0; t < nT; t++) {
        const uint32_t* idx = m.tri_vertex_indices + 3 * t;
        const double* M = &Mt[6 * t];
        double x0 = x[idx[0]], x1 = x[idx[1]], x2 = x[idx[2]];
        y[idx[0]] += M[0] * x0 + M[1] * x1 + M[2] * x2;
        y[idx[1]] += M[1] * x0 + M[3] * x1 + M[4] * x2;
        y[idx[2]] += M[2] * x0 + M[4] * x1 + M[5] * x2;
      }
      for (const LsEdge& E : edges) {
        double al[3], be[3];
        ls_edge_coeffs(E, al, be);
        const double a1 = al[0] * x[E.i] + al[1] * x[E.j] + al[2] * x[E.p];
        const double a2 = be[0] * x[E.i] + be[1] * x[E.j] + be[2] * x[E.q];
        const double g1 = w * E.W * (a1 - E.c * a2), g2 = w * E.W * (a2 - E.c * a1);   // R x = half the gradient of x^T R x = W (a1^2 + a2^2 - 2 c a1 a2)
        y[E.i] += g1 * al[0] + g2 * be[0]; y[E.j] += g1 * al[1] + g2 * be[1];
        y[E.p] += g1 * al[2]; y[E.q] += g2 * be[2];
      }
    };
    for (uint64_t t = 0; t < nT; t++) {
      const uint32_t* idx = m.tri_vertex_indices + 3 * t;
      diag[idx[0]] += Mt[6 * t + 0]; diag[idx[1]] += Mt[6 * t + 3]; diag[idx[2]] += Mt[6 * t + 5];
    }
    // decision #7: a vertex with zero lumped mass (no sample on any incident triangle) gets
    // M_vv = 1, b_v = 0 — it is anchored at 0 like the averaging filter leaves it, and the
    // system stays well conditioned when whole regions are unsampled (< 1 sample/triangle).
    std::vector<uint8_t> fixed(nV, 0);
    for (uint64_t v = 0; v < nV; v++)
      if (!(diag[v] > 0.0)) { fixed[v] = 1; diag[v] = 1.0; b[v] = 0.0; }
    for (const LsEdge& E : edges) {
      double al[3], be[3];
      ls_edge_coeffs(E, al, be);
      diag[E.i] += w * E.W * (al[0] * al[0] + be[0] * be[0] - 2.0 * E.c * al[0] * be[0]);
      diag[E.j] += w * E.W * (al[1] * al[1] + be[1] * be[1] - 2.0 * E.c * al[1] * be[1]);
      diag[E.p] += w * E.W * al[2] * al[2];
      diag[E.q] += w * E.W * be[2] * be[2];
    }
    std::vector<double> x(nV, 0.0), r(b), z(nV), p(nV), Ap(nV);
    double bnorm = 0.0;
    for (uint64_t v = 0; v < nV; v++) bnorm += b[v] * b[v];
    bnorm = std::sqrt(bnorm);
    if (check_only) {
      for (uint64_t v = 0; v < nV; v++) x[v] = vertex_ao[inst][v];
      apply(x, Ap);
      double rn = 0.0;
      for (uint64_t v = 0; v < nV; v++) {
        const double rv = b[v] - (Ap[v] + (fixed[v] ? x[v] : 0.0));
        rn += rv * rv;
      }
      if (relres && bnorm > 0.0) *relres = std::max(*relres, std::sqrt(rn) / bnorm);
      base += per_instance[inst];
      continue;
    }
    int it = 0;
    if (bnorm > 0.0) {
      double rz = 0.0;
      for (uint64_t v = 0; v < nV; v++) { z[v] = r[v] / diag[v]; p[v] = z[v]; rz += r[v] * z[v]; }
      for (; it < max_iter; it++) {
        double rn = 0.0;
        for (uint64_t v = 0; v < nV; v++) rn += r[v] * r[v];
        if (std::sqrt(rn) <= tol * bnorm) break;
        apply(p, Ap);
        for (uint64_t v = 0; v < nV; v++) if (fixed[v]) Ap[v] += p[v];
        double pAp = 0.0;
        for (uint64_t v = 0; v < nV; v++) pAp += p[v] * Ap[v];
        if (!(pAp > 0.0)) break;
        double alpha = rz / pAp;
        double rz2 = 0.0;
        for (uint64_t v = 0; v < nV; v++) {
          x[v] += alpha * p[v]; r[v] -= alpha * Ap[v]; z[v] = r[v] / diag[v]; rz2 += r[v] * z[v];
        }
        double beta = rz2 / rz;
        rz = rz2;
        for (uint64_t v = 0; v < nV; v++) p[v] = z[v] + beta * p[v];
      }
    }
    for (uint64_t v = 0; v < nV; v++) vertex_ao[inst][v] = (float)x[v];
    total_iters += it;
    base += per_instance[inst];
  }
  return total_iters;
}

int ao_oracle_filter_least_squares(const OrScene* sc, const uint64_t* per_instance, const OrSamples* S,
                                   const float* ao, float weight, double tol, int max_iter,
                                   float** vertex_ao, int energy) {
  return ls_impl(sc, per_instance, S, ao, weight, tol, max_iter, vertex_ao, energy, false, nullptr);
}

// Relative residual of a candidate least-squares solution under the oracle's operator (see ls_impl).
double ao_oracle_ls_residual(const OrScene* sc, const uint64_t* per_instance, const OrSamples* S, const float* ao, float weight,
                             float** vertex_x, int energy) {
  double res = 0.0;
  ls_impl(sc, per_instance, S, ao, weight, 0.0, 0, vertex_x, energy, true, &res);
  return res;
}

}  // extern "C"
