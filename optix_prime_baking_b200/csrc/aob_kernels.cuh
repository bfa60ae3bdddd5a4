// aob_kernels.cuh — every __global__ kernel of libaobake.so (sm_100a).
//   build   : world triangles + boxes, centroid bounds, Morton keys, LBVH hierarchy/refit,
//             8-wide collapse, leaf-order gather
//   sample  : triangle areas, fixed-shape area sums, per-triangle counts, placement
//   trace   : explicit-ray any-hit, ray dump, fused raygen+traverse+accumulate (two variants)
//   filter  : area-weighted vertex map, least-squares (matrix-free PCG) pieces
#pragma once
#include <cuda_runtime.h>
#include "aob_bvh.cuh"

namespace aob {

struct Xf12 {
  float m[12];
};

// ---------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// largest i in [0, n) with a[i] <= v  (a non-decreasing, a[0] <= v)
template <typename T>
__device__ __forceinline__ uint64_t upper_slot(const T* __restrict__ a, uint64_t n, T v) {
  uint64_t lo = 0, hi = n;  // invariant: a[lo] <= v, (hi == n or a[hi] > v)
  while (hi - lo > 1) {
    uint64_t mid = (lo + hi) >> 1;
    if (a[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------
// BVH build
// ---------------------------------------------------------------------------------------
// One thread per triangle: (optionally) transform to world space with the exact fp32 formula
// (aob::xf_point), write the 3 x float4 soup record and the primitive box.
__global__ void k_make_tris(const float* __restrict__ verts, const uint32_t* __restrict__ tris, uint32_t nT,
                            Xf12 xf, int identity, uint32_t out_offset, F4* __restrict__ soup,
                            F4* __restrict__ plo, F4* __restrict__ phi) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nT) return;
  V3 w[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t vi = tris[3ull * t + k];
    V3 p = v3(verts[3ull * vi], verts[3ull * vi + 1], verts[3ull * vi + 2]);
    w[k] = identity ? p : xf_point(xf.m, p);
  }
  const uint32_t o = out_offset + t;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    F4 f; f.x = w[k].x; f.y = w[k].y; f.z = w[k].z; f.w = __uint_as_float(o);
    soup[3ull * o + k] = f;
  }
  F4 lo, hi;
  lo.x = fminf(w[0].x, fminf(w[1].x, w[2].x)); lo.y = fminf(w[0].y, fminf(w[1].y, w[2].y)); lo.z = fminf(w[0].z, fminf(w[1].z, w[2].z)); lo.w = 0.f;
  hi.x = fmaxf(w[0].x, fmaxf(w[1].x, w[2].x)); hi.y = fmaxf(w[0].y, fmaxf(w[1].y, w[2].y)); hi.z = fmaxf(w[0].z, fmaxf(w[1].z, w[2].z)); hi.w = 0.f;
  plo[o] = lo;
  phi[o] = hi;
}

__global__ void k_init_bounds(int* b6) {
  if (threadIdx.x < 3) b6[threadIdx.x] = 0x7fffffff;
  else if (threadIdx.x < 6) b6[threadIdx.x] = (int)0x80000000;
}
// centroid bounds + full box bounds (b6 = centroid min/max, f6 = box min/max), ordered-int atomics
__global__ void k_bounds(const F4* __restrict__ plo, const F4* __restrict__ phi, uint32_t n, int* b6, int* f6) {
  float cmin[3] = {3.0e38f, 3.0e38f, 3.0e38f}, cmax[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  float fmin_[3] = {3.0e38f, 3.0e38f, 3.0e38f}, fmax_[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const F4 lo = plo[i], hi = phi[i];
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    const float l[3] = {lo.x, lo.y, lo.z}, h[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      cmin[k] = fminf(cmin[k], c[k]); cmax[k] = fmaxf(cmax[k], c[k]);
      fmin_[k] = fminf(fmin_[k], l[k]); fmax_[k] = fmaxf(fmax_[k], h[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    for (int o = 16; o > 0; o >>= 1) {
      cmin[k] = fminf(cmin[k], __shfl_xor_sync(0xffffffffu, cmin[k], o));
      cmax[k] = fmaxf(cmax[k], __shfl_xor_sync(0xffffffffu, cmax[k], o));
      fmin_[k] = fminf(fmin_[k], __shfl_xor_sync(0xffffffffu, fmin_[k], o));
      fmax_[k] = fmaxf(fmax_[k], __shfl_xor_sync(0xffffffffu, fmax_[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(&b6[k], float_to_ordered(cmin[k]));
      atomicMax(&b6[3 + k], float_to_ordered(cmax[k]));
      atomicMin(&f6[k], float_to_ordered(fmin_[k]));
      atomicMax(&f6[3 + k], float_to_ordered(fmax_[k]));
    }
  }
}
// world-space box of one instance: every vertex of its mesh through the instance transform
// (tighter than transforming the 8 corners of the object-space box, which for a rotated mesh can
// inflate the TLAS leaf several-fold).  b6 must be initialised with k_init_bounds.
__global__ void k_instance_bounds(const float* __restrict__ verts, uint32_t nV, Xf12 xf, int* b6) {
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nV; i += gridDim.x * blockDim.x) {
    const V3 w = xf_point(xf.m, v3(verts[3ull * i], verts[3ull * i + 1], verts[3ull * i + 2]));
    lo[0] = fminf(lo[0], w.x); lo[1] = fminf(lo[1], w.y); lo[2] = fminf(lo[2], w.z);
    hi[0] = fmaxf(hi[0], w.x); hi[1] = fmaxf(hi[1], w.y); hi[2] = fmaxf(hi[2], w.z);
  }
#pragma unroll
  for (int k = 0; k < 3; k++)
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(&b6[k], float_to_ordered(lo[k]));
      atomicMax(&b6[3 + k], float_to_ordered(hi[k]));
    }
  }
}
// max squared distance of an instance's transformed vertices from `centre` (non-negative floats
// order like their bit patterns, so atomicMax on the bits works)
__global__ void k_instance_radius2(const float* __restrict__ verts, uint32_t nV, Xf12 xf, float cx, float cy, float cz, uint32_t* r2bits) {
  float m = 0.0f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nV; i += gridDim.x * blockDim.x) {
    const V3 w = xf_point(xf.m, v3(verts[3ull * i], verts[3ull * i + 1], verts[3ull * i + 2]));
    const float dx = w.x - cx, dy = w.y - cy, dz = w.z - cz;
    m = fmaxf(m, dx * dx + dy * dy + dz * dz);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(r2bits, __float_as_uint(m));
}
__global__ void k_decode_bounds_n(const int* f6, float* out6, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out6[i] = ordered_to_float(f6[i]);
}
__global__ void k_init_bounds_n(int* b6, uint32_t n_boxes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * n_boxes) b6[i] = (i % 6 < 3) ? 0x7fffffff : (int)0x80000000;
}
__global__ void k_decode_bounds(const int* f6, float* out6) {
  if (threadIdx.x < 6) out6[threadIdx.x] = ordered_to_float(f6[threadIdx.x]);
}
__global__ void k_morton(const F4* __restrict__ plo, const F4* __restrict__ phi, uint32_t n, const int* __restrict__ b6,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 cmin = v3(ordered_to_float(b6[0]), ordered_to_float(b6[1]), ordered_to_float(b6[2]));
  const V3 cmax = v3(ordered_to_float(b6[3]), ordered_to_float(b6[4]), ordered_to_float(b6[5]));
  // one scale for all axes (cubical Morton cells): a thin axis — the height of a terrain — only
  // gets split once the cells have shrunk to its extent, instead of wasting the top-level
  // splits that per-axis normalisation would spend on it
  const float ext = fmaxf(cmax.x - cmin.x, fmaxf(cmax.y - cmin.y, cmax.z - cmin.z));
  const float inv = ext > 0.0f ? 1.0f / ext : 0.0f;
  const V3 cinv = v3(inv, inv, inv);
  keys[i] = morton63(plo[i], phi[i], cmin, cinv);
  vals[i] = i;
}
// ---- oversized primitives (a ground-plane quad under a fine mesh, a huge instance in a TLAS) ----------
// A primitive whose box spans a large part of the scene poisons the top of an LBVH: every node that
// contains it has a scene-sized box, the 8-bit grids of those nodes are hundreds of leaf-sizes coarse, and
// the children that hold the real geometry are so inflated by the quantisation that every ray visits all of
// them (config 5: 15.1 node visits per ray instead of 6.9 for the same mesh without its ground plane).
// Such primitives are kept out of the tree: flagged here, sorted behind everything else (key = ~0), and
// attached as leaf slots of one extra root node whose only internal child is the root of the tree over the
// rest — one coarse node on top, precise grids everywhere below.
__global__ void k_flag_big(const F4* __restrict__ plo, const F4* __restrict__ phi, uint32_t n, const int* __restrict__ f6,
                           uint8_t* __restrict__ big, uint32_t* __restrict__ count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sx = ordered_to_float(f6[3]) - ordered_to_float(f6[0]), sy = ordered_to_float(f6[4]) - ordered_to_float(f6[1]),
              sz = ordered_to_float(f6[5]) - ordered_to_float(f6[2]);
  const float ext = fmaxf(sx, fmaxf(sy, sz));
  const bool b = box_is_oversized(plo[i], phi[i], ext);
  big[i] = b ? 1 : 0;
  if (b) atomicAdd(count, 1u);
}
// centroid bounds (b6) and box bounds (f6) of the primitives that are NOT flagged
__global__ void k_bounds_small(const F4* __restrict__ plo, const F4* __restrict__ phi, const uint8_t* __restrict__ big, uint32_t n, int* b6, int* f6) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  float c[3] = {3.0e38f, 3.0e38f, 3.0e38f}, C[3] = {-3.0e38f, -3.0e38f, -3.0e38f}, l[3] = {3.0e38f, 3.0e38f, 3.0e38f}, h[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  if (i < n && !big[i]) {
    const F4 lo = plo[i], hi = phi[i];
    c[0] = C[0] = 0.5f * (lo.x + hi.x); c[1] = C[1] = 0.5f * (lo.y + hi.y); c[2] = C[2] = 0.5f * (lo.z + hi.z);
    l[0] = lo.x; l[1] = lo.y; l[2] = lo.z; h[0] = hi.x; h[1] = hi.y; h[2] = hi.z;
  }
#pragma unroll
  for (int k = 0; k < 3; k++)
    for (int o = 16; o > 0; o >>= 1) {
      c[k] = fminf(c[k], __shfl_xor_sync(0xffffffffu, c[k], o)); C[k] = fmaxf(C[k], __shfl_xor_sync(0xffffffffu, C[k], o));
      l[k] = fminf(l[k], __shfl_xor_sync(0xffffffffu, l[k], o)); h[k] = fmaxf(h[k], __shfl_xor_sync(0xffffffffu, h[k], o));
    }
  if ((threadIdx.x & 31) == 0 && c[0] < 3.0e38f) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(&b6[k], float_to_ordered(c[k])); atomicMax(&b6[3 + k], float_to_ordered(C[k]));
      atomicMin(&f6[k], float_to_ordered(l[k])); atomicMax(&f6[3 + k], float_to_ordered(h[k]));
    }
  }
}
__global__ void k_morton_small(const F4* __restrict__ plo, const F4* __restrict__ phi, const uint8_t* __restrict__ big, uint32_t n,
                               const int* __restrict__ b6, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const V3 cmin = v3(ordered_to_float(b6[0]), ordered_to_float(b6[1]), ordered_to_float(b6[2]));
  const V3 cmax = v3(ordered_to_float(b6[3]), ordered_to_float(b6[4]), ordered_to_float(b6[5]));
  const float ext = fmaxf(cmax.x - cmin.x, fmaxf(cmax.y - cmin.y, cmax.z - cmin.z));
  const float inv = ext > 0.0f ? 1.0f / ext : 0.0f;
  keys[i] = big[i] ? ~0ull : morton63(plo[i], phi[i], cmin, v3(inv, inv, inv));   // 63-bit codes: ~0 sorts behind all of them
  vals[i] = i;
}
// The extra root: slot 0 = the tree over the ordinary primitives (node `main_root`, box f6_small), the
// following slots = the oversized primitives (sorted positions [n_small, n), `per_slot` per slot), which
// also complete the leaf order.  One thread.
__global__ void k_super_root(Node8* __restrict__ nodes, uint32_t node_index, uint32_t main_root, const int* __restrict__ f6_all,
                             const int* __restrict__ f6_small, const F4* __restrict__ plo, const F4* __restrict__ phi,
                             const uint32_t* __restrict__ sorted_prims, uint32_t n_small, uint32_t n_big, uint32_t per_slot,
                             uint32_t prim_offset, uint32_t* __restrict__ leaf_prims) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  F4 nlo, nhi, slo, shi;
  nlo.x = ordered_to_float(f6_all[0]); nlo.y = ordered_to_float(f6_all[1]); nlo.z = ordered_to_float(f6_all[2]); nlo.w = 0.f;
  nhi.x = ordered_to_float(f6_all[3]); nhi.y = ordered_to_float(f6_all[4]); nhi.z = ordered_to_float(f6_all[5]); nhi.w = 0.f;
  slo.x = ordered_to_float(f6_small[0]); slo.y = ordered_to_float(f6_small[1]); slo.z = ordered_to_float(f6_small[2]); slo.w = 0.f;
  shi.x = ordered_to_float(f6_small[3]); shi.y = ordered_to_float(f6_small[4]); shi.z = ordered_to_float(f6_small[5]); shi.w = 0.f;
  super_root_body(nodes + node_index, main_root, nlo, nhi, slo, shi, plo, phi, sorted_prims, n_small, n_big, per_slot, prim_offset, leaf_prims);
}
__global__ void k_hierarchy(Lbvh L) { lbvh_hierarchy_body(blockIdx.x * blockDim.x + threadIdx.x, L); }
__global__ void k_refit(Lbvh L) { lbvh_refit_body(blockIdx.x * blockDim.x + threadIdx.x, L); }
__global__ void k_collapse(CollapseArgs A, uint32_t lb, uint32_t le) {
  const uint32_t w = lb + blockIdx.x * blockDim.x + threadIdx.x;
  if (w < le) collapse_body(w, A);
}
__global__ void k_set_u32(uint32_t* p, uint32_t v) { *p = v; }
// Input validation for set_scene: flags a mesh whose index buffer points past its vertex array, so
// that a malformed scene is an error instead of an out-of-bounds read in k_make_tris.
__global__ void k_check_indices(const uint32_t* __restrict__ tris, uint64_t n_indices, uint32_t nV, uint32_t* __restrict__ flag) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_indices && tris[i] >= nV) *flag = 1u;
}
__global__ void k_gather_tris(const F4* __restrict__ soup, const uint32_t* __restrict__ leaf_prims, uint32_t n,
                              uint32_t soup_offset, F4* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t s = 3ull * (soup_offset + leaf_prims[i]);
  out[3ull * i] = soup[s];
  out[3ull * i + 1] = soup[s + 1];
  out[3ull * i + 2] = soup[s + 2];
}
// empty-tree root: no children hit, ever
__global__ void k_empty_node(Node8* nd) {
  Node8 n;
  memset(&n, 0, sizeof(n));
  for (int k = 0; k < 8; k++) for (int a = 0; a < 3; a++) n.q[a][k][0] = 255;
  *nd = n;
}

// ---------------------------------------------------------------------------------------
// Sampling (bake_sample.cpp)
// ---------------------------------------------------------------------------------------
struct InstDesc {
  float xf[12];
  float inv[12];
  const float* verts;     // packed xyz
  const float* normals;   // packed xyz or null
  const uint32_t* tris;
  uint64_t tri_begin;     // offset of this instance's triangles in the flattened element space
  uint64_t num_tris;
  uint64_t block_begin;   // offset of its 1024-element blocks
  uint64_t sample_begin;  // offset of its samples
  uint64_t num_samples;   // N_i
  uint32_t seed;          // instance seed (= instance index, decision #10)
  uint32_t pad;
};

__device__ __forceinline__ uint32_t find_instance(const InstDesc* __restrict__ inst, uint32_t n_inst, uint64_t e) {
  uint32_t lo = 0, hi = n_inst;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (inst[mid].tri_begin <= e) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ void load_tri_obj(const InstDesc& I, uint64_t t, V3* p, uint32_t* idx) {
  idx[0] = I.tris[3 * t]; idx[1] = I.tris[3 * t + 1]; idx[2] = I.tris[3 * t + 2];
#pragma unroll
  for (int k = 0; k < 3; k++) p[k] = v3(I.verts[3ull * idx[k]], I.verts[3ull * idx[k] + 1], I.verts[3ull * idx[k] + 2]);
}

__global__ void k_tri_areas(const InstDesc* __restrict__ inst, uint32_t n_inst, uint64_t total_e, double* __restrict__ area) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_e) return;
  const InstDesc& I = inst[find_instance(inst, n_inst, e)];
  V3 p[3];
  uint32_t idx[3];
  load_tri_obj(I, e - I.tri_begin, p, idx);
  area[e] = tri_area(xf_point(I.xf, p[0]), xf_point(I.xf, p[1]), xf_point(I.xf, p[2]));
}
// fixed-shape sum, level 1: one thread per 1024-element block, sequential (decision #3)
__global__ void k_block_sums(const InstDesc* __restrict__ inst, uint32_t n_inst, uint64_t total_blocks,
                             const double* __restrict__ area, double* __restrict__ bsum) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= total_blocks) return;
  uint32_t lo = 0, hi = n_inst;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (inst[mid].block_begin <= b) lo = mid; else hi = mid;
  }
  const InstDesc& I = inst[lo];
  const uint64_t j = b - I.block_begin;
  const uint64_t s = j * 1024, e = min(I.num_tris, s + 1024);
  double acc = 0.0;
  for (uint64_t i = s; i < e; i++) acc = ex::dadd(acc, area[I.tri_begin + i]);
  bsum[b] = acc;
}
// level 2: one thread per instance, sequential over its block sums
__global__ void k_inst_totals(const InstDesc* __restrict__ inst, uint32_t n_inst, const double* __restrict__ bsum,
                              double* __restrict__ total) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inst) return;
  const uint64_t nb = (inst[i].num_tris + 1023) / 1024;
  double acc = 0.0;
  for (uint64_t b = 0; b < nb; b++) acc = ex::dadd(acc, bsum[inst[i].block_begin + b]);
  total[i] = acc;
}
__global__ void k_tri_counts(const InstDesc* __restrict__ inst, uint32_t n_inst, uint64_t total_e, const double* __restrict__ area,
                             const double* __restrict__ total, uint64_t min_per_tri, uint64_t* __restrict__ counts) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_e) return;
  const uint32_t ii = find_instance(inst, n_inst, e);
  const InstDesc& I = inst[ii];
  const uint64_t Na = I.num_samples - min_per_tri * I.num_tris;
  const double T = total[ii];
  uint64_t c = min_per_tri;
  if (Na > 0 && T > 0.0) c += (uint64_t)ex::ddiv(ex::dmul((double)Na, area[e]), T);
  counts[e] = c;
}
// per instance: leftover L_i = N_i - assigned_i; status[0] set to 1 if negative.
__global__ void k_inst_leftover(const InstDesc* __restrict__ inst, uint32_t n_inst, const uint64_t* __restrict__ offs,
                                const uint64_t* __restrict__ counts, long long* __restrict__ leftover, int* status) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_inst) return;
  const InstDesc& I = inst[i];
  if (I.num_tris == 0) { leftover[i] = 0; if (I.num_samples) atomicExch(status, 2); return; }
  const uint64_t e1 = I.tri_begin + I.num_tris - 1;
  const uint64_t assigned = offs[e1] + counts[e1] - offs[I.tri_begin];
  const long long L = (long long)I.num_samples - (long long)assigned;
  leftover[i] = L;
  if (L < 0) atomicExch(status, 1);
}
// final per-triangle sample offsets: the leftover sweep adds +1 to triangles 0,1,2,... (wrapping)
__global__ void k_final_offsets(const InstDesc* __restrict__ inst, uint32_t n_inst, uint64_t total_e, const uint64_t* __restrict__ offs,
                                const uint64_t* __restrict__ counts, const long long* __restrict__ leftover,
                                uint64_t* __restrict__ final_off, uint32_t* __restrict__ final_cnt, uint64_t total_samples) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == total_e) final_off[e] = total_samples;
  if (e >= total_e) return;
  const uint32_t ii = find_instance(inst, n_inst, e);
  const InstDesc& I = inst[ii];
  const uint64_t t = e - I.tri_begin;
  const uint64_t L = (uint64_t)max(0ll, leftover[ii]);
  const uint64_t per = L / I.num_tris, rem = L % I.num_tris;
  final_off[e] = I.sample_begin + (offs[e] - offs[I.tri_begin]) + t * per + min(t, rem);
  final_cnt[e] = (uint32_t)(counts[e] + per + (t < rem ? 1 : 0));
}
// one thread per sample (bake_sample.cpp sample_triangle)
__global__ void k_place_samples(const InstDesc* __restrict__ inst, uint32_t n_inst, uint64_t total_e,
                                const uint64_t* __restrict__ final_off, const uint32_t* __restrict__ final_cnt,
                                const double* __restrict__ area, uint64_t total_samples, float* __restrict__ pos,
                                float* __restrict__ nrm, float* __restrict__ fnrm, AoSampleInfo* __restrict__ info) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_samples) return;
  const uint64_t e = upper_slot<uint64_t>(final_off, total_e, g);
  const InstDesc& I = inst[find_instance(inst, n_inst, e)];
  const uint64_t t = e - I.tri_begin;
  const uint32_t k = (uint32_t)(g - final_off[e]);
  const uint32_t c = final_cnt[e];
  V3 p[3];
  uint32_t idx[3];
  load_tri_obj(I, t, p, idx);
  const V3 fn = normalize(cross(sub(p[1], p[0]), sub(p[2], p[0])));
  V3 n0 = fn, n1 = fn, n2 = fn;
  if (I.normals) {
    n0 = v3(I.normals[3ull * idx[0]], I.normals[3ull * idx[0] + 1], I.normals[3ull * idx[0] + 2]);
    n1 = v3(I.normals[3ull * idx[1]], I.normals[3ull * idx[1] + 1], I.normals[3ull * idx[1] + 2]);
    n2 = v3(I.normals[3ull * idx[2]], I.normals[3ull * idx[2] + 1], I.normals[3ull * idx[2] + 2]);
    if (dot(n0, fn) < 0.0f) n0 = neg(n0);
    if (dot(n1, fn) < 0.0f) n1 = neg(n1);
    if (dot(n2, fn) < 0.0f) n2 = neg(n2);
  }
  const V3 fnw = normalize(xf_normal(I.inv, fn));
  uint32_t seed = tea<4>(I.seed, (uint32_t)t);
  const float ox = rnd(seed), oy = rnd(seed);
  float r1 = ex::add(ox, halton(k + 1, 2)); r1 = ex::sub(r1, floorf(r1));
  float r2 = ex::add(oy, halton(k + 1, 3)); r2 = ex::sub(r2, floorf(r2));
  const float s = ex::sqrt(r1);
  const float b0 = ex::sub(1.0f, s), b1 = ex::mul(r2, s), b2 = ex::sub(ex::sub(1.0f, b0), b1);
  const V3 po = v3(ex::add(ex::add(ex::mul(b0, p[0].x), ex::mul(b1, p[1].x)), ex::mul(b2, p[2].x)),
                   ex::add(ex::add(ex::mul(b0, p[0].y), ex::mul(b1, p[1].y)), ex::mul(b2, p[2].y)),
                   ex::add(ex::add(ex::mul(b0, p[0].z), ex::mul(b1, p[1].z)), ex::mul(b2, p[2].z)));
  const V3 pw = xf_point(I.xf, po);
  const V3 no = v3(ex::add(ex::add(ex::mul(b0, n0.x), ex::mul(b1, n1.x)), ex::mul(b2, n2.x)),
                   ex::add(ex::add(ex::mul(b0, n0.y), ex::mul(b1, n1.y)), ex::mul(b2, n2.y)),
                   ex::add(ex::add(ex::mul(b0, n0.z), ex::mul(b1, n1.z)), ex::mul(b2, n2.z)));
  const V3 nw = normalize(xf_normal(I.inv, no));
  pos[3 * g] = pw.x; pos[3 * g + 1] = pw.y; pos[3 * g + 2] = pw.z;
  nrm[3 * g] = nw.x; nrm[3 * g + 1] = nw.y; nrm[3 * g + 2] = nw.z;
  fnrm[3 * g] = fnw.x; fnrm[3 * g + 1] = fnw.y; fnrm[3 * g + 2] = fnw.z;
  AoSampleInfo si;
  si.tri_idx = (uint32_t)t;
  si.bary[0] = b0; si.bary[1] = b1; si.bary[2] = b2;
  si.dA = (float)ex::ddiv(area[e], (double)c);
  info[g] = si;
}

// ---------------------------------------------------------------------------------------
// Trace
// ---------------------------------------------------------------------------------------
// H2 = false (AoBakeParams::node_test = 1) compiles the fp32 node test only.
template <bool H2>
__global__ void k_trace_rays(BvhView bvh, const float* __restrict__ rays, uint64_t n, uint8_t* __restrict__ hit) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 a = reinterpret_cast<const float4*>(rays)[2 * i], b = reinterpret_cast<const float4*>(rays)[2 * i + 1];
  U2 stack[kStackSize];
  hit[i] = trace_any_hit<false, H2>(bvh, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), a.w, b.w, stack, nullptr) ? 1 : 0;
}

struct SampleView {
  const float* pos;
  const float* nrm;
  const float* fnrm;
};

__global__ void k_dump_rays(SampleView S, uint64_t begin, uint64_t end, int q, float offset, float maxdist, float* __restrict__ out) {
  const uint64_t q2 = (uint64_t)q * q;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (end - begin) * q2) return;
  const uint64_t g = begin + i / q2;
  const uint32_t pass = (uint32_t)(i % q2);
  const V3 p = v3(S.pos[3 * g], S.pos[3 * g + 1], S.pos[3 * g + 2]), n = v3(S.nrm[3 * g], S.nrm[3 * g + 1], S.nrm[3 * g + 2]),
           fn = v3(S.fnrm[3 * g], S.fnrm[3 * g + 1], S.fnrm[3 * g + 2]);
  const Onb onb = make_onb(n);
  const V3 o = ao_ray_origin(p, n, offset);
  const V3 d = ao_ray_dir((uint32_t)g, pass, q, n, fn, onb);
  float4* r = reinterpret_cast<float4*>(out) + 2 * i;
  r[0] = make_float4(o.x, o.y, o.z, 0.0f);
  r[1] = make_float4(d.x, d.y, d.z, maxdist);
}

// The packed-fp16 node test assumes unit-length world rays, i.e. unit-length shading normals
// (cosine_dir builds the direction from n and an orthonormal pair derived from it).  Samples made by
// k_place_samples are normalised; caller-provided ones (aobake_set_samples) are checked here and a
// sample set with any other normal is traced with the fp32 test.
__global__ void k_check_unit_normals(const float* __restrict__ nrm, uint64_t begin, uint64_t end, uint32_t* __restrict__ flag) {
  const uint64_t g = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= end) return;
  const float x = nrm[3 * g], y = nrm[3 * g + 1], z = nrm[3 * g + 2];
  const float d2 = x * x + y * y + z * z;
  if (!(fabsf(d2 - 1.0f) <= 1.0e-4f)) *flag = 1u;
}

// Variant 1 (simple): one warp per (32-sample block, strata chunk); lane = sample; each lane
// walks its strata sequentially with a private local-memory stack.
template <bool STATS, bool H2>
__global__ void __launch_bounds__(256) k_ao_simple(BvhView bvh, SampleView S, uint64_t begin, uint64_t end, int q, float offset,
                                                   float maxdist, uint32_t n_chunks, uint32_t* __restrict__ hits,
                                                   unsigned long long* __restrict__ stats) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t n_blocks = (end - begin + 31) / 32;
  if (warp >= n_blocks * n_chunks) return;
  const uint64_t sblock = warp / n_chunks;
  const uint32_t chunk = (uint32_t)(warp % n_chunks);
  const uint64_t g = begin + sblock * 32 + lane;
  if (g >= end) return;
  const uint32_t q2 = (uint32_t)(q * q);
  const uint32_t p0 = (uint32_t)(((uint64_t)chunk * q2) / n_chunks), p1 = (uint32_t)(((uint64_t)(chunk + 1) * q2) / n_chunks);
  const V3 p = v3(S.pos[3 * g], S.pos[3 * g + 1], S.pos[3 * g + 2]), n = v3(S.nrm[3 * g], S.nrm[3 * g + 1], S.nrm[3 * g + 2]),
           fn = v3(S.fnrm[3 * g], S.fnrm[3 * g + 1], S.fnrm[3 * g + 2]);
  const Onb onb = make_onb(n);
  const V3 o = ao_ray_origin(p, n, offset);
  U2 stack[kStackSize];
  TraceCounters cnt = {0, 0, 0};
  uint32_t h = 0;
  for (uint32_t pass = p0; pass < p1; pass++) {
    const V3 d = ao_ray_dir((uint32_t)g, pass, q, n, fn, onb);
    h += trace_any_hit<STATS, H2>(bvh, o, d, 0.0f, maxdist, stack, &cnt) ? 1u : 0u;
  }
  if (n_chunks > 1) atomicAdd(&hits[g - begin], h);
  else hits[g - begin] = h;
  if (STATS) {
    atomicAdd(&stats[0], (unsigned long long)cnt.nodes);
    atomicAdd(&stats[1], (unsigned long long)cnt.tris);
    atomicAdd(&stats[2], (unsigned long long)cnt.insts);
  }
}

// Variant 0 (default): persistent warps with per-lane ray refill.
//   * grid = resident CTAs only (multiple of the SM count); every lane owns one work item
//     (sample, strata chunk) at a time, fetched with a warp-aggregated atomicAdd on a global
//     counter, and walks its strata one ray at a time;
//   * lanes whose ray terminates go idle; when fewer than `refill_below` lanes of the warp are
//     still traversing, the idle lanes generate their next rays together (so ray generation
//     runs converged) and traversal resumes — the warp-level compaction/refill of north_star;
//   * the traversal stack is a per-thread local-memory array (L1 resident); a shared-memory head
//     ([depth][thread] columns, AOB_SM_STACK entries) is available but measured slower;
//   * the Woop shear constants are computed lazily, only by lanes that reach a triangle.
constexpr int kAoBlock = 128;
#ifndef AOB_LOOKAHEAD
#define AOB_LOOKAHEAD 1
#endif
constexpr int kLookahead = AOB_LOOKAHEAD;   // rays a lane keeps queued (generated converged at a refill)
// Entries of the traversal stack kept in shared memory ([depth][thread] columns); the rest lives in
// local memory.  Measured on B200 (profiles/r1/sweep_stack_placement.log): 0 — the whole stack in
// L1-resident local memory — is fastest (config 2: 12.87 vs 12.24 Grays/s with 6, config 3: 7.13 vs
// 6.88): AO rays push and pop rarely, the shared variant pays index arithmetic and a branch per
// access, and every KB of shared memory is a KB less L1 for BVH nodes.
#ifndef AOB_SM_STACK
#define AOB_SM_STACK 0
#endif
constexpr int kSmStack = AOB_SM_STACK;

// Rays the packed-fp16 node test cannot take (a direction component below 2^-12 in the space being
// traversed: ~5 in 10^4) are not traced by the fused kernel: it appends (sample, stratum) to `list`
// and k_ao_deferred traces them afterwards with the fp32 test.  count > capacity means entries were
// dropped; the host then repeats the launch with the fp32 kernels.
struct DeferredRays {
  U2* list;            // (sample index relative to `begin`, stratum)
  uint32_t* count;
  uint32_t capacity;
};
AOB_D void defer_ray(const DeferredRays& D, uint32_t rel, uint32_t pass) {
  const uint32_t i = atomicAdd(D.count, 1u);
  if (i < D.capacity) { U2 e; e.x = rel; e.y = pass; D.list[i] = e; }
}

template <bool STATS, bool TWO_LEVEL, bool CLAMP_TMAX, bool H2, bool PACKET>
__global__ void __launch_bounds__(kAoBlock, 7) k_ao_persistent(BvhView bvh, SampleView S, uint64_t begin, uint32_t n, int q, float offset,
                                                             float maxdist, uint32_t n_chunks, uint32_t refill_below, uint32_t tri_batch,
                                                             uint32_t part, uint32_t num_parts, uint32_t sb_blocks, uint32_t n_local_blocks,
                                                             uint32_t* __restrict__ hits, unsigned long long* __restrict__ counter,
                                                             unsigned long long* __restrict__ stats, DeferredRays deferred) {
#if AOB_SM_STACK > 0
  __shared__ U2 s_stack[kSmStack][kAoBlock];
#endif
  U2 l_stack[kStackSize - kSmStack];
  // The fp16 test is used for flattened scenes only: under a TLAS it measured slower than fp32 (config 4:
  // 3.98 vs 4.23 Grays/s), and object-space rays of scaled instances are not unit length.
  static_assert(!(H2 && TWO_LEVEL), "the packed-fp16 node test is instantiated for flattened scenes only");
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t q2 = (uint32_t)(q * q);
  auto push = [&](int& sp, U2 v) {
#if AOB_SM_STACK > 0
    if (sp < kSmStack) s_stack[sp][threadIdx.x] = v;
    else l_stack[sp - kSmStack] = v;
#else
    l_stack[sp] = v;
#endif
    sp++;
  };
  auto pop = [&](int& sp) -> U2 {
    sp--;
#if AOB_SM_STACK > 0
    return sp < kSmStack ? s_stack[sp][threadIdx.x] : l_stack[sp - kSmStack];
#else
    return l_stack[sp];
#endif
  };

  const NodeConsts nc = make_node_consts();
  // per-lane item state
  __shared__ float s_la[kLookahead][6][kAoBlock];  // queued (lookahead) rays per thread: direction + slab reciprocals
  uint32_t la_count = 0, la_head = 0;              // rays queued; slot of the oldest
  // PACKET (stratum-major ray order): the warp, not the lane, owns the work item (32 consecutive samples x a strata
  // chunk) and deals its rays out in the order (stratum, sample): ray k of the item is stratum k / ns of sample k % ns,
  // and whichever lane has room takes the next one.  No lane is tied to a sample any more, so there is no per-lane
  // item state to set up, no sample whose expensive rays hold one lane back while the others run ahead, and ray
  // generation runs with nearly the whole warp (27.8 lanes against 19.5); the 32 rays in flight start on neighbouring
  // samples and come from one or two strata.  Measured +7 % / +13 % / +9 % on configs 2 / 3 / 1 (profiles/r2/
  // sweep_ray_order.log); it is the DEALING that pays — turning the strata so that one stratum is one world direction for
  // the whole warp, or making the whole GPU shoot into one sector at a time, added nothing.  Hits are
  // counted per (item buffer, sample) in shared memory (16-bit halves: q^2 <= 65025) and written out when the buffer is
  // retired; two items can be in flight (rays of the previous item still traversing while the next is dealt out).
  static_assert(!PACKET || kLookahead == 1, "stratum-major order keeps one queued ray per lane");
  __shared__ uint32_t s_pk_hits[PACKET ? kAoBlock / 32 : 1][2][16];
  uint32_t pk_rel0_a = 0, pk_rel0_b = 0, pk_ns_a = 0, pk_ns_b = 0;   // warp-uniform: first sample / samples of the item in buffer 0 / 1 (ns == 0: free)
  uint32_t pk_buf = 0, pk_next = 0, pk_end = 0, pk_pass0 = 0;   // warp-uniform: buffer being dealt out, next ray, rays, first stratum
  uint32_t cur_tag = 0, la_tag = 0;   // (buffer << 5) | sample slot of the lane's active / queued ray
  V3 la_org = v3(0, 0, 0);            // origin of the queued ray (the active one's is `org`)
  if (PACKET) {
    if (lane < 16u) { s_pk_hits[threadIdx.x >> 5][0][lane] = 0u; s_pk_hits[threadIdx.x >> 5][1][lane] = 0u; }
    __syncwarp();
  }
  // retires item buffer b: its per-sample hit counts go to hits[] (every ray of the item has ended: the caller checked)
  auto pk_flush = [&](uint32_t b) {
    __syncwarp();
    const uint32_t ns = b ? pk_ns_b : pk_ns_a, rel0 = b ? pk_rel0_b : pk_rel0_a;
    if (ns != 0u) {
      uint32_t* cnt = s_pk_hits[threadIdx.x >> 5][b];
      if (lane < ns) {
        const uint32_t v = (cnt[lane >> 1] >> (16u * (lane & 1u))) & 0xffffu;
        if (n_chunks > 1) atomicAdd(&hits[rel0 + lane], v);
        else hits[rel0 + lane] = v;
      }
      __syncwarp();
      if (lane < 16u) cnt[lane] = 0u;
      __syncwarp();
      if (b) pk_ns_b = 0u; else pk_ns_a = 0u;
    }
  };
  bool have_item = false, ray_active = false, exhausted = false;
  uint32_t rel = 0, pass = 0, pass_end = 0, nh = 0;
  uint32_t supply_next = 0, supply_left = 0, supply_chunk = 0;  // warp-uniform
  V3 org = v3(0, 0, 0);   // the normals are re-read per generated ray (L1): 6 registers matter more
  Onb onb;
  onb.t = onb.b = v3(0, 0, 0);
  // per-lane ray state
  RayState r;
  r.tmin = 0.0f; r.tmax = maxdist;
  r.org = org; r.dir = org; r.idir = org; r.wide = false;
  V3 wdir = v3(0, 0, 0);  // world-space direction (two-level: restored after a BLAS)
  bool in_blas = !TWO_LEVEL;
  U2 G, T;   // node group being walked; primitive group of the last node step (kept while the lane is paused)
  G.x = 0; G.y = 0;
  T.x = 0; T.y = 0;
  bool paused = false;   // the lane holds leaf hits (T) and waits for the warp's next triangle block
  uint32_t waited = 0;   // warp-uniform: iterations since the first of the currently paused lanes paused
  const uint32_t tri_wait = tri_batch >> 8;   // (packed by the host: low byte = lanes, next byte = iterations)
  tri_batch &= 0xffu;
  int sp = 0;
  uint32_t c_nodes = 0, c_tris = 0, c_insts = 0;
  // Multi-GPU interleave: this launch owns the super-blocks sb (of sb_blocks 32-sample blocks) with
  // sb % num_parts == part; local block k maps to global block ((k / sb_blocks) * num_parts + part)
  // * sb_blocks + k % sb_blocks.  num_parts == 1 is the identity.
  const unsigned long long n_blocks = n_local_blocks;
  const unsigned long long n_global_blocks = ((unsigned long long)n + 31ull) / 32ull;
  const unsigned long long total_items = n_blocks * n_chunks;  // item = (block of 32 samples, strata chunk)
  auto start_queued = [&]() {
    const float(*la)[kAoBlock] = s_la[la_head];
    wdir = v3(la[0][threadIdx.x], la[1][threadIdx.x], la[2][threadIdx.x]);
    la_head = la_head + 1 == kLookahead ? 0u : la_head + 1;
    la_count--;
    if (PACKET) { org = la_org; cur_tag = la_tag; }
    r.org = org; r.dir = wdir;
    r.idir = v3(la[3][threadIdx.x], la[4][threadIdx.x], la[5][threadIdx.x]);
    in_blas = !TWO_LEVEL;
    G.x = bvh.root;
    G.y = (1u << 24) | 1u;
    sp = 0;
    ray_active = true;
  };

  while (true) {
    // ------------------------------ refill ------------------------------
    if (PACKET) {
      // every lane with an empty queue takes the next ray of the warp's item, in (stratum, sample) order
      bool want = la_count == 0u;
      while (true) {
        const uint32_t want_mask = __ballot_sync(0xffffffffu, want);
        if (want_mask == 0u) break;
        if (pk_next == pk_end) {
          if (exhausted) break;
          // the next item goes into the other buffer; rays of the item before the current one may still be in flight
          // there (only if one of them outlived a whole item: wait for it, correctness before the last percent)
          const uint32_t nb = pk_buf ^ 1u;
          const bool mine = (ray_active && (cur_tag >> 5) == nb) || (la_count != 0u && (la_tag >> 5) == nb);
          if (__any_sync(0xffffffffu, mine)) break;
          pk_flush(nb);
          unsigned long long w = 0;
          if (lane == 0) w = atomicAdd(counter, 1ull);
          w = __shfl_sync(0xffffffffu, w, 0);
          if (w >= total_items) { exhausted = true; break; }
          const uint32_t chunk = (uint32_t)(w / n_blocks);  // chunk-major: concurrent warps work on neighbouring blocks
          unsigned long long blk = w - (unsigned long long)chunk * n_blocks;
          if (num_parts > 1) blk = ((blk / sb_blocks) * num_parts + part) * sb_blocks + blk % sb_blocks;
          if (blk >= n_global_blocks) continue;  // padding block of a partial last super-block
          const uint32_t ns = (uint32_t)min(32ull, (unsigned long long)n - blk * 32ull);
          pk_pass0 = (uint32_t)(((uint64_t)chunk * q2) / n_chunks);
          const uint32_t pe = (uint32_t)(((uint64_t)(chunk + 1) * q2) / n_chunks);
          if (nb) { pk_rel0_b = (uint32_t)(blk * 32ull); pk_ns_b = ns; }
          else { pk_rel0_a = (uint32_t)(blk * 32ull); pk_ns_a = ns; }
          pk_buf = nb;
          pk_next = 0u;
          pk_end = ns * (pe - pk_pass0);
          if (pk_end == 0u) continue;
        }
        const uint32_t my = (uint32_t)__popc(want_mask & lt_mask);
        const uint32_t take = min((uint32_t)__popc(want_mask), pk_end - pk_next);
        if (want && my < take) {
          const uint32_t k = pk_next + my;
          const uint32_t ns = pk_buf ? pk_ns_b : pk_ns_a;
          const uint32_t st = ns == 32u ? k >> 5 : k / ns;
          const uint32_t sl = k - st * ns;
          const uint32_t rl = (pk_buf ? pk_rel0_b : pk_rel0_a) + sl, ps = pk_pass0 + st;
          const uint64_t gf = 3ull * (begin + rl);
          const V3 p = v3(__ldg(S.pos + gf), __ldg(S.pos + gf + 1), __ldg(S.pos + gf + 2));
          const V3 nrm = v3(__ldg(S.nrm + gf), __ldg(S.nrm + gf + 1), __ldg(S.nrm + gf + 2));
          const V3 fnrm = v3(__ldg(S.fnrm + gf), __ldg(S.fnrm + gf + 1), __ldg(S.fnrm + gf + 2));
          const Onb ob = make_onb(nrm);
          const V3 d = ao_ray_dir((uint32_t)(begin + rl), ps, q, nrm, fnrm, ob);
          const V3 id = v3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
          if (H2 && !(fmaxf(fmaxf(fabsf(id.x), fabsf(id.y)), fabsf(id.z)) <= kH2MaxIdir)) {
            defer_ray(deferred, rl, ps);
          } else {
            float(*la)[kAoBlock] = s_la[0];
            la[0][threadIdx.x] = d.x; la[1][threadIdx.x] = d.y; la[2][threadIdx.x] = d.z;
            la[3][threadIdx.x] = id.x; la[4][threadIdx.x] = id.y; la[5][threadIdx.x] = id.z;
            la_org = ao_ray_origin(p, nrm, offset);
            la_tag = (pk_buf << 5) | sl;
            la_head = 0u;
            la_count = 1u;
          }
          want = false;   // (a lane whose ray was deferred takes its next one at the next refill)
        }
        pk_next += take;
      }
      if (!ray_active && la_count != 0u) start_queued();
      if (!__any_sync(0xffffffffu, ray_active)) {
        // nothing in flight: either rays are left to deal out (deferred ones used up this round's), or this is the end
        if (exhausted && pk_next == pk_end) break;
        continue;
      }
    } else {
    if (!ray_active && la_count == 0u && have_item && pass == pass_end) {
      if (n_chunks > 1) atomicAdd(&hits[rel], nh);
      else hits[rel] = nh;
      have_item = false;
    }
    // Lanes that finished their item take the next sample of the warp's *supply block*: a block
    // of 32 consecutive samples (x one strata chunk) fetched with one atomicAdd by the warp.
    // Consecutive samples sit on the same few triangles (samples are emitted in triangle
    // order), so whatever the lanes' progress the warp's rays stay spatially coherent, and no
    // lane waits for the slowest sample of a block.
    bool need = !have_item && !exhausted;
    while (true) {
      const uint32_t need_mask = __ballot_sync(0xffffffffu, need);
      if (need_mask == 0u) break;
      if (supply_left == 0u) {
        unsigned long long w = 0;
        if (lane == 0) w = atomicAdd(counter, 1ull);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total_items) { exhausted = true; break; }
        supply_chunk = (uint32_t)(w / n_blocks);  // chunk-major: concurrent warps work on neighbouring blocks
        unsigned long long blk = w - (unsigned long long)supply_chunk * n_blocks;
        if (num_parts > 1) blk = ((blk / sb_blocks) * num_parts + part) * sb_blocks + blk % sb_blocks;
        if (blk >= n_global_blocks) continue;  // padding block of a partial last super-block
        supply_next = (uint32_t)(blk * 32ull);
        supply_left = (uint32_t)min(32ull, (unsigned long long)n - blk * 32ull);
      }
      const uint32_t my = (uint32_t)__popc(need_mask & lt_mask);
      const uint32_t take = min((uint32_t)__popc(need_mask), supply_left);
      if (need && my < take) {
        rel = supply_next + my;
        pass = (uint32_t)(((uint64_t)supply_chunk * q2) / n_chunks);
        pass_end = (uint32_t)(((uint64_t)(supply_chunk + 1) * q2) / n_chunks);
        const uint64_t g = begin + rel;
        const V3 p = v3(S.pos[3 * g], S.pos[3 * g + 1], S.pos[3 * g + 2]);
        const V3 nrm = v3(S.nrm[3 * g], S.nrm[3 * g + 1], S.nrm[3 * g + 2]);
        onb = make_onb(nrm);
        org = ao_ray_origin(p, nrm, offset);
        nh = 0;
        have_item = true;
        need = false;
      }
      supply_next += take;
      supply_left -= take;
    }
    // Lookahead: at a refill *every* lane whose queue has room generates rays for it (idle lanes and lanes still
    // traversing alike), so ray generation runs with most of the warp converged instead of only the few idle lanes; a
    // lane whose ray ends inside the traversal loop starts its oldest queued ray at once, without a refill.  A queue of
    // kLookahead = 2 halves the number of refills: a lane goes idle only after using up both, by which time most of
    // the warp has room for at least one.
    for (int rep = 0; rep < kLookahead; rep++) {
      const bool gen = have_item && la_count < (uint32_t)kLookahead && pass < pass_end;
      if (kLookahead > 1 && !__any_sync(0xffffffffu, gen)) break;
      if (gen) {
        const uint64_t gf = 3ull * (begin + rel);
        const V3 fnrm = v3(__ldg(S.fnrm + gf), __ldg(S.fnrm + gf + 1), __ldg(S.fnrm + gf + 2));
        const V3 nrm = v3(__ldg(S.nrm + gf), __ldg(S.nrm + gf + 1), __ldg(S.nrm + gf + 2));
        const V3 d = ao_ray_dir((uint32_t)(begin + rel), pass, q, nrm, fnrm, onb);
        const V3 id = v3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
        if (H2 && !(fmaxf(fmaxf(fabsf(id.x), fabsf(id.y)), fabsf(id.z)) <= kH2MaxIdir)) {
          defer_ray(deferred, rel, pass);   // world rays are unit length (the host checks the normals): only the reciprocal can disqualify them
        } else {
          uint32_t slot = la_head + la_count;
          if (slot >= (uint32_t)kLookahead) slot -= (uint32_t)kLookahead;
          float(*la)[kAoBlock] = s_la[slot];
          la[0][threadIdx.x] = d.x; la[1][threadIdx.x] = d.y; la[2][threadIdx.x] = d.z;
          la[3][threadIdx.x] = id.x; la[4][threadIdx.x] = id.y; la[5][threadIdx.x] = id.z;
          la_count++;
        }
        pass++;
      }
    }
    if (!ray_active && la_count != 0u) start_queued();
    if (!__any_sync(0xffffffffu, ray_active)) {
      if (exhausted && !__any_sync(0xffffffffu, have_item)) break;  // nothing in flight and nothing left
      continue;
    }
    }   // (!PACKET)

    // ------------------------------ traverse ------------------------------
    // Triangle tests are batched across the warp.  A lane whose node test reports leaf hits does not
    // test them at once: it PAUSES (keeps its primitive group T, takes no further node step), and the
    // triangle block runs only when at least `tri_batch` lanes are paused or no lane can still take a
    // node step.  Run in place, the block costs the warp ~200 instructions in EVERY iteration for the
    // two or three lanes that happen to have reached a leaf in that iteration (30-40 % of all issued
    // instructions on configs 3 and 4, profiles/r2/srcprof_base_*.txt); batched it runs every few
    // iterations for `tri_batch` lanes.  Any-hit semantics are untouched: a paused ray does no
    // speculative work and resumes (or ends) with the result of its own test.
    uint32_t act = __ballot_sync(0xffffffffu, ray_active);
    while (true) {
      bool hit = false;
      if (ray_active && !paused) {
        T.x = 0; T.y = 0;
        if (G.y & 0xff000000u) {
          const int bit = 31 - __clz((int)G.y);
          G.y &= ~(1u << bit);
          const uint32_t slot = (uint32_t)bit - 24u;
          const uint32_t node = G.x + (uint32_t)__popc(G.y & 0xffu & ((1u << slot) - 1u));
          if (G.y & 0xff000000u) push(sp, G);
          uint32_t cb, pb, im, mlo, mhi;
          const uint32_t hb = H2 ? intersect_node8_h2_raw<CLAMP_TMAX>(bvh.nodes, node, r, &cb, &pb, &im, &mlo, &mhi)
                                 : intersect_node8_raw<CLAMP_TMAX>(bvh.nodes, node, r, nc, &cb, &pb, &im, &mlo, &mhi);
          if (STATS) c_nodes++;
          G.x = cb; G.y = ((hb & im) << 24) | im;
          T.x = node; T.y = hb & ~im;   // (node, hit leaf slots): expanded to primitives when the block runs
        } else if (TWO_LEVEL) {
          T = G;  // a postponed TLAS primitive group (node, leaf slots left)
          G.x = 0; G.y = 0;
        }
        // leaf hits (triangles inside a BLAS / a flattened scene, instances in a TLAS) are not processed at once: the
        // lane pauses and the warp handles them in one block (below)
        if (T.y) paused = true;
      }
      const uint32_t pm = __ballot_sync(0xffffffffu, paused);
      if (pm != 0u) {
        // run the block when enough lanes wait, when the oldest has waited tri_wait iterations (rare leaf hits must
        // not idle a lane for longer than a ray lives), or when no lane can take a node step any more
        waited++;
        if ((uint32_t)__popc(pm) >= tri_batch || waited >= tri_wait || pm == act) {
          waited = 0;
          if (paused) {
            paused = false;
            if (in_blas) {
              uint32_t tested = 0, pbase;
              const uint32_t pmask = leaf_slots_to_prims(bvh.nodes, T.x, T.y, &pbase);
              hit = test_tri_group(bvh.tris, pbase, pmask, r.org, r.dir, 0.0f, maxdist, &tested);
              if (STATS) c_tris += tested;
            } else if (TWO_LEVEL) {
              // instances of the group (one per TLAS leaf slot): skip those whose bounding sphere the world ray
              // cannot touch, enter the first one it can — save the TLAS continuation, switch to object space
              // (the node's primitive base and meta bytes are re-read per slot, L1 hits, rather than held in registers:
              //  this kernel has none to spare)
              const uint32_t* tn1 = reinterpret_cast<const uint32_t*>(bvh.nodes + 5ull * T.x + 1);
              while (T.y) {
                const int sl = __ffs((int)T.y) - 1;
                T.y &= T.y - 1u;
                const uint32_t b = (__ldg(tn1 + 2 + (sl >> 2)) >> (8 * (sl & 3))) & 31u;   // offset of the slot's instance
                const F4* rec = bvh.insts + (uint64_t)kInstF4 * ((uint64_t)__ldg(tn1 + 1) + b);
                if (!sphere_may_hit(org, wdir, ld_f4(rec + 4))) continue;
                if (T.y) push(sp, T);
                if (G.y & 0xff000000u) push(sp, G);
                U2 sen;
                sen.x = kSentinel; sen.y = 0;
                push(sp, sen);
                const F4 r0 = ld_f4(rec), r1 = ld_f4(rec + 1), r2 = ld_f4(rec + 2), r3 = ld_f4(rec + 3);
                const float m[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
                if (STATS) c_insts++;
                r.org = xf_point(m, org);
                r.dir = xf_vector(m, wdir);
                r.idir = v3(safe_rcp(r.dir.x), safe_rcp(r.dir.y), safe_rcp(r.dir.z));
                in_blas = true;
                G.x = __float_as_uint(r3.x);
                G.y = (1u << 24) | 1u;
                break;
              }
            }
          }
        }
      }
      // Ray end and restart are written once, straight-line: lanes ending on a hit, lanes ending
      // on an empty stack and lanes that merely pop all run the same short sequence instead of
      // three serialised divergent copies.
      if (ray_active && !paused) {
        bool done = hit;
        if (!hit && (G.y & 0xff000000u) == 0u) {
          while (true) {
            if (sp == 0) { done = true; break; }
            G = pop(sp);
            if (TWO_LEVEL && G.x == kSentinel && G.y == 0u) {
              r.org = org; r.dir = wdir;
              r.idir = v3(safe_rcp(wdir.x), safe_rcp(wdir.y), safe_rcp(wdir.z));
              in_blas = false;
              continue;
            }
            break;
          }
        }
        if (PACKET) {
          if (hit) atomicAdd(&s_pk_hits[threadIdx.x >> 5][cur_tag >> 5][(cur_tag & 31u) >> 1], 1u << (16u * (cur_tag & 1u)));
        } else {
          nh += hit ? 1u : 0u;
        }
        if (done) {
          ray_active = false;
          if (la_count != 0u) start_queued();
        }
      }
      act = __ballot_sync(0xffffffffu, ray_active);   // (unchanged until the end of the next iteration: reused by the pause test)
      if (act == 0u) break;
      if ((uint32_t)__popc(act) < refill_below) {
        // leave only if some idle lane can actually take a new ray
        const bool can = !ray_active && (PACKET ? (pk_next < pk_end || !exhausted)
                                                : ((have_item && pass < pass_end) || !exhausted));  // (a queued ray would already have started)
        if (__any_sync(0xffffffffu, can)) break;
      }
    }
  }
  if (PACKET) { pk_flush(0u); pk_flush(1u); }
  if (STATS) {
    atomicAdd(&stats[0], (unsigned long long)c_nodes);
    atomicAdd(&stats[1], (unsigned long long)c_tris);
    atomicAdd(&stats[2], (unsigned long long)c_insts);
  }
}

// The rays k_ao_persistent<.., H2 = true> set aside, traced with the fp32 node test; runs after it on
// the same stream and adds to the same hit counters.
__global__ void __launch_bounds__(128) k_ao_deferred(BvhView bvh, SampleView S, uint64_t begin, int q, float offset, float maxdist,
                                                     DeferredRays D, uint32_t* __restrict__ hits) {
  const uint32_t count = *D.count;
  const uint32_t n = count < D.capacity ? count : D.capacity;
  U2 stack[kStackSize];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const U2 e = D.list[i];
    const uint64_t g = begin + e.x;
    const V3 p = v3(S.pos[3 * g], S.pos[3 * g + 1], S.pos[3 * g + 2]);
    const V3 nrm = v3(S.nrm[3 * g], S.nrm[3 * g + 1], S.nrm[3 * g + 2]);
    const V3 fnrm = v3(S.fnrm[3 * g], S.fnrm[3 * g + 1], S.fnrm[3 * g + 2]);
    const Onb onb = make_onb(nrm);
    const V3 d = ao_ray_dir((uint32_t)g, e.y, q, nrm, fnrm, onb);
    if (trace_any_hit<false, false>(bvh, ao_ray_origin(p, nrm, offset), d, 0.0f, maxdist, stack, nullptr)) atomicAdd(&hits[e.x], 1u);
  }
}

// ao = 1 - hits/q^2 for the samples this launch owns; 0 for the others (interleaved multi-GPU
// partition), so that a sum all-reduce assembles the full array exactly (x + 0 + ... + 0 = x).
__global__ void k_ao_finalize(const uint32_t* __restrict__ hits, uint64_t n, float denom, float* __restrict__ ao, uint32_t part,
                              uint32_t num_parts, uint32_t block_samples) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool owned = num_parts <= 1 || (i / block_samples) % num_parts == part;
  ao[i] = owned ? ex::sub(1.0f, ex::div((float)hits[i], denom)) : 0.0f;
}

// ---------------------------------------------------------------------------------------
// Vertex maps
// ---------------------------------------------------------------------------------------
__global__ void k_area_final(const double* __restrict__ num, const double* __restrict__ wgt, uint64_t nV, float* __restrict__ out) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  out[v] = wgt[v] > 0.0 ? (float)(num[v] / wgt[v]) : 0.0f;
}

// ---- least squares (bake_filter_least_squares.cpp; Kavan et al. 2011) ---------------------
// half-edge keys: (min(a,b) << 32 | max(a,b)), value = 3*tri + e; degenerate edges get key ~0
__global__ void k_ls_halfedges(const uint32_t* __restrict__ tris, uint64_t nT, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * nT) return;
  const uint64_t t = i / 3;
  const uint32_t e = (uint32_t)(i % 3);
  const uint32_t a = tris[3 * t + e], b = tris[3 * t + (e + 1) % 3];
  keys[i] = (a == b) ? ~0ull : (((uint64_t)min(a, b) << 32) | max(a, b));
  vals[i] = (uint32_t)i;
}
// One interior edge (i, j) with opposite vertices p, q (global vertex numbers): s = foot parameter of the
// opposite vertex along the edge, h = its altitude, c = m1 . m2 (the unit in-plane edge normals pointing at
// p and q; -1 for a flat pair), W = energy weight.  Energy = W (a1^2 + a2^2 - 2 c a1 a2) with
// a1 = (x_p - s1 x_j - (1 - s1) x_i) / h1, a2 likewise with q — the squared jump of the 3-D gradient of the
// piecewise-linear interpolant across the edge times (A1 + A2) (see oracle/ao_oracle.cpp).
struct LsEdge {
  uint32_t i, j, p, q;
  double s1, h1, s2, h2, c, W;   // W == 0: degenerate edge, contributes nothing
};
AOB_D void ls_edge_coeffs(const LsEdge& E, double* al, double* be) {
  al[0] = -(1.0 - E.s1) / E.h1; al[1] = -E.s1 / E.h1; al[2] = 1.0 / E.h1;
  be[0] = -(1.0 - E.s2) / E.h2; be[1] = -E.s2 / E.h2; be[2] = 1.0 / E.h2;
}
// ---- batched (block-diagonal over instances) least-squares assembly --------------------------
// All instances are solved as ONE system: global vertex = instance vertex offset + mesh vertex.
// A scene of 1000 instances then costs ~100 CG iterations of a few launches each instead of
// 1000 separate solves (config 4: 5.8 s -> see DESIGN.md §4.4).
struct LsInst {
  float xf[12];
  const uint32_t* tris;   // mesh triangles (mesh-local vertex indices)
  const float* verts;
  const uint32_t* topo;   // interior-edge topology of the mesh: (i, j, p, q) per edge, mesh-local
  uint64_t sample_begin;
  uint64_t tri_begin;     // first global triangle of the instance
  uint64_t edge_begin;    // first global edge
  uint32_t vert_begin;    // first global vertex
  uint32_t num_tris;
  uint32_t num_edges;
  uint32_t pad;
};
template <typename KeyFn>
__device__ __forceinline__ uint32_t ls_find(uint32_t n_inst, uint64_t key, KeyFn begin_of) {
  uint32_t lo = 0, hi = n_inst;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (begin_of(mid) <= key) lo = mid; else hi = mid;
  }
  return lo;
}
// bake_filter.cpp filter_mesh for all instances at once: scatter ao*bary*dA and bary*dA to the three
// (global) vertices of the sample's triangle with fp64 atomics; k_area_final divides.
__global__ void k_area_scatter_b(const AoSampleInfo* __restrict__ info, const float* __restrict__ ao, uint64_t n_samples,
                                 const LsInst* __restrict__ inst, uint32_t n_inst, double* __restrict__ num, double* __restrict__ wgt) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_samples) return;
  const LsInst& I = inst[ls_find(n_inst, s, [&](uint32_t m) { return inst[m].sample_begin; })];
  const AoSampleInfo si = info[s];
  const double dA = si.dA, val = (double)ao[s] * dA;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const uint32_t v = I.tris[3ull * si.tri_idx + c] + I.vert_begin;
    atomicAdd(&num[v], (double)si.bary[c] * val);
    atomicAdd(&wgt[v], (double)si.bary[c] * dA);
  }
}
__global__ void k_ls_gtris(const LsInst* __restrict__ inst, uint32_t n_inst, uint64_t NT, uint32_t* __restrict__ gtris) {
  const uint64_t gt = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gt >= NT) return;
  const LsInst& I = inst[ls_find(n_inst, gt, [&](uint32_t m) { return inst[m].tri_begin; })];
  const uint64_t tl = gt - I.tri_begin;
#pragma unroll
  for (int c = 0; c < 3; c++) gtris[3 * gt + c] = I.tris[3 * tl + c] + I.vert_begin;
}
__global__ void k_ls_mass_b(const AoSampleInfo* __restrict__ info, const float* __restrict__ ao, uint64_t n_samples,
                            const LsInst* __restrict__ inst, uint32_t n_inst, const uint32_t* __restrict__ gtris,
                            double* __restrict__ Mt, double* __restrict__ rhs) {
  const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_samples) return;
  const LsInst& I = inst[ls_find(n_inst, s, [&](uint32_t m) { return inst[m].sample_begin; })];
  const AoSampleInfo si = info[s];
  const uint64_t gt = I.tri_begin + si.tri_idx;
  const double dA = si.dA, a = ao[s], b0 = si.bary[0], b1 = si.bary[1], b2 = si.bary[2];
  double* M = Mt + 6 * gt;
  atomicAdd(&M[0], dA * b0 * b0); atomicAdd(&M[1], dA * b0 * b1); atomicAdd(&M[2], dA * b0 * b2);
  atomicAdd(&M[3], dA * b1 * b1); atomicAdd(&M[4], dA * b1 * b2); atomicAdd(&M[5], dA * b2 * b2);
  const uint32_t* idx = gtris + 3 * gt;
  atomicAdd(&rhs[idx[0]], dA * a * b0); atomicAdd(&rhs[idx[1]], dA * a * b1); atomicAdd(&rhs[idx[2]], dA * a * b2);
}
// interior-edge topology of one mesh from its sorted half-edges: the first of a run of >= 2 equal
// keys emits (i, j, p, q); non-manifold edges pair their first two triangles (decision #6)
__global__ void k_ls_topo(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t nH,
                          const uint32_t* __restrict__ tris, uint32_t* __restrict__ topo, uint32_t* __restrict__ count) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k + 1 >= nH) return;
  const uint64_t key = keys[k];
  if (key == ~0ull) return;
  if (k > 0 && keys[k - 1] == key) return;
  if (keys[k + 1] != key) return;
  const uint32_t h0 = vals[k], h1 = vals[k + 1];
  const uint32_t e = atomicAdd(count, 1u);
  topo[4ull * e] = (uint32_t)(key >> 32);
  topo[4ull * e + 1] = (uint32_t)(key & 0xffffffffu);
  topo[4ull * e + 2] = tris[3ull * (h0 / 3) + (h0 % 3 + 2) % 3];
  topo[4ull * e + 3] = tris[3ull * (h1 / 3) + (h1 % 3 + 2) % 3];
}
// coefficients of the co-normal-derivative jump for every (instance, interior edge), world space
__global__ void k_ls_edge_coeffs(const LsInst* __restrict__ inst, uint32_t n_inst, uint64_t NE, int energy, LsEdge* __restrict__ edges) {
  const uint64_t ge = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ge >= NE) return;
  const LsInst& I = inst[ls_find(n_inst, ge, [&](uint32_t m) { return inst[m].edge_begin; })];
  const uint32_t* tp = I.topo + 4 * (ge - I.edge_begin);
  LsEdge E;
  const uint32_t li = tp[0], lj = tp[1], lp = tp[2], lq = tp[3];
  E.i = li + I.vert_begin; E.j = lj + I.vert_begin; E.p = lp + I.vert_begin; E.q = lq + I.vert_begin;
  E.s1 = E.s2 = 0.0; E.h1 = E.h2 = 1.0; E.c = -1.0; E.W = 0.0;
  auto W = [&](uint32_t v) { return xf_point(I.xf, v3(I.verts[3ull * v], I.verts[3ull * v + 1], I.verts[3ull * v + 2])); };
  const V3 pi = W(li), pj = W(lj), pp = W(lp), pq = W(lq);
  const double ex_ = (double)pj.x - pi.x, ey_ = (double)pj.y - pi.y, ez_ = (double)pj.z - pi.z;
  const double L2 = ex_ * ex_ + ey_ * ey_ + ez_ * ez_;
  if (L2 > 0.0) {
    double s[2], h[2], A[2], r[2][3];
    const V3 opp[2] = {pp, pq};
#pragma unroll
    for (int m = 0; m < 2; m++) {
      const double ox = (double)opp[m].x - pi.x, oy = (double)opp[m].y - pi.y, oz = (double)opp[m].z - pi.z;
      s[m] = (ox * ex_ + oy * ey_ + oz * ez_) / L2;
      r[m][0] = ox - s[m] * ex_; r[m][1] = oy - s[m] * ey_; r[m][2] = oz - s[m] * ez_;
      h[m] = sqrt(r[m][0] * r[m][0] + r[m][1] * r[m][1] + r[m][2] * r[m][2]);
      A[m] = 0.5 * sqrt(L2) * h[m];
    }
    if (h[0] > 0.0 && h[1] > 0.0) {
      E.s1 = s[0]; E.h1 = h[0]; E.s2 = s[1]; E.h2 = h[1];
      if (energy == 1) { E.c = -1.0; E.W = (A[0] + A[1]) * (A[0] + A[1]); }
      else { E.c = (r[0][0] * r[1][0] + r[0][1] * r[1][1] + r[0][2] * r[1][2]) / (h[0] * h[1]); E.W = A[0] + A[1]; }
    }
  }
  edges[ge] = E;   // degenerate edges keep W = 0: a no-op in y = (M + wR) x
}

// diagonal of the sampled mass matrix (lumped-mass test and the Jacobi preconditioner)
__global__ void k_ls_diag_mass(const uint32_t* __restrict__ tris, uint64_t nT, const double* __restrict__ Mt, double* __restrict__ diag) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nT) return;
  const double a = Mt[6 * i], b = Mt[6 * i + 3], c = Mt[6 * i + 5];
  if (a != 0.0) atomicAdd(&diag[tris[3 * i]], a);
  if (b != 0.0) atomicAdd(&diag[tris[3 * i + 1]], b);
  if (c != 0.0) atomicAdd(&diag[tris[3 * i + 2]], c);
}
// decision #7: vertices with zero lumped mass get M_vv = 1, rhs 0
__global__ void k_ls_fix(double* __restrict__ diag, double* __restrict__ rhs, uint8_t* __restrict__ fixed, uint64_t nV) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const bool f = !(diag[v] > 0.0);
  fixed[v] = f ? 1 : 0;
  if (f) { diag[v] = 1.0; rhs[v] = 0.0; }
}
__global__ void k_ls_diag_edges(const LsEdge* __restrict__ edges, uint32_t nE, double w, double* __restrict__ diag) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nE) return;
  const LsEdge E = edges[i];
  if (E.W == 0.0) return;
  double al[3], be[3];
  ls_edge_coeffs(E, al, be);
  const double ww = w * E.W;
  atomicAdd(&diag[E.i], ww * (al[0] * al[0] + be[0] * be[0] - 2.0 * E.c * al[0] * be[0]));
  atomicAdd(&diag[E.j], ww * (al[1] * al[1] + be[1] * be[1] - 2.0 * E.c * al[1] * be[1]));
  atomicAdd(&diag[E.p], ww * al[2] * al[2]);
  atomicAdd(&diag[E.q], ww * be[2] * be[2]);
}
// y += (M + w R) x   (y zeroed by the caller); one thread per triangle and per edge
__global__ void k_ls_apply(const uint32_t* __restrict__ tris, uint64_t nT, const double* __restrict__ Mt, const LsEdge* __restrict__ edges,
                           uint32_t nE, double w, const double* __restrict__ x, double* __restrict__ y) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nT) {
    const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
    const double* M = Mt + 6 * i;
    const double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3], m4 = M[4], m5 = M[5];
    if (m0 != 0.0 || m3 != 0.0 || m5 != 0.0) {
      const double x0 = x[a], x1 = x[b], x2 = x[c];
      atomicAdd(&y[a], m0 * x0 + m1 * x1 + m2 * x2);
      atomicAdd(&y[b], m1 * x0 + m3 * x1 + m4 * x2);
      atomicAdd(&y[c], m2 * x0 + m4 * x1 + m5 * x2);
    }
  }
  if (i < nE) {
    const LsEdge E = edges[i];
    if (E.W != 0.0) {
      double al[3], be[3];
      ls_edge_coeffs(E, al, be);
      const double xi = x[E.i], xj = x[E.j];
      const double a1 = al[0] * xi + al[1] * xj + al[2] * x[E.p], a2 = be[0] * xi + be[1] * xj + be[2] * x[E.q];
      const double g1 = w * E.W * (a1 - E.c * a2), g2 = w * E.W * (a2 - E.c * a1);
      atomicAdd(&y[E.i], g1 * al[0] + g2 * be[0]); atomicAdd(&y[E.j], g1 * al[1] + g2 * be[1]);
      atomicAdd(&y[E.p], g1 * al[2]); atomicAdd(&y[E.q], g2 * be[2]);
    }
  }
}
// ---- Jacobi-PCG pieces.  Scalars live on the device: S[0] = |b|^2, S[1] = r.z of the previous iteration, and two
// banks S[4 + 4 k ...] (k = iteration parity) of {p.Ap, r.z, |r|^2} accumulators; the bank of the NEXT iteration is
// zeroed by k_ls_dir, so an iteration is four launches with no memset, no copy and no host round trip (the host reads
// |r|^2 every few iterations only). ----
AOB_D double block_sum(double acc, double* sh) {   // all threads of a 256-thread block; result valid in thread 0
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
  if (threadIdx.x < 32)
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}
// out[slot] += sum a[i]*b[i]  (block reduce + one fp64 atomic per block)
__global__ void k_dot(const double* __restrict__ a, const double* __restrict__ b, uint64_t n, double* __restrict__ out) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) acc += a[i] * b[i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(out, acc);
}
// anchored rows (decision #7): Ap += p there; then bank[0] += p.Ap
__global__ void k_ls_pap(const uint8_t* __restrict__ fixed, const double* __restrict__ p, double* __restrict__ Ap, uint64_t nV, double* __restrict__ bank) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += (uint64_t)gridDim.x * blockDim.x) {
    const double pv = p[v];
    double a = Ap[v];
    if (fixed[v]) { a += pv; Ap[v] = a; }
    acc += pv * a;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(&bank[0], acc);
}
// z = r / diag ; p = z (init) ; x = 0
__global__ void k_ls_init(const double* __restrict__ rhs, const double* __restrict__ diag, double* __restrict__ r, double* __restrict__ z,
                          double* __restrict__ p, double* __restrict__ x, double* __restrict__ Ap, uint64_t nV) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  x[v] = 0.0;
  r[v] = rhs[v];
  z[v] = rhs[v] / diag[v];
  p[v] = z[v];
  Ap[v] = 0.0;
}
// x += alpha p ; r -= alpha Ap ; z = r/diag ; bank[1] += r.z ; bank[2] += |r|^2     (alpha = S[1] / bank[0])
__global__ void k_ls_update(const double* __restrict__ S, const double* __restrict__ p, const double* __restrict__ Ap,
                            const double* __restrict__ diag, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, uint64_t nV,
                            double* bank) {
  __shared__ double sh[32];
  const double alpha = S[1] / bank[0];   // (bank[0] is complete: k_ls_pap ran before this launch)
  double rz = 0.0, rr = 0.0;
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += (uint64_t)gridDim.x * blockDim.x) {
    x[v] += alpha * p[v];
    const double rv = r[v] - alpha * Ap[v];
    r[v] = rv;
    const double zv = rv / diag[v];
    z[v] = zv;
    rz += rv * zv;
    rr += rv * rv;
  }
  rz = block_sum(rz, sh);
  __syncthreads();
  rr = block_sum(rr, sh);
  if (threadIdx.x == 0) { atomicAdd(&bank[1], rz); atomicAdd(&bank[2], rr); }
}
// p = z + beta p ; Ap = 0 for the next product     (beta = bank[1] / S[1]); the last block to finish rotates the
// scalars: S[1] <- bank[1], S[2] <- bank[2] (|r|^2 for the host), S[3] <- bank[0] (p.Ap, breakdown check), and zeroes the other bank
__global__ void k_ls_dir(double* __restrict__ S, double* __restrict__ bank, double* __restrict__ other_bank, const double* __restrict__ z,
                         double* __restrict__ p, double* __restrict__ Ap, uint64_t nV, unsigned int* __restrict__ done_blocks) {
  const double beta = bank[1] / S[1];
  for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nV; v += (uint64_t)gridDim.x * blockDim.x) {
    p[v] = z[v] + beta * p[v];
    Ap[v] = 0.0;
  }
  __syncthreads();   // every thread of the block has read S[1] and bank[1]
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done_blocks, 1u) == gridDim.x - 1) {   // all blocks have read the scalars
      S[3] = bank[0]; S[2] = bank[2]; S[1] = bank[1];
      other_bank[0] = other_bank[1] = other_bank[2] = 0.0;
      *done_blocks = 0u;
      __threadfence();
    }
  }
}
// ---- explicit matrix for the PCG product (round 2) --------------------------------------------------------------
// The matrix-free product above scatters 4 fp64 atomics per edge and 3 per triangle (180 M atomics per iteration on
// config 5) and re-reads every 64-byte edge record: ~1.07 ms per iteration for 10 M unknowns.  The few thousand
// iterations the faithful regulariser needs (decision 6) make it worth assembling A = M + wR once:
//   1. every owned row gets a small open-addressing hash table (kLsHashCap columns); triangles add their 3 x 3 mass
//      block, edges their 4 x 4 block w W [(al al^T + be be^T) - c (al be^T + be al^T)], anchored rows +1 on the
//      diagonal — fp64 atomics once, not once per iteration;
//   2. the tables are compacted into sliced ELL (slices of 32 rows, padded to the slice's longest row, column-major
//      inside a slice: coalesced loads, ~14 entries per row on a triangle mesh);
//   3. the product is one thread per row, no atomics, deterministic, with the anchored rows and the p.Ap partial
//      sums folded in.
// A row with more distinct columns than the table holds (a vertex of valence > ~14: the pole of a UV sphere, the hub
// of a fan) is flagged; flagged rows are left empty in the matrix and multiplied matrix-free from the (few) triangles
// and edges that touch them (k_ls_flag_over_items, k_ls_apply_rows with a row mask, k_ls_pap_list).
constexpr uint32_t kLsHashCap = 32;
constexpr uint32_t kLsEmpty = 0xffffffffu;
AOB_D void ls_hash_add(uint32_t* __restrict__ keys, double* __restrict__ vals, uint64_t row_rel, uint32_t col, double v, uint8_t* __restrict__ row_over) {
  uint32_t* k = keys + row_rel * kLsHashCap;
  double* d = vals + row_rel * kLsHashCap;
  uint32_t s = (col * 2654435761u) >> 27;   // top 5 bits
  for (uint32_t probe = 0; probe < kLsHashCap; probe++, s = (s + 1u) & (kLsHashCap - 1u)) {
    const uint32_t old = atomicCAS(&k[s], kLsEmpty, col);
    if (old == kLsEmpty || old == col) { atomicAdd(&d[s], v); return; }
  }
  row_over[row_rel] = 1;   // more distinct columns than the table holds: this ROW is multiplied matrix-free
}
// one thread per listed (or every) triangle / edge: rows outside [v0, v1) are skipped
__global__ void k_ls_assemble(const uint32_t* __restrict__ tri_list, uint64_t n_tri, const uint32_t* __restrict__ edge_list, uint64_t n_edge,
                              const uint32_t* __restrict__ tris, const double* __restrict__ Mt, const LsEdge* __restrict__ edges, double w,
                              uint32_t v0, uint32_t v1, uint32_t* __restrict__ keys, double* __restrict__ vals, uint8_t* __restrict__ row_over) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto add = [&](uint32_t r, uint32_t c, double v) { if (r >= v0 && r < v1 && v != 0.0) ls_hash_add(keys, vals, r - v0, c, v, row_over); };
  if (k < n_tri) {
    const uint64_t i = tri_list ? tri_list[k] : k;
    const double* M = Mt + 6 * i;
    const double m[6] = {M[0], M[1], M[2], M[3], M[4], M[5]};
    if (m[0] != 0.0 || m[3] != 0.0 || m[5] != 0.0) {
      const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
      add(a, a, m[0]); add(a, b, m[1]); add(a, c, m[2]);
      add(b, a, m[1]); add(b, b, m[3]); add(b, c, m[4]);
      add(c, a, m[2]); add(c, b, m[4]); add(c, c, m[5]);
    }
  }
  if (k < n_edge) {
    const LsEdge E = edges[edge_list ? edge_list[k] : k];
    if (E.W != 0.0) {
      double al3[3], be3[3];
      ls_edge_coeffs(E, al3, be3);
      const uint32_t v[4] = {E.i, E.j, E.p, E.q};
      const double al[4] = {al3[0], al3[1], al3[2], 0.0}, be[4] = {be3[0], be3[1], 0.0, be3[2]};
      const double ww = w * E.W;
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) add(v[r], v[c], ww * (al[r] * al[c] + be[r] * be[c] - E.c * (al[r] * be[c] + be[r] * al[c])));
    }
  }
}
// anchored rows (decision #7) carry +1 on the diagonal; every row gets its diagonal slot so that no row is empty
__global__ void k_ls_assemble_diag(const uint8_t* __restrict__ fixed, uint32_t v0, uint32_t v1, uint32_t* __restrict__ keys, double* __restrict__ vals,
                                   uint8_t* __restrict__ row_over) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < v1 - v0) ls_hash_add(keys, vals, r, v0 + (uint32_t)r, fixed[v0 + r] ? 1.0 : 0.0, row_over);
}
// items that touch a flagged row (row_over is indexed by v - v0 for v in [v0, v1))
__global__ void k_ls_flag_over_items(const uint32_t* __restrict__ tri_list, uint64_t n_tri, const uint32_t* __restrict__ edge_list, uint64_t n_edge,
                                     const uint32_t* __restrict__ tris, const LsEdge* __restrict__ edges, uint32_t v0, uint32_t v1,
                                     const uint8_t* __restrict__ row_over, uint8_t* __restrict__ tri_flag, uint8_t* __restrict__ edge_flag) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto over = [&](uint32_t v) { return v >= v0 && v < v1 && row_over[v - v0] != 0; };
  if (k < n_tri) {
    const uint64_t i = tri_list ? tri_list[k] : k;
    tri_flag[k] = (over(tris[3 * i]) || over(tris[3 * i + 1]) || over(tris[3 * i + 2])) ? 1 : 0;
  }
  if (k < n_edge) {
    const LsEdge E = edges[edge_list ? edge_list[k] : k];
    edge_flag[k] = (E.W != 0.0 && (over(E.i) || over(E.j) || over(E.p) || over(E.q))) ? 1 : 0;
  }
}
// out[k] = list ? list[sel[k]] : sel[k]   (positions selected in a list -> the item ids themselves)
__global__ void k_compose_u32(const uint32_t* __restrict__ list, const uint32_t* __restrict__ sel, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = list ? list[sel[k]] : sel[k];
}
// flagged rows: anchored-row term and their share of p.Ap (one block)
__global__ void k_ls_pap_list(const uint32_t* __restrict__ rows_rel, uint32_t n, uint32_t v0, const uint8_t* __restrict__ fixed, const double* __restrict__ p,
                              double* __restrict__ Ap, double* __restrict__ bank) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {
    const uint32_t v = v0 + rows_rel[k];
    double a = Ap[v];
    if (fixed[v]) { a += p[v]; Ap[v] = a; }
    acc += p[v] * a;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(&bank[0], acc);
}
__global__ void k_u32_to_u64(const uint32_t* __restrict__ a, uint64_t n, uint64_t* __restrict__ out) {   // out[n] = 0 (scan sentinel)
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n) out[i] = i < n ? a[i] : 0ull;
}
// entries per row and, per slice of 32 rows, the padded width
__global__ void k_ls_row_widths(const uint32_t* __restrict__ keys, const uint8_t* __restrict__ row_over, uint64_t n_rows, uint32_t* __restrict__ row_nnz,
                                uint32_t* __restrict__ slice_width) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n = 0;
  if (r < n_rows && !row_over[r])
    for (uint32_t s = 0; s < kLsHashCap; s++) n += keys[r * kLsHashCap + s] != kLsEmpty ? 1u : 0u;
  if (r < n_rows) row_nnz[r] = n;
  uint32_t m = n;
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && r < n_rows) slice_width[r >> 5] = m;
}
// sliced ELL: entry k of row r sits at slice_off[r / 32] * 32 + k * 32 + r % 32; padding = (column r, value 0)
__global__ void k_ls_fill_sell(const uint32_t* __restrict__ keys, const double* __restrict__ vals, const uint8_t* __restrict__ row_over, uint64_t n_rows, uint32_t v0,
                               const uint64_t* __restrict__ slice_off, const uint32_t* __restrict__ slice_width, uint32_t* __restrict__ cols,
                               double* __restrict__ A) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint64_t base = slice_off[r >> 5] * 32ull + (r & 31);
  const uint32_t width = slice_width[r >> 5];
  uint32_t k = 0;
  for (uint32_t s = 0; s < kLsHashCap && !row_over[r]; s++) {
    const uint32_t c = keys[r * kLsHashCap + s];
    if (c != kLsEmpty) { cols[base + 32ull * k] = c; A[base + 32ull * k] = vals[r * kLsHashCap + s]; k++; }
  }
  for (; k < width; k++) { cols[base + 32ull * k] = v0 + (uint32_t)r; A[base + 32ull * k] = 0.0; }
}
// Ap[v0 + r] = (A p)[row r]; bank[0] += p.Ap over these rows.  256-thread blocks (8 slices).
__global__ void k_ls_spmv_sell(const uint64_t* __restrict__ slice_off, const uint32_t* __restrict__ slice_width, const uint32_t* __restrict__ cols,
                               const double* __restrict__ A, uint64_t n_rows, uint32_t v0, const double* __restrict__ p, double* __restrict__ Ap,
                               double* __restrict__ bank) {
  __shared__ double sh[32];
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0, y = 0.0;
  if (r < n_rows) {
    const uint64_t base = slice_off[r >> 5] * 32ull + (r & 31);
    const uint32_t width = slice_width[r >> 5];
    for (uint32_t k = 0; k < width; k++) y += A[base + 32ull * k] * __ldg(&p[cols[base + 32ull * k]]);
    Ap[v0 + r] = y;
    acc = p[v0 + r] * y;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(&bank[0], acc);
}

// ---- row-partitioned PCG across ranks (one big system, e.g. the single 10 M-vertex instance of config 5) ----
// Rank r owns the vertex rows [v0, v1).  It walks only the triangles / edges that touch one of its rows
// (flagged once, compacted into item lists) and adds only into its own rows, so the product needs no
// reduction across ranks; what it needs from the others is p at the vertices its items reference — the
// BOUNDARY vertices (those sharing an item with a vertex of another owner), exchanged once per iteration.
__global__ void k_ls_flag_items(const uint32_t* __restrict__ gtris, uint64_t NT, const LsEdge* __restrict__ edges, uint64_t NE, uint32_t v0, uint32_t v1,
                                uint32_t rows_per_rank, uint8_t* __restrict__ tri_mine, uint8_t* __restrict__ edge_mine, uint8_t* __restrict__ boundary) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto mine = [&](uint32_t v) { return v >= v0 && v < v1; };
  if (i < NT) {
    const uint32_t a = gtris[3 * i], b = gtris[3 * i + 1], c = gtris[3 * i + 2];
    tri_mine[i] = (mine(a) || mine(b) || mine(c)) ? 1 : 0;
    const uint32_t oa = a / rows_per_rank, ob = b / rows_per_rank, oc = c / rows_per_rank;
    if (oa != ob || oa != oc) { boundary[a] = 1; boundary[b] = 1; boundary[c] = 1; }
  }
  if (i < NE) {
    const LsEdge E = edges[i];
    const bool live = E.W != 0.0;
    edge_mine[i] = (live && (mine(E.i) || mine(E.j) || mine(E.p) || mine(E.q))) ? 1 : 0;
    const uint32_t o = E.i / rows_per_rank;
    if (live && (E.j / rows_per_rank != o || E.p / rows_per_rank != o || E.q / rows_per_rank != o)) {
      boundary[E.i] = 1; boundary[E.j] = 1; boundary[E.p] = 1; boundary[E.q] = 1;
    }
  }
}
// y[rows v0..v1) += (M + w R) x over this rank's item lists
__global__ void k_ls_apply_rows(const uint32_t* __restrict__ tri_list, uint32_t n_tri, const uint32_t* __restrict__ edge_list, uint32_t n_edge,
                                const uint32_t* __restrict__ tris, const double* __restrict__ Mt, const LsEdge* __restrict__ edges, double w,
                                uint32_t v0, uint32_t v1, const uint8_t* __restrict__ row_mask /*nullable: only rows with row_mask[v - v0]*/,
                                const double* __restrict__ x, double* __restrict__ y) {
  const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto add = [&](uint32_t v, double val) { if (v >= v0 && v < v1 && (!row_mask || row_mask[v - v0])) atomicAdd(&y[v], val); };
  if (k < n_tri) {
    const uint64_t i = tri_list[k];
    const uint32_t a = tris[3 * i], b = tris[3 * i + 1], c = tris[3 * i + 2];
    const double* M = Mt + 6 * i;
    const double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3], m4 = M[4], m5 = M[5];
    if (m0 != 0.0 || m3 != 0.0 || m5 != 0.0) {
      const double x0 = x[a], x1 = x[b], x2 = x[c];
      add(a, m0 * x0 + m1 * x1 + m2 * x2);
      add(b, m1 * x0 + m3 * x1 + m4 * x2);
      add(c, m2 * x0 + m4 * x1 + m5 * x2);
    }
  }
  if (k < n_edge) {
    const LsEdge E = edges[edge_list[k]];
    double al[3], be[3];
    ls_edge_coeffs(E, al, be);
    const double xi = x[E.i], xj = x[E.j];
    const double a1 = al[0] * xi + al[1] * xj + al[2] * x[E.p], a2 = be[0] * xi + be[1] * xj + be[2] * x[E.q];
    const double g1 = w * E.W * (a1 - E.c * a2), g2 = w * E.W * (a2 - E.c * a1);
    add(E.i, g1 * al[0] + g2 * be[0]); add(E.j, g1 * al[1] + g2 * be[1]);
    add(E.p, g1 * al[2]); add(E.q, g2 * be[2]);
  }
}
__global__ void k_iota_u32(uint32_t* __restrict__ a, uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (uint32_t)i;
}
// first index k with list[k] >= bound[s], for every s (list sorted ascending)
__global__ void k_lower_bounds(const uint32_t* __restrict__ list, uint32_t n, const uint32_t* __restrict__ bound, uint32_t n_bounds, uint32_t* __restrict__ out) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_bounds) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (list[mid] < bound[s]) lo = mid + 1; else hi = mid;
  }
  out[s] = lo;
}
__global__ void k_gather_d(const double* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t begin, uint32_t end, double* __restrict__ buf) {
  const uint32_t k = begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (k < end) buf[k] = src[idx[k]];
}
// dst[idx[k]] = buf[k] for k outside [skip_begin, skip_end) (the caller's own segment)
__global__ void k_scatter_d(double* __restrict__ dst, const uint32_t* __restrict__ idx, uint32_t n, uint32_t skip_begin, uint32_t skip_end,
                            const double* __restrict__ buf) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && (k < skip_begin || k >= skip_end)) dst[idx[k]] = buf[k];
}
// out[v] = (float)x[v] for rows [v0, v1), 0 elsewhere (a sum all-reduce then assembles the vector)
__global__ void k_d2f_rows(const double* __restrict__ x, float* __restrict__ out, uint64_t n, uint32_t v0, uint32_t v1) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) out[v] = (v >= v0 && v < v1) ? (float)x[v] : 0.0f;
}
__global__ void k_d2f(const double* __restrict__ x, float* __restrict__ out, uint64_t n) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) out[v] = (float)x[v];
}

}  // namespace aob
