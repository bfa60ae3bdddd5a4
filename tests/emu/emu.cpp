// emu.cpp — host emulation harness (TEST ONLY): compiles the kernel bodies of
// optix_prime_baking_b200/csrc/*.cuh with plain g++ and runs them serially, so that the
// LBVH build, the 8-wide collapse, the quantised traversal, sample placement and ray
// generation can be checked against the oracle in a container without a GPU.  Never
// loaded by the product; the product path is CUDA only.
#define AOB_HOST_EMU 1
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>
#include "../../optix_prime_baking_b200/csrc/aob_bvh.cuh"

using namespace aob;

static int g_node_test_fp32 = 0;              // 1: fp32 node test for every ray (cross-check of the fp16 test)
static uint32_t g_max_leaf_override = 0;   // experiments: leaf slot capacity (0 = builder default)
static int g_no_oversized_split = 0;       // 1: keep oversized primitives in the tree (AoBakeParams::no_oversized_split)

struct EmuBvh {
  std::vector<Node8> nodes;
  std::vector<F4> tris;
  std::vector<F4> insts;
  uint32_t root = 0;
  bool two_level = false;
};

// Builds one BVH segment over `n` boxes appended at nodes.size(); returns root index and
// fills leaf_prims (leaf order -> primitive id).
static uint32_t build_tree(const std::vector<F4>& plo, const std::vector<F4>& phi, const std::vector<uint32_t>& subset, uint32_t max_leaf,
                           uint32_t prim_offset, std::vector<Node8>& nodes, std::vector<uint32_t>& leaf_prims);

// Builds one BVH segment over `n` boxes appended at nodes.size(); returns root index and fills leaf_prims (leaf order ->
// primitive id).  Mirrors build_segment of aobake.cu, including the split of oversized primitives onto an extra root.
static uint32_t build_segment(const std::vector<F4>& plo, const std::vector<F4>& phi, uint32_t max_leaf,
                              uint32_t prim_offset, std::vector<Node8>& nodes, std::vector<uint32_t>& leaf_prims) {
  const uint32_t n = (uint32_t)plo.size();
  std::vector<uint32_t> small, big;
  F4 alo, ahi, slo, shi;
  alo.x = alo.y = alo.z = slo.x = slo.y = slo.z = 3.0e38f; ahi.x = ahi.y = ahi.z = shi.x = shi.y = shi.z = -3.0e38f;
  alo.w = ahi.w = slo.w = shi.w = 0.f;
  for (uint32_t i = 0; i < n; i++) {
    alo.x = fminf(alo.x, plo[i].x); alo.y = fminf(alo.y, plo[i].y); alo.z = fminf(alo.z, plo[i].z);
    ahi.x = fmaxf(ahi.x, phi[i].x); ahi.y = fmaxf(ahi.y, phi[i].y); ahi.z = fmaxf(ahi.z, phi[i].z);
  }
  const float ext = n ? fmaxf(ahi.x - alo.x, fmaxf(ahi.y - alo.y, ahi.z - alo.z)) : 0.f;
  for (uint32_t i = 0; i < n; i++) (box_is_oversized(plo[i], phi[i], ext) ? big : small).push_back(i);
  const uint32_t per_slot = std::max(1u, std::min(max_leaf, 3u));
  const bool split = !g_no_oversized_split && !big.empty() && big.size() < n && big.size() <= 7u * per_slot;
  if (!split) {
    std::vector<uint32_t> all(n);
    std::iota(all.begin(), all.end(), 0u);
    return build_tree(plo, phi, all, max_leaf, prim_offset, nodes, leaf_prims);
  }
  for (uint32_t i : small) {
    slo.x = fminf(slo.x, plo[i].x); slo.y = fminf(slo.y, plo[i].y); slo.z = fminf(slo.z, plo[i].z);
    shi.x = fmaxf(shi.x, phi[i].x); shi.y = fmaxf(shi.y, phi[i].y); shi.z = fmaxf(shi.z, phi[i].z);
  }
  const uint32_t main_root = build_tree(plo, phi, small, max_leaf, prim_offset, nodes, leaf_prims);
  std::vector<uint32_t> sorted(small);          // positions [n_small, n) = the oversized primitives, in index order (stable sort of key ~0)
  sorted.insert(sorted.end(), big.begin(), big.end());
  leaf_prims.resize(n);
  Node8 nd;
  super_root_body(&nd, main_root, alo, ahi, slo, shi, plo.data(), phi.data(), sorted.data(), (uint32_t)small.size(), (uint32_t)big.size(), per_slot,
                  prim_offset, leaf_prims.data());
  nodes.push_back(nd);
  return (uint32_t)nodes.size() - 1;
}

// One LBVH -> 8-wide tree over the primitives listed in `subset`; leaf_prims gets subset.size() entries.
static uint32_t build_tree(const std::vector<F4>& plo, const std::vector<F4>& phi, const std::vector<uint32_t>& subset, uint32_t max_leaf,
                           uint32_t prim_offset, std::vector<Node8>& nodes, std::vector<uint32_t>& leaf_prims) {
  const uint32_t n = (uint32_t)subset.size();
  const uint32_t node_offset = (uint32_t)nodes.size();
  leaf_prims.assign(n, 0);
  if (n == 0) {
    Node8 nd;
    memset(&nd, 0, sizeof(nd));
    for (int k = 0; k < 8; k++) for (int a = 0; a < 3; a++) nd.q[a][k][0] = 255;
    nodes.push_back(nd);
    return node_offset;
  }
  V3 cmin = v3(1e30f, 1e30f, 1e30f), cmax = v3(-1e30f, -1e30f, -1e30f);
  for (uint32_t i : subset) {
    float cx = 0.5f * (plo[i].x + phi[i].x), cy = 0.5f * (plo[i].y + phi[i].y), cz = 0.5f * (plo[i].z + phi[i].z);
    cmin = v3(fminf(cmin.x, cx), fminf(cmin.y, cy), fminf(cmin.z, cz));
    cmax = v3(fmaxf(cmax.x, cx), fmaxf(cmax.y, cy), fmaxf(cmax.z, cz));
  }
  const float ext = fmaxf(cmax.x - cmin.x, fmaxf(cmax.y - cmin.y, cmax.z - cmin.z));
  const float inv = ext > 0.0f ? 1.0f / ext : 0.0f;
  V3 cinv = v3(inv, inv, inv);
  std::vector<uint64_t> keys(n);
  std::vector<uint32_t> pos(n), order(n);
  for (uint32_t k = 0; k < n; k++) keys[k] = morton63(plo[subset[k]], phi[subset[k]], cmin, cinv);
  std::iota(pos.begin(), pos.end(), 0u);
  std::stable_sort(pos.begin(), pos.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  std::vector<uint64_t> skeys(n);
  for (uint32_t k = 0; k < n; k++) { skeys[k] = keys[pos[k]]; order[k] = subset[pos[k]]; }
  const uint32_t ni = n > 1 ? n - 1 : 1;
  std::vector<uint32_t> left(ni), right(ni), first(ni), last(ni), pint(ni), pleaf(n), flags(ni, 0), wide2bin(n), count(ni, 0);
  std::vector<F4> ilo(ni), ihi(ni);
  Lbvh L;
  L.keys = skeys.data(); L.prim = order.data(); L.plo = plo.data(); L.phi = phi.data();
  L.left = left.data(); L.right = right.data(); L.first = first.data(); L.last = last.data();
  L.parent_int = pint.data(); L.parent_leaf = pleaf.data(); L.ilo = ilo.data(); L.ihi = ihi.data();
  L.flags = flags.data(); L.count = count.data(); L.n = n;
  const uint32_t root_ref = (n == 1) ? (0u | kLeafBit) : 0u;
  for (uint32_t i = 0; i + 1 < n; i++) lbvh_hierarchy_body(i, L);
  for (uint32_t i = 0; i < n; i++) lbvh_refit_body(i, L);
  nodes.resize(node_offset + n);
  uint32_t node_count = 1, prim_count = 0;
  CollapseArgs A;
  A.L = L; A.nodes = nodes.data(); A.wide2bin = wide2bin.data(); A.leaf_prims = leaf_prims.data();
  A.node_count = &node_count; A.prim_count = &prim_count; A.node_offset = node_offset; A.prim_offset = prim_offset;
  A.max_leaf = (g_max_leaf_override && max_leaf > 1) ? g_max_leaf_override : max_leaf;
  wide2bin[0] = root_ref;
  uint32_t lb = 0, le = 1;
  while (lb < le) {
    for (uint32_t w = lb; w < le; w++) collapse_body(w, A);
    lb = le;
    le = node_count;
  }
  if (prim_count != n) fprintf(stderr, "emu: prim_count %u != n %u\n", prim_count, n);
  nodes.resize(node_offset + node_count);
  return node_offset;
}

static void tri_boxes(const float* v9, uint32_t n, std::vector<F4>& lo, std::vector<F4>& hi) {
  lo.resize(n); hi.resize(n);
  for (uint32_t t = 0; t < n; t++) {
    const float* p = v9 + 9 * (size_t)t;
    lo[t].x = fminf(p[0], fminf(p[3], p[6])); lo[t].y = fminf(p[1], fminf(p[4], p[7])); lo[t].z = fminf(p[2], fminf(p[5], p[8])); lo[t].w = 0;
    hi[t].x = fmaxf(p[0], fmaxf(p[3], p[6])); hi[t].y = fmaxf(p[1], fmaxf(p[4], p[7])); hi[t].z = fmaxf(p[2], fmaxf(p[5], p[8])); hi[t].w = 0;
  }
}
static void append_tris(const float* v9, const std::vector<uint32_t>& leaf_prims, std::vector<F4>& tris) {
  for (uint32_t id : leaf_prims) {
    const float* p = v9 + 9 * (size_t)id;
    for (int k = 0; k < 3; k++) { F4 f; f.x = p[3 * k]; f.y = p[3 * k + 1]; f.z = p[3 * k + 2]; f.w = as_float(id); tris.push_back(f); }
  }
}

extern "C" {

void* emu_bvh_create_flat(const float* world_tris9, uint32_t n) {
  EmuBvh* B = new EmuBvh();
  std::vector<F4> lo, hi;
  tri_boxes(world_tris9, n, lo, hi);
  std::vector<uint32_t> leaf_prims;
  B->root = build_segment(lo, hi, 3, 0, B->nodes, leaf_prims);
  append_tris(world_tris9, leaf_prims, B->tris);
  return B;
}

// Two-level: meshes given as object-space soups; instances as (mesh id, xform16 row-major, inv12).
void* emu_bvh_create_two_level(uint32_t num_meshes, const float* const* mesh_tris9, const uint32_t* mesh_ntris,
                               uint32_t num_inst, const uint32_t* inst_mesh, const float* inst_xform16,
                               const float* inst_inv12) {
  EmuBvh* B = new EmuBvh();
  B->two_level = true;
  std::vector<uint32_t> roots(num_meshes);
  std::vector<F4> rlo(num_meshes), rhi(num_meshes);
  for (uint32_t m = 0; m < num_meshes; m++) {
    std::vector<F4> lo, hi;
    tri_boxes(mesh_tris9[m], mesh_ntris[m], lo, hi);
    std::vector<uint32_t> leaf_prims;
    roots[m] = build_segment(lo, hi, 3, (uint32_t)(B->tris.size() / 3), B->nodes, leaf_prims);
    append_tris(mesh_tris9[m], leaf_prims, B->tris);
    F4 a, b;
    a.x = a.y = a.z = 1e30f; b.x = b.y = b.z = -1e30f;
    for (uint32_t t = 0; t < mesh_ntris[m]; t++) {
      a.x = fminf(a.x, lo[t].x); a.y = fminf(a.y, lo[t].y); a.z = fminf(a.z, lo[t].z);
      b.x = fmaxf(b.x, hi[t].x); b.y = fmaxf(b.y, hi[t].y); b.z = fmaxf(b.z, hi[t].z);
    }
    rlo[m] = a; rhi[m] = b;
  }
  std::vector<F4> ilo(num_inst), ihi(num_inst);
  for (uint32_t i = 0; i < num_inst; i++) {
    const float* xf = inst_xform16 + 16 * (size_t)i;
    const F4 a = rlo[inst_mesh[i]], b = rhi[inst_mesh[i]];
    F4 lo, hi;
    lo.x = lo.y = lo.z = 1e30f; hi.x = hi.y = hi.z = -1e30f; lo.w = hi.w = 0;
    for (int c = 0; c < 8; c++) {
      V3 p = xf_point(xf, v3((c & 1) ? b.x : a.x, (c & 2) ? b.y : a.y, (c & 4) ? b.z : a.z));
      lo.x = fminf(lo.x, p.x); lo.y = fminf(lo.y, p.y); lo.z = fminf(lo.z, p.z);
      hi.x = fmaxf(hi.x, p.x); hi.y = fmaxf(hi.y, p.y); hi.z = fmaxf(hi.z, p.z);
    }
    const float px = 3.8e-6f * fmaxf(fabsf(lo.x), fabsf(hi.x)), py = 3.8e-6f * fmaxf(fabsf(lo.y), fabsf(hi.y)),
                pz = 3.8e-6f * fmaxf(fabsf(lo.z), fabsf(hi.z));
    lo.x -= px; lo.y -= py; lo.z -= pz; hi.x += px; hi.y += py; hi.z += pz;
    ilo[i] = lo; ihi[i] = hi;
  }
  std::vector<uint32_t> leaf_insts;
  B->root = build_segment(ilo, ihi, 1, 0, B->nodes, leaf_insts);
  for (uint32_t id : leaf_insts) {
    const float* inv = inst_inv12 + 12 * (size_t)id;
    for (int r = 0; r < 3; r++) { F4 f; f.x = inv[4 * r]; f.y = inv[4 * r + 1]; f.z = inv[4 * r + 2]; f.w = inv[4 * r + 3]; B->insts.push_back(f); }
    F4 f; f.x = as_float(roots[inst_mesh[id]]); f.y = as_float(id); f.z = 0; f.w = 0;
    B->insts.push_back(f);
    // bounding sphere about the centre of the (unpadded) world box of the transformed vertices
    {
      const float* xf = inst_xform16 + 16 * (size_t)id;
      const float* v9 = mesh_tris9[inst_mesh[id]];
      const uint32_t nv = 3 * mesh_ntris[inst_mesh[id]];
      V3 lo = v3(1e30f, 1e30f, 1e30f), hi = v3(-1e30f, -1e30f, -1e30f);
      for (uint32_t q = 0; q < nv; q++) {
        V3 w = xf_point(xf, v3(v9[3 * q], v9[3 * q + 1], v9[3 * q + 2]));
        lo = v3(fminf(lo.x, w.x), fminf(lo.y, w.y), fminf(lo.z, w.z)); hi = v3(fmaxf(hi.x, w.x), fmaxf(hi.y, w.y), fmaxf(hi.z, w.z));
      }
      F4 s; s.x = 0.5f * (lo.x + hi.x); s.y = 0.5f * (lo.y + hi.y); s.z = 0.5f * (lo.z + hi.z);
      float r2 = 0.f;
      for (uint32_t q = 0; q < nv; q++) {
        V3 w = xf_point(xf, v3(v9[3 * q], v9[3 * q + 1], v9[3 * q + 2]));
        r2 = fmaxf(r2, (w.x - s.x) * (w.x - s.x) + (w.y - s.y) * (w.y - s.y) + (w.z - s.z) * (w.z - s.z));
      }
      s.w = r2 * 1.0002f + 1e-30f;
      B->insts.push_back(s);
    }
  }
  return B;
}

void emu_set_max_leaf(uint32_t m) { g_max_leaf_override = m; }
void emu_set_node_test_fp32(int on) { g_node_test_fp32 = on; }
void emu_set_no_oversized_split(int on) { g_no_oversized_split = on; }
void emu_bvh_destroy(void* h) { delete static_cast<EmuBvh*>(h); }
uint64_t emu_bvh_num_nodes(void* h) { return static_cast<EmuBvh*>(h)->nodes.size(); }

// returns total node visits; hit[i] = 1 if occluded
uint64_t emu_trace(void* h, const float* rays, uint64_t n, uint8_t* hit, uint64_t* tri_tests) {
  EmuBvh* B = static_cast<EmuBvh*>(h);
  BvhView v;
  v.nodes = reinterpret_cast<const U4*>(B->nodes.data());
  v.tris = B->tris.data();
  v.insts = B->insts.data();
  v.root = B->root;
  v.two_level = B->two_level ? 1u : 0u;
  uint64_t nodes = 0, tris = 0;
  for (uint64_t i = 0; i < n; i++) {
    const float* p = rays + 8 * i;
    U2 stack[kStackSize];
    TraceCounters c = {0, 0, 0};
    hit[i] = (g_node_test_fp32 ? trace_any_hit<true, false>(v, v3(p[0], p[1], p[2]), v3(p[4], p[5], p[6]), p[3], p[7], stack, &c)
                               : trace_any_hit<true, true>(v, v3(p[0], p[1], p[2]), v3(p[4], p[5], p[6]), p[3], p[7], stack, &c)) ? 1 : 0;
    nodes += c.nodes; tris += c.tris;
  }
  if (tri_tests) *tri_tests = tris;
  return nodes;
}

// select-based vs axis-specialised Woop test must agree bit for bit
int emu_woop_variants_agree(const float* rays, uint64_t n, const float* tris9, uint64_t nt) {
  int bad = 0;
  for (uint64_t i = 0; i < n; i++) {
    const float* p = rays + 8 * i;
    const V3 o = v3(p[0], p[1], p[2]), d = v3(p[4], p[5], p[6]);
    const Shear sh = make_shear(d);
    for (uint64_t t = 0; t < nt; t++) {
      const float* q = tris9 + 9 * t;
      const V3 a = v3(q[0], q[1], q[2]), b = v3(q[3], q[4], q[5]), c = v3(q[6], q[7], q[8]);
      if (woop_hit(o, sh, p[3], p[7], a, b, c) != woop_hit_sel(o, sh, p[3], p[7], a, b, c)) bad++;
    }
  }
  return bad;
}

// ---- the two box tests in isolation (property tests: conservative against an exact fp64 slab test) ----
// node80: one Node8 record; rays: n x (o.xyz, tmin, d.xyz, tmax); variant 0 = packed fp16, 1 = fp32.
// masks[i] = traversal mask of ray i ([31:24] internal slots | [23:0] leaf primitive bits).
void emu_node_test(const void* node80, const float* rays, uint64_t n, int variant, uint32_t* masks, uint8_t* wide) {
  const U4* nodes = reinterpret_cast<const U4*>(node80);
  const NodeConsts nc = make_node_consts();
  for (uint64_t i = 0; i < n; i++) {
    const float* p = rays + 8 * i;
    RayState r;
    ray_setup(r, v3(p[0], p[1], p[2]), v3(p[4], p[5], p[6]), p[3], p[7]);
    uint32_t cb, pb, im;
    wide[i] = r.wide ? 1 : 0;
    masks[i] = variant == 0 ? intersect_node8_h2(nodes, 0, r, &cb, &pb, &im) : intersect_node8<true>(nodes, 0, r, nc, &cb, &pb, &im);
  }
}
// fp16 emulation shims against an independent implementation (numpy float16) in the tests
uint32_t emu_h2_pack_sat(float hi, float lo) { return h2_pack_sat(hi, lo); }
uint32_t emu_h2_fma(uint32_t a, uint32_t b, uint32_t c) { return h2_fma(a, b, c); }
uint32_t emu_h2_min(uint32_t a, uint32_t b) { return h2_min(a, b); }
uint32_t emu_h2_lane_sum(uint32_t a) { return h2_lane_sum(a); }

// ---- math parity hooks ----
uint32_t emu_tea(uint32_t rounds, uint32_t v0, uint32_t v1) {
  switch (rounds) { case 2: return tea<2>(v0, v1); case 4: return tea<4>(v0, v1); case 16: return tea<16>(v0, v1); default: return 0; }
}
float emu_halton(uint32_t i, uint32_t b) { return halton(i, b); }
void emu_sincos2pi(float u, float* c, float* s) { sincos2pi(u, c, s); }
int emu_sqrt_rays(int r) { return sqrt_rays(r); }

// rays for samples [begin,end) exactly as the fused kernel generates them
void emu_generate_rays(const float* pos, const float* nrm, const float* fnrm, uint64_t begin, uint64_t end,
                       int rays_per_sample, float offset, float maxdist, float* out) {
  const int q = sqrt_rays(rays_per_sample);
  for (uint64_t g = begin; g < end; g++) {
    V3 p = v3(pos[3 * g], pos[3 * g + 1], pos[3 * g + 2]), n = v3(nrm[3 * g], nrm[3 * g + 1], nrm[3 * g + 2]),
       fn = v3(fnrm[3 * g], fnrm[3 * g + 1], fnrm[3 * g + 2]);
    Onb onb = make_onb(n);
    V3 o = ao_ray_origin(p, n, offset);
    for (int pass = 0; pass < q * q; pass++) {
      V3 d = ao_ray_dir((uint32_t)g, (uint32_t)pass, q, n, fn, onb);
      float* r = out + ((g - begin) * (uint64_t)(q * q) + (uint64_t)pass) * 8;
      r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = 0.0f; r[4] = d.x; r[5] = d.y; r[6] = d.z; r[7] = maxdist;
    }
  }
}
}
