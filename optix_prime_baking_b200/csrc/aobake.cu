// aobake.cu — host orchestrator + C-ABI of libaobake.so (see include/aobake.h).
//
// Everything that touches geometry, samples, rays or AO runs in the CUDA kernels of
// aob_kernels.cuh; this file owns device memory, launch order, CUDA events and the
// status/last-error convention.  There is no CPU fallback: aobake_create fails without a
// CUDA device, and every compute entry point requires a live context.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "aobake.h"
#include "aob_kernels.cuh"

using namespace aob;

namespace {

thread_local std::string g_create_error;
// Stream-ordered allocations (cudaMallocAsync) on the stream of the API call in progress: the
// BVH build and the filters allocate dozens of scratch buffers, and the driver's pool makes
// those microseconds instead of the milliseconds of cudaMalloc/cudaFree.
thread_local cudaStream_t g_alloc_stream = nullptr;

// ---- NCCL, bound at run time: no link-time dependency, and inside a process that already loaded
// an NCCL (e.g. torch's bundled one) dlopen returns that same library ----
struct NcclId { char internal[128]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*CommAbort)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
  template <typename F>
  bool sym(F& f, const char* name) {
    f = reinterpret_cast<F>(dlsym(handle, name));
    if (!f) error = std::string("libnccl lacks the symbol ") + name;
    return f != nullptr;
  }
  bool load() {
    if (handle) return true;
    // AOBAKE_NCCL_LIB names the library to bind instead of the default sonames (deployment override;
    // the tests use it to force a load failure)
    const char* env = getenv("AOBAKE_NCCL_LIB");
    const char* names[] = {env && *env ? env : "libnccl.so.2", env && *env ? env : "libnccl.so"};
    std::string tried;
    for (const char* n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
      const char* e = dlerror();   // dlerror() clears the error state: read it exactly once per failure
      tried = std::string("dlopen(") + n + "): " + (e ? e : "not found");
    }
    if (!handle) { error = tried; return false; }
    if (!sym(GetUniqueId, "ncclGetUniqueId") || !sym(CommInitRank, "ncclCommInitRank") || !sym(CommDestroy, "ncclCommDestroy") ||
        !sym(CommAbort, "ncclCommAbort") || !sym(AllReduce, "ncclAllReduce") || !sym(AllGather, "ncclAllGather") ||
        !sym(Broadcast, "ncclBroadcast") || !sym(GroupStart, "ncclGroupStart") || !sym(GroupEnd, "ncclGroupEnd") ||
        !sym(GetErrorString, "ncclGetErrorString")) {
      dlclose(handle);
      handle = nullptr;
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat = 7, kNcclDouble = 8, kNcclInt32 = 2, kNcclUint8 = 1, kNcclSum = 0, kNcclMax = 2;   // ncclFloat32/64, ncclInt32, ncclUint8; ncclSum, ncclMax

template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), g_alloc_stream);
  }
  void release() {
    if (p) cudaFreeAsync(p, g_alloc_stream);
    p = nullptr;
    n = 0;
  }
  void swap(DBuf& o) { std::swap(p, o.p); std::swap(n, o.n); }
};

struct DeviceMesh {
  uint64_t nV = 0, nT = 0;
  DBuf<float> verts;    // packed xyz
  DBuf<float> normals;  // packed xyz or empty
  DBuf<uint32_t> tris;
};
struct HostInstance {
  float xf[12];
  float inv[12];
  uint32_t mesh;
  uint64_t storage_id;
};

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// inverse of the affine part in fp64 by cofactors, rounded to fp32 once (BASELINE.md §4.2;
// the reference's Matrix4x4::inverse in bake_sample.cpp).  Host code of this file is built with
// -ffp-contract=off so the result is reproducible.
void affine_inverse(const float* m, float* inv) {
  const double a00 = m[0], a01 = m[1], a02 = m[2], t0 = m[3];
  const double a10 = m[4], a11 = m[5], a12 = m[6], t1 = m[7];
  const double a20 = m[8], a21 = m[9], a22 = m[10], t2 = m[11];
  const double c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
  const double c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
  const double det = (a00 * c00 + a01 * c10) + a02 * c20;
  const double i[9] = {c00 / det, c01 / det, c02 / det, c10 / det, c11 / det, c12 / det, c20 / det, c21 / det, c22 / det};
  for (int r = 0; r < 3; r++) {
    inv[4 * r + 0] = (float)i[3 * r + 0];
    inv[4 * r + 1] = (float)i[3 * r + 1];
    inv[4 * r + 2] = (float)i[3 * r + 2];
    inv[4 * r + 3] = (float)(-((i[3 * r + 0] * t0 + i[3 * r + 1] * t1) + i[3 * r + 2] * t2));
  }
}

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

struct AoBake {
  AoBakeParams params;
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int sm_count = 148;

  // scene
  bool have_scene = false;
  bool have_bvh = false;                   // false after aobake_set_scene_geometry: nothing to trace against
  std::vector<DeviceMesh> meshes;          // scene meshes (blockers are only needed for the BVH)
  std::vector<HostInstance> insts;         // scene instances
  std::vector<uint64_t> inst_num_verts;
  DBuf<InstDesc> d_inst;                   // sampling descriptors (refreshed per sample call)
  DBuf<double> d_tri_area, d_bsum, d_inst_area;
  std::vector<double> inst_area;           // host copy
  uint64_t total_tris = 0, total_blocks = 0;
  bool areas_ready = false;

  // BVH
  DBuf<Node8> d_nodes;
  DBuf<F4> d_tris;
  DBuf<F4> d_insts;
  uint32_t root = 0;
  bool two_level = false;
  float scene_diag = 0.f;   // diagonal of the world box of everything in the BVH
  float main_diag = 0.f;    // diagonal of the box of the top-level tree without its oversized primitives (== scene_diag if none)
  AoStats stats{};

  // samples + AO
  uint64_t num_samples = 0;
  DBuf<float> d_pos, d_nrm, d_fnrm;
  DBuf<AoSampleInfo> d_info;
  std::vector<uint64_t> per_instance;
  DBuf<float> d_ao;
  DBuf<uint32_t> d_hits;
  DBuf<unsigned long long> d_stats;
  DBuf<unsigned long long> d_counter;
  DBuf<U2> d_deferred;               // rays the fp16 node test set aside (k_ao_deferred)
  DBuf<uint32_t> d_deferred_count;
  bool have_ao = false;
  bool have_infos = false;           // sample_infos (tri_idx, bary, dA) are resident — needed by the vertex maps
  bool unit_normals = true;          // every resident shading normal is unit length (the fp16 node test relies on it)
  bool samples_sharded = false;      // only this rank's super-blocks of pos/nrm/fnrm are resident (set_samples_distributed)

  AoTimings timings{};

  // native multi-GPU exchange
  void* nccl_comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  DBuf<int> d_comm_status;           // one int: the status the ranks agree on before a collective

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return ctx->fail(AOBAKE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define CKL() CK(cudaGetLastError())

namespace {

constexpr uint32_t kDefaultBlockSamples = 16384;   // super-block of the interleaved multi-GPU partition (config 3 at 8 ranks: 76 or 77 per rank)
#ifndef AOB_ITEMS_PER_WARP
#define AOB_ITEMS_PER_WARP 32
#endif
constexpr uint32_t kItemsPerWarp = AOB_ITEMS_PER_WARP;
constexpr uint32_t kTriBatchFlat = 8 | (6 << 8), kTriBatchFlatPacket = 16 | (6 << 8), kTriBatchTwoLevel = 12 | (6 << 8);   // default AoBakeParams::tri_batch (sweep: profiles/r2/sweep_tri_batch.log)

// Makes a local status collective: every rank of the communicator calls this once at the same point
// with its own status; all of them return non-zero if any rank failed, so that no rank enters the
// following collective alone (a rank blocked in ncclAllReduce would otherwise hang the job).
int comm_agree(AoBake* ctx, int rc) {
  if (ctx->comm_size <= 1 || !ctx->nccl_comm) return rc;
  cudaStream_t st = ctx->stream;
  int mine = rc, agreed = rc;
  bool ok = cudaMemcpyAsync(ctx->d_comm_status.p, &mine, sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess;
  ok = ok && g_nccl.AllReduce(ctx->d_comm_status.p, ctx->d_comm_status.p, 1, kNcclInt32, kNcclMax, ctx->nccl_comm, st) == 0;
  ok = ok && cudaMemcpyAsync(&agreed, ctx->d_comm_status.p, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(st) == cudaSuccess;
  if (rc) return rc;   // keep the local error text
  if (!ok) return ctx->fail(AOBAKE_ERR_COMM, "status exchange between the ranks failed");
  if (agreed) return ctx->fail(AOBAKE_ERR_COMM, "another rank failed with status %d before the collective", agreed);
  return AOBAKE_OK;
}

// Host -> device copy of `bytes` bytes.  shard = true (multi-GPU set_scene): every rank holds the same
// host array, so rank r copies only slice r over its PCIe link and one in-place ncclAllGather over
// NVLink completes the array on every rank; `dst` must have room for padded_bytes(bytes, nranks).
size_t shard_chunk(size_t bytes, int nranks) {
  const size_t c = (bytes + (size_t)nranks - 1) / (size_t)nranks;
  return (c + 255) & ~(size_t)255;
}
size_t padded_bytes(size_t bytes, int nranks) { return nranks > 1 ? shard_chunk(bytes, nranks) * (size_t)nranks : bytes; }
int upload_bytes(AoBake* ctx, void* dst, const void* src, size_t bytes, bool shard) {
  if (!bytes) return AOBAKE_OK;
  cudaStream_t st = ctx->stream;
  if (!shard || ctx->comm_size <= 1) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return AOBAKE_OK;
  }
  const size_t chunk = shard_chunk(bytes, ctx->comm_size);
  const size_t b = std::min(bytes, chunk * (size_t)ctx->comm_rank), e = std::min(bytes, b + chunk);
  if (e > b) CK(cudaMemcpyAsync((char*)dst + b, (const char*)src + b, e - b, cudaMemcpyHostToDevice, st));
  const int nrc = g_nccl.AllGather((char*)dst + chunk * (size_t)ctx->comm_rank, dst, chunk, kNcclUint8, ctx->nccl_comm, st);
  if (nrc != 0) return ctx->fail(AOBAKE_ERR_COMM, "ncclAllGather: %s", g_nccl.GetErrorString(nrc));
  return AOBAKE_OK;
}

struct Segment {
  uint32_t root = 0;
  uint32_t node_count = 0;
  int levels = 1;  // depth of the 8-wide tree (collapse rounds) — bounds the traversal stack
  float box[6] = {0, 0, 0, 0, 0, 0};
  float main_box[6] = {0, 0, 0, 0, 0, 0};   // box of the ordinary primitives (== box when nothing was split off)
};

// Builds one 8-wide BVH over n primitive boxes into d_nodes[node_offset...]; d_leaf_prims[n]
// receives the leaf order (relative primitive ids).
int build_segment(AoBake* ctx, const F4* d_plo, const F4* d_phi, uint32_t n, uint32_t max_leaf, uint32_t node_offset,
                  uint32_t prim_offset, Node8* d_nodes, uint32_t* d_leaf_prims, Segment* out) {
  cudaStream_t st = ctx->stream;
  out->root = node_offset;
  if (n == 0) {
    k_empty_node<<<1, 1, 0, st>>>(d_nodes + node_offset);
    CKL();
    out->node_count = 1;
    return AOBAKE_OK;
  }
  DBuf<int> d_b;  // 6 centroid + 6 box bounds
  DBuf<float> d_box;
  DBuf<uint64_t> keys, keys_s;
  DBuf<uint32_t> vals, vals_s, left, right, first, last, pint, pleaf, flags, wide2bin, counters, count;
  DBuf<F4> ilo, ihi;
  const uint32_t ni = n > 1 ? n - 1 : 1;
  CK(d_b.alloc(12)); CK(d_box.alloc(6));
  CK(keys.alloc(n)); CK(keys_s.alloc(n)); CK(vals.alloc(n)); CK(vals_s.alloc(n));
  CK(left.alloc(ni)); CK(right.alloc(ni)); CK(first.alloc(ni)); CK(last.alloc(ni)); CK(pint.alloc(ni)); CK(pleaf.alloc(n));
  CK(flags.alloc(ni)); CK(count.alloc(ni)); CK(wide2bin.alloc(n)); CK(counters.alloc(2)); CK(ilo.alloc(ni)); CK(ihi.alloc(ni));
  k_init_bounds<<<1, 32, 0, st>>>(d_b.p);
  k_init_bounds<<<1, 32, 0, st>>>(d_b.p + 6);
  k_bounds<<<std::min<unsigned>(grid_for(n, 256), 148u * 8u), 256, 0, st>>>(d_plo, d_phi, n, d_b.p, d_b.p + 6);
  k_decode_bounds<<<1, 32, 0, st>>>(d_b.p + 6, d_box.p);
  // oversized primitives stay out of the tree (see k_flag_big): at most 7 leaf slots of the extra root
  DBuf<uint8_t> big;
  DBuf<uint32_t> big_count;
  DBuf<int> d_bs;   // centroid + box bounds of the ordinary primitives
  CK(big.alloc(n)); CK(big_count.alloc(1)); CK(d_bs.alloc(12));
  CK(cudaMemsetAsync(big_count.p, 0, sizeof(uint32_t), st));
  k_flag_big<<<grid_for(n, 256), 256, 0, st>>>(d_plo, d_phi, n, d_b.p + 6, big.p, big_count.p);
  CKL();
  uint32_t n_big = 0;
  CK(cudaMemcpyAsync(&n_big, big_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const uint32_t per_slot = std::max(1u, std::min(max_leaf, 3u));
  const bool split = ctx->params.no_oversized_split == 0 && n_big > 0 && n_big < n && n_big <= 7u * per_slot;
  const uint32_t n_small = split ? n - n_big : n;
  if (split) {
    k_init_bounds<<<1, 32, 0, st>>>(d_bs.p);
    k_init_bounds<<<1, 32, 0, st>>>(d_bs.p + 6);
    k_bounds_small<<<grid_for(n, 256), 256, 0, st>>>(d_plo, d_phi, big.p, n, d_bs.p, d_bs.p + 6);
    k_morton_small<<<grid_for(n, 256), 256, 0, st>>>(d_plo, d_phi, big.p, n, d_bs.p, keys.p, vals.p);
  } else {
    k_morton<<<grid_for(n, 256), 256, 0, st>>>(d_plo, d_phi, n, d_b.p, keys.p, vals.p);
  }
  CKL();
  {
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)n, 0, 64, st));
    DBuf<uint8_t> tmp;
    CK(tmp.alloc(tmp_bytes));
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (int)n, 0, 64, st));
    CK(cudaStreamSynchronize(st));
  }
  const uint32_t n_all = n;
  n = n_small;   // the tree below is built over the ordinary primitives only (all of them when nothing was split off)
  Lbvh L;
  L.keys = keys_s.p; L.prim = vals_s.p; L.plo = d_plo; L.phi = d_phi;
  L.left = left.p; L.right = right.p; L.first = first.p; L.last = last.p;
  L.parent_int = pint.p; L.parent_leaf = pleaf.p; L.ilo = ilo.p; L.ihi = ihi.p; L.flags = flags.p; L.count = count.p; L.n = n;
  if (n >= 2) {
    CK(cudaMemsetAsync(flags.p, 0, ni * sizeof(uint32_t), st));
    k_hierarchy<<<grid_for(n - 1, 256), 256, 0, st>>>(L);
    k_refit<<<grid_for(n, 256), 256, 0, st>>>(L);
    CKL();
  }
  // level-synchronous collapse into 8-wide nodes
  CollapseArgs A;
  A.L = L; A.nodes = d_nodes; A.wide2bin = wide2bin.p; A.leaf_prims = d_leaf_prims;
  A.node_count = counters.p; A.prim_count = counters.p + 1; A.node_offset = node_offset; A.prim_offset = prim_offset;
  A.max_leaf = max_leaf;
  k_set_u32<<<1, 1, 0, st>>>(counters.p, 1u);
  k_set_u32<<<1, 1, 0, st>>>(counters.p + 1, 0u);
  k_set_u32<<<1, 1, 0, st>>>(wide2bin.p, n == 1 ? (0u | kLeafBit) : 0u);
  uint32_t lb = 0, le = 1;
  uint32_t hc[2] = {1, 0};
  int levels = 0;
  while (lb < le) {
    k_collapse<<<grid_for(le - lb, 128), 128, 0, st>>>(A, lb, le);
    CKL();
    CK(cudaMemcpyAsync(hc, counters.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    lb = le;
    le = hc[0];
    if (++levels > 4096) return ctx->fail(AOBAKE_ERR_CUDA, "BVH collapse did not terminate");
  }
  if (hc[1] != n) return ctx->fail(AOBAKE_ERR_CUDA, "BVH collapse emitted %u of %u primitives", hc[1], n);
  out->node_count = hc[0];
  out->levels = levels;
  if (split) {
    // one extra root on top: [tree over the ordinary primitives | oversized primitives as leaf slots]
    k_super_root<<<1, 32, 0, st>>>(d_nodes, node_offset + hc[0], node_offset, d_b.p + 6, d_bs.p + 6, d_plo, d_phi, vals_s.p, n_small, n_all - n_small,
                                   per_slot, prim_offset, d_leaf_prims);
    CKL();
    out->root = node_offset + hc[0];
    out->node_count = hc[0] + 1;
    out->levels = levels + 1;
  }
  CK(cudaMemcpyAsync(out->box, d_box.p, sizeof(out->box), cudaMemcpyDeviceToHost, st));
  if (split) {
    k_decode_bounds<<<1, 32, 0, st>>>(d_bs.p + 6, d_box.p);
    CKL();
    CK(cudaStreamSynchronize(st));   // out->box has arrived before d_box is reused
    CK(cudaMemcpyAsync(out->main_box, d_box.p, sizeof(out->main_box), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (!split) memcpy(out->main_box, out->box, sizeof(out->box));
  return AOBAKE_OK;
}

// Device arrays of one mesh: alloc_mesh sizes them (padded for the sharded upload), copy_mesh fills them.
int alloc_mesh(AoBake* ctx, const AoMesh& m, DeviceMesh& dm, bool want_normals, bool shard) {
  dm.nV = m.num_vertices;
  dm.nT = m.num_triangles;
  if (m.num_vertices > 0xfffffff0ull || m.num_triangles > 0x7ffffff0ull)
    return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "mesh too large for 32-bit indices");
  if ((m.num_vertices && !m.vertices) || (m.num_triangles && !m.tri_vertex_indices))
    return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "mesh has null vertex or index pointer");
  const int nr = shard ? ctx->comm_size : 1;
  CK(dm.verts.alloc(padded_bytes(12 * dm.nV, nr) / sizeof(float)));
  CK(dm.tris.alloc(padded_bytes(12 * dm.nT, nr) / sizeof(uint32_t)));
  if (want_normals && m.normals) CK(dm.normals.alloc(padded_bytes(12 * dm.nV, nr) / sizeof(float)));
  return AOBAKE_OK;
}
int copy_mesh(AoBake* ctx, const AoMesh& m, DeviceMesh& dm, bool shard) {
  const uint32_t vs = m.vertex_stride_bytes ? m.vertex_stride_bytes : 12u;
  const uint32_t ns = m.normal_stride_bytes ? m.normal_stride_bytes : 12u;
  int rc;
  auto copy_strided = [&](const float* src, uint32_t stride, float* dst) -> int {
    if (dm.nV == 0) return AOBAKE_OK;
    if (stride == 12) return upload_bytes(ctx, dst, src, 12 * dm.nV, shard);
    CK(cudaMemcpy2DAsync(dst, 12, src, stride, 12, dm.nV, cudaMemcpyHostToDevice, ctx->stream));   // strided arrays: every rank copies all of it
    return AOBAKE_OK;
  };
  if ((rc = copy_strided(m.vertices, vs, dm.verts.p))) return rc;
  if (dm.normals.p && (rc = copy_strided(m.normals, ns, dm.normals.p))) return rc;
  if ((rc = upload_bytes(ctx, dm.tris.p, m.tri_vertex_indices, 12 * dm.nT, shard))) return rc;
  return AOBAKE_OK;
}

int check_scene(AoBake* ctx, const AoScene* s, const char* what) {
  if (!s) return AOBAKE_OK;
  if ((s->num_meshes && !s->meshes) || (s->num_instances && !s->instances))
    return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "%s: null meshes/instances pointer", what);
  for (uint64_t i = 0; i < s->num_instances; i++)
    if (s->instances[i].mesh_index >= s->num_meshes)
      return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "%s: instance %llu references mesh %u of %llu", what,
                       (unsigned long long)i, s->instances[i].mesh_index, (unsigned long long)s->num_meshes);
  return AOBAKE_OK;
}

int ensure_areas(AoBake* ctx) {
  if (ctx->areas_ready) return AOBAKE_OK;
  cudaStream_t st = ctx->stream;
  const uint32_t ni = (uint32_t)ctx->insts.size();
  std::vector<InstDesc> h(ni);
  uint64_t e = 0, b = 0;
  for (uint32_t i = 0; i < ni; i++) {
    const HostInstance& I = ctx->insts[i];
    const DeviceMesh& m = ctx->meshes[I.mesh];
    memcpy(h[i].xf, I.xf, sizeof(I.xf));
    memcpy(h[i].inv, I.inv, sizeof(I.inv));
    h[i].verts = m.verts.p; h[i].normals = m.normals.p; h[i].tris = m.tris.p;
    h[i].tri_begin = e; h[i].num_tris = m.nT; h[i].block_begin = b;
    h[i].sample_begin = 0; h[i].num_samples = 0; h[i].seed = i; h[i].pad = 0;
    e += m.nT;
    b += (m.nT + 1023) / 1024;
  }
  ctx->total_tris = e;
  ctx->total_blocks = b;
  CK(ctx->d_inst.alloc(std::max<uint32_t>(ni, 1)));
  CK(ctx->d_tri_area.alloc(std::max<uint64_t>(e, 1)));
  CK(ctx->d_bsum.alloc(std::max<uint64_t>(b, 1)));
  CK(ctx->d_inst_area.alloc(std::max<uint32_t>(ni, 1)));
  ctx->inst_area.assign(ni, 0.0);
  if (ni) {
    CK(cudaMemcpyAsync(ctx->d_inst.p, h.data(), ni * sizeof(InstDesc), cudaMemcpyHostToDevice, st));
    if (e) k_tri_areas<<<grid_for(e, 256), 256, 0, st>>>(ctx->d_inst.p, ni, e, ctx->d_tri_area.p);
    if (b) k_block_sums<<<grid_for(b, 128), 128, 0, st>>>(ctx->d_inst.p, ni, b, ctx->d_tri_area.p, ctx->d_bsum.p);
    k_inst_totals<<<grid_for(ni, 128), 128, 0, st>>>(ctx->d_inst.p, ni, ctx->d_bsum.p, ctx->d_inst_area.p);
    CKL();
    CK(cudaMemcpyAsync(ctx->inst_area.data(), ctx->d_inst_area.p, ni * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  ctx->areas_ready = true;
  return AOBAKE_OK;
}

int alloc_samples(AoBake* ctx, uint64_t n) {
  ctx->num_samples = n;
  ctx->have_ao = false;
  ctx->have_infos = false;
  ctx->unit_normals = true;
  ctx->samples_sharded = false;
  const uint64_t m = std::max<uint64_t>(n, 1);
  CK(ctx->d_pos.alloc(3 * m)); CK(ctx->d_nrm.alloc(3 * m)); CK(ctx->d_fnrm.alloc(3 * m)); CK(ctx->d_info.alloc(m));
  CK(ctx->d_ao.alloc(m)); CK(ctx->d_hits.alloc(m));
  CK(cudaMemsetAsync(ctx->d_ao.p, 0, m * sizeof(float), ctx->stream));
  CK(cudaMemsetAsync(ctx->d_hits.p, 0, m * sizeof(uint32_t), ctx->stream));
  return AOBAKE_OK;
}

BvhView bvh_view(const AoBake* ctx) {
  BvhView v;
  v.nodes = reinterpret_cast<const U4*>(ctx->d_nodes.p);
  v.tris = ctx->d_tris.p;
  v.insts = ctx->d_insts.p;
  v.root = ctx->root;
  v.two_level = ctx->two_level ? 1u : 0u;
  return v;
}

struct ScopedTimer {
  AoBake* c;
  double t0;
  explicit ScopedTimer(AoBake* ctx) : c(ctx), t0(now_ms()) { g_alloc_stream = ctx->stream; }
  ~ScopedTimer() { c->timings.host_total_ms = (float)(now_ms() - t0); }
};

#ifndef AOB_RAY_ORDER_DEFAULT
#define AOB_RAY_ORDER_DEFAULT 2
#endif
using AoKernelT = void (*)(BvhView, SampleView, uint64_t, uint32_t, int, float, float, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t,
                           uint32_t, uint32_t*, unsigned long long*, unsigned long long*, DeferredRays);
template <bool TWO_LEVEL, bool CLAMP, bool H2, bool PACKET>
AoKernelT pick_ao_kernel_s(bool stats) {
  return stats ? (AoKernelT)k_ao_persistent<true, TWO_LEVEL, CLAMP, H2, PACKET> : (AoKernelT)k_ao_persistent<false, TWO_LEVEL, CLAMP, H2, PACKET>;
}
template <bool PACKET>
AoKernelT pick_ao_kernel_p(bool stats, bool two_level, bool clamp, bool h2) {
  // the packed-fp16 node test exists for flattened scenes only
  if (h2) return clamp ? pick_ao_kernel_s<false, true, true, PACKET>(stats) : pick_ao_kernel_s<false, false, true, PACKET>(stats);
  if (two_level) return clamp ? pick_ao_kernel_s<true, true, false, PACKET>(stats) : pick_ao_kernel_s<true, false, false, PACKET>(stats);
  return clamp ? pick_ao_kernel_s<false, true, false, PACKET>(stats) : pick_ao_kernel_s<false, false, false, PACKET>(stats);
}
AoKernelT pick_ao_kernel(bool stats, bool two_level, bool clamp, bool h2, bool packet) {
  return packet ? pick_ao_kernel_p<true>(stats, two_level, clamp, h2) : pick_ao_kernel_p<false>(stats, two_level, clamp, h2);
}

}  // namespace

// =========================================================================================
// C-ABI
// =========================================================================================
extern "C" {

int aobake_default_params(AoBakeParams* p) {
  if (!p) return AOBAKE_ERR_INVALID_ARGUMENT;
  memset(p, 0, sizeof(*p));
  p->device = 0;
  p->instancing_mode = AOBAKE_INSTANCING_AUTO;
  p->cg_max_iterations = 20000;
  p->cg_tolerance = 1e-6f;
  p->trace_kernel = 0;
  p->collect_stats = 0;
  p->refill_below = 0;
  p->leaf_tris = 0;
  p->node_test = 0;
  p->deferred_capacity = 0;
  p->tri_batch = 0;
  p->no_oversized_split = 0;
  p->ls_energy = 0;
  p->ls_matrix_free = 0;
  p->ray_order = 0;
  return AOBAKE_OK;
}

const char* aobake_last_error(const AoBake* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int aobake_create(const AoBakeParams* params, AoBake** out) {
  if (!out) return AOBAKE_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  AoBakeParams p;
  aobake_default_params(&p);
  if (params) p = *params;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libaobake has no CPU fallback)";
    return AOBAKE_ERR_NO_DEVICE;
  }
  if (p.device < 0 || p.device >= count) {
    g_create_error = "device ordinal out of range";
    return AOBAKE_ERR_INVALID_ARGUMENT;
  }
  if (p.ray_order < 0 || p.ray_order > 2 || p.trace_kernel < 0 || p.trace_kernel > 2) {
    g_create_error = "ray_order and trace_kernel must be 0, 1 or 2";
    return AOBAKE_ERR_INVALID_ARGUMENT;
  }
  if ((e = cudaSetDevice(p.device)) != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return AOBAKE_ERR_CUDA;
  }
  AoBake* ctx = new AoBake();
  ctx->params = p;
  ctx->device = p.device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, p.device);
  {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, p.device) == cudaSuccess) {
      unsigned long long keep = ~0ull;  // keep freed scratch in the pool between calls
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess || ((g_alloc_stream = ctx->own_stream), false) ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      ctx->d_stats.alloc(4) != cudaSuccess || ctx->d_counter.alloc(1) != cudaSuccess || ctx->d_deferred_count.alloc(1) != cudaSuccess) {
    g_create_error = std::string("context setup: ") + cudaGetErrorString(cudaGetLastError());
    delete ctx;
    return AOBAKE_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return AOBAKE_OK;
}

void aobake_destroy(AoBake* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->nccl_comm && g_nccl.handle) { g_nccl.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  g_alloc_stream = ctx->own_stream;
  cudaStream_t own = ctx->own_stream;
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  delete ctx;  // frees every buffer stream-ordered on `own`
  if (own) { cudaStreamSynchronize(own); cudaStreamDestroy(own); }
  g_alloc_stream = nullptr;
}

int aobake_set_stream(AoBake* ctx, void* s) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = s ? reinterpret_cast<cudaStream_t>(s) : ctx->own_stream;
  CK(cudaStreamSynchronize(ctx->stream));
  g_alloc_stream = ctx->stream;
  return AOBAKE_OK;
}
int aobake_synchronize(AoBake* ctx) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  return AOBAKE_OK;
}

static int set_scene_impl(AoBake* ctx, const AoScene* scene, const AoScene* blockers, bool shard, bool build_bvh = true) {
  if (!ctx || !scene) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (shard && ctx->comm_size > 1 && !ctx->nccl_comm) return ctx->fail(AOBAKE_ERR_STATE, "aobake_comm_init has not been called");
  shard = shard && ctx->comm_size > 1;
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  if (blockers && blockers->num_instances == 0) blockers = nullptr;
  int rc;
  if ((rc = check_scene(ctx, scene, "scene")) || (rc = check_scene(ctx, blockers, "blockers"))) return rc;
  ctx->have_scene = false;
  ctx->have_bvh = false;
  ctx->areas_ready = false;
  ctx->num_samples = 0;
  ctx->have_ao = false;
  ctx->per_instance.clear();
  ctx->meshes.clear();
  ctx->insts.clear();
  ctx->inst_num_verts.clear();

  // ---- upload ----
  CK(cudaEventRecord(ctx->ev0, st));
  ctx->meshes.resize(scene->num_meshes);
  std::vector<DeviceMesh> bmeshes(blockers ? blockers->num_meshes : 0);
  rc = AOBAKE_OK;
  for (uint64_t m = 0; m < scene->num_meshes && !rc; m++) rc = alloc_mesh(ctx, scene->meshes[m], ctx->meshes[m], true, shard);
  for (size_t m = 0; m < bmeshes.size() && !rc; m++) rc = alloc_mesh(ctx, blockers->meshes[m], bmeshes[m], false, shard);
  // sharded upload: no rank starts the all-gathers unless every rank could allocate
  if (shard) rc = comm_agree(ctx, rc);
  if (rc) return rc;
  for (uint64_t m = 0; m < scene->num_meshes; m++)
    if ((rc = copy_mesh(ctx, scene->meshes[m], ctx->meshes[m], shard))) return rc;
  for (size_t m = 0; m < bmeshes.size(); m++)
    if ((rc = copy_mesh(ctx, blockers->meshes[m], bmeshes[m], shard))) return rc;
  std::vector<HostInstance> binsts;
  auto add_insts = [&](const AoScene* s, std::vector<HostInstance>& dst) {
    for (uint64_t i = 0; i < s->num_instances; i++) {
      HostInstance h;
      memcpy(h.xf, s->instances[i].xform, sizeof(h.xf));
      affine_inverse(h.xf, h.inv);
      h.mesh = s->instances[i].mesh_index;
      h.storage_id = s->instances[i].storage_identifier;
      dst.push_back(h);
    }
  };
  add_insts(scene, ctx->insts);
  if (blockers) add_insts(blockers, binsts);
  for (const HostInstance& I : ctx->insts) ctx->inst_num_verts.push_back(ctx->meshes[I.mesh].nV);
  for (const std::vector<HostInstance>* v : {&ctx->insts, &binsts})
    for (const HostInstance& I : *v)
      for (int k = 0; k < 12; k++)
        if (!std::isfinite(I.xf[k]) || !std::isfinite(I.inv[k]))
          return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "instance %zu of the %s has a singular or non-finite transform", (size_t)(&I - v->data()),
                           v == &ctx->insts ? "scene" : "blockers");
  // every triangle index must address a vertex of its mesh
  const size_t n_all_meshes = ctx->meshes.size() + bmeshes.size();
  DBuf<uint32_t> d_bad;
  std::vector<uint32_t> h_bad(std::max<size_t>(n_all_meshes, 1), 0u);
  CK(d_bad.alloc(h_bad.size()));
  CK(cudaMemsetAsync(d_bad.p, 0, h_bad.size() * sizeof(uint32_t), st));
  for (size_t m = 0; m < n_all_meshes; m++) {
    const DeviceMesh& dm = m < ctx->meshes.size() ? ctx->meshes[m] : bmeshes[m - ctx->meshes.size()];
    if (dm.nT) k_check_indices<<<grid_for(3 * dm.nT, 256), 256, 0, st>>>(dm.tris.p, 3 * dm.nT, (uint32_t)dm.nV, d_bad.p + m);
  }
  CKL();
  CK(cudaMemcpyAsync(h_bad.data(), d_bad.p, h_bad.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  for (size_t m = 0; m < n_all_meshes; m++)
    if (h_bad[m])
      return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "%s mesh %zu: a triangle index is >= num_vertices", m < ctx->meshes.size() ? "scene" : "blocker",
                       m < ctx->meshes.size() ? m : m - ctx->meshes.size());
  CK(cudaEventElapsedTime(&ctx->timings.upload_ms, ctx->ev0, ctx->ev1));

  if (!build_bvh) {
    // geometry only (aobake_set_scene_geometry): enough for distribute/sample_instances and the vertex maps
    ctx->d_nodes.release(); ctx->d_tris.release(); ctx->d_insts.release();
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->timings.bvh_build_ms = 0.f;
    ctx->have_bvh = false;
    ctx->have_scene = true;
    return AOBAKE_OK;
  }
  // ---- instancing mode (decision #12) ----
  bool two_level = ctx->params.instancing_mode == AOBAKE_INSTANCING_TWO_LEVEL;
  if (ctx->params.instancing_mode == AOBAKE_INSTANCING_AUTO) {
    std::vector<uint32_t> refs(ctx->meshes.size(), 0), brefs(bmeshes.size(), 0);
    for (const HostInstance& I : ctx->insts) if (++refs[I.mesh] > 1) two_level = true;
    for (const HostInstance& I : binsts) if (++brefs[I.mesh] > 1) two_level = true;
  }
  ctx->two_level = two_level;

  // ---- BVH build ----
  // Triangles per leaf slot.  Fewer triangles per leaf = more nodes but fewer triangle tests, and a
  // triangle test costs a warp far more than a node test because so few lanes are in one at a time.
  // Measured (profiles/r1/sweep_leaf_tris.log): 2 is best for flattened scenes (config 2 +1.9 %,
  // config 3 +2.8 % over 3); 1 is best for a BLAS, which is small and cache resident (config 4 +13.6 %).
  const uint32_t leaf_tris = (ctx->params.leaf_tris >= 1 && ctx->params.leaf_tris <= 3) ? (uint32_t)ctx->params.leaf_tris : (two_level ? 1u : 2u);
  CK(cudaEventRecord(ctx->ev0, st));
  struct Ref { const DeviceMesh* mesh; const HostInstance* inst; };
  std::vector<Ref> all;
  for (const HostInstance& I : ctx->insts) all.push_back({&ctx->meshes[I.mesh], &I});
  for (const HostInstance& I : binsts) all.push_back({&bmeshes[I.mesh], &I});
  memset(&ctx->stats, 0, sizeof(ctx->stats));
  ctx->stats.two_level = two_level ? 1 : 0;
  if (!two_level) {
    uint64_t n64 = 0;
    for (const Ref& r : all) n64 += r.mesh->nT;
    if (n64 > 0x7ffffff0ull) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "flattened scene has too many triangles (%llu)", (unsigned long long)n64);
    const uint32_t n = (uint32_t)n64;
    DBuf<F4> soup, plo, phi;
    DBuf<uint32_t> leaf_prims;
    DBuf<Node8> nodes;
    CK(soup.alloc(3ull * std::max(n, 1u))); CK(plo.alloc(std::max(n, 1u))); CK(phi.alloc(std::max(n, 1u)));
    CK(leaf_prims.alloc(std::max(n, 1u))); CK(nodes.alloc(std::max(n, 1u)));
    uint32_t off = 0;
    for (const Ref& r : all) {
      if (!r.mesh->nT) continue;
      Xf12 xf;
      memcpy(xf.m, r.inst->xf, sizeof(xf.m));
      k_make_tris<<<grid_for(r.mesh->nT, 256), 256, 0, st>>>(r.mesh->verts.p, r.mesh->tris.p, (uint32_t)r.mesh->nT, xf, 0, off,
                                                            soup.p, plo.p, phi.p);
      off += (uint32_t)r.mesh->nT;
    }
    CKL();
    Segment seg;
    if ((rc = build_segment(ctx, plo.p, phi.p, n, leaf_tris, 0, 0, nodes.p, leaf_prims.p, &seg))) return rc;
    // every visited node pushes at most one stack entry: the stack never holds more than the tree depth
    if (seg.levels + 2 > kStackSize)
      return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "BVH depth %d exceeds the traversal stack (%d entries)", seg.levels, kStackSize);
    ctx->stats.reserved[0] = seg.levels;
    CK(ctx->d_tris.alloc(3ull * std::max(n, 1u)));
    if (n) k_gather_tris<<<grid_for(n, 256), 256, 0, st>>>(soup.p, leaf_prims.p, n, 0, ctx->d_tris.p);
    CKL();
    CK(ctx->d_nodes.alloc(seg.node_count));
    CK(cudaMemcpyAsync(ctx->d_nodes.p, nodes.p, seg.node_count * sizeof(Node8), cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    ctx->d_insts.release();
    ctx->root = seg.root;
    ctx->scene_diag = n ? sqrtf((seg.box[3] - seg.box[0]) * (seg.box[3] - seg.box[0]) + (seg.box[4] - seg.box[1]) * (seg.box[4] - seg.box[1]) +
                                (seg.box[5] - seg.box[2]) * (seg.box[5] - seg.box[2])) : 0.f;
    ctx->main_diag = n ? sqrtf((seg.main_box[3] - seg.main_box[0]) * (seg.main_box[3] - seg.main_box[0]) + (seg.main_box[4] - seg.main_box[1]) * (seg.main_box[4] - seg.main_box[1]) +
                               (seg.main_box[5] - seg.main_box[2]) * (seg.main_box[5] - seg.main_box[2])) : 0.f;
    ctx->stats.num_bvh_nodes = seg.node_count;
    ctx->stats.num_bvh_triangles = n;
  } else {
    // one BLAS per mesh (scene meshes, then blocker meshes), object space
    std::vector<const DeviceMesh*> ml;
    for (const DeviceMesh& m : ctx->meshes) ml.push_back(&m);
    for (const DeviceMesh& m : bmeshes) ml.push_back(&m);
    uint64_t tri_total = 0, node_cap = 0;
    for (const DeviceMesh* m : ml) { tri_total += m->nT; node_cap += std::max<uint64_t>(m->nT, 1); }
    node_cap += std::max<uint64_t>(all.size(), 1);
    if (tri_total > 0x7ffffff0ull || node_cap > 0x7ffffff0ull) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "scene too large");
    DBuf<Node8> nodes;
    CK(nodes.alloc(node_cap));
    CK(ctx->d_tris.alloc(3ull * std::max<uint64_t>(tri_total, 1)));
    std::vector<Segment> segs(ml.size());
    uint32_t node_off = 0, prim_off = 0;
    for (size_t mi = 0; mi < ml.size(); mi++) {
      const DeviceMesh* m = ml[mi];
      const uint32_t n = (uint32_t)m->nT;
      DBuf<F4> soup, plo, phi;
      DBuf<uint32_t> leaf_prims;
      CK(soup.alloc(3ull * std::max(n, 1u))); CK(plo.alloc(std::max(n, 1u))); CK(phi.alloc(std::max(n, 1u))); CK(leaf_prims.alloc(std::max(n, 1u)));
      Xf12 idn;
      memset(&idn, 0, sizeof(idn));
      if (n) k_make_tris<<<grid_for(n, 256), 256, 0, st>>>(m->verts.p, m->tris.p, n, idn, 1, 0, soup.p, plo.p, phi.p);
      CKL();
      if ((rc = build_segment(ctx, plo.p, phi.p, n, leaf_tris, node_off, prim_off, nodes.p, leaf_prims.p, &segs[mi]))) return rc;
      if (n) k_gather_tris<<<grid_for(n, 256), 256, 0, st>>>(soup.p, leaf_prims.p, n, 0, ctx->d_tris.p + 3ull * prim_off);
      CKL();
      CK(cudaStreamSynchronize(st));
      node_off += segs[mi].node_count;
      prim_off += n;
    }
    // TLAS over instance world boxes: every mesh vertex through the instance transform (exact
    // fp32 formula), padded by a few ulp against the rounding of the inverse transform.
    const uint32_t nI = (uint32_t)all.size();
    std::vector<F4> hlo(std::max(nI, 1u)), hhi(std::max(nI, 1u));
    std::vector<uint32_t> inst_blas(nI);
    std::vector<float> inst_r2, inst_center;
    {
      DBuf<int> d_ib;
      DBuf<float> d_fb;
      CK(d_ib.alloc(6ull * std::max(nI, 1u)));
      CK(d_fb.alloc(6ull * std::max(nI, 1u)));
      if (nI) k_init_bounds_n<<<grid_for(6ull * nI, 256), 256, 0, st>>>(d_ib.p, nI);
      for (uint32_t i = 0; i < nI; i++) {
        const DeviceMesh* m = all[i].mesh;
        if (!m->nV) continue;
        Xf12 xf;
        memcpy(xf.m, all[i].inst->xf, sizeof(xf.m));
        k_instance_bounds<<<std::min<unsigned>(grid_for(m->nV, 256), 64u), 256, 0, st>>>(m->verts.p, (uint32_t)m->nV, xf, d_ib.p + 6ull * i);
      }
      if (nI) k_decode_bounds_n<<<grid_for(6ull * nI, 256), 256, 0, st>>>(d_ib.p, d_fb.p, 6 * nI);
      CKL();
      std::vector<float> hb(6ull * std::max(nI, 1u));
      if (nI) CK(cudaMemcpyAsync(hb.data(), d_fb.p, 6ull * nI * sizeof(float), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      // bounding sphere about the centre of the world box: max vertex distance (second pass)
      DBuf<uint32_t> d_r2;
      CK(d_r2.alloc(std::max(nI, 1u)));
      CK(cudaMemsetAsync(d_r2.p, 0, std::max(nI, 1u) * sizeof(uint32_t), st));
      for (uint32_t i = 0; i < nI; i++) {
        const DeviceMesh* m = all[i].mesh;
        if (!m->nV) continue;
        Xf12 xf;
        memcpy(xf.m, all[i].inst->xf, sizeof(xf.m));
        const float* b = &hb[6ull * i];
        k_instance_radius2<<<std::min<unsigned>(grid_for(m->nV, 256), 64u), 256, 0, st>>>(m->verts.p, (uint32_t)m->nV, xf, 0.5f * (b[0] + b[3]),
                                                                                         0.5f * (b[1] + b[4]), 0.5f * (b[2] + b[5]), d_r2.p + i);
      }
      CKL();
      inst_r2.assign(std::max(nI, 1u), 0.f);
      if (nI) CK(cudaMemcpyAsync(inst_r2.data(), d_r2.p, nI * sizeof(float), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      inst_center.assign(3ull * std::max(nI, 1u), 0.f);
      for (uint32_t i = 0; i < nI; i++)
        for (int k = 0; k < 3; k++) inst_center[3ull * i + k] = 0.5f * (hb[6ull * i + k] + hb[6ull * i + 3 + k]);
      for (uint32_t i = 0; i < nI; i++) {
        const bool is_blocker = i >= ctx->insts.size();
        const uint32_t mi = all[i].inst->mesh + (is_blocker ? (uint32_t)ctx->meshes.size() : 0u);
        inst_blas[i] = mi;
        F4 lo, hi;
        lo.w = hi.w = 0.f;
        if (ml[mi]->nT == 0 || ml[mi]->nV == 0) { lo.x = lo.y = lo.z = 0.f; hi = lo; }
        else {
          const float* b = &hb[6ull * i];
          lo.x = b[0]; lo.y = b[1]; lo.z = b[2]; hi.x = b[3]; hi.y = b[4]; hi.z = b[5];
          const float px = 3.8e-6f * fmaxf(fabsf(lo.x), fabsf(hi.x)), py = 3.8e-6f * fmaxf(fabsf(lo.y), fabsf(hi.y)),
                      pz = 3.8e-6f * fmaxf(fabsf(lo.z), fabsf(hi.z));
          lo.x -= px; lo.y -= py; lo.z -= pz; hi.x += px; hi.y += py; hi.z += pz;
        }
        hlo[i] = lo; hhi[i] = hi;
      }
    }
    DBuf<F4> plo, phi;
    DBuf<uint32_t> leaf_insts;
    CK(plo.alloc(std::max(nI, 1u))); CK(phi.alloc(std::max(nI, 1u))); CK(leaf_insts.alloc(std::max(nI, 1u)));
    if (nI) {
      CK(cudaMemcpyAsync(plo.p, hlo.data(), nI * sizeof(F4), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(phi.p, hhi.data(), nI * sizeof(F4), cudaMemcpyHostToDevice, st));
    }
    Segment tl;
    if ((rc = build_segment(ctx, plo.p, phi.p, nI, 1, node_off, 0, nodes.p, leaf_insts.p, &tl))) return rc;
    std::vector<uint32_t> order(nI);
    if (nI) CK(cudaMemcpyAsync(order.data(), leaf_insts.p, nI * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    std::vector<F4> recs((size_t)kInstF4 * std::max(nI, 1u));
    for (uint32_t k = 0; k < nI; k++) {
      const HostInstance* I = all[order[k]].inst;
      for (int r = 0; r < 3; r++) { F4 f; f.x = I->inv[4 * r]; f.y = I->inv[4 * r + 1]; f.z = I->inv[4 * r + 2]; f.w = I->inv[4 * r + 3]; recs[(size_t)kInstF4 * k + r] = f; }
      F4 f;
      uint32_t rootidx = segs[inst_blas[order[k]]].root, id = order[k];
      memcpy(&f.x, &rootidx, 4); memcpy(&f.y, &id, 4); f.z = 0.f; f.w = 0.f;
      recs[(size_t)kInstF4 * k + 3] = f;
      // bounding sphere, radius^2 padded against the rounding of the test and of the transforms
      F4 sph;
      sph.x = inst_center[3ull * order[k]]; sph.y = inst_center[3ull * order[k] + 1]; sph.z = inst_center[3ull * order[k] + 2];
      sph.w = inst_r2[order[k]] * 1.0002f + 1e-30f;
      recs[(size_t)kInstF4 * k + 4] = sph;
    }
    CK(ctx->d_insts.alloc(recs.size()));
    CK(cudaMemcpyAsync(ctx->d_insts.p, recs.data(), recs.size() * sizeof(F4), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    {
      int blas_levels = 0;
      for (const Segment& sg : segs) blas_levels = std::max(blas_levels, sg.levels);
      // TLAS entries (node group + primitive group per level) + sentinel + BLAS entries
      if (2 * tl.levels + blas_levels + 3 > kStackSize)
        return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "TLAS depth %d + BLAS depth %d exceed the traversal stack (%d entries)", tl.levels, blas_levels, kStackSize);
      ctx->stats.reserved[0] = tl.levels;
      ctx->stats.reserved[1] = blas_levels;
    }
    const uint32_t total_nodes = node_off + tl.node_count;
    CK(ctx->d_nodes.alloc(total_nodes));
    CK(cudaMemcpyAsync(ctx->d_nodes.p, nodes.p, total_nodes * sizeof(Node8), cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    ctx->root = tl.root;
    ctx->scene_diag = nI ? sqrtf((tl.box[3] - tl.box[0]) * (tl.box[3] - tl.box[0]) + (tl.box[4] - tl.box[1]) * (tl.box[4] - tl.box[1]) +
                                 (tl.box[5] - tl.box[2]) * (tl.box[5] - tl.box[2])) : 0.f;
    ctx->main_diag = nI ? sqrtf((tl.main_box[3] - tl.main_box[0]) * (tl.main_box[3] - tl.main_box[0]) + (tl.main_box[4] - tl.main_box[1]) * (tl.main_box[4] - tl.main_box[1]) +
                                (tl.main_box[5] - tl.main_box[2]) * (tl.main_box[5] - tl.main_box[2])) : 0.f;
    ctx->stats.num_bvh_nodes = total_nodes;
    ctx->stats.num_bvh_triangles = tri_total;
    ctx->stats.num_tlas_instances = nI;
  }
  ctx->stats.bvh_bytes = ctx->stats.num_bvh_nodes * sizeof(Node8) + ctx->stats.num_bvh_triangles * 48 + ctx->stats.num_tlas_instances * 16 * kInstF4;
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->timings.bvh_build_ms, ctx->ev0, ctx->ev1));
  ctx->have_scene = true;
  ctx->have_bvh = true;
  return AOBAKE_OK;
}

int aobake_set_scene(AoBake* ctx, const AoScene* scene, const AoScene* blockers) { return set_scene_impl(ctx, scene, blockers, false); }

int aobake_set_scene_distributed(AoBake* ctx, const AoScene* scene, const AoScene* blockers) {
  return set_scene_impl(ctx, scene, blockers, true);
}

int aobake_set_scene_geometry(AoBake* ctx, const AoScene* scene) { return set_scene_impl(ctx, scene, nullptr, false, false); }

int aobake_distribute_samples(AoBake* ctx, size_t min_per_tri, size_t requested, size_t* per_instance, size_t* total) {
  if (!ctx || !per_instance) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!ctx->have_scene) return ctx->fail(AOBAKE_ERR_STATE, "distribute_samples before set_scene");
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure_areas(ctx))) return rc;
  // distribute_samples_generic over instances (bake_sample_internal.h), host side:
  // the per-instance areas come from the device's fixed-shape sums.
  const size_t n = ctx->insts.size();
  std::vector<uint64_t> mins(n);
  uint64_t summin = 0;
  for (size_t i = 0; i < n; i++) { mins[i] = (uint64_t)min_per_tri * ctx->meshes[ctx->insts[i].mesh].nT; summin += mins[i]; }
  const uint64_t N = std::max<uint64_t>(requested, summin);
  double total_area = 0.0;
  for (size_t b = 0; b < n; b += 1024) {
    double s = 0.0;
    for (size_t i = b; i < std::min(n, b + 1024); i++) s = s + ctx->inst_area[i];
    total_area = total_area + s;
  }
  const uint64_t Na = N - summin;
  uint64_t assigned = 0;
  for (size_t i = 0; i < n; i++) {
    uint64_t c = mins[i];
    if (Na > 0 && total_area > 0.0) c += (uint64_t)(((double)Na * ctx->inst_area[i]) / total_area);
    per_instance[i] = c;
    assigned += c;
  }
  if (assigned > N) return ctx->fail(AOBAKE_ERR_SAMPLE_OVERFLOW, "instance budget floors exceed the total");
  uint64_t left = N - assigned;
  if (n == 0 && left) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "samples requested for an empty scene");
  for (size_t i = 0; left > 0; i = (i + 1) % n, left--) per_instance[i] += 1;
  if (total) *total = N;
  return AOBAKE_OK;
}

int aobake_sample_instances(AoBake* ctx, const size_t* per_instance, size_t min_per_tri, AoSamples* host_out) {
  if (!ctx || !per_instance) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!ctx->have_scene) return ctx->fail(AOBAKE_ERR_STATE, "sample_instances before set_scene");
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int rc;
  if ((rc = ensure_areas(ctx))) return rc;
  const uint32_t ni = (uint32_t)ctx->insts.size();
  uint64_t total = 0;
  std::vector<InstDesc> h(ni);
  if (ni) CK(cudaMemcpyAsync(h.data(), ctx->d_inst.p, ni * sizeof(InstDesc), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (uint32_t i = 0; i < ni; i++) {
    if (per_instance[i] < min_per_tri * h[i].num_tris)
      return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "instance %u: %llu samples < minimum %llu", i, (unsigned long long)per_instance[i],
                       (unsigned long long)(min_per_tri * h[i].num_tris));
    if (per_instance[i] && !h[i].num_tris) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "instance %u has no triangles to sample", i);
    h[i].sample_begin = total;
    h[i].num_samples = per_instance[i];
    total += per_instance[i];
  }
  if (host_out && host_out->num_samples != total)
    return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "host_out->num_samples %llu != sum(per_instance) %llu",
                     (unsigned long long)host_out->num_samples, (unsigned long long)total);
  if ((rc = alloc_samples(ctx, total))) return rc;
  ctx->per_instance.assign(per_instance, per_instance + ni);
  CK(cudaEventRecord(ctx->ev0, st));
  const uint64_t E = ctx->total_tris;
  if (total && E) {
    CK(cudaMemcpyAsync(ctx->d_inst.p, h.data(), ni * sizeof(InstDesc), cudaMemcpyHostToDevice, st));
    DBuf<uint64_t> counts, offs, final_off;
    DBuf<uint32_t> final_cnt;
    DBuf<long long> leftover;
    DBuf<int> status;
    CK(counts.alloc(E)); CK(offs.alloc(E)); CK(final_off.alloc(E + 1)); CK(final_cnt.alloc(E)); CK(leftover.alloc(ni)); CK(status.alloc(1));
    CK(cudaMemsetAsync(status.p, 0, sizeof(int), st));
    k_tri_counts<<<grid_for(E, 256), 256, 0, st>>>(ctx->d_inst.p, ni, E, ctx->d_tri_area.p, ctx->d_inst_area.p, (uint64_t)min_per_tri, counts.p);
    CKL();
    {
      size_t tmp_bytes = 0;
      CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.p, offs.p, (long long)E, st));
      DBuf<uint8_t> tmp;
      CK(tmp.alloc(tmp_bytes));
      CK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, counts.p, offs.p, (long long)E, st));
      CK(cudaStreamSynchronize(st));
    }
    k_inst_leftover<<<grid_for(ni, 128), 128, 0, st>>>(ctx->d_inst.p, ni, offs.p, counts.p, leftover.p, status.p);
    k_final_offsets<<<grid_for(E + 1, 256), 256, 0, st>>>(ctx->d_inst.p, ni, E, offs.p, counts.p, leftover.p, final_off.p, final_cnt.p, total);
    CKL();
    int hs = 0;
    CK(cudaMemcpyAsync(&hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hs) return ctx->fail(AOBAKE_ERR_SAMPLE_OVERFLOW, "area-proportional floors exceeded an instance budget (status %d)", hs);
    k_place_samples<<<grid_for(total, 256), 256, 0, st>>>(ctx->d_inst.p, ni, E, final_off.p, final_cnt.p, ctx->d_tri_area.p, total, ctx->d_pos.p,
                                                         ctx->d_nrm.p, ctx->d_fnrm.p, ctx->d_info.p);
    CKL();
    ctx->have_infos = true;
    CK(cudaStreamSynchronize(st));
  }
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->timings.sample_ms, ctx->ev0, ctx->ev1));
  if (host_out && total) {
    CK(cudaMemcpyAsync(host_out->sample_positions, ctx->d_pos.p, 12 * total, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(host_out->sample_normals, ctx->d_nrm.p, 12 * total, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(host_out->sample_face_normals, ctx->d_fnrm.p, 12 * total, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(host_out->sample_infos, ctx->d_info.p, sizeof(AoSampleInfo) * total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return AOBAKE_OK;
}

static int set_samples_impl(AoBake* ctx, const AoSamples* s, const size_t* per_instance, bool shard) {
  if (!ctx || !s) return AOBAKE_ERR_INVALID_ARGUMENT;
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  const uint64_t n = s->num_samples;
  if (n && (!s->sample_positions || !s->sample_normals || !s->sample_face_normals))
    return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "null sample arrays");
  shard = shard && ctx->comm_size > 1;
  int rc;
  if ((rc = alloc_samples(ctx, n))) return rc;
  ctx->per_instance.clear();
  if (per_instance) {
    uint64_t sum = 0;
    for (size_t i = 0; i < ctx->insts.size(); i++) sum += per_instance[i];
    if (sum != n) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "sum(per_instance) != num_samples");
    ctx->per_instance.assign(per_instance, per_instance + ctx->insts.size());
  }
  if (n) {
    cudaStream_t st = ctx->stream;
    DBuf<uint32_t> d_flag;
    CK(d_flag.alloc(1));
    CK(cudaMemsetAsync(d_flag.p, 0, sizeof(uint32_t), st));
    // sharded: only the super-blocks this rank traces (aobake_compute_ao_distributed) cross its PCIe link
    const uint64_t bs = kDefaultBlockSamples;
    const uint64_t step = shard ? bs * (uint64_t)ctx->comm_size : n;
    for (uint64_t b = shard ? bs * (uint64_t)ctx->comm_rank : 0; b < n; b += step) {
      const uint64_t e = shard ? std::min(n, b + bs) : n;
      CK(cudaMemcpyAsync(ctx->d_pos.p + 3 * b, s->sample_positions + 3 * b, 12 * (e - b), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->d_nrm.p + 3 * b, s->sample_normals + 3 * b, 12 * (e - b), cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(ctx->d_fnrm.p + 3 * b, s->sample_face_normals + 3 * b, 12 * (e - b), cudaMemcpyHostToDevice, st));
      k_check_unit_normals<<<grid_for(e - b, 256), 256, 0, st>>>(ctx->d_nrm.p, b, e, d_flag.p);
    }
    CKL();
    if (s->sample_infos) {
      CK(cudaMemcpyAsync(ctx->d_info.p, s->sample_infos, sizeof(AoSampleInfo) * n, cudaMemcpyHostToDevice, st));
      ctx->have_infos = true;
    }
    uint32_t flag = 0;
    CK(cudaMemcpyAsync(&flag, d_flag.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->unit_normals = flag == 0;
    ctx->samples_sharded = shard;
  }
  return AOBAKE_OK;
}

int aobake_set_samples(AoBake* ctx, const AoSamples* s, const size_t* per_instance) { return set_samples_impl(ctx, s, per_instance, false); }

int aobake_set_samples_distributed(AoBake* ctx, const AoSamples* s, const size_t* per_instance) {
  if (ctx && ctx->comm_size > 1 && !ctx->nccl_comm) return ctx->fail(AOBAKE_ERR_STATE, "aobake_comm_init has not been called");
  return set_samples_impl(ctx, s, per_instance, true);
}

size_t aobake_num_samples(const AoBake* ctx) { return ctx ? ctx->num_samples : 0; }

static int compute_ao_impl(AoBake* ctx, size_t begin, size_t end, int rays_per_sample, float offset, float maxdist, float* host_ao,
                           uint32_t part, uint32_t num_parts, uint32_t block_samples, bool force_fp32 = false) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!ctx->have_scene || !ctx->have_bvh) return ctx->fail(AOBAKE_ERR_STATE, "compute_ao before set_scene (aobake_set_scene_geometry builds no BVH)");
  if (begin > end || end > ctx->num_samples) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "sample range [%zu,%zu) outside [0,%llu)", begin, end, (unsigned long long)ctx->num_samples);
  if (rays_per_sample < 1) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "rays_per_sample must be >= 1");
  if (ctx->samples_sharded && !(num_parts == (uint32_t)ctx->comm_size && part == (uint32_t)ctx->comm_rank && begin == 0 && end == ctx->num_samples &&
                                (block_samples == 0 || block_samples == kDefaultBlockSamples)))
    return ctx->fail(AOBAKE_ERR_STATE, "only this rank's super-blocks are resident (aobake_set_samples_distributed): use aobake_compute_ao_distributed");
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int q = sqrt_rays(rays_per_sample);
  if (q < 1 || q > 255) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "sqrt(rays_per_sample) must be in [1,255]");
  const uint64_t n = end - begin;
  // interleaved partition: which 32-sample blocks this call owns
  const uint64_t n_global_blocks = (n + 31) / 32;
  uint32_t sb_blocks = 1;
  uint64_t n_local_blocks = n_global_blocks, owned_samples = n;
  if (num_parts > 1) {
    if (part >= num_parts) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "part %u of %u", part, num_parts);
    if (block_samples == 0) block_samples = kDefaultBlockSamples;
    if (block_samples % 32) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "block_samples must be a multiple of 32");
    sb_blocks = block_samples / 32;
    const uint64_t n_sb = (n_global_blocks + sb_blocks - 1) / sb_blocks;
    const uint64_t owned_sb = n_sb / num_parts + ((n_sb % num_parts) > part ? 1 : 0);
    n_local_blocks = owned_sb * sb_blocks;
    owned_samples = 0;
    for (uint64_t sb = part; sb < n_sb; sb += num_parts) {
      const uint64_t s0 = sb * (uint64_t)block_samples, s1 = std::min<uint64_t>(n, s0 + block_samples);
      owned_samples += s1 - s0;
    }
  }
  ctx->timings.rays_traced = owned_samples * (uint64_t)q * q;
  ctx->timings.trace_ms = 0.f;
  if (n == 0) {
    if (ctx->num_samples == 0) ctx->have_ao = true;  // nothing to trace: the (empty) AO array is complete
    return AOBAKE_OK;
  }
  if (num_parts > 1) {
    // everything this part does not own reads as zero, so that an all-reduce (sum) assembles ao[]
    CK(cudaMemsetAsync(ctx->d_hits.p + begin, 0, n * sizeof(uint32_t), st));
  }
  const bool stats = ctx->params.collect_stats != 0;
  SampleView S{ctx->d_pos.p, ctx->d_nrm.p, ctx->d_fnrm.p};
  const BvhView bvh = bvh_view(ctx);
  if (stats) CK(cudaMemsetAsync(ctx->d_stats.p, 0, 4 * sizeof(unsigned long long), st));
  CK(cudaEventRecord(ctx->ev0, st));
  const uint32_t q2 = (uint32_t)(q * q);
  int launches = 0;
  bool use_h2 = false;
  // trace_kernel: 0 = auto (persistent, except for launches too small to amortise its work
  // distribution: < 32 M rays), 1 = simple, 2 = persistent
  const bool use_simple = num_parts == 1 && (ctx->params.trace_kernel == 1 || (ctx->params.trace_kernel == 0 && n * (uint64_t)q2 < (32ull << 20)));
  if (use_simple) {
    // simple variant: enough (sample block, strata chunk) items to fill the machine
    const uint64_t n_blocks = (n + 31) / 32;
    const uint64_t want = (uint64_t)ctx->sm_count * 64ull * 4ull;
    uint32_t n_chunks = 1;
    while (n_blocks * n_chunks < want && n_chunks * 2 <= q2) n_chunks *= 2;
    if (n_chunks > 1) CK(cudaMemsetAsync(ctx->d_hits.p + begin, 0, n * sizeof(uint32_t), st));
    const uint64_t warps = n_blocks * n_chunks;
    const unsigned grid = grid_for(warps * 32, 256);
    const bool h2 = AOB_H2 != 0 && ctx->params.node_test != 1;   // per ray: trace_any_hit checks direction range and length itself
    auto kern = h2 ? (stats ? k_ao_simple<true, true> : k_ao_simple<false, true>) : (stats ? k_ao_simple<true, false> : k_ao_simple<false, false>);
    kern<<<grid, 256, 0, st>>>(bvh, S, begin, end, q, offset, maxdist, n_chunks, ctx->d_hits.p + begin, ctx->d_stats.p);
    CKL();
    launches++;
  } else {
    // persistent variant: one resident wave of CTAs (a multiple of the SM count), dynamic work fetch
    if (n > 0xfffffff0ull) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "more than 2^32 samples in one range");
    using KernelT = AoKernelT;
    // node test: packed fp16 (two planes per instruction) for flattened scenes, unless asked otherwise
    // or a previous attempt overflowed the deferred-ray list; fp32 under a TLAS (measured faster there)
    use_h2 = AOB_H2 != 0 && !force_fp32 && ctx->params.node_test != 1 && !ctx->two_level && ctx->unit_normals;
    // the far clamp of the fp32 node test is only needed when maxdist can actually cull inside the scene
    // (measured against the tree over the ORDINARY primitives: an oversized blocker on the extra root — the ground plane,
    // 100 x the scene — must not switch the clamp on for the whole traversal; culling by tmax is an optimisation only)
    const bool clamp = !(maxdist > 1.01f * ctx->main_diag + fabsf(offset));
    // ray order within a work item: stratum-major (the warp's rays share origin neighbourhood AND direction) or sample-major
    const bool packet = (ctx->params.ray_order == 0 ? AOB_RAY_ORDER_DEFAULT : ctx->params.ray_order) == 2;
    KernelT kern = pick_ao_kernel(stats, ctx->two_level, clamp, use_h2, packet);
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAoBlock, 0));
    if (per_sm < 1) per_sm = 1;
    const uint64_t resident_threads = (uint64_t)per_sm * ctx->sm_count * kAoBlock;
    // work items (32-sample block x strata chunk) per resident warp: below ~32 the last items of the launch leave most
    // warps idle (a rank's share of config 3 at 8 GPUs is 9 blocks per warp: 13 % tail) — split the strata
    uint32_t n_chunks = 1;
    while (owned_samples * n_chunks < (uint64_t)kItemsPerWarp * resident_threads && n_chunks * 2 <= q2) n_chunks *= 2;
    // stratum-major order: an item costs the warp one atomic and one flush of 32 counters whatever its size, so the tail can be
    // halved again — 64 items per warp — as long as an item keeps at least 8 strata (256 rays).  One rank's share of config 3 at
    // 8 GPUs: +2.1 % -> +1.5 % over an eighth of the whole pass, config 2 +0.5 % (profiles/r2/part_probe_c3_items_per_warp.log)
    if (packet)
      while (owned_samples * n_chunks < 2ull * kItemsPerWarp * resident_threads && q2 / (n_chunks * 2) >= 8) n_chunks *= 2;
    unsigned grid = (unsigned)(per_sm * ctx->sm_count);
    const uint64_t items = std::max<uint64_t>(owned_samples, 1) * n_chunks;
    if ((uint64_t)grid * kAoBlock > items) grid = (unsigned)((items + kAoBlock - 1) / kAoBlock);
    if (n_chunks > 1 && num_parts == 1) CK(cudaMemsetAsync(ctx->d_hits.p + begin, 0, n * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(ctx->d_counter.p, 0, sizeof(unsigned long long), st));
    DeferredRays deferred{nullptr, nullptr, 0u};
    if (use_h2) {
      // expected: ~7e-4 of the rays; room for 1/128
      const uint64_t want = ctx->params.deferred_capacity > 0
                                ? (uint64_t)ctx->params.deferred_capacity
                                : std::min<uint64_t>(std::max<uint64_t>(owned_samples * q2 / 128ull, 1ull << 18), 1ull << 28);
      if (ctx->d_deferred.n < want || ctx->params.deferred_capacity > 0) CK(ctx->d_deferred.alloc(want));
      CK(cudaMemsetAsync(ctx->d_deferred_count.p, 0, sizeof(uint32_t), st));
      deferred.list = ctx->d_deferred.p; deferred.count = ctx->d_deferred_count.p; deferred.capacity = (uint32_t)ctx->d_deferred.n;
    }
    const uint32_t refill = ctx->params.refill_below > 0 ? (uint32_t)ctx->params.refill_below : 28u;
    // lanes that must hold leaf hits before the warp runs its triangle block (1 = test at once)
    // tri_batch = lanes | iterations << 8 (the longest a paused lane waits); iterations 0 => no limit
    uint32_t tri_batch = ctx->params.tri_batch > 0 ? (uint32_t)ctx->params.tri_batch
                                                   : (ctx->two_level ? kTriBatchTwoLevel : (packet ? kTriBatchFlatPacket : kTriBatchFlat));
    {
      const uint32_t lanes = std::min(std::max(tri_batch & 0xffu, 1u), 32u), wait = (tri_batch >> 8) & 0xffu;
      tri_batch = lanes | ((wait ? wait : 255u) << 8);
    }
    kern<<<grid, kAoBlock, 0, st>>>(bvh, S, (uint64_t)begin, (uint32_t)n, q, offset, maxdist, n_chunks, refill, tri_batch, part, num_parts, sb_blocks,
                                    (uint32_t)n_local_blocks, ctx->d_hits.p + begin, ctx->d_counter.p, ctx->d_stats.p, deferred);
    CKL();
    launches++;
    if (use_h2) {
      k_ao_deferred<<<(unsigned)ctx->sm_count * 4u, 128, 0, st>>>(bvh, S, (uint64_t)begin, q, offset, maxdist, deferred, ctx->d_hits.p + begin);
      CKL();
      launches++;
    }
  }
  k_ao_finalize<<<grid_for(n, 256), 256, 0, st>>>(ctx->d_hits.p + begin, n, (float)(q * q), ctx->d_ao.p + begin, part, num_parts, sb_blocks * 32u);
  CKL();
  launches++;
  ctx->timings.kernel_launches = launches;
  CK(cudaEventRecord(ctx->ev1, st));
  uint32_t n_deferred = 0;
  if (use_h2) CK(cudaMemcpyAsync(&n_deferred, ctx->d_deferred_count.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  if (host_ao) CK(cudaMemcpyAsync(host_ao, ctx->d_ao.p + begin, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (use_h2 && n_deferred > ctx->d_deferred.n)   // entries were dropped: this result is incomplete, trace again in fp32
    return compute_ao_impl(ctx, begin, end, rays_per_sample, offset, maxdist, host_ao, part, num_parts, block_samples, true);
  ctx->stats.reserved[2] = (int32_t)std::min<uint32_t>(n_deferred, 0x7fffffffu);
  CK(cudaEventElapsedTime(&ctx->timings.trace_ms, ctx->ev0, ctx->ev1));
  if (stats) {
    unsigned long long hs[4];
    CK(cudaMemcpyAsync(hs, ctx->d_stats.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ctx->stats.node_visits = hs[0]; ctx->stats.triangle_tests = hs[1]; ctx->stats.instance_entries = hs[2];
    ctx->stats.rays = ctx->timings.rays_traced;
  }
  ctx->have_ao = true;
  return AOBAKE_OK;
}

int aobake_compute_ao_range(AoBake* ctx, size_t begin, size_t end, int rays_per_sample, float offset, float maxdist, float* host_ao) {
  return compute_ao_impl(ctx, begin, end, rays_per_sample, offset, maxdist, host_ao, 0, 1, 0);
}

int aobake_compute_ao_interleaved(AoBake* ctx, uint32_t part, uint32_t num_parts, uint32_t block_samples, int rays_per_sample, float offset,
                                  float maxdist) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (num_parts == 0) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "num_parts must be >= 1");
  return compute_ao_impl(ctx, 0, ctx->num_samples, rays_per_sample, offset, maxdist, nullptr, part, num_parts, block_samples);
}

int aobake_comm_unique_id(void* id128) {
  if (!id128) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!g_nccl.load()) { g_create_error = g_nccl.error; return AOBAKE_ERR_COMM; }
  NcclId id;
  const int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(rc); return AOBAKE_ERR_COMM; }
  memcpy(id128, id.internal, sizeof(id.internal));
  return AOBAKE_OK;
}

int aobake_comm_init(AoBake* ctx, int rank, int nranks, const void* id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!g_nccl.load()) return ctx->fail(AOBAKE_ERR_COMM, "%s", g_nccl.error.c_str());
  CK(cudaSetDevice(ctx->device));
  if (ctx->nccl_comm) { g_nccl.CommDestroy(ctx->nccl_comm); ctx->nccl_comm = nullptr; }
  NcclId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  const int rc = g_nccl.CommInitRank(&ctx->nccl_comm, nranks, id, rank);
  if (rc != 0) return ctx->fail(AOBAKE_ERR_COMM, "ncclCommInitRank: %s", g_nccl.GetErrorString(rc));
  ctx->comm_rank = rank;
  ctx->comm_size = nranks;
  g_alloc_stream = ctx->stream;
  CK(ctx->d_comm_status.alloc(1));
  return AOBAKE_OK;
}

int aobake_comm_destroy(AoBake* ctx) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (ctx->nccl_comm && g_nccl.handle) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy(ctx->nccl_comm);
  }
  ctx->nccl_comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_size = 1;
  return AOBAKE_OK;
}

int aobake_compute_ao_distributed(AoBake* ctx, int rays_per_sample, float offset, float maxdist, float* host_ao) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (ctx->comm_size > 1 && !ctx->nccl_comm) return ctx->fail(AOBAKE_ERR_STATE, "aobake_comm_init has not been called");
  int rc = compute_ao_impl(ctx, 0, ctx->num_samples, rays_per_sample, offset, maxdist, nullptr, (uint32_t)ctx->comm_rank,
                           (uint32_t)ctx->comm_size, 0);
  if ((rc = comm_agree(ctx, rc))) return rc;   // a rank that failed locally must not leave the others in the all-reduce
  cudaStream_t st = ctx->stream;
  if (ctx->comm_size > 1 && ctx->num_samples) {
    // every rank holds exact zeros outside its super-blocks: the sum assembles ao[] bit for bit
    const int nrc = g_nccl.AllReduce(ctx->d_ao.p, ctx->d_ao.p, ctx->num_samples, kNcclFloat, kNcclSum, ctx->nccl_comm, st);
    if (nrc != 0) return ctx->fail(AOBAKE_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString(nrc));
  }
  if (host_ao && ctx->num_samples) CK(cudaMemcpyAsync(host_ao, ctx->d_ao.p, ctx->num_samples * sizeof(float), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return AOBAKE_OK;
}

int aobake_compute_ao(AoBake* ctx, int rays_per_sample, float offset, float maxdist, float* host_ao) {
  if (!ctx) return AOBAKE_ERR_INVALID_ARGUMENT;
  return aobake_compute_ao_range(ctx, 0, ctx->num_samples, rays_per_sample, offset, maxdist, host_ao);
}

int aobake_get_ao_device(AoBake* ctx, float** d_ao, size_t* n) {
  if (!ctx || !d_ao) return AOBAKE_ERR_INVALID_ARGUMENT;
  *d_ao = ctx->d_ao.p;
  if (n) *n = ctx->num_samples;
  return AOBAKE_OK;
}
int aobake_set_ao(AoBake* ctx, const float* host_ao) {
  if (!ctx || !host_ao) return AOBAKE_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  if (ctx->num_samples) CK(cudaMemcpyAsync(ctx->d_ao.p, host_ao, ctx->num_samples * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->have_ao = true;
  return AOBAKE_OK;
}
int aobake_get_hit_counts(AoBake* ctx, uint32_t* host_counts) {
  if (!ctx || !host_counts) return AOBAKE_ERR_INVALID_ARGUMENT;
  CK(cudaSetDevice(ctx->device));
  if (ctx->num_samples) CK(cudaMemcpyAsync(host_counts, ctx->d_hits.p, ctx->num_samples * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return AOBAKE_OK;
}

}  // extern "C"

namespace {

// Global numbering shared by both vertex maps: instance i owns vertices [vbase[i], vbase[i+1]),
// triangles, samples and (least squares) interior edges in consecutive global ranges.
// The numbering is relative to the instance range [ib, ie) being solved (the whole scene, or one
// rank's share of it: the systems are block diagonal, so instances split freely across GPUs).
void build_filter_instances(AoBake* ctx, uint32_t ib, uint32_t ie, const std::vector<DBuf<uint32_t>>* topo,
                            const std::vector<uint32_t>* topo_count, std::vector<LsInst>& h, std::vector<uint64_t>& vbase, uint64_t* NV,
                            uint64_t* NT, uint64_t* NE, uint64_t* NS, uint64_t* sample0) {
  const uint32_t nI = ie - ib;
  h.assign(std::max(nI, 1u), LsInst{});
  vbase.assign(nI + 1, 0);
  uint64_t s0 = 0;
  for (uint32_t i = 0; i < ib; i++) s0 += ctx->per_instance[i];
  *sample0 = s0;
  uint64_t nv = 0, nt = 0, ne = 0, ns = 0;
  for (uint32_t i = 0; i < nI; i++) {
    const HostInstance& I = ctx->insts[ib + i];
    const DeviceMesh& M = ctx->meshes[I.mesh];
    LsInst& L = h[i];
    memcpy(L.xf, I.xf, sizeof(L.xf));
    L.tris = M.tris.p; L.verts = M.verts.p; L.topo = topo ? (*topo)[I.mesh].p : nullptr;
    L.sample_begin = ns; L.tri_begin = nt; L.edge_begin = ne; L.vert_begin = (uint32_t)nv;
    L.num_tris = (uint32_t)M.nT; L.num_edges = topo_count ? (*topo_count)[I.mesh] : 0u; L.pad = 0;
    vbase[i] = nv;
    ns += ctx->per_instance[ib + i]; nt += M.nT; ne += L.num_edges; nv += M.nV;
  }
  vbase[nI] = nv;
  *NV = nv; *NT = nt; *NE = ne; *NS = ns;
}

// bake_filter.cpp filter / filter_mesh for all instances in one pass.
int area_filter_batched(AoBake* ctx, uint32_t ib, uint32_t ie, DBuf<float>& d_out) {
  cudaStream_t st = ctx->stream;
  const uint32_t nI = ie - ib;
  std::vector<LsInst> h;
  std::vector<uint64_t> vbase;
  uint64_t NV = 0, NT = 0, NE = 0, NS = 0, S0 = 0;
  build_filter_instances(ctx, ib, ie, nullptr, nullptr, h, vbase, &NV, &NT, &NE, &NS, &S0);
  if (NV > 0xfffffff0ull) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "too many vertices for 32-bit indices");
  if (S0 + NS > ctx->num_samples) return ctx->fail(AOBAKE_ERR_STATE, "per-instance sample counts exceed the resident samples");
  const uint64_t v1 = std::max<uint64_t>(NV, 1);
  DBuf<LsInst> d_inst;
  DBuf<double> num, wgt;
  CK(d_inst.alloc(std::max(nI, 1u))); CK(num.alloc(v1)); CK(wgt.alloc(v1)); CK(d_out.alloc(v1));
  if (nI) CK(cudaMemcpyAsync(d_inst.p, h.data(), nI * sizeof(LsInst), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(num.p, 0, v1 * sizeof(double), st));
  CK(cudaMemsetAsync(wgt.p, 0, v1 * sizeof(double), st));
  if (NS) k_area_scatter_b<<<grid_for(NS, 256), 256, 0, st>>>(ctx->d_info.p + S0, ctx->d_ao.p + S0, NS, d_inst.p, nI, num.p, wgt.p);
  if (NV) k_area_final<<<grid_for(NV, 256), 256, 0, st>>>(num.p, wgt.p, NV, d_out.p);
  CKL();
  CK(cudaStreamSynchronize(st));
  return AOBAKE_OK;
}

// bake_filter_least_squares.cpp: (M + w R) x = b, fp64, for ALL instances as one block-diagonal
// system, matrix-free Jacobi-PCG.
// rows = true (multi-GPU, fewer instances than ranks): every rank assembles the whole system [ib, ie) = all instances
// and the PCG is partitioned by vertex ROWS across the ranks of the communicator (see k_ls_flag_items); d_out then
// holds this rank's rows and zeros elsewhere.
int ls_filter_batched(AoBake* ctx, float weight, uint32_t ib, uint32_t ie, DBuf<float>& d_out, bool rows = false) {
  cudaStream_t st = ctx->stream;
  const uint32_t nI = ie - ib;
  const double w = weight;
  // ---- interior-edge topology, once per mesh that is instanced ----
  const size_t nM = ctx->meshes.size();
  std::vector<DBuf<uint32_t>> topo(nM);
  std::vector<uint32_t> topo_count(nM, 0);
  std::vector<char> used(nM, 0);
  for (uint32_t i = ib; i < ie; i++) used[ctx->insts[i].mesh] = 1;
  if (weight != 0.0f) {
    for (size_t m = 0; m < nM; m++) {
      const DeviceMesh& M = ctx->meshes[m];
      if (!used[m] || !M.nT) continue;
      const uint64_t nH = 3 * M.nT;
      DBuf<uint64_t> keys, keys_s;
      DBuf<uint32_t> vals, vals_s, cnt;
      CK(keys.alloc(nH)); CK(keys_s.alloc(nH)); CK(vals.alloc(nH)); CK(vals_s.alloc(nH)); CK(cnt.alloc(1));
      CK(topo[m].alloc(4 * (nH / 2 + 1)));
      CK(cudaMemsetAsync(cnt.p, 0, sizeof(uint32_t), st));
      k_ls_halfedges<<<grid_for(nH, 256), 256, 0, st>>>(M.tris.p, M.nT, keys.p, vals.p);
      CKL();
      size_t tmp_bytes = 0;
      CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (long long)nH, 0, 64, st));
      DBuf<uint8_t> tmp;
      CK(tmp.alloc(tmp_bytes));
      CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys_s.p, vals.p, vals_s.p, (long long)nH, 0, 64, st));
      k_ls_topo<<<grid_for(nH, 256), 256, 0, st>>>(keys_s.p, vals_s.p, nH, M.tris.p, topo[m].p, cnt.p);
      CKL();
      CK(cudaMemcpyAsync(&topo_count[m], cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
  }
  // ---- instance descriptors and global numbering ----
  std::vector<LsInst> h;
  std::vector<uint64_t> vbase;
  uint64_t NV = 0, NT = 0, NE = 0, NS = 0, S0 = 0;
  build_filter_instances(ctx, ib, ie, &topo, &topo_count, h, vbase, &NV, &NT, &NE, &NS, &S0);
  if (NV > 0xfffffff0ull || NE > 0xfffffff0ull) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "least-squares system too large for 32-bit indices");
  if (S0 + NS > ctx->num_samples) return ctx->fail(AOBAKE_ERR_STATE, "per-instance sample counts exceed the resident samples");
  const uint64_t v1 = std::max<uint64_t>(NV, 1), t1 = std::max<uint64_t>(NT, 1);
  DBuf<LsInst> d_inst;
  DBuf<uint32_t> gtris;
  DBuf<double> Mt, rhs, diag, x, r, z, p, Ap, scal;
  DBuf<uint8_t> fixed;
  DBuf<LsEdge> edges;
  CK(d_inst.alloc(std::max(nI, 1u))); CK(gtris.alloc(3 * t1)); CK(Mt.alloc(6 * t1)); CK(rhs.alloc(v1)); CK(diag.alloc(v1));
  CK(x.alloc(v1)); CK(r.alloc(v1)); CK(z.alloc(v1)); CK(p.alloc(v1)); CK(Ap.alloc(v1)); CK(scal.alloc(12)); CK(fixed.alloc(v1));
  CK(edges.alloc(std::max<uint64_t>(NE, 1))); CK(d_out.alloc(v1));
  if (nI) CK(cudaMemcpyAsync(d_inst.p, h.data(), nI * sizeof(LsInst), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(Mt.p, 0, Mt.n * sizeof(double), st));
  CK(cudaMemsetAsync(rhs.p, 0, v1 * sizeof(double), st));
  CK(cudaMemsetAsync(diag.p, 0, v1 * sizeof(double), st));
  if (NT) k_ls_gtris<<<grid_for(NT, 256), 256, 0, st>>>(d_inst.p, nI, NT, gtris.p);
  if (NS) k_ls_mass_b<<<grid_for(NS, 256), 256, 0, st>>>(ctx->d_info.p + S0, ctx->d_ao.p + S0, NS, d_inst.p, nI, gtris.p, Mt.p, rhs.p);
  if (NE) k_ls_edge_coeffs<<<grid_for(NE, 256), 256, 0, st>>>(d_inst.p, nI, NE, ctx->params.ls_energy, edges.p);
  if (NT) k_ls_diag_mass<<<grid_for(NT, 256), 256, 0, st>>>(gtris.p, NT, Mt.p, diag.p);
  if (NV) k_ls_fix<<<grid_for(NV, 256), 256, 0, st>>>(diag.p, rhs.p, fixed.p, NV);
  if (NE) k_ls_diag_edges<<<grid_for(NE, 256), 256, 0, st>>>(edges.p, (uint32_t)NE, w, diag.p);
  if (NV) k_ls_init<<<grid_for(NV, 256), 256, 0, st>>>(rhs.p, diag.p, r.p, z.p, p.p, x.p, Ap.p, NV);
  CKL();
  // scal: [0] = |b|^2, [1] = r.z of the previous iteration, [2] = |r|^2 and [3] = p.Ap of the last finished iteration,
  // [4..6] and [8..10] = the two accumulator banks {p.Ap, r.z, |r|^2} (iteration parity)
  const int nranks = rows ? ctx->comm_size : 1;
  const uint32_t rows_per_rank = (uint32_t)((NV + (uint64_t)nranks - 1) / (uint64_t)nranks);
  const uint32_t r0 = rows ? (uint32_t)std::min<uint64_t>((uint64_t)ctx->comm_rank * rows_per_rank, NV) : 0u;
  const uint32_t r1 = rows ? (uint32_t)std::min<uint64_t>((uint64_t)r0 + rows_per_rank, NV) : (uint32_t)NV;
  const uint64_t n_rows = r1 - r0;
  // ---- row partition: this rank's item lists and the boundary vertices all ranks exchange ----
  DBuf<uint32_t> tri_list, edge_list, bidx, d_boff;
  DBuf<double> hbuf;
  uint32_t n_tri_mine = 0, n_edge_mine = 0, n_boundary = 0;
  std::vector<uint32_t> boff(nranks + 1, 0);
  if (rows) {
    DBuf<uint8_t> tri_mine, edge_mine, boundary;
    DBuf<uint32_t> counts, d_bound;
    CK(tri_mine.alloc(t1)); CK(edge_mine.alloc(std::max<uint64_t>(NE, 1))); CK(boundary.alloc(v1)); CK(counts.alloc(3));
    CK(tri_list.alloc(t1)); CK(edge_list.alloc(std::max<uint64_t>(NE, 1))); CK(bidx.alloc(v1)); CK(d_boff.alloc(nranks + 1)); CK(d_bound.alloc(nranks + 1));
    CK(cudaMemsetAsync(boundary.p, 0, v1, st));
    const uint64_t nitems = std::max<uint64_t>(NT, NE);
    if (nitems) k_ls_flag_items<<<grid_for(nitems, 256), 256, 0, st>>>(gtris.p, NT, edges.p, NE, r0, r1, std::max(rows_per_rank, 1u), tri_mine.p, edge_mine.p, boundary.p);
    CKL();
    auto select = [&](const uint8_t* flags, uint64_t n, uint32_t* out, uint32_t* d_count) -> int {
      if (!n) return AOBAKE_OK;
      cub::CountingInputIterator<uint32_t> iota(0u);
      size_t tmp_bytes = 0;
      CK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, flags, out, d_count, (long long)n, st));
      DBuf<uint8_t> tmp;
      CK(tmp.alloc(tmp_bytes));
      CK(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, iota, flags, out, d_count, (long long)n, st));
      return AOBAKE_OK;
    };
    CK(cudaMemsetAsync(counts.p, 0, 3 * sizeof(uint32_t), st));
    int rc;
    if ((rc = select(tri_mine.p, NT, tri_list.p, counts.p)) || (rc = select(edge_mine.p, NE, edge_list.p, counts.p + 1)) ||
        (rc = select(boundary.p, NV, bidx.p, counts.p + 2)))
      return rc;
    uint32_t hc[3];
    CK(cudaMemcpyAsync(hc, counts.p, sizeof(hc), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    n_tri_mine = hc[0]; n_edge_mine = hc[1]; n_boundary = hc[2];
    std::vector<uint32_t> bound(nranks + 1);
    for (int s2 = 0; s2 <= nranks; s2++) bound[s2] = (uint32_t)std::min<uint64_t>((uint64_t)s2 * rows_per_rank, NV);
    CK(cudaMemcpyAsync(d_bound.p, bound.data(), (nranks + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    k_lower_bounds<<<1, 64, 0, st>>>(bidx.p, n_boundary, d_bound.p, (uint32_t)(nranks + 1), d_boff.p);
    CKL();
    CK(cudaMemcpyAsync(boff.data(), d_boff.p, (nranks + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(hbuf.alloc(std::max(n_boundary, 1u)));
  }
  // a scalar every rank needs the same value of: summed over the ranks (the result of an all-reduce is identical on
  // all of them, so they take the same number of iterations and stay in step)
  auto all_sum = [&](double* d, size_t count) -> int {
    if (!rows) return AOBAKE_OK;
    const int nrc = g_nccl.AllReduce(d, d, count, kNcclDouble, kNcclSum, ctx->nccl_comm, st);
    return nrc == 0 ? AOBAKE_OK : ctx->fail(AOBAKE_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString(nrc));
  };
  // p at the boundary vertices: every owner broadcasts its segment of the (sorted) boundary list
  auto exchange_boundary = [&]() -> int {
    if (!rows || !n_boundary) return AOBAKE_OK;
    const uint32_t mb = boff[ctx->comm_rank], me = boff[ctx->comm_rank + 1];
    if (me > mb) k_gather_d<<<grid_for(me - mb, 256), 256, 0, st>>>(p.p, bidx.p, mb, me, hbuf.p);
    int nrc = g_nccl.GroupStart();
    for (int s2 = 0; s2 < nranks && nrc == 0; s2++)
      if (boff[s2 + 1] > boff[s2]) nrc = g_nccl.Broadcast(hbuf.p + boff[s2], hbuf.p + boff[s2], boff[s2 + 1] - boff[s2], kNcclDouble, s2, ctx->nccl_comm, st);
    const int nrc2 = g_nccl.GroupEnd();
    if (nrc != 0 || nrc2 != 0) return ctx->fail(AOBAKE_ERR_COMM, "ncclBroadcast (boundary exchange): %s", g_nccl.GetErrorString(nrc ? nrc : nrc2));
    k_scatter_d<<<grid_for(n_boundary, 256), 256, 0, st>>>(p.p, bidx.p, n_boundary, mb, me, hbuf.p);
    return AOBAKE_OK;
  };
  // ---- A = M + wR assembled once into sliced ELL (k_ls_assemble ... k_ls_spmv_sell); the matrix-free product stays as
  // the fallback for rows with more columns than the assembly tables hold, and as the A/B switch ls_matrix_free ----
  bool use_matrix = ctx->params.ls_matrix_free == 0 && n_rows > 0;
  DBuf<uint32_t> sell_cols, slice_width, over_rows, over_tris, over_edges;
  DBuf<uint64_t> slice_off;
  DBuf<double> sell_vals;
  DBuf<uint8_t> row_over;
  uint32_t n_over_rows = 0, n_over_tris = 0, n_over_edges = 0;
  if (use_matrix) {
    const uint64_t n_slices = (n_rows + 31) / 32;
    DBuf<uint32_t> hkeys, row_nnz;
    DBuf<double> hvals;
    DBuf<uint64_t> width64;
    CK(hkeys.alloc(n_rows * kLsHashCap)); CK(hvals.alloc(n_rows * kLsHashCap)); CK(row_nnz.alloc(n_rows)); CK(row_over.alloc(n_rows));
    CK(slice_width.alloc(n_slices)); CK(slice_off.alloc(n_slices + 1)); CK(width64.alloc(n_slices + 1));
    CK(cudaMemsetAsync(hkeys.p, 0xff, n_rows * kLsHashCap * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(hvals.p, 0, n_rows * kLsHashCap * sizeof(double), st));
    CK(cudaMemsetAsync(row_over.p, 0, n_rows, st));
    const uint64_t a_tri = rows ? n_tri_mine : NT, a_edge = rows ? n_edge_mine : NE;
    const uint32_t* a_tri_list = rows ? tri_list.p : nullptr;
    const uint32_t* a_edge_list = rows ? edge_list.p : nullptr;
    if (std::max(a_tri, a_edge))
      k_ls_assemble<<<grid_for(std::max(a_tri, a_edge), 256), 256, 0, st>>>(a_tri_list, a_tri, a_edge_list, a_edge, gtris.p, Mt.p, edges.p, w, r0, r1, hkeys.p,
                                                                           hvals.p, row_over.p);
    k_ls_assemble_diag<<<grid_for(n_rows, 256), 256, 0, st>>>(fixed.p, r0, r1, hkeys.p, hvals.p, row_over.p);
    k_ls_row_widths<<<grid_for(n_slices * 32, 256), 256, 0, st>>>(hkeys.p, row_over.p, n_rows, row_nnz.p, slice_width.p);
    k_u32_to_u64<<<grid_for(n_slices + 1, 256), 256, 0, st>>>(slice_width.p, n_slices, width64.p);
    CKL();
    {
      size_t tmp_bytes = 0;
      CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, width64.p, slice_off.p, (long long)(n_slices + 1), st));
      DBuf<uint8_t> tmp;
      CK(tmp.alloc(tmp_bytes));
      CK(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, width64.p, slice_off.p, (long long)(n_slices + 1), st));
    }
    // rows whose table overflowed, and the items that touch them (multiplied matrix-free every iteration)
    auto select_u8 = [&](const uint8_t* flags, uint64_t n, uint32_t* out, uint32_t* d_count) -> int {
      cub::CountingInputIterator<uint32_t> iota(0u);
      size_t tmp_bytes = 0;
      CK(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, flags, out, d_count, (long long)n, st));
      DBuf<uint8_t> tmp;
      CK(tmp.alloc(tmp_bytes));
      CK(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, iota, flags, out, d_count, (long long)n, st));
      return AOBAKE_OK;
    };
    DBuf<uint32_t> d_cnt;
    CK(d_cnt.alloc(3));
    CK(cudaMemsetAsync(d_cnt.p, 0, 3 * sizeof(uint32_t), st));
    CK(over_rows.alloc(n_rows));
    int rc2;
    if ((rc2 = select_u8(row_over.p, n_rows, over_rows.p, d_cnt.p))) return rc2;
    uint32_t h_cnt[3] = {0, 0, 0};
    uint64_t total_width = 0;
    CK(cudaMemcpyAsync(h_cnt, d_cnt.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&total_width, slice_off.p + n_slices, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    n_over_rows = h_cnt[0];
    if (n_over_rows) {
      DBuf<uint8_t> tflag, eflag;
      DBuf<uint32_t> tsel, esel;
      CK(tflag.alloc(std::max<uint64_t>(a_tri, 1))); CK(eflag.alloc(std::max<uint64_t>(a_edge, 1)));
      CK(tsel.alloc(std::max<uint64_t>(a_tri, 1))); CK(esel.alloc(std::max<uint64_t>(a_edge, 1)));
      k_ls_flag_over_items<<<grid_for(std::max<uint64_t>(std::max(a_tri, a_edge), 1), 256), 256, 0, st>>>(a_tri_list, a_tri, a_edge_list, a_edge, gtris.p, edges.p, r0,
                                                                                                    r1, row_over.p, tflag.p, eflag.p);
      CKL();
      if (a_tri && (rc2 = select_u8(tflag.p, a_tri, tsel.p, d_cnt.p + 1))) return rc2;
      if (a_edge && (rc2 = select_u8(eflag.p, a_edge, esel.p, d_cnt.p + 2))) return rc2;
      CK(cudaMemcpyAsync(h_cnt, d_cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      n_over_tris = h_cnt[1]; n_over_edges = h_cnt[2];
      CK(over_tris.alloc(std::max(n_over_tris, 1u))); CK(over_edges.alloc(std::max(n_over_edges, 1u)));
      if (n_over_tris) k_compose_u32<<<grid_for(n_over_tris, 256), 256, 0, st>>>(a_tri_list, tsel.p, n_over_tris, over_tris.p);
      if (n_over_edges) k_compose_u32<<<grid_for(n_over_edges, 256), 256, 0, st>>>(a_edge_list, esel.p, n_over_edges, over_edges.p);
      CKL();
    }
    CK(sell_cols.alloc(std::max<uint64_t>(total_width * 32, 1))); CK(sell_vals.alloc(std::max<uint64_t>(total_width * 32, 1)));
    k_ls_fill_sell<<<grid_for(n_rows, 256), 256, 0, st>>>(hkeys.p, hvals.p, row_over.p, n_rows, r0, slice_off.p, slice_width.p, sell_cols.p, sell_vals.p);
    CKL();
    CK(cudaStreamSynchronize(st));   // the tables are released here
  }
  ctx->stats.reserved[3] = use_matrix ? 1 : 0;
  ctx->stats.reserved[4] = (int32_t)n_over_rows;
  const unsigned vec_grid = std::min<unsigned>(grid_for(std::max<uint64_t>(n_rows, 1), 256), (unsigned)ctx->sm_count * 8u);
  const uint64_t nwork = rows ? std::max<uint64_t>(n_tri_mine, n_edge_mine) : std::max<uint64_t>(NT, NE);
  DBuf<unsigned int> done_blocks;
  CK(done_blocks.alloc(1));
  CK(cudaMemsetAsync(done_blocks.p, 0, sizeof(unsigned int), st));
  CK(cudaMemsetAsync(scal.p, 0, 12 * sizeof(double), st));
  int rc;
  if (n_rows) {
    k_dot<<<vec_grid, 256, 0, st>>>(rhs.p + r0, rhs.p + r0, n_rows, scal.p + 0);
    k_dot<<<vec_grid, 256, 0, st>>>(r.p + r0, z.p + r0, n_rows, scal.p + 1);
  }
  if ((rc = all_sum(scal.p, 2))) return rc;
  double hs[4];
  CK(cudaMemcpyAsync(hs, scal.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const double bnorm2 = hs[0];
  int it = 0;
  if (NV && bnorm2 > 0.0) {
    const double tol2 = (double)ctx->params.cg_tolerance * (double)ctx->params.cg_tolerance * bnorm2;
    double rr = bnorm2;
    // the host looks at |r|^2 every kCheck iterations: up to kCheck - 1 iterations past the tolerance, no round trip per iteration
    constexpr int kCheck = 8;
    while (it < ctx->params.cg_max_iterations && rr > tol2) {
      const int burst = std::min(kCheck, ctx->params.cg_max_iterations - it);
      for (int k = 0; k < burst; k++, it++) {
        double* bank = scal.p + 4 + 4 * (it & 1);
        double* other = scal.p + 4 + 4 * ((it + 1) & 1);
        if (use_matrix) {
          k_ls_spmv_sell<<<grid_for(n_rows, 256), 256, 0, st>>>(slice_off.p, slice_width.p, sell_cols.p, sell_vals.p, n_rows, r0, p.p, Ap.p, bank);
          if (n_over_rows) {   // the flagged rows (empty in the matrix): matrix-free from the items that touch them
            const uint32_t nw = std::max(n_over_tris, n_over_edges);
            if (nw) k_ls_apply_rows<<<grid_for(nw, 256), 256, 0, st>>>(over_tris.p, n_over_tris, over_edges.p, n_over_edges, gtris.p, Mt.p, edges.p, w, r0, r1, row_over.p, p.p, Ap.p);
            k_ls_pap_list<<<1, 256, 0, st>>>(over_rows.p, n_over_rows, r0, fixed.p, p.p, Ap.p, bank);
          }
        } else {
          if (nwork) {
            if (rows) k_ls_apply_rows<<<grid_for(nwork, 256), 256, 0, st>>>(tri_list.p, n_tri_mine, edge_list.p, n_edge_mine, gtris.p, Mt.p, edges.p, w, r0, r1, nullptr, p.p, Ap.p);
            else k_ls_apply<<<grid_for(nwork, 256), 256, 0, st>>>(gtris.p, NT, Mt.p, edges.p, (uint32_t)NE, w, p.p, Ap.p);
          }
          k_ls_pap<<<vec_grid, 256, 0, st>>>(fixed.p + r0, p.p + r0, Ap.p + r0, n_rows, bank);
        }
        if ((rc = all_sum(bank, 1))) return rc;
        k_ls_update<<<vec_grid, 256, 0, st>>>(scal.p, p.p + r0, Ap.p + r0, diag.p + r0, x.p + r0, r.p + r0, z.p + r0, n_rows, bank);
        if ((rc = all_sum(bank + 1, 2))) return rc;
        k_ls_dir<<<vec_grid, 256, 0, st>>>(scal.p, bank, other, z.p + r0, p.p + r0, Ap.p + r0, n_rows, done_blocks.p);
        if ((rc = exchange_boundary())) return rc;
      }
      CKL();
      CK(cudaMemcpyAsync(hs, scal.p, sizeof(hs), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (!(hs[3] > 0.0)) return ctx->fail(AOBAKE_ERR_SOLVER, "CG breakdown: p.Ap = %g near iteration %d", hs[3], it);
      rr = hs[2];
    }
    if (rr > tol2) return ctx->fail(AOBAKE_ERR_SOLVER, "CG did not reach %g in %d iterations (|r|/|b| = %g)", (double)ctx->params.cg_tolerance, it, sqrt(rr / bnorm2));
  } else if (NV) {
    CK(cudaMemsetAsync(x.p, 0, NV * sizeof(double), st));
  }
  ctx->timings.cg_iterations = it;
  if (NV) {
    if (rows) k_d2f_rows<<<grid_for(NV, 256), 256, 0, st>>>(x.p, d_out.p, NV, r0, r1);
    else k_d2f<<<grid_for(NV, 256), 256, 0, st>>>(x.p, d_out.p, NV);
  }
  CKL();
  CK(cudaStreamSynchronize(st));
  return AOBAKE_OK;
}

}  // namespace

extern "C" {

static int map_ao_impl(AoBake* ctx, int mode, float weight, float* const* host_vertex_ao, bool distributed) {
  if (!ctx || !host_vertex_ao) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!ctx->have_scene || !ctx->have_ao) return ctx->fail(AOBAKE_ERR_STATE, "map_ao_to_vertices needs a scene and AO values");
  if (ctx->per_instance.size() != ctx->insts.size()) return ctx->fail(AOBAKE_ERR_STATE, "per-instance sample counts unknown (pass them to set_samples)");
  if (ctx->num_samples && !ctx->have_infos) return ctx->fail(AOBAKE_ERR_STATE, "sample_infos were not uploaded (set_samples was given a null sample_infos)");
  if (mode != AOBAKE_FILTER_AREA_BASED && mode != AOBAKE_FILTER_LEAST_SQUARES) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "invalid filter mode %d", mode);
  const bool dist = distributed && ctx->comm_size > 1;
  if (dist && !ctx->nccl_comm) return ctx->fail(AOBAKE_ERR_STATE, "aobake_comm_init has not been called");
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  CK(cudaEventRecord(ctx->ev0, st));
  ctx->timings.cg_iterations = 0;
  const uint32_t nI = (uint32_t)ctx->insts.size();
  // global vertex offsets of all instances
  std::vector<uint64_t> vb(nI + 1, 0);
  for (uint32_t i = 0; i < nI; i++) vb[i + 1] = vb[i] + ctx->meshes[ctx->insts[i].mesh].nV;
  const uint64_t NVall = vb[nI];
  // this rank's contiguous share of the instances, balanced by vertex count
  uint32_t ib = 0, ie = nI;
  if (dist) {
    auto cut = [&](int r) -> uint32_t {
      const uint64_t target = NVall * (uint64_t)r / (uint64_t)ctx->comm_size;
      return (uint32_t)(std::lower_bound(vb.begin(), vb.end(), target) - vb.begin());
    };
    ib = std::min(cut(ctx->comm_rank), nI);
    ie = ctx->comm_rank + 1 == ctx->comm_size ? nI : std::min(cut(ctx->comm_rank + 1), nI);
    if (ie < ib) ie = ib;
  }
  // fewer instances than ranks (the single big mesh of configs 3/5): one system, its rows split over the ranks
  const bool rows = dist && mode == AOBAKE_FILTER_LEAST_SQUARES && nI < (uint32_t)ctx->comm_size && NVall > 0;
  if (rows) { ib = 0; ie = nI; }
  DBuf<float> d_sub;
  int rc = mode == AOBAKE_FILTER_LEAST_SQUARES ? ls_filter_batched(ctx, weight, ib, ie, d_sub, rows) : area_filter_batched(ctx, ib, ie, d_sub);
  if (dist) rc = comm_agree(ctx, rc);   // e.g. a CG breakdown on one rank must not strand the others in the all-reduce
  if (rc) return rc;
  const float* d_all = d_sub.p;   // vertex AO of every instance, global numbering
  DBuf<float> d_full;
  if (dist) {
    // block-diagonal systems: every rank solved its own instances; zeros elsewhere + sum = gather
    CK(d_full.alloc(std::max<uint64_t>(NVall, 1)));
    CK(cudaMemsetAsync(d_full.p, 0, std::max<uint64_t>(NVall, 1) * sizeof(float), st));
    const uint64_t nsub = vb[ie] - vb[ib];
    if (nsub) CK(cudaMemcpyAsync(d_full.p + vb[ib], d_sub.p, nsub * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (NVall) {
      const int nrc = g_nccl.AllReduce(d_full.p, d_full.p, NVall, kNcclFloat, kNcclSum, ctx->nccl_comm, st);
      if (nrc != 0) return ctx->fail(AOBAKE_ERR_COMM, "ncclAllReduce: %s", g_nccl.GetErrorString(nrc));
    }
    d_all = d_full.p;
  }
  for (uint32_t i = 0; i < nI; i++) {
    const uint64_t nv = vb[i + 1] - vb[i];
    if (nv && host_vertex_ao[i]) CK(cudaMemcpyAsync(host_vertex_ao[i], d_all + vb[i], nv * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaEventRecord(ctx->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->timings.filter_ms, ctx->ev0, ctx->ev1));
  return AOBAKE_OK;
}

int aobake_map_ao_to_vertices(AoBake* ctx, int mode, float weight, float* const* host_vertex_ao) {
  return map_ao_impl(ctx, mode, weight, host_vertex_ao, false);
}

int aobake_map_ao_to_vertices_distributed(AoBake* ctx, int mode, float weight, float* const* host_vertex_ao) {
  return map_ao_impl(ctx, mode, weight, host_vertex_ao, true);
}

int aobake_make_ground_plane(const float bbox_min[3], const float bbox_max[3], int upaxis, float scale_factor, float offset_factor,
                             float* verts, uint32_t* tris) {
  if (!bbox_min || !bbox_max || !verts || !tris || upaxis < 0 || upaxis > 5) return AOBAKE_ERR_INVALID_ARGUMENT;
  const int axis = upaxis % 3, a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
  const bool flip = upaxis >= 3;
  float ext = 0.0f;
  for (int k = 0; k < 3; k++) ext = std::max(ext, bbox_max[k] - bbox_min[k]);
  const float h = flip ? bbox_max[axis] + offset_factor * ext : bbox_min[axis] - offset_factor * ext;
  const float c1 = 0.5f * (bbox_min[a1] + bbox_max[a1]), c2 = 0.5f * (bbox_min[a2] + bbox_max[a2]);
  const float h1 = 0.5f * scale_factor * (bbox_max[a1] - bbox_min[a1]), h2 = 0.5f * scale_factor * (bbox_max[a2] - bbox_min[a2]);
  const float s1[4] = {-1, 1, 1, -1}, s2[4] = {-1, -1, 1, 1};
  for (int i = 0; i < 4; i++) {
    verts[3 * i + axis] = h;
    verts[3 * i + a1] = c1 + s1[i] * h1;
    verts[3 * i + a2] = c2 + s2[i] * h2;
  }
  const uint32_t up[6] = {0, 1, 2, 0, 2, 3}, dn[6] = {0, 2, 1, 0, 3, 2};
  for (int i = 0; i < 6; i++) tris[i] = flip ? dn[i] : up[i];
  return AOBAKE_OK;
}

int aobake_trace_rays(AoBake* ctx, const float* rays, size_t n, uint8_t* hit) {
  if (!ctx || (n && (!rays || !hit))) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (!ctx->have_scene || !ctx->have_bvh) return ctx->fail(AOBAKE_ERR_STATE, "trace_rays before set_scene (aobake_set_scene_geometry builds no BVH)");
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  if (!n) return AOBAKE_OK;
  cudaStream_t st = ctx->stream;
  DBuf<float> d_rays;
  DBuf<uint8_t> d_hit;
  CK(d_rays.alloc(8 * n)); CK(d_hit.alloc(n));
  CK(cudaMemcpyAsync(d_rays.p, rays, 32 * n, cudaMemcpyHostToDevice, st));
  if (AOB_H2 != 0 && ctx->params.node_test != 1) k_trace_rays<true><<<grid_for(n, 128), 128, 0, st>>>(bvh_view(ctx), d_rays.p, n, d_hit.p);
  else k_trace_rays<false><<<grid_for(n, 128), 128, 0, st>>>(bvh_view(ctx), d_rays.p, n, d_hit.p);
  CKL();
  CK(cudaMemcpyAsync(hit, d_hit.p, n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return AOBAKE_OK;
}

int aobake_dump_rays(AoBake* ctx, size_t begin, size_t end, int rays_per_sample, float offset, float maxdist, float* out) {
  if (!ctx || !out) return AOBAKE_ERR_INVALID_ARGUMENT;
  if (begin > end || end > ctx->num_samples) return ctx->fail(AOBAKE_ERR_INVALID_ARGUMENT, "sample range outside the resident samples");
  ScopedTimer tm(ctx);
  CK(cudaSetDevice(ctx->device));
  const int q = sqrt_rays(rays_per_sample);
  const uint64_t n = (uint64_t)(end - begin) * q * q;
  if (!n) return AOBAKE_OK;
  cudaStream_t st = ctx->stream;
  DBuf<float> d_rays;
  CK(d_rays.alloc(8 * n));
  SampleView S{ctx->d_pos.p, ctx->d_nrm.p, ctx->d_fnrm.p};
  k_dump_rays<<<grid_for(n, 256), 256, 0, st>>>(S, begin, end, q, offset, maxdist, d_rays.p);
  CKL();
  CK(cudaMemcpyAsync(out, d_rays.p, 32 * n, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return AOBAKE_OK;
}

int aobake_get_timings(AoBake* ctx, AoTimings* out) {
  if (!ctx || !out) return AOBAKE_ERR_INVALID_ARGUMENT;
  *out = ctx->timings;
  return AOBAKE_OK;
}
int aobake_get_stats(AoBake* ctx, AoStats* out) {
  if (!ctx || !out) return AOBAKE_ERR_INVALID_ARGUMENT;
  *out = ctx->stats;
  return AOBAKE_OK;
}

}  // extern "C"
