/* aobake.h — C-ABI of libaobake.so: the B200-native drop-in for the optix_prime_baking
 * ambient-occlusion bake path.
 *
 * Boundary replaced: the four free functions and the POD types of the reference's
 * bake_api.h (namespace bake; SURVEY.md §8(a) rows a1–a5, a7, a9–a16 and §8(b)).  The
 * pre-deprecation sources are absent from /root/reference, so file names
 * below are as recalled and carry no line numbers.
 *
 *   reference (bake_api.h / bake_api.cpp)            this header
 *   ------------------------------------------------ ---------------------------------
 *   struct bake::Mesh / Instance / Scene             AoMesh / AoInstance / AoScene
 *   struct bake::SampleInfo / AOSamples              AoSampleInfo / AoSamples
 *   enum   bake::VertexFilterMode                    AoVertexFilterMode
 *   bake::distributeSamples  (bake_sample.cpp)       aobake_distribute_samples
 *   bake::sampleInstances    (bake_sample.cpp)       aobake_sample_instances
 *   bake::computeAO          (bake_ao_optix_prime.cpp + bake_kernels.cu + OptiX Prime
 *                             rtpModel.., rtpQuery..) aobake_set_scene + aobake_compute_ao
 *   bake::mapAOToVertices    (bake_filter.cpp,
 *                             bake_filter_least_squares.cpp)  aobake_map_ao_to_vertices
 *   make_ground_plane        (main.cpp)              aobake_make_ground_plane
 *   allocate/destroy_ao_samples (bake_util.cpp)      caller-owned host arrays, unchanged
 *
 * Differences from the reference API, all additive: a context (AoBake*) keeps the scene,
 * BVH, samples and AO resident in HBM between calls (the reference rebuilt the OptiX Prime
 * context inside every computeAO); every call returns a status (the reference returned void
 * and asserted); two parity hooks (aobake_trace_rays, aobake_dump_rays) expose the explicit
 * ray sets the oracle comparison needs; range/device entry points serve multi-GPU sharding.
 * include/bake_api.hpp re-creates the exact bake:: signatures on top of this header.
 *
 * All pointers are HOST memory unless a parameter name starts with d_.  Plain C, no CUDA or
 * torch types in any signature.  There is no CPU fallback: without a CUDA device
 * aobake_create fails with AOBAKE_ERR_NO_DEVICE.
 */
#ifndef AOBAKE_H_
#define AOBAKE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AOBAKE_VERSION 1

/* ---- POD scene / sample types (field-for-field bake_api.h) ------------------------ */
typedef struct AoMesh {
  uint64_t num_vertices;
  const float* vertices;          /* xyz, stride vertex_stride_bytes (0 => 12) */
  uint32_t vertex_stride_bytes;
  const float* normals;           /* nullable: face normals are used */
  uint32_t normal_stride_bytes;   /* 0 => 12 */
  uint64_t num_triangles;
  const uint32_t* tri_vertex_indices; /* 3 per triangle */
  float bbox_min[3];
  float bbox_max[3];
} AoMesh;

typedef struct AoInstance {
  float xform[16];                /* row-major 4x4, affine (bottom row ignored) */
  uint64_t storage_identifier;
  uint32_t mesh_index;
  float bbox_min[3];              /* world; informational */
  float bbox_max[3];
} AoInstance;

typedef struct AoScene {
  const AoMesh* meshes;
  uint64_t num_meshes;
  const AoInstance* instances;
  uint64_t num_instances;
} AoScene;

typedef struct AoSampleInfo {
  uint32_t tri_idx;
  float bary[3];
  float dA;
} AoSampleInfo;                   /* 20 bytes */

typedef struct AoSamples {
  uint64_t num_samples;
  float* sample_positions;        /* 3 * num_samples */
  float* sample_normals;          /* 3 * num_samples */
  float* sample_face_normals;     /* 3 * num_samples */
  AoSampleInfo* sample_infos;     /* num_samples */
} AoSamples;

typedef enum AoVertexFilterMode {
  AOBAKE_FILTER_AREA_BASED = 0,
  AOBAKE_FILTER_LEAST_SQUARES = 1,
  AOBAKE_FILTER_INVALID = 2
} AoVertexFilterMode;

typedef enum AoInstancingMode {
  AOBAKE_INSTANCING_AUTO = 0,      /* flatten unless some mesh has more than one instance */
  AOBAKE_INSTANCING_FLATTEN = 1,   /* one world-space BVH over every instance's triangles */
  AOBAKE_INSTANCING_TWO_LEVEL = 2  /* TLAS over instances, one BLAS per mesh */
} AoInstancingMode;

typedef enum AoStatus {
  AOBAKE_OK = 0,
  AOBAKE_ERR_INVALID_ARGUMENT = 1,
  AOBAKE_ERR_CUDA = 2,
  AOBAKE_ERR_STATE = 3,           /* call order violated (e.g. compute before set_scene) */
  AOBAKE_ERR_SAMPLE_OVERFLOW = 4, /* area-proportional floors exceeded the budget */
  AOBAKE_ERR_NO_DEVICE = 5,
  AOBAKE_ERR_SOLVER = 6,
  AOBAKE_ERR_COMM = 7             /* NCCL missing or a collective failed */
} AoStatus;

typedef struct AoBakeParams {
  int32_t device;                 /* CUDA ordinal */
  int32_t instancing_mode;        /* AoInstancingMode */
  int32_t cg_max_iterations;      /* least-squares filter */
  float   cg_tolerance;           /* relative residual */
  int32_t trace_kernel;           /* 0 = auto (persistent refilling kernel; simple kernel below 32 M rays); 1 = simple; 2 = persistent */
  int32_t collect_stats;          /* 1: count node visits / triangle tests in aobake_compute_ao */
  int32_t refill_below;           /* persistent kernel: refill a warp when fewer lanes are traversing (0 = default 28) */
  int32_t leaf_tris;              /* triangles per leaf slot of the 8-wide BVH, 1..3 (0 = default: 2 flattened, 1 per BLAS) */
  int32_t node_test;              /* box test of the fused kernel: 0 = auto (packed fp16, two planes per instruction; the few
                                     rays outside its range are traced by a second small launch in fp32), 1 = fp32 only */
  int32_t deferred_capacity;      /* entries of the deferred-ray list (0 = auto: 1/128 of a launch's rays); an overflow makes
                                     aobake_compute_ao repeat the launch with the fp32 kernels — settable so that tests can force it */
  int32_t tri_batch;              /* fused kernel: low byte = lanes of a warp that must hold leaf hits before the warp runs its
                                     triangle block (paused lanes take no node steps meanwhile; 1 = test at once); next byte =
                                     the most iterations a paused lane waits (0 = no limit); 0 = default (16 lanes flattened — 8 with
                                     ray_order 1 — 12 under a TLAS; 6 iterations) */
  int32_t no_oversized_split;     /* BVH build: 0 = primitives (or TLAS instances) spanning more than a quarter of the scene (a ground
                                     plane under a fine mesh) are kept out of the tree and hang off one extra root node; 1 = build
                                     one tree over everything (A/B switch; both give the same hits) */
  int32_t ls_energy;              /* least-squares regulariser per interior edge: 0 = (A1 + A2) |grad(T1) - grad(T2)|^2, the 3-D
                                     gradient jump of Kavan et al. 2011; 1 = (A1 + A2)^2 x (jump of the co-normal
                                     derivative)^2, the scale-free variant of round 1 (same null space, far better conditioned) */
  int32_t ls_matrix_free;         /* least-squares PCG product: 0 = A = M + wR assembled once (sliced ELL, no atomics per iteration;
                                     rows of vertices with too many neighbours stay matrix-free), 1 = matrix-free scatter */
  int32_t ray_order;              /* fused kernel, the order in which a warp traces the rays of its 32-sample work item: 0 = default
                                     (AOB_RAY_ORDER_DEFAULT of the build), 1 = sample-major (a lane owns a sample and walks its strata),
                                     2 = stratum-major (the warp deals out the item's rays stratum by stratum: its 32 rays start on
                                     neighbouring samples and come from one or two strata; whichever lane is free takes the next ray).
                                     Hit counts are identical either way. */
} AoBakeParams;

typedef struct AoTimings {        /* milliseconds, device-timed with CUDA events unless noted */
  float upload_ms;                /* set_scene: host->device copies */
  float bvh_build_ms;             /* set_scene: BVH construction */
  float sample_ms;                /* last sample_instances */
  float trace_ms;                 /* last compute_ao: fused raygen+traverse+accumulate kernel(s) */
  float filter_ms;                /* last map_ao_to_vertices */
  float host_total_ms;            /* wall clock of the last API call */
  uint64_t rays_traced;           /* last compute_ao */
  int32_t cg_iterations;          /* last least-squares solve (sum over instances) */
  int32_t kernel_launches;        /* kernels launched inside the last compute_ao's timed region */
  int32_t reserved[6];
} AoTimings;

typedef struct AoStats {
  uint64_t num_bvh_nodes;         /* 80-byte 8-wide nodes, all levels */
  uint64_t num_bvh_triangles;     /* 48-byte triangle records */
  uint64_t num_tlas_instances;    /* 0 in flatten mode */
  uint64_t bvh_bytes;
  uint64_t node_visits;           /* valid after a compute_ao with collect_stats = 1 */
  uint64_t triangle_tests;
  uint64_t instance_entries;
  uint64_t rays;
  int32_t two_level;
  int32_t reserved[7];            /* [0] = depth of the top-level 8-wide tree, [1] = deepest BLAS (two-level),
                                     [2] = rays the last compute_ao handed to the deferred fp32 launch,
                                     [3] = 1 if the last least-squares solve used the assembled matrix,
                                     [4] = rows of that solve multiplied matrix-free (more columns than the assembly holds) */
} AoStats;

typedef struct AoBake AoBake;

/* ---- lifecycle -------------------------------------------------------------------- */
int aobake_default_params(AoBakeParams* params);
int aobake_create(const AoBakeParams* params /*nullable*/, AoBake** out);
void aobake_destroy(AoBake* ctx);
/* Last error text for ctx (or for a failed aobake_create when ctx == NULL). */
const char* aobake_last_error(const AoBake* ctx);
/* Run all subsequent work on an existing CUDA stream (a cudaStream_t passed as void*);
 * NULL restores the context's own stream. */
int aobake_set_stream(AoBake* ctx, void* cuda_stream);
int aobake_synchronize(AoBake* ctx);

/* ---- the bake path ---------------------------------------------------------------- */
/* Uploads scene + blockers (nullable) and builds the BVH (bake_ao_optix_prime.cpp:
 * rtpModelSetTriangles/SetInstances/Update).  The reference asserts on bad input; here a
 * mesh index out of range, a triangle index >= num_vertices, or a singular / non-finite
 * instance transform returns AOBAKE_ERR_INVALID_ARGUMENT and leaves the context without a
 * scene.  Only the affine 3x4 part of xform is used (the fourth row is ignored). */
int aobake_set_scene(AoBake* ctx, const AoScene* scene, const AoScene* blockers);
/* Uploads the scene WITHOUT building a BVH: all that bake::distributeSamples, bake::sampleInstances and
 * bake::mapAOToVertices need (the one-shot bake:: shims of bake_api.hpp use it).  aobake_compute_ao and
 * aobake_trace_rays then return AOBAKE_ERR_STATE. */
int aobake_set_scene_geometry(AoBake* ctx, const AoScene* scene);

/* bake::distributeSamples.  per_instance has scene->num_instances entries. */
int aobake_distribute_samples(AoBake* ctx, size_t min_samples_per_triangle, size_t requested_num_samples,
                              size_t* per_instance, size_t* total);

/* bake::sampleInstances on the device; host_out nullable (samples stay resident either way).
 * When given, host_out->num_samples must equal sum(per_instance). */
int aobake_sample_instances(AoBake* ctx, const size_t* per_instance, size_t min_samples_per_triangle,
                            AoSamples* host_out);

/* Upload caller-made samples instead (what bake::computeAO receives).  per_instance nullable;
 * needed only for a later aobake_map_ao_to_vertices. */
int aobake_set_samples(AoBake* ctx, const AoSamples* host_samples, const size_t* per_instance);

/* bake::computeAO: ao[g] = 1 - occluded(g) / q^2, q = round(sqrt(rays_per_sample)).
 * host_ao nullable (num_samples floats); the result always stays resident. */
int aobake_compute_ao(AoBake* ctx, int rays_per_sample, float scene_offset, float scene_maxdistance,
                      float* host_ao);
/* Same for the global sample range [begin, end) only — the multi-GPU shard.  RNG streams are
 * functions of the global sample index, so shards reproduce the 1-GPU result bit for bit.
 * host_ao nullable, (end - begin) floats. */
int aobake_compute_ao_range(AoBake* ctx, size_t begin, size_t end, int rays_per_sample, float scene_offset,
                            float scene_maxdistance, float* host_ao);
/* Interleaved multi-GPU partition of the whole sample set: this call traces the super-blocks of
 * block_samples samples (multiple of 32; 0 => 16384) whose index % num_parts == part, and sets the
 * resident ao[] of every other sample to 0 — a sum all-reduce over the parts then assembles the
 * full array exactly.  Interleaving evens out regions of different traversal cost (the contiguous
 * ranges of aobake_compute_ao_range can differ by 10-15 % on a terrain). */
int aobake_compute_ao_interleaved(AoBake* ctx, uint32_t part, uint32_t num_parts, uint32_t block_samples,
                                  int rays_per_sample, float scene_offset, float scene_maxdistance);
/* ---- native multi-GPU exchange (NCCL, resolved at run time with dlopen("libnccl.so.2")) ----
 * One process (or thread) per GPU.  Rank 0 calls aobake_comm_unique_id and ships the 128 bytes to
 * the other ranks by any means; every rank then calls aobake_comm_init on its own context.
 * aobake_compute_ao_distributed = aobake_compute_ao_interleaved(rank, nranks) + one in-place
 * ncclAllReduce(sum) over the resident ao[] on the context's stream: afterwards every rank holds
 * the full AO array, bit-identical to a single-GPU bake.  host_ao nullable (num_samples floats). */
#define AOBAKE_COMM_ID_BYTES 128
int aobake_comm_unique_id(void* id128);
int aobake_comm_init(AoBake* ctx, int rank, int nranks, const void* id128);
int aobake_comm_destroy(AoBake* ctx);
int aobake_compute_ao_distributed(AoBake* ctx, int rays_per_sample, float scene_offset, float scene_maxdistance,
                                  float* host_ao);
/* The host-buffer forms of aobake_set_scene / aobake_set_samples for N ranks that all hold the same
 * host arrays (what bake::computeAO receives on every rank).  set_scene_distributed: rank r copies only
 * slice r of every vertex / index array over its own PCIe link and one in-place ncclAllGather per
 * array completes them over NVLink (N uploads of the whole scene become one); every rank then builds
 * the same BVH.  set_samples_distributed: only the super-blocks of 16384 samples this rank traces in
 * aobake_compute_ao_distributed are copied (positions / normals / face normals; sample_infos whole,
 * the vertex maps need them); afterwards only aobake_compute_ao_distributed may trace.  With one
 * rank both are the plain calls.  A rank that fails before a collective makes every rank return
 * AOBAKE_ERR_COMM instead of leaving the others blocked. */
int aobake_set_scene_distributed(AoBake* ctx, const AoScene* scene, const AoScene* blockers);
int aobake_set_samples_distributed(AoBake* ctx, const AoSamples* host_samples, const size_t* per_instance);
/* bake::mapAOToVertices across the ranks of aobake_comm_init: the per-instance systems are independent
 * (block diagonal), so each rank filters a contiguous share of the instances (balanced by vertex
 * count) and one ncclAllReduce over zero-padded arrays gathers the result; every rank receives the
 * vertex AO of every instance.  Requires the full AO array on every rank (compute_ao_distributed). */
int aobake_map_ao_to_vertices_distributed(AoBake* ctx, int mode /*AoVertexFilterMode*/, float regularization_weight,
                                          float* const* host_vertex_ao);

/* Device pointer to the resident ao[num_samples] array (for an NCCL all-gather by the caller). */
int aobake_get_ao_device(AoBake* ctx, float** d_ao, size_t* num_samples);
/* Replace the resident AO array from the host (num_samples floats). */
int aobake_set_ao(AoBake* ctx, const float* host_ao);

/* bake::mapAOToVertices.  host_vertex_ao[i] has meshes[instances[i].mesh_index].num_vertices
 * floats. */
int aobake_map_ao_to_vertices(AoBake* ctx, int mode /*AoVertexFilterMode*/, float regularization_weight,
                              float* const* host_vertex_ao);

/* make_ground_plane (main.cpp): upaxis 0..5 = +X,+Y,+Z,-X,-Y,-Z; writes 4 vertices (12
 * floats) and 2 triangles (6 indices).  Pure host helper. */
int aobake_make_ground_plane(const float bbox_min[3], const float bbox_max[3], int upaxis, float scale_factor,
                             float offset_factor, float* vertices12, uint32_t* indices6);

/* ---- parity hooks ----------------------------------------------------------------- */
/* Any-hit for explicit rays: n x 8 floats (o.xyz, tmin, d.xyz, tmax); hit[i] = 1 iff occluded. */
int aobake_trace_rays(AoBake* ctx, const float* rays, size_t n, uint8_t* hit);
/* The rays compute_ao would trace for samples [begin,end): (end-begin) * q*q * 8 floats,
 * stratum-major (px*q+py) inside a sample. */
int aobake_dump_rays(AoBake* ctx, size_t sample_begin, size_t sample_end, int rays_per_sample,
                     float scene_offset, float scene_maxdistance, float* rays_out);
/* Per-sample occluded-ray counts of the last compute_ao (num_samples uint32). */
int aobake_get_hit_counts(AoBake* ctx, uint32_t* host_counts);

/* ---- introspection ---------------------------------------------------------------- */
int aobake_get_timings(AoBake* ctx, AoTimings* out);
int aobake_get_stats(AoBake* ctx, AoStats* out);
size_t aobake_num_samples(const AoBake* ctx);

#ifdef __cplusplus
}
#endif
#endif /* AOBAKE_H_ */
