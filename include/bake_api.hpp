// bake_api.hpp — the reference's C++ bake API (namespace bake of bake_api.h)
// re-created header-only on top of the C-ABI of aobake.h, so code written against
// optix_prime_baking's bake_api.h compiles against libaobake.so with an include swap.
//
//   reference (recalled; sources absent)            here
//   bake::Mesh / Instance / Scene / SampleInfo / AOSamples     layout-compatible aliases
//   bake::distributeSamples / sampleInstances / computeAO /    same names, same argument order
//   mapAOToVertices
//   allocate_ao_samples / destroy_ao_samples (bake_util.h)     same
//
// Error behaviour: the reference's functions return void and abort through assert/C++
// exceptions from optix_primepp; here every failure throws bake::Error carrying the C-ABI
// status and message.  `bake::Context` is the resident alternative (scene/BVH/samples stay
// in HBM across calls); the free functions each create and destroy a context, exactly as the
// reference rebuilt its OptiX Prime context per computeAO call.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "aobake.h"

namespace bake {

using Mesh = ::AoMesh;
using Instance = ::AoInstance;
using Scene = ::AoScene;
using SampleInfo = ::AoSampleInfo;
using AOSamples = ::AoSamples;

enum VertexFilterMode {
  VERTEX_FILTER_AREA_BASED = AOBAKE_FILTER_AREA_BASED,
  VERTEX_FILTER_LEAST_SQUARES = AOBAKE_FILTER_LEAST_SQUARES,
  VERTEX_FILTER_INVALID = AOBAKE_FILTER_INVALID
};

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error("aobake status " + std::to_string(s) + ": " + m), status(s) {}
};

class Context {
 public:
  explicit Context(int device = 0, const AoBakeParams* params = nullptr) {
    AoBakeParams p;
    aobake_default_params(&p);
    if (params) p = *params;
    else p.device = device;
    int rc = aobake_create(&p, &ctx_);
    if (rc != AOBAKE_OK) throw Error(rc, aobake_last_error(nullptr));
  }
  ~Context() { aobake_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  AoBake* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != AOBAKE_OK) throw Error(rc, aobake_last_error(ctx_));
  }

 private:
  AoBake* ctx_ = nullptr;
};

// bake_util.cpp
inline void allocate_ao_samples(AOSamples& s, size_t n) {
  s.num_samples = n;
  s.sample_positions = new float[3 * n];
  s.sample_normals = new float[3 * n];
  s.sample_face_normals = new float[3 * n];
  s.sample_infos = new SampleInfo[n];
}
inline void destroy_ao_samples(AOSamples& s) {
  delete[] s.sample_positions;
  delete[] s.sample_normals;
  delete[] s.sample_face_normals;
  delete[] s.sample_infos;
  s = AOSamples{};
}

// bake_sample.cpp
inline size_t distributeSamples(const Scene& scene, size_t min_samples_per_triangle, size_t requested_num_samples,
                                size_t* num_samples_per_instance) {
  Context c;
  c.check(aobake_set_scene_geometry(c.get(), &scene));   // no ray is traced here: no BVH
  size_t total = 0;
  c.check(aobake_distribute_samples(c.get(), min_samples_per_triangle, requested_num_samples, num_samples_per_instance, &total));
  return total;
}
inline void sampleInstances(const Scene& scene, const size_t* num_samples_per_instance, size_t min_samples_per_triangle,
                            AOSamples& ao_samples) {
  Context c;
  c.check(aobake_set_scene_geometry(c.get(), &scene));
  c.check(aobake_sample_instances(c.get(), num_samples_per_instance, min_samples_per_triangle, &ao_samples));
}
// bake_ao_optix_prime.cpp
inline void computeAO(const Scene& scene, const Scene& blockers, const AOSamples& ao_samples, int rays_per_sample,
                      float scene_offset, float scene_maxdistance, float* ao_values) {
  Context c;
  c.check(aobake_set_scene(c.get(), &scene, blockers.num_instances ? &blockers : nullptr));
  c.check(aobake_set_samples(c.get(), &ao_samples, nullptr));
  c.check(aobake_compute_ao(c.get(), rays_per_sample, scene_offset, scene_maxdistance, ao_values));
}
// bake_filter.cpp / bake_filter_least_squares.cpp
inline void mapAOToVertices(const Scene& scene, const size_t* num_samples_per_instance, const AOSamples& ao_samples,
                            const float* ao_values, VertexFilterMode mode, float regularization_weight, float** vertex_ao) {
  Context c;
  c.check(aobake_set_scene_geometry(c.get(), &scene));
  c.check(aobake_set_samples(c.get(), &ao_samples, num_samples_per_instance));
  c.check(aobake_set_ao(c.get(), ao_values));
  c.check(aobake_map_ao_to_vertices(c.get(), (int)mode, regularization_weight, vertex_ao));
}

}  // namespace bake
