// test_bake_api.cpp — a caller written against the reference's bake:: API (bake_api.h order of
// calls in main.cpp: distributeSamples -> sampleInstances -> computeAO -> mapAOToVertices),
// compiled with plain g++ against include/bake_api.hpp + libaobake.so.  Checks the analytic
// known answer AO(n) = (1 + n.up)/2 for a sphere over a huge ground plane.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "bake_api.hpp"

int main() {
  const int stacks = 32, slices = 32;
  std::vector<float> v, n;
  std::vector<unsigned> t;
  auto push = [&](double x, double y, double z) { v.push_back((float)x); v.push_back((float)y); v.push_back((float)z); n.push_back((float)x); n.push_back((float)y); n.push_back((float)z); };
  push(0, 1, 0);
  for (int k = 1; k < stacks; k++)
    for (int j = 0; j < slices; j++) {
      double th = M_PI * k / stacks, ph = 2 * M_PI * j / slices;
      push(sin(th) * cos(ph), cos(th), sin(th) * sin(ph));
    }
  push(0, -1, 0);
  auto ring = [&](int k) { return 1 + k * slices; };
  for (int j = 0; j < slices; j++) { t.push_back(0); t.push_back(ring(0) + (j + 1) % slices); t.push_back(ring(0) + j); }
  for (int k = 0; k < stacks - 2; k++)
    for (int j = 0; j < slices; j++) {
      unsigned a = ring(k) + j, b = ring(k) + (j + 1) % slices, c = ring(k + 1) + j, d = ring(k + 1) + (j + 1) % slices;
      t.insert(t.end(), {a, b, d, a, d, c});
    }
  unsigned south = (unsigned)(v.size() / 3 - 1);
  for (int j = 0; j < slices; j++) { t.push_back(south); t.push_back(ring(stacks - 2) + j); t.push_back(ring(stacks - 2) + (j + 1) % slices); }

  bake::Mesh mesh{};
  mesh.num_vertices = v.size() / 3; mesh.vertices = v.data(); mesh.vertex_stride_bytes = 12;
  mesh.normals = n.data(); mesh.normal_stride_bytes = 12;
  mesh.num_triangles = t.size() / 3; mesh.tri_vertex_indices = t.data();
  bake::Instance inst{};
  const float I[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(inst.xform, I, sizeof(I));
  inst.mesh_index = 0;
  bake::Scene scene{&mesh, 1, &inst, 1};

  float gv[12]; unsigned gt[6];
  const float lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
  if (aobake_make_ground_plane(lo, hi, 1, 1.0e6f, 0.03f, gv, gt) != AOBAKE_OK) return 2;
  bake::Mesh gm{};
  gm.num_vertices = 4; gm.vertices = gv; gm.num_triangles = 2; gm.tri_vertex_indices = gt;
  bake::Instance gi = inst;
  bake::Scene blockers{&gm, 1, &gi, 1};

  try {
    size_t per_instance[1];
    const size_t total = bake::distributeSamples(scene, 3, 0, per_instance);
    bake::AOSamples samples{};
    bake::allocate_ao_samples(samples, total);
    bake::sampleInstances(scene, per_instance, 3, samples);
    std::vector<float> ao(total);
    bake::computeAO(scene, blockers, samples, 256, 0.02f, 1.0e7f, ao.data());
    std::vector<float> vao(mesh.num_vertices), vls(mesh.num_vertices);
    float* out[1] = {vao.data()};
    bake::mapAOToVertices(scene, per_instance, samples, ao.data(), bake::VERTEX_FILTER_AREA_BASED, 0.1f, out);
    out[0] = vls.data();
    bake::mapAOToVertices(scene, per_instance, samples, ao.data(), bake::VERTEX_FILTER_LEAST_SQUARES, 0.1f, out);
    double err = 0, maxerr = 0, verr = 0;
    for (size_t i = 0; i < total; i++) {
      double e = std::fabs(ao[i] - (1.0 + samples.sample_normals[3 * i + 1]) / 2.0);
      err += e; maxerr = std::max(maxerr, e);
    }
    for (size_t i = 0; i < mesh.num_vertices; i++) verr = std::max(verr, (double)std::fabs(vao[i] - (1.0 + n[3 * i + 1]) / 2.0));
    double lsdiff = 0;
    for (size_t i = 0; i < mesh.num_vertices; i++) lsdiff = std::max(lsdiff, (double)std::fabs(vao[i] - vls[i]));
    printf("samples %zu mean|err| %.4f max|err| %.4f vertex max|err| %.4f |ls-area| %.4f\n", total, err / total, maxerr, verr, lsdiff);
    bake::destroy_ao_samples(samples);
    if (total != 3 * mesh.num_triangles || err / total > 8e-3 || maxerr > 0.1 || verr > 0.06 || lsdiff > 0.1) return 1;
    // error behaviour: a bad mesh index throws
    bake::Instance bad = inst;
    bad.mesh_index = 7;
    bake::Scene broken{&mesh, 1, &bad, 1};
    bool threw = false;
    try { bake::distributeSamples(broken, 3, 0, per_instance); } catch (const bake::Error& e) { threw = e.status == AOBAKE_ERR_INVALID_ARGUMENT; }
    if (!threw) return 3;
  } catch (const bake::Error& e) {
    fprintf(stderr, "bake::Error: %s\n", e.what());
    return 4;
  }
  printf("bake_api ok\n");
  return 0;
}
