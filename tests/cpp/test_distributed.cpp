// test_distributed.cpp — the multi-GPU bake with no Python anywhere: one host thread per GPU,
// each with its own AoBake context; libaobake.so does the sharding (interleaved super-blocks) and
// the NCCL all-reduce itself.  Checks that every rank ends up with the AO array of a single-GPU bake,
// bit for bit.  usage: test_distributed [num_gpus]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "aobake.h"

static void make_sphere(std::vector<float>& v, std::vector<float>& n, std::vector<unsigned>& t, int stacks, int slices) {
  auto push = [&](double x, double y, double z) { v.insert(v.end(), {(float)x, (float)y, (float)z}); n.insert(n.end(), {(float)x, (float)y, (float)z}); };
  push(0, 1, 0);
  for (int k = 1; k < stacks; k++)
    for (int j = 0; j < slices; j++) { double th = M_PI * k / stacks, ph = 2 * M_PI * j / slices; push(sin(th) * cos(ph), cos(th), sin(th) * sin(ph)); }
  push(0, -1, 0);
  auto ring = [&](int k) { return 1 + k * slices; };
  for (int j = 0; j < slices; j++) t.insert(t.end(), {0u, (unsigned)(ring(0) + (j + 1) % slices), (unsigned)(ring(0) + j)});
  for (int k = 0; k < stacks - 2; k++)
    for (int j = 0; j < slices; j++) {
      unsigned a = ring(k) + j, b = ring(k) + (j + 1) % slices, c = ring(k + 1) + j, d = ring(k + 1) + (j + 1) % slices;
      t.insert(t.end(), {a, b, d, a, d, c});
    }
  unsigned south = (unsigned)(v.size() / 3 - 1);
  for (int j = 0; j < slices; j++) t.insert(t.end(), {south, (unsigned)(ring(stacks - 2) + j), (unsigned)(ring(stacks - 2) + (j + 1) % slices)});
}

int main(int argc, char** argv) {
  const int world = argc > 1 ? atoi(argv[1]) : 2;
  std::vector<float> v, n;
  std::vector<unsigned> t;
  make_sphere(v, n, t, 96, 96);
  AoMesh mesh{};
  mesh.num_vertices = v.size() / 3; mesh.vertices = v.data(); mesh.normals = n.data();
  mesh.num_triangles = t.size() / 3; mesh.tri_vertex_indices = t.data();
  std::vector<AoInstance> insts(3);
  for (int i = 0; i < 3; i++) {
    memset(&insts[i], 0, sizeof(AoInstance));
    const float I4[16] = {1, 0, 0, 2.5f * i, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(insts[i].xform, I4, sizeof(I4));
  }
  AoScene scene{&mesh, 1, insts.data(), 3};
  float gv[12]; unsigned gt[6];
  const float lo[3] = {-1, -1, -1}, hi[3] = {6, 1, 1};
  aobake_make_ground_plane(lo, hi, 1, 100.f, 0.03f, gv, gt);
  AoMesh gm{}; gm.num_vertices = 4; gm.vertices = gv; gm.num_triangles = 2; gm.tri_vertex_indices = gt;
  AoInstance gi = insts[0]; gi.xform[3] = 0.f;
  AoScene blockers{&gm, 1, &gi, 1};

  char id[AOBAKE_COMM_ID_BYTES];
  if (aobake_comm_unique_id(id) != AOBAKE_OK) { fprintf(stderr, "unique id: %s\n", aobake_last_error(nullptr)); return 2; }

  std::vector<std::vector<float>> ao(world);
  std::vector<int> status(world, -1);
  std::vector<size_t> totals(world, 0);
  auto worker = [&](int rank) {
    AoBakeParams p; aobake_default_params(&p); p.device = rank;
    AoBake* ctx = nullptr;
    if (aobake_create(&p, &ctx) != AOBAKE_OK) { fprintf(stderr, "rank %d: %s\n", rank, aobake_last_error(nullptr)); return; }
    auto ok = [&](int rc, const char* what) { if (rc != AOBAKE_OK) fprintf(stderr, "rank %d %s: %s\n", rank, what, aobake_last_error(ctx)); return rc == AOBAKE_OK; };
    size_t per[3], total = 0;
    if (ok(aobake_set_scene(ctx, &scene, &blockers), "set_scene") && ok(aobake_distribute_samples(ctx, 3, 200003, per, &total), "distribute") &&
        ok(aobake_sample_instances(ctx, per, 3, nullptr), "sample") && ok(aobake_comm_init(ctx, rank, world, id), "comm_init")) {
      ao[rank].resize(total);
      totals[rank] = total;
      if (ok(aobake_compute_ao_distributed(ctx, 64, 0.07f, 70.f, ao[rank].data()), "compute_ao_distributed")) status[rank] = 0;
    }
    aobake_destroy(ctx);
  };
  std::vector<std::thread> th;
  for (int r = 0; r < world; r++) th.emplace_back(worker, r);
  for (auto& x : th) x.join();
  for (int r = 0; r < world; r++) if (status[r] != 0) { fprintf(stderr, "rank %d failed\n", r); return 1; }

  // single-GPU reference on device 0
  AoBake* ref = nullptr;
  aobake_create(nullptr, &ref);
  size_t per[3], total = 0;
  aobake_set_scene(ref, &scene, &blockers);
  aobake_distribute_samples(ref, 3, 200003, per, &total);
  aobake_sample_instances(ref, per, 3, nullptr);
  std::vector<float> want(total);
  aobake_compute_ao(ref, 64, 0.07f, 70.f, want.data());
  aobake_destroy(ref);
  for (int r = 0; r < world; r++) {
    if (totals[r] != total || memcmp(ao[r].data(), want.data(), total * sizeof(float)) != 0) { fprintf(stderr, "rank %d: AO differs from the single-GPU bake\n", r); return 1; }
  }
  double mean = 0;
  for (float x : want) mean += x;
  printf("distributed ok: %d ranks, %zu samples, mean AO %.5f\n", world, total, mean / total);
  return 0;
}
