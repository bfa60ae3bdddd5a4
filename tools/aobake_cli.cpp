// aobake_cli.cpp — the main.cpp of the reference minus the viewer (SURVEY.md §8f "next" rows 1-3):
// load an OBJ scene (or a built-in sphere), optionally replicate it in a grid (-i), add the
// ground-plane blocker, run distributeSamples -> sampleInstances -> computeAO ->
// mapAOToVertices through the C-ABI of libaobake.so, print the stage timers the sample printed,
// and optionally write the raw per-instance vertex-AO dump (-o).
//
// Flags and defaults follow the reference's Config (main.cpp, recalled):
//   -f/--file <obj>  -o/--outfile <raw>  -i/--instances n (1)  -r/--rays n (64)
//   -s/--samples n (0 = per-face minimum only)  -t/--samples_per_face n (3)
//   -d/--ray_distance_scale s (0.01)  -m/--hit_distance_scale s (10)
//   --ray_distance d  --hit_distance d  -g/--ground_setup axis scale offset (1 100 0.03)
//   --no_ground_plane  -w/--regularization_weight w (0.1)  --no_least_squares
//   --flip_orientation   (--cpu is rejected: there is no CPU path)   --no_viewer accepted, ignored
//   --gpus n (1): one host thread per GPU; the AO pass is sharded (interleaved super-blocks) and
//                 all-reduced over NCCL inside libaobake.so, the vertex map runs on GPU 0
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "aobake.h"

namespace {

struct HostMesh {
  std::vector<float> v, n;
  std::vector<unsigned> t;
};

double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Minimal Wavefront OBJ reader: v / vn / f (v, v/vt, v//vn, v/vt/vn; polygons fan-triangulated;
// negative indices).  Normals are used only if every corner references the normal with the
// same index as its position would allow a per-vertex array; otherwise they are dropped and the
// baker falls back to face normals, as the reference does for meshes without normals.
bool load_obj(const std::string& path, HostMesh& m, bool flip) {
  std::ifstream in(path);
  if (!in) return false;
  std::vector<float> vn;
  std::vector<int> corner_v, corner_n;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    std::string tag;
    ss >> tag;
    if (tag == "v") { float x, y, z; ss >> x >> y >> z; m.v.insert(m.v.end(), {x, y, z}); }
    else if (tag == "vn") { float x, y, z; ss >> x >> y >> z; vn.insert(vn.end(), {x, y, z}); }
    else if (tag == "f") {
      std::vector<int> fv, fn;
      std::string c;
      while (ss >> c) {
        int vi = 0, ni = 0;
        size_t s1 = c.find('/');
        vi = std::atoi(c.substr(0, s1).c_str());
        if (s1 != std::string::npos) {
          size_t s2 = c.find('/', s1 + 1);
          if (s2 != std::string::npos && s2 + 1 < c.size()) ni = std::atoi(c.substr(s2 + 1).c_str());
        }
        const int nv = (int)(m.v.size() / 3), nn = (int)(vn.size() / 3);
        fv.push_back(vi < 0 ? nv + vi : vi - 1);
        fn.push_back(ni == 0 ? -1 : (ni < 0 ? nn + ni : ni - 1));
      }
      for (size_t k = 1; k + 1 < fv.size(); k++) {
        const size_t a = 0, b = flip ? k + 1 : k, c2 = flip ? k : k + 1;
        corner_v.insert(corner_v.end(), {fv[a], fv[b], fv[c2]});
        corner_n.insert(corner_n.end(), {fn[a], fn[b], fn[c2]});
      }
    }
  }
  const size_t nv = m.v.size() / 3, nn = vn.size() / 3;
  // validate every index BEFORE it is used (a face may reference vertices or normals declared later in the
  // file, so this cannot be done while parsing): a malformed OBJ is "could not load", not a heap overrun
  for (int c : corner_v)
    if (c < 0 || (size_t)c >= nv) return false;
  bool ok_normals = nn > 0;
  for (int n : corner_n) {
    if (n < 0) { ok_normals = false; continue; }   // a corner without a normal: fall back to face normals
    if ((size_t)n >= nn) return false;
  }
  if (ok_normals) {
    // per-vertex normal = normalised sum of the normals its corners reference (creases are averaged)
    m.n.assign(m.v.size(), 0.0f);
    for (size_t i = 0; i < corner_v.size(); i++)
      for (int k = 0; k < 3; k++) m.n[3 * (size_t)corner_v[i] + k] += (flip ? -1.0f : 1.0f) * vn[3 * (size_t)corner_n[i] + k];
    for (size_t v = 0; v < nv; v++) {
      float* p = &m.n[3 * v];
      const float l = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
      if (l > 0) { p[0] /= l; p[1] /= l; p[2] /= l; }
    }
  } else {
    m.n.clear();
  }
  for (int c : corner_v) m.t.push_back((unsigned)c);
  return !m.v.empty() && !m.t.empty();
}

void make_sphere(HostMesh& m, int stacks, int slices) {
  auto push = [&](double x, double y, double z) { m.v.insert(m.v.end(), {(float)x, (float)y, (float)z}); m.n.insert(m.n.end(), {(float)x, (float)y, (float)z}); };
  push(0, 1, 0);
  for (int k = 1; k < stacks; k++)
    for (int j = 0; j < slices; j++) {
      const double th = M_PI * k / stacks, ph = 2 * M_PI * j / slices;
      push(std::sin(th) * std::cos(ph), std::cos(th), std::sin(th) * std::sin(ph));
    }
  push(0, -1, 0);
  auto ring = [&](int k) { return 1 + k * slices; };
  for (int j = 0; j < slices; j++) m.t.insert(m.t.end(), {0u, (unsigned)(ring(0) + (j + 1) % slices), (unsigned)(ring(0) + j)});
  for (int k = 0; k < stacks - 2; k++)
    for (int j = 0; j < slices; j++) {
      const unsigned a = ring(k) + j, b = ring(k) + (j + 1) % slices, c = ring(k + 1) + j, d = ring(k + 1) + (j + 1) % slices;
      m.t.insert(m.t.end(), {a, b, d, a, d, c});
    }
  const unsigned south = (unsigned)(m.v.size() / 3 - 1);
  for (int j = 0; j < slices; j++) m.t.insert(m.t.end(), {south, (unsigned)(ring(stacks - 2) + j), (unsigned)(ring(stacks - 2) + (j + 1) % slices)});
}

struct Config {
  std::string file, outfile;
  int instances = 1, rays = 64, samples_per_face = 3;
  size_t samples = 0;
  float ray_distance_scale = 0.01f, hit_distance_scale = 10.0f, ray_distance = -1.0f, hit_distance = -1.0f;
  int ground_axis = 1;
  float ground_scale = 100.0f, ground_offset = 0.03f, weight = 0.1f;
  bool ground = true, least_squares = true, flip = false;
  int gpus = 1;
};

int usage(const char* argv0) {
  fprintf(stderr,
          "usage: %s [-f scene.obj] [-o out.raw] [-i n] [-r rays] [-s samples] [-t samples_per_face]\n"
          "          [-d ray_distance_scale] [-m hit_distance_scale] [--ray_distance d] [--hit_distance d]\n"
          "          [-g axis scale offset] [--no_ground_plane] [-w weight] [--no_least_squares] [--flip_orientation] [--gpus n]\n",
          argv0);
  return 2;
}

}  // namespace

int main(int argc, char** argv) {
  Config cfg;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&](const char* what) -> const char* {
      if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", what); exit(usage(argv[0])); }
      return argv[++i];
    };
    if (a == "-f" || a == "--file") cfg.file = next("-f");
    else if (a == "-o" || a == "--outfile") cfg.outfile = next("-o");
    else if (a == "-i" || a == "--instances") cfg.instances = std::max(1, atoi(next("-i")));
    else if (a == "-r" || a == "--rays") cfg.rays = atoi(next("-r"));
    else if (a == "-s" || a == "--samples") cfg.samples = (size_t)atoll(next("-s"));
    else if (a == "-t" || a == "--samples_per_face") cfg.samples_per_face = atoi(next("-t"));
    else if (a == "-d" || a == "--ray_distance_scale") cfg.ray_distance_scale = (float)atof(next("-d"));
    else if (a == "-m" || a == "--hit_distance_scale") cfg.hit_distance_scale = (float)atof(next("-m"));
    else if (a == "--ray_distance") cfg.ray_distance = (float)atof(next("--ray_distance"));
    else if (a == "--hit_distance") cfg.hit_distance = (float)atof(next("--hit_distance"));
    else if (a == "-g" || a == "--ground_setup") { cfg.ground_axis = atoi(next("-g")); cfg.ground_scale = (float)atof(next("-g")); cfg.ground_offset = (float)atof(next("-g")); }
    else if (a == "--no_ground_plane") cfg.ground = false;
    else if (a == "-w" || a == "--regularization_weight") cfg.weight = (float)atof(next("-w"));
    else if (a == "--no_least_squares") cfg.least_squares = false;
    else if (a == "--flip_orientation") cfg.flip = true;
    else if (a == "--gpus") cfg.gpus = std::max(1, atoi(next("--gpus")));
    else if (a == "--no_viewer" || a == "--conserve_memory") {}
    else if (a == "--cpu") { fprintf(stderr, "--cpu: libaobake has no CPU path (B200 only)\n"); return 2; }
    else if (a == "-h" || a == "--help") return usage(argv[0]);
    else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return usage(argv[0]); }
  }

  // ---- load ----
  double t0 = now_ms();
  HostMesh hm;
  if (!cfg.file.empty()) {
    if (!load_obj(cfg.file, hm, cfg.flip)) { fprintf(stderr, "could not load %s\n", cfg.file.c_str()); return 1; }
  } else {
    make_sphere(hm, 64, 64);
  }
  AoMesh mesh{};
  mesh.num_vertices = hm.v.size() / 3; mesh.vertices = hm.v.data(); mesh.vertex_stride_bytes = 12;
  mesh.normals = hm.n.empty() ? nullptr : hm.n.data(); mesh.normal_stride_bytes = 12;
  mesh.num_triangles = hm.t.size() / 3; mesh.tri_vertex_indices = hm.t.data();
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (size_t i = 0; i < mesh.num_vertices; i++)
    for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], hm.v[3 * i + k]); hi[k] = std::max(hi[k], hm.v[3 * i + k]); }
  memcpy(mesh.bbox_min, lo, sizeof(lo)); memcpy(mesh.bbox_max, hi, sizeof(hi));
  // -i n: replicate the mesh on an n-instance grid (loaders' num_instances_per_mesh)
  std::vector<AoInstance> insts(cfg.instances);
  const int side = (int)std::ceil(std::cbrt((double)cfg.instances));
  float slo[3] = {1e30f, 1e30f, 1e30f}, shi[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = 0; i < cfg.instances; i++) {
    AoInstance& I = insts[i];
    memset(&I, 0, sizeof(I));
    const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    memcpy(I.xform, I4, sizeof(I4));
    const int g[3] = {i % side, (i / side) % side, i / (side * side)};
    for (int k = 0; k < 3; k++) {
      I.xform[4 * k + 3] = 1.1f * (hi[k] - lo[k]) * (float)g[k];
      I.bbox_min[k] = lo[k] + I.xform[4 * k + 3]; I.bbox_max[k] = hi[k] + I.xform[4 * k + 3];
      slo[k] = std::min(slo[k], I.bbox_min[k]); shi[k] = std::max(shi[k], I.bbox_max[k]);
    }
    I.storage_identifier = (uint64_t)i;
    I.mesh_index = 0;
  }
  AoScene scene{&mesh, 1, insts.data(), (uint64_t)insts.size()};
  float extent = 0.0f;
  for (int k = 0; k < 3; k++) extent = std::max(extent, shi[k] - slo[k]);
  const float scene_offset = cfg.ray_distance >= 0 ? cfg.ray_distance : cfg.ray_distance_scale * extent;
  const float scene_maxdist = cfg.hit_distance >= 0 ? cfg.hit_distance : cfg.hit_distance_scale * extent;
  fprintf(stderr, "Load scene ... %.2f ms\n", now_ms() - t0);
  fprintf(stderr, "Number of meshes: 1\nNumber of instances: %d\nUninstanced vertices: %llu\nUninstanced triangles: %llu\n", cfg.instances,
          (unsigned long long)mesh.num_vertices, (unsigned long long)mesh.num_triangles);

  // ---- ground plane blocker ----
  float gv[12];
  unsigned gt[6];
  AoMesh gmesh{};
  AoInstance ginst = insts[0];
  const float I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  memcpy(ginst.xform, I4, sizeof(I4));
  AoScene blockers{nullptr, 0, nullptr, 0};
  if (cfg.ground) {
    if (aobake_make_ground_plane(slo, shi, cfg.ground_axis, cfg.ground_scale, cfg.ground_offset, gv, gt) != AOBAKE_OK) { fprintf(stderr, "bad ground setup\n"); return 2; }
    gmesh.num_vertices = 4; gmesh.vertices = gv; gmesh.num_triangles = 2; gmesh.tri_vertex_indices = gt;
    blockers = AoScene{&gmesh, 1, &ginst, 1};
  }

  AoBake* ctx = nullptr;
  if (aobake_create(nullptr, &ctx) != AOBAKE_OK) { fprintf(stderr, "%s\n", aobake_last_error(nullptr)); return 1; }
  auto ck = [&](int rc, const char* what) {
    if (rc != AOBAKE_OK) { fprintf(stderr, "%s failed: %s\n", what, aobake_last_error(ctx)); aobake_destroy(ctx); exit(1); }
  };
  t0 = now_ms();
  ck(aobake_set_scene(ctx, &scene, cfg.ground ? &blockers : nullptr), "set_scene");
  fprintf(stderr, "Upload scene + build BVH ... %.2f ms\n", now_ms() - t0);

  t0 = now_ms();
  std::vector<size_t> per(cfg.instances);
  size_t total = 0;
  ck(aobake_distribute_samples(ctx, (size_t)cfg.samples_per_face, cfg.samples, per.data(), &total), "distribute_samples");
  ck(aobake_sample_instances(ctx, per.data(), (size_t)cfg.samples_per_face, nullptr), "sample_instances");
  fprintf(stderr, "Minimum samples per face: %d\nTotal samples: %zu\nGenerate sample points ... %.2f ms\n", cfg.samples_per_face, total, now_ms() - t0);

  const int q = (int)(std::sqrt((float)cfg.rays) + 0.5f);
  fprintf(stderr, "Rays per sample: %d\nTotal rays: %zu\n", q * q, total * (size_t)q * q);
  t0 = now_ms();
  if (cfg.gpus > 1) {
    // helper ranks 1..n-1: same scene and samples (replicated), their share of the AO pass; rank 0 is ctx
    char id[AOBAKE_COMM_ID_BYTES];
    if (aobake_comm_unique_id(id) != AOBAKE_OK) { fprintf(stderr, "NCCL: %s\n", aobake_last_error(nullptr)); aobake_destroy(ctx); return 1; }
    std::vector<int> rcs(cfg.gpus, AOBAKE_OK);
    // Two phases, so that one failing rank cannot strand the others inside ncclCommInitRank: every helper
    // first creates its context, uploads the scene and samples; only if ALL of them succeeded does anyone
    // enter the communicator set-up.  (Inside the library a rank that fails later makes the collective
    // entry points return AOBAKE_ERR_COMM on every rank instead of blocking.)
    std::mutex mu;
    std::condition_variable cv;
    int prepared = 0;
    bool go = false, abort_all = false;
    auto helper = [&](int rank) {
      AoBakeParams p;
      aobake_default_params(&p);
      p.device = rank;
      AoBake* c = nullptr;
      int rc = aobake_create(&p, &c);
      if (rc == AOBAKE_OK) rc = aobake_set_scene(c, &scene, cfg.ground ? &blockers : nullptr);
      if (rc == AOBAKE_OK) rc = aobake_sample_instances(c, per.data(), (size_t)cfg.samples_per_face, nullptr);
      if (rc != AOBAKE_OK) fprintf(stderr, "rank %d: %s\n", rank, c ? aobake_last_error(c) : aobake_last_error(nullptr));
      {
        std::unique_lock<std::mutex> lk(mu);
        prepared++;
        if (rc != AOBAKE_OK) abort_all = true;
        cv.notify_all();
        cv.wait(lk, [&] { return go; });
      }
      if (!abort_all) {
        rc = aobake_comm_init(c, rank, cfg.gpus, id);
        if (rc == AOBAKE_OK) rc = aobake_compute_ao_distributed(c, cfg.rays, scene_offset, scene_maxdist, nullptr);
        if (rc != AOBAKE_OK) fprintf(stderr, "rank %d: %s\n", rank, aobake_last_error(c));
      }
      rcs[rank] = rc;
      if (c) aobake_destroy(c);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < cfg.gpus; r++) th.emplace_back(helper, r);
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return prepared == cfg.gpus - 1; });
      go = true;
      cv.notify_all();
    }
    int rc0 = AOBAKE_OK;
    if (!abort_all) {
      rc0 = aobake_comm_init(ctx, 0, cfg.gpus, id);
      if (rc0 == AOBAKE_OK) rc0 = aobake_compute_ao_distributed(ctx, cfg.rays, scene_offset, scene_maxdist, nullptr);
      if (rc0 != AOBAKE_OK) fprintf(stderr, "rank 0: %s\n", aobake_last_error(ctx));
    }
    for (auto& x : th) x.join();
    bool failed = abort_all || rc0 != AOBAKE_OK;
    for (int r = 1; r < cfg.gpus; r++) failed = failed || rcs[r] != AOBAKE_OK;
    if (failed) { fprintf(stderr, "multi-GPU bake failed\n"); aobake_destroy(ctx); return 1; }
  } else {
    ck(aobake_compute_ao(ctx, cfg.rays, scene_offset, scene_maxdist, nullptr), "compute_ao");
  }
  AoTimings tm;
  aobake_get_timings(ctx, &tm);
  fprintf(stderr, "Compute AO ... %.2f ms   (fused raygen + query + accumulate kernel %.2f ms, %.1f Mrays/s)\n", now_ms() - t0, tm.trace_ms,
          tm.trace_ms > 0 ? (double)tm.rays_traced / tm.trace_ms / 1e3 : 0.0);

  t0 = now_ms();
  std::vector<std::vector<float>> vao(cfg.instances, std::vector<float>(mesh.num_vertices));
  std::vector<float*> vptr(cfg.instances);
  for (int i = 0; i < cfg.instances; i++) vptr[i] = vao[i].data();
  ck(aobake_map_ao_to_vertices(ctx, cfg.least_squares ? AOBAKE_FILTER_LEAST_SQUARES : AOBAKE_FILTER_AREA_BASED, cfg.weight, vptr.data()), "map_ao_to_vertices");
  aobake_get_timings(ctx, &tm);
  fprintf(stderr, "Map AO to vertices (%s) ... %.2f ms", cfg.least_squares ? "least squares" : "area based", now_ms() - t0);
  if (cfg.least_squares) fprintf(stderr, "   (%d CG iterations)", tm.cg_iterations);
  fprintf(stderr, "\n");
  double mean = 0;
  for (auto& v : vao) for (float x : v) mean += x;
  fprintf(stderr, "Mean vertex AO: %.6f\n", mean / ((double)cfg.instances * mesh.num_vertices));

  // ---- raw result file: u64 num_instances, u64 num_vertices_total, then per instance
  //      {u64 storage_identifier, u64 vertex_offset, u64 num_vertices}, then all floats ----
  if (!cfg.outfile.empty()) {
    FILE* f = fopen(cfg.outfile.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", cfg.outfile.c_str()); aobake_destroy(ctx); return 1; }
    const uint64_t ni = (uint64_t)cfg.instances, nvt = ni * mesh.num_vertices;
    fwrite(&ni, 8, 1, f); fwrite(&nvt, 8, 1, f);
    for (uint64_t i = 0; i < ni; i++) {
      const uint64_t rec[3] = {insts[i].storage_identifier, i * mesh.num_vertices, mesh.num_vertices};
      fwrite(rec, 8, 3, f);
    }
    for (auto& v : vao) fwrite(v.data(), 4, v.size(), f);
    fclose(f);
    fprintf(stderr, "Wrote %s\n", cfg.outfile.c_str());
  }
  aobake_destroy(ctx);
  return 0;
}
