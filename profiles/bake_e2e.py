#!/usr/bin/env python
"""End-to-end bake seconds (BASELINE.json metric, second half): set_scene (upload + BVH build)
-> distribute -> sample -> compute AO (sharded over ranks, NCCL gather) -> vertex map.
usage: [torchrun ...] bake_e2e.py <c1|c2|c3|c4|c5> [area|ls]"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from optix_prime_baking_b200 import api, scenes  # noqa: E402
from optix_prime_baking_b200.multi_gpu import DistributedBaker  # noqa: E402

w = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else ("ls" if w == "c5" else "area")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.perf_counter()
if w == "c5":
    scene, blockers = scenes.config3_bigmesh(with_ground=True)
    min_per, requested, rays = 0, 10_000_000, 1024
else:
    scene, blockers, min_per, requested, _ = bench.make_workload(w)
    rays = bench.RAYS[w]
off, maxd = scenes.default_distances(scene)
t_gen = time.perf_counter() - t0
out = {"workload": w, "filter": mode, "world": world, "scene_gen_s": t_gen}
# one context per rank for the whole process; the NCCL communicator is process set-up (like
# dist.init_process_group), not part of a bake: ncclCommInitRank alone takes ~2 s at 8 ranks
bk = api.Baker(device=local, cg_tolerance=1e-6, cg_max_iterations=5000)
if world > 1:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(api.Baker.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, src=0)
    bk.comm_init(rank, world, idt.cpu().numpy().tobytes())
for rep in range(2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bk.set_scene(scene, blockers)
    tm = bk.timings()
    t1 = time.perf_counter()
    total, per = bk.distribute_samples(min_per, requested)
    bk.sample_instances(per, min_per, download=False)
    t2 = time.perf_counter()
    if world > 1:
        # native path: sharding, NCCL all-reduce of ao[] and the instance-split vertex map all inside libaobake.so
        bk.compute_ao_distributed(rays, off, maxd, download=False)
    else:
        bk.compute_ao(rays, off, maxd, download=False)
    t3 = time.perf_counter()
    trace_ms = bk.timings().trace_ms
    v = bk.map_ao_to_vertices(api.FILTER_LEAST_SQUARES if mode == "ls" else api.FILTER_AREA_BASED, 0.1, distributed=world > 1)
    t4 = time.perf_counter()
    tf = bk.timings()
    st = bk.stats()
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    q = int(round(rays ** 0.5))
    res = {"bake_seconds": t5 - t0, "set_scene_s": t1 - t0, "upload_ms": tm.upload_ms, "bvh_build_ms": tm.bvh_build_ms,
           "sample_s": t2 - t1, "compute_ao_s": t3 - t2, "trace_kernel_ms": trace_ms, "vertex_map_s": t4 - t3,
           "cg_iterations": tf.cg_iterations, "samples": int(total), "rays": int(total) * q * q,
           "Mrays_per_s_whole_bake": int(total) * q * q / (t5 - t0) / 1e6, "bvh_nodes": st.num_bvh_nodes,
           "vertex_ao_mean": float(np.mean([x.mean() for x in v])), "vertex_ao_min": float(min(x.min() for x in v)),
           "vertex_ao_max": float(max(x.max() for x in v))}
    out[f"run{rep}"] = res
if rank == 0 and "--oracle" in sys.argv:
    # the same bake on the CPU oracle (host cores), phase by phase — BASELINE.md §2
    from tests.oracle_binding import Oracle, lib
    orc = Oracle(scene, blockers)
    t0 = time.perf_counter()
    ototal, oper = orc.distribute_samples(min_per, requested)
    osb = orc.sample_instances(oper, min_per)
    t1 = time.perf_counter()
    _ = orc.tracer
    t2 = time.perf_counter()
    oao, _h = orc.compute_ao(osb, rays, off, maxd)
    t3 = time.perf_counter()
    ov = orc.filter_least_squares(osb, oao, 0.1, tol=1e-6) if mode == "ls" else orc.filter_area(osb, oao)
    t4 = time.perf_counter()
    q = int(round(rays ** 0.5))
    out["oracle_cpu"] = {"cores": int(lib().ao_oracle_num_threads()), "sample_s": t1 - t0, "bvh_build_s": t2 - t1, "trace_s": t3 - t2,
                         "trace_Mrays_per_s": ototal * q * q / (t3 - t2) / 1e6, "vertex_map_s": t4 - t3, "bake_seconds": t4 - t0,
                         "max_abs_vertex_ao_diff_vs_gpu": float(max(np.abs(a - b).max() for a, b in zip(ov, v)))}
if rank == 0:
    print(json.dumps(out))
bk.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
