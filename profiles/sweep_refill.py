import os, sys, time, numpy as np
sys.path.insert(0,'.')
from optix_prime_baking_b200 import api, scenes
w=sys.argv[1]
import bench
scene, blockers, min_per, requested, desc = bench.make_workload(w)
rays=bench.RAYS[w]
off,maxd=scenes.default_distances(scene)
for tk, rb, lt in [(2,30,0)]:
    with api.Baker(trace_kernel=tk, refill_below=rb, leaf_tris=lt, node_test=int(os.environ.get('NODE_TEST', '0'))) as bk:
        bk.set_scene(scene, blockers)
        total, per = bk.distribute_samples(min_per, requested)
        bk.sample_instances(per, min_per, download=False)
        n = total if w!='c4' else total//8
        ts=[]
        for i in range(3):
            bk.compute_ao(rays, off, maxd, download=False, begin=0, end=n)
            ts.append(bk.timings().trace_ms)
        t=min(ts); q2=int(round(rays**0.5))**2
        print(w, "kernel",tk,"refill_below",rb,"leaf_tris",lt, "ms %.2f"%t, "Grays/s %.2f"%(n*q2/t/1e6), "node_test", os.environ.get("NODE_TEST", "0"), "deferred rays", bk.stats().reserved[2], flush=True)
