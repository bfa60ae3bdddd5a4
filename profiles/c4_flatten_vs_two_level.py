import sys, numpy as np
sys.path.insert(0,'.')
import bench
from optix_prime_baking_b200 import api, scenes
scene, blockers, min_per, requested, desc = bench.make_workload("c4")
off,maxd=scenes.default_distances(scene)
res={}
for mode in (2,1):
    with api.Baker(instancing_mode=mode, collect_stats=False) as bk:
        bk.set_scene(scene, blockers)
        tm=bk.timings(); st=bk.stats()
        total, per = bk.distribute_samples(3, 0)
        bk.sample_instances(per, 3, download=False)
        n=total//8; b=int((total-n)*0.37)
        for i in range(2):
            ao=bk.compute_ao(256, off, maxd, begin=b, end=b+n)
            t=bk.timings().trace_ms
        res[mode]=ao
        print("mode",mode,"two_level",st.two_level,"nodes",st.num_bvh_nodes,"bvh MB %.0f"%(st.bvh_bytes/1e6),"build_ms %.1f"%tm.bvh_build_ms,"ms %.1f"%t,"Grays/s %.2f"%(n*256/t/1e6), flush=True)
d=np.abs(res[1]-res[2]); print("AO flatten vs two-level: max diff %.4f, differing samples %d of %d"%(d.max(), (d>0).sum(), len(d)))
