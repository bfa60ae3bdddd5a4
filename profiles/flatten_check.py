import sys, numpy as np
sys.path.insert(0,'.')
from optix_prime_baking_b200 import api, scenes
for grid, st in [(3,40),(4,60),(5,80),(6,100),(8,100),(10,100)]:
    scene, blockers = scenes.config4_instanced(grid, st, st, seed=4)
    off,maxd=scenes.default_distances(scene)
    res={}
    for mode in (2,1):
        with api.Baker(instancing_mode=mode) as bk:
            bk.set_scene(scene, blockers)
            total, per = bk.distribute_samples(1, 0)
            bk.sample_instances(per, 1, download=False)
            res[mode]=bk.hit_counts() if False else bk.compute_ao(16, off, maxd)
            nodes=bk.stats().num_bvh_nodes
    d=np.abs(res[1]-res[2])
    print("grid",grid,"tris",scene.num_triangles,"samples",total,"flat nodes",nodes,"differing samples",(d>0).sum(),"max diff %.3f"%d.max(), "diff>0.1:",(d>0.1).sum(), flush=True)
