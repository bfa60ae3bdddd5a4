import sys, numpy as np
sys.path.insert(0,'.')
from optix_prime_baking_b200 import api, scenes
from tests.oracle_binding import Oracle
scene, blockers = scenes.config4_instanced(5, 80, 80, seed=4)
off,maxd=scenes.default_distances(scene)
res={}; bks={}
for mode in (2,1):
    bk=api.Baker(instancing_mode=mode); bks[mode]=bk
    bk.set_scene(scene, blockers)
    total, per = bk.distribute_samples(1, 0)
    bk.sample_instances(per, 1, download=False)
    res[mode]=bk.compute_ao(16, off, maxd)
d=np.abs(res[1]-res[2]); idx=np.nonzero(d>0)[0]
o1=Oracle(scene,blockers,1); o2=Oracle(scene,blockers,2)
for g in idx:
    g=int(g)
    rays=bks[2].dump_rays(g,g+1,16,off,maxd).reshape(-1,8)
    g1=bks[1].trace_rays(rays); g2=bks[2].trace_rays(rays); c1=o1.trace_rays(rays); c2=o2.trace_rays(rays); b1=o1.trace_rays(rays,brute=True)
    m=o2.ray_margin(rays)
    dis=np.nonzero(g1!=g2)[0]
    print("sample",g,"hits gpu-flat",g1.sum(),"oracle-flat",c1.sum(),"oracle-flat-brute",b1.sum(),"| gpu-2level",g2.sum(),"oracle-2level",c2.sum(),"| margins of disagreeing rays",m[dis], flush=True)
