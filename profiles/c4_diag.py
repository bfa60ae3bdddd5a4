import sys, numpy as np
sys.path.insert(0,'.')
import bench
from optix_prime_baking_b200 import api, scenes
from tests.oracle_binding import Oracle
scene, blockers, min_per, requested, desc = bench.make_workload("c4")
off,maxd=scenes.default_distances(scene)
bks={}
res={}
for mode in (2,1):
    bk=api.Baker(instancing_mode=mode); bks[mode]=bk
    bk.set_scene(scene, blockers)
    total, per = bk.distribute_samples(3, 0)
    bk.sample_instances(per, 3, download=False)
    n=total//8; b=int((total-n)*0.37)
    res[mode]=bk.compute_ao(256, off, maxd, begin=b, end=b+n)
    print("mode",mode,"done",flush=True)
d=np.abs(res[1]-res[2]); idx=np.nonzero(d>0)[0]
print("differing",len(idx),"max",d.max())
order=idx[np.argsort(-d[idx])][:6]
orc=Oracle(scene, blockers, 2)
for k in order:
    g=int(b+k)
    rays=bks[2].dump_rays(g,g+1,256,off,maxd).reshape(-1,8)
    h2=bks[2].trace_rays(rays); h1=bks[1].trace_rays(rays); ho=orc.trace_rays(rays); hb=orc.trace_rays(rays,brute=True)
    inst=np.searchsorted(np.cumsum(per), g, side='right')
    print("sample",g,"inst",inst,"ao two-level %.4f flat %.4f"%(res[2][k],res[1][k]),"| trace_rays hits: two-level",h2.sum(),"flat",h1.sum(),"oracle bvh",ho.sum(),"oracle brute",hb.sum(), "| origin",rays[0,:3], flush=True)
