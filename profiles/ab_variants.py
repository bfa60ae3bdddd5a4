#!/usr/bin/env python
"""Kernel A/B harness: builds variants of libaobake.so with extra -D flags HERE (nvcc cross-compiles
without a GPU) into variants_tmp/ (travels to the GPU box with the snapshot; gpurun_out/ does not), and
prints the one-line gpurun command that times every variant on the sweep workloads.

usage: ab_variants.py name=-DFLAG=1[,-DOTHER=2] ...      e.g.  ab_variants.py corner=-DAOB_H2_CORNER_FRAME=1
Then run the printed command; each line of its log is `<variant> <workload> ... Grays/s`.
Delete variants_tmp/ afterwards (it is git-ignored)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optix_prime_baking_b200 import build  # noqa: E402

os.makedirs(os.path.join(ROOT, "variants_tmp"), exist_ok=True)
names = ["head"]
build.build(out=os.path.join(ROOT, "variants_tmp", "libaobake_head.so"))
for arg in sys.argv[1:]:
    name, flags = arg.split("=", 1)
    build.build(out=os.path.join(ROOT, "variants_tmp", f"libaobake_{name}.so"), extra_flags=flags.split(","))
    names.append(name)
runs = " ".join(f'for w in c2 c4 c3; do echo -n "{n} "; AOBAKE_LIB=variants_tmp/libaobake_{n}.so timeout 150 python profiles/sweep_refill.py $w | grep Grays; done;'
                for n in names)
print("built:", ", ".join(names))
print(f"/usr/local/graft/bin/gpurun --timeout 900 -- '({runs}) > gpurun_out/ab.log 2>&1; cat gpurun_out/ab.log'")
