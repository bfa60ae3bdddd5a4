"""Regenerates tests/golden/oracle_golden.json from the oracle.  The reference shipped no
golden vectors and cannot be run, so these pin the oracle against itself
across refactors; independent anchors are the analytic tests in test_oracle_cpu.py."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_binding as ob  # noqa: E402

out = {"tea": [], "lcg": [], "halton": []}
for rounds, v0, v1 in [(2, 0, 0), (2, 1, 2), (4, 0, 0), (4, 0, 1), (4, 7, 79599), (4, 999, 123456), (16, 3, 5),
                       (2, (63 << 16) | 63, 238799), (2, (1023 << 16) | 1023, 9999999)]:
    out["tea"].append([rounds, v0, v1, hex(ob.tea(rounds, v0, v1))])
for seed in [0, 1, 0xDEADBEEF]:
    out["lcg"].append([seed, [hex(x) for x in ob.lcg_stream(seed, 6)]])
for i, b in [(1, 2), (2, 2), (3, 2), (7, 2), (1000, 2), (1, 3), (2, 3), (5, 3), (26, 3), (1000, 3)]:
    out["halton"].append([i, b, hex(int(np.float32(ob.halton(i, b)).view(np.uint32)))])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["tea"]), "tea,", len(out["lcg"]), "lcg,", len(out["halton"]), "halton vectors")
