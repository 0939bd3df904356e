#!/usr/bin/env python
"""Quick A/B of the tiled at-output operator against the two-pass (gather, every group Jt[Jp]) operator on
shape_from_shading: same problem, GN 2 x 30, final costs and ms per solve.   python scripts/try_jp_all.py [size]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from thallo_b200 import workloads as wl
from thallo_b200.api import ThalloSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
out = {}
for tag, kw in (("tiled", {}), ("two_pass", dict(schedule="gather", define_kwargs=dict(jp_all=True)))):
    p = wl.sfs_params(wl.sfs_inputs(n, n))
    dp = [torch.from_numpy(np.ascontiguousarray(x)).cuda() if np.asarray(x).size > 1 else x for x in p]
    s = ThalloSolver([n, n], "shape_from_shading", "gauss_newton", timing=2, **kw)
    s.set_parameters(nIterations=2, lIterations=30)
    ms = []
    for rep in range(3):
        for i, x in enumerate(p):
            if np.asarray(x).size > 1:
                dp[i].copy_(torch.from_numpy(np.ascontiguousarray(x)))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cost = s.solve(dp)
        e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    out[tag] = dict(final_cost=cost, ms_per_solve=min(ms[1:]), kernels=s.kernel_times() if hasattr(s, "kernel_times") else None)
    s.close()
print(json.dumps(out))
