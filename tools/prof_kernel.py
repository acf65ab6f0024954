"""Launch the hot kernels a few times on the bench workload (for ncu captures; no timing here)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rome_b200 as rb  # noqa: E402

family = sys.argv[1] if len(sys.argv) > 1 else "pose2pose2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = bench.build_workload(1)
F, N, Np = len(w["ip"]), bench.NPART, rb.npad(bench.NPART)
sets = []
for s in range(4):
    c = rb.Context(0)
    c.use_torch_stream()
    c.set_particles(rb.POSE2, w["poses"] + 1e-4 * s)
    c.set_factors_pose2pose2(w["ip"], w["iq"], w["mu"], w["cov"])
    sets.append((c, torch.zeros((F, Np, 3), device="cuda"), torch.zeros((F, 16), device="cuda"),
                 torch.randn((F, Np, 3), device="cuda") * 0.05))
torch.cuda.synchronize()
for k in range(reps):
    for c, res, st, meas in sets:
        c.eval(rb.POSE2POSE2, rb.SAMPLE | rb.RESIDUAL | rb.STATS, seed=1, stream_id=k, res=res, stats=st)
    for c, res, st, meas in sets:
        c.eval(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS, meas=meas, res=res, stats=st)
torch.cuda.synchronize()
print("done")
