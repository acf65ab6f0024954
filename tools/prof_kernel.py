"""Launch one family's hot kernels a few times on its BASELINE config (for ncu captures; no timing here).

    python tools/prof_kernel.py pose2pose2|bearingrange|pose3pose3 [reps]
pose2pose2: the bench workload (10k-pose SE(2) graph, N=100); bearingrange: Beehive-shaped C4 (N=200);
pose3pose3: C5 SE(3) chain + loops (N=100).  Each rep launches the fused-sample kernel and the
supplied-measurement kernel on 4 rotating working-set copies."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import rome_b200 as rb  # noqa: E402

family = sys.argv[1] if len(sys.argv) > 1 else "pose2pose2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sets = []
if family == "pose2pose2":
    fam, N = rb.POSE2POSE2, bench.NPART
    w = bench.build_workload(1)
    for s in range(4):
        c = rb.Context(0)
        c.use_torch_stream()
        c.set_particles(rb.POSE2, w["poses"] + 1e-4 * s)
        c.set_factors_pose2pose2(w["ip"], w["iq"], w["mu"], w["cov"])
        sets.append(c)
else:
    if family == "bearingrange":
        fam, N = rb.BEARINGRANGE, 200
        fg = rb.generateGraph_Beehive(10000, N=N)
        rb.seed_particles(fg, N=N, seed=3)
    elif family == "pose3pose3":
        fam, N = rb.POSE3POSE3, 100
        fg = rb.generateGraph_Pose3Chain(10000, loops=1000)
        rb.seed_particles(fg, N=N, seed=4)
    else:
        raise SystemExit(f"unknown family {family}")
    for s in range(4):
        dg = rb.DeviceGraph(fg, ctx=rb.Context(0), N=N)
        dg.ctx.use_torch_stream()
        sets.append(dg.ctx)
vt0, vt1, dm, dr, ns, dj, dfwd, dbwd = rb.FAMILY[fam]
Np = rb.npad(N)
F = sets[0].num_factors(fam)
bufs = [(torch.zeros((F, Np, dr), device="cuda"), torch.zeros((F, ns), device="cuda"),
         torch.randn((F, Np, dm), device="cuda") * 0.01) for _ in sets]
torch.cuda.synchronize()
for k in range(reps):
    for c, (res, st, meas) in zip(sets, bufs):
        c.eval(fam, rb.SAMPLE | rb.RESIDUAL | rb.STATS, seed=1, stream_id=k, res=res, stats=st)
    for c, (res, st, meas) in zip(sets, bufs):
        c.eval(fam, rb.RESIDUAL | rb.STATS, meas=meas, res=res, stats=st)
torch.cuda.synchronize()
print("done", family, F, N)
