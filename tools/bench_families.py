#!/usr/bin/env python
"""Per-family kernel timings on the BASELINE configs (extra to bench.py's headline line).

  C3  examples/manhattan.g2o (tests/golden/manhattan_g2o.npz): 3500 Pose2, 5453 Pose2Pose2 + prior, N=100
  C4  Beehive-shaped: P poses + lattice landmarks, one Pose2Point2BearingRange per pose + odometry, N=200
  C5  SE(3) helix chain: P Pose3, chain + loop Pose3Pose3 + PriorPose3, N=100
Each kernel is timed alone: K launches from one CUDA graph rotating over `sets` working-set copies (> 2x L2 when the
config is big enough), fused-sample mode (SAMPLE|RESIDUAL|STATS) and supplied-measurement mode, with and without
ROME_B200_INDEPENDENT overlap.  Prints one JSON object per (config, family)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rome_b200 as rb  # noqa: E402

PEAK = 6556.8
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def family_arrays(fg, family):
    dg_f = sorted((f for f in fg.factors.values() if f.fnc.family == family), key=lambda f: f.index)
    return len(dg_f)


def time_family(name, fg, family, N, steps, sets_n):
    Np = rb.npad(N)
    vt0, vt1, dm, dr, ns, dj, dfwd, dbwd = rb.FAMILY[family]
    stream = torch.cuda.Stream()
    sets = []
    with torch.cuda.stream(stream):
        for s in range(sets_n):
            dg = rb.DeviceGraph(fg, ctx=rb.Context(0), N=N)
            dg.ctx.use_torch_stream()
            nF = dg.ctx.num_factors(family)
            sets.append((dg.ctx, torch.zeros((nF, Np, dr), device="cuda"), torch.zeros((nF, ns), device="cuda"),
                         torch.randn((nF, Np, dm), device="cuda") * 0.01))
        stream.synchronize()
        out = dict(config=name, family=family, factors=nF, N=N, evals_per_launch=nF * N)
        for mode, flags, bpe in (("fused_sample", rb.SAMPLE | rb.RESIDUAL | rb.STATS, rb.BYTES_PER_EVAL_SAMPLED[family]),
                                 ("supplied_meas", rb.RESIDUAL | rb.STATS, rb.BYTES_PER_EVAL[family])):
            for indep in (True, False):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    for c, _, _, _ in sets:
                        c.use_torch_stream()
                    for k in range(steps):
                        c, res, st, meas = sets[k % sets_n]
                        kw = dict(res=res, stats=st)
                        if not flags & rb.SAMPLE:
                            kw["meas"] = meas
                        c.eval(family, flags | (rb.INDEPENDENT if indep else 0), seed=1, stream_id=k, **kw)
                for c, _, _, _ in sets:
                    c.use_torch_stream()
                g.replay()
                stream.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                g.replay()
                e1.record(stream)
                stream.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / steps
                gbs = nF * N * bpe / (us * 1e-6) / 1e9
                out[f"{mode}{'' if indep else '_serialized'}"] = dict(us_per_launch=us, evals_per_s=nF * N / (us * 1e-6),
                                                                     gbs=gbs, frac_hbm=gbs / PEAK, bytes_per_eval=bpe)
    for c, *_ in sets:
        c.close()
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--poses", type=int, default=10000)
    args = ap.parse_args()
    golden = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "manhattan_g2o.npz")
    z = np.load(golden)
    fg = rb.graphFromEdgeArrays(z["ids"], z["mu"], z["info"])
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal(np.zeros(3), np.diag([0.1, 0.1, 0.05]) ** 2)))
    odo = {(a, b): m for (a, b), m in zip(map(tuple, z["ids"]), z["mu"]) if b == a + 1}
    pose = np.zeros(3)
    fg["x0"].simulated = pose.copy()
    for i in range(3499):
        c, s = np.cos(pose[2]), np.sin(pose[2])
        m = odo[(i, i + 1)]
        pose = np.array([pose[0] + c * m[0] - s * m[1], pose[1] + s * m[0] + c * m[1], pose[2] + m[2]])
        fg[f"x{i+1}"].simulated = pose.copy()
    rb.seed_particles(fg, seed=1, N=100)
    time_family("C3 manhattan.g2o", fg, rb.POSE2POSE2, 100, args.steps, 24)
    bh = rb.generateGraph_Beehive(args.poses, N=200)
    rb.seed_particles(bh, N=200, seed=3)
    time_family(f"C4 beehive {args.poses} poses", bh, rb.BEARINGRANGE, 200, args.steps, 12)
    time_family(f"C4 beehive {args.poses} poses", bh, rb.POSE2POSE2, 200, args.steps, 8)
    p3 = rb.generateGraph_Pose3Chain(args.poses, loops=args.poses // 10)
    rb.seed_particles(p3, seed=4, N=100)
    time_family(f"C5 se3 chain {args.poses} poses", p3, rb.POSE3POSE3, 100, args.steps, 6)


if __name__ == "__main__":
    main()
