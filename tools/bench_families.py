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


def next_row_graph(family, poses, N, rng):
    """synthetic workload for a next-row family: `poses` factors over `poses` + landmark variables laid
    out like the Manhattan-shaped walk (unit steps), landmarks 5 m off the path"""
    fg = rb.initfg(rb.SolverParams(N=N))
    walk = np.cumsum(rng.choice([-1.0, 1.0], size=(poses, 2)), 0)
    vt0, vt1 = rb.FAMILY[family][0], rb.FAMILY[family][1]
    cls = {rb.POSE2: rb.Pose2, rb.POINT2: rb.Point2, rb.POSE3: rb.Pose3, rb.POINT3: rb.Point3}
    sig = {rb.POSE2: [0.1, 0.12, 0.02], rb.POINT2: [0.3, 0.3], rb.POSE3: [0.1] * 3 + [0.02] * 3, rb.POINT3: [0.3] * 3}

    def truth(t, i):
        d = rb.VAR_DIM[t]
        x = np.zeros(d)
        x[:2] = walk[i]
        if t in (rb.POINT2, rb.POINT3):
            x[:2] += 5.0
        if t == rb.POSE2:
            x[2] = rng.uniform(-3, 3)
        if t == rb.POSE3:
            x[3:] = rng.normal(size=3) * 0.5
        return x
    names0 = [f"a{i}" for i in range(poses)]
    for i, l in enumerate(names0):
        v = rb.addVariable(fg, l, cls[vt0])
        v.val = truth(vt0, i)[None] + rng.normal(size=(N, rb.VAR_DIM[vt0])) * sig[vt0]
    if vt1 is not None and vt1 != vt0:
        for i in range(poses):
            v = rb.addVariable(fg, f"b{i}", cls[vt1])
            v.val = truth(vt1, i)[None] + rng.normal(size=(N, rb.VAR_DIM[vt1])) * sig[vt1]
    dm = rb.FAMILY[family][2]
    mk = {rb.PRIORPOINT2: lambda: rb.PriorPoint2(rb.MvNormal(np.zeros(2), np.eye(2) * 0.01)),
          rb.POINT2POINT2: lambda: rb.Point2Point2(rb.MvNormal(np.array([1.0, 1.0]), np.eye(2) * 0.01)),
          rb.POSE2POINT2: lambda: rb.Pose2Point2(rb.MvNormal(np.array([5.0, 5.0]), np.eye(2) * 0.01)),
          rb.POSE2POINT2RANGE: lambda: rb.Pose2Point2Range(rb.Normal(7.0, 0.1)),
          rb.POINT2POINT2RANGE: lambda: rb.Point2Point2Range(rb.Normal(1.4, 0.1)),
          rb.POSE2POINT2BEARING: lambda: rb.Pose2Point2Bearing(rb.Normal(0.7, 0.05)),
          rb.PRIORPOINT3: lambda: rb.PriorPoint3(rb.MvNormal(np.zeros(3), np.eye(3) * 0.01)),
          rb.POINT3POINT3: lambda: rb.Point3Point3(rb.MvNormal(np.array([1.0, 1.0, 0.0]), np.eye(3) * 0.01)),
          rb.POSE3POSE3XYYAW: lambda: rb.Pose3Pose3XYYaw(rb.MvNormal(np.array([1.0, 1.0, 0.1]), np.eye(3) * 0.01)),
          rb.POSE3POSE3ROTATION: lambda: rb.Pose3Pose3Rotation(rb.MvNormal(np.zeros(3), np.eye(3) * 1e-4)),
          rb.POSE3POSE3UNITTRANS: lambda: rb.Pose3Pose3UnitTrans()}[family]
    for i in range(poses):
        if vt1 is None:
            rb.addFactor(fg, [names0[i]], mk())
        elif vt1 == vt0:
            if i + 1 < poses:
                rb.addFactor(fg, [names0[i], names0[i + 1]], mk())
        else:
            rb.addFactor(fg, [names0[i], f"b{i}"], mk())
    return fg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--poses", type=int, default=10000)
    ap.add_argument("--next-rows", action="store_true", help="time the next-row families (SURVEY 8f N1) instead")
    args = ap.parse_args()
    if args.next_rows:
        rng = np.random.default_rng(6)
        names = {rb.PRIORPOINT2: "PriorPoint2", rb.POINT2POINT2: "Point2Point2", rb.POSE2POINT2: "Pose2Point2",
                 rb.POSE2POINT2RANGE: "Pose2Point2Range", rb.POINT2POINT2RANGE: "Point2Point2Range",
                 rb.POSE2POINT2BEARING: "Pose2Point2Bearing", rb.PRIORPOINT3: "PriorPoint3", rb.POINT3POINT3: "Point3Point3",
                 rb.POSE3POSE3XYYAW: "Pose3Pose3XYYaw", rb.POSE3POSE3ROTATION: "Pose3Pose3Rotation",
                 rb.POSE3POSE3UNITTRANS: "Pose3Pose3UnitTrans"}
        for fam, nm in names.items():
            fg = next_row_graph(fam, args.poses, 100, rng)
            time_family(f"N1 {nm} synthetic {args.poses}", fg, fam, 100, args.steps, 8)
        return
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
