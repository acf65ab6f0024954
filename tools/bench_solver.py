#!/usr/bin/env python
"""Device-resident sweeps (convolutions + products of proposals) on the bench workload: time per sweep,
split into the convolution kernels and the product kernels.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rome_b200 as rb  # noqa: E402


def main(poses=10000, N=100, sweeps=3):
    fg = rb.generateGraph_ManhattanShaped(poses, seed=2, N=N)
    rb.seed_particles(fg, seed=1)
    truth0 = np.stack([v.simulated for v in fg.variables.values()])
    est0 = np.stack([v.val.mean(0) for v in fg.variables.values()])
    err_before = float(np.abs(est0[:, :2] - truth0[:, :2]).mean())
    dg = rb.DeviceGraph(fg, ctx=rb.Context(0), N=N)
    gs = rb.GibbsSolver(dg)
    c = dg.ctx
    gs.sweep(0)
    c.synchronize()
    t0 = time.perf_counter()
    for s in range(sweeps):
        gs.sweep(1 + s)
    c.synchronize()
    dt = time.perf_counter() - t0
    dg.download_particles()  # the state after the sweeps (the timing loops below go on changing it)
    truth = np.stack([v.simulated for v in fg.variables.values()])
    est = np.stack([v.val.mean(0) for v in fg.variables.values()])
    # the two halves of a sweep, each timed in its OWN loop between synchronisations (no subtraction of wall clocks):
    # the convolution launches are ~40 us per sweep, so that loop repeats them often enough to last milliseconds
    reps = 50
    c.synchronize()
    t1 = time.perf_counter()
    for s in range(reps):
        gs.convolve(100 + s)
    c.synchronize()
    dconv = (time.perf_counter() - t1) / reps
    t2 = time.perf_counter()
    for s in range(sweeps):
        gs.update(200 + s)
    c.synchronize()
    dprod = (time.perf_counter() - t2) / sweeps
    k = np.diff(gs.plans[rb.POSE2][0])
    nfac = sum(c.num_factors(f) for f in gs.families)
    out = dict(workload=f"manhattan_shaped_{poses}_se2_N{N}", sweeps=sweeps, ms_per_sweep=1e3 * dt / sweeps,
               ms_product_per_sweep=1e3 * dprod, ms_convolution_per_sweep=1e3 * dconv,
               variables=int(len(k)), factors=int(nfac), proposals_per_variable_mean=float(k.mean()),
               proposals_per_variable_max=int(k.max()),
               convolved_particles_per_sweep=int(k.sum()) * N,
               convolved_particles_per_s=float(k.sum()) * N / (dt / sweeps))
    out["mean_abs_translation_error_m"] = float(np.abs(est[:, :2] - truth[:, :2]).mean())
    out["mean_abs_translation_error_before_m"] = err_before  # the seeded particles the sweeps start from: simulated truth + 0.1 m spread, i.e. BETTER than the noisy measurements support -- the sweeps move the beliefs to what the measurements say
    print(json.dumps(out))
    gs.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10000)
