"""Full-size parity on the device: the BASELINE workloads at the sizes bench.py times, every residual
row compared with the float64 oracle -- not a sampled subset, not the emulator.

  T   10 000 Pose2, 11 999 Pose2Pose2 + PriorPose2, N = 100 (the bench workload)
  C5  10 000 Pose3, 10 999 Pose3Pose3 + PriorPose3, N = 100
  C4  Beehive: 10 000 Pose2 + landmarks, Pose2Pose2 + Pose2Point2BearingRange, N = 200
fused getSample with the samples written back (ROME_B200_SAMPLE | WRITE_MEAS | RESIDUAL | STATS), default arithmetic
(float32 per particle): |gpu - ref| <= 1e-5 max(|ref|, 0.1) against the oracle on the caller's ORIGINAL Float64 particles;
the Float64 chain (ROME_B200_PRECISE) on the same samples: floor 1e-2.  Plus a fixture with heading spreads that cross
+-pi (reference edge case test/testBasicPose2Conv.jl:25-56)."""
import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O
from rome_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


def _check(ctx, w, fam, sweep, angle_cols=()):
    f, N = w["families"][fam], w["N"]
    fl = rb.SAMPLE | rb.WRITE_MEAS | rb.RESIDUAL | rb.STATS
    out = ctx.alloc_host_outputs(fam, fl)
    ctx.eval_host(fam, fl, seed=21, stream_id=4, **out)
    mu = np.column_stack([f["a"][:, 0], f["b"][:, 0]]) if fam == rb.BEARINGRANGE else np.asarray(f["a"])
    meas = rb.offsets_to_meas(out["meas_out"], mu, N)
    ref = sweep(meas)
    res = rb.rows_to_particle_major(out["res"], N)

    def worst(r, floor):
        d = r - ref
        for c in angle_cols:
            d[..., c] = O.np_wrap(d[..., c])
        return float((np.abs(d) / np.maximum(np.abs(ref), floor)).max()), float(np.abs(d).max())
    rel, ab = worst(res, 1e-1)
    assert rel < 1e-5, (fam, rel, ab)
    st = out["stats"]
    dr = res.shape[-1]
    assert np.allclose(st[:, :dr], res.sum(1), rtol=1e-3, atol=2e-3)
    # the Float64 chain on the same samples
    o64 = ctx.alloc_host_outputs(fam, rb.RESIDUAL | rb.PRECISE)
    ctx.eval_host(fam, rb.RESIDUAL | rb.PRECISE, meas=out["meas_out"], **o64)
    rel64, _ = worst(rb.rows_to_particle_major(o64["res"], N), 1e-2)
    assert rel64 < 1e-5, (fam, rel64)
    return ab


def test_workload_T_every_row(ctx):
    w = W.manhattan_arrays(10000, seed=2, N=100, particle_seed=1)
    P = w["particles"][rb.POSE2]
    ctx.set_particles(rb.POSE2, P)
    for fam in (rb.POSE2POSE2, rb.PRIORPOSE2):
        f = w["families"][fam]
        W.upload_family(ctx, fam, f["i0"], f["i1"], f["a"], f["b"])
    f = w["families"][rb.POSE2POSE2]
    ab = _check(ctx, w, rb.POSE2POSE2, lambda m: O.sweep_pose2pose2(f["i0"], f["i1"], P, m), angle_cols=(2,))
    assert ab < 5e-7   # 1.2e6 evaluations x 3 components: the float32 per-particle arithmetic stays below half a micrometre
    g = w["families"][rb.PRIORPOSE2]
    _check(ctx, w, rb.PRIORPOSE2, lambda m: O.sweep_priorpose2(g["i0"], P, m), angle_cols=(2,))


def test_config5_se3_chain_every_row(ctx):
    fg = rb.generateGraph_Pose3Chain(10000, loops=1000)
    rb.seed_particles(fg, N=100, seed=4)
    w = W.graph_arrays(fg, 100)
    P = w["particles"][rb.POSE3]
    ctx.set_particles(rb.POSE3, P)
    for fam in (rb.POSE3POSE3, rb.PRIORPOSE3):
        f = w["families"][fam]
        W.upload_family(ctx, fam, f["i0"], f["i1"], f["a"], f["b"])
    f = w["families"][rb.POSE3POSE3]
    assert len(f["i0"]) == 10999
    _check(ctx, w, rb.POSE3POSE3, lambda m: O.sweep_pose3pose3(f["i0"], f["i1"], P, m))
    g = w["families"][rb.PRIORPOSE3]
    _check(ctx, w, rb.PRIORPOSE3, lambda m: O.sweep_priorpose3(g["i0"], P, m))


def test_config4_beehive_every_row(ctx):
    fg = rb.generateGraph_Beehive(10000, N=200)
    rb.seed_particles(fg, N=200, seed=3)
    w = W.graph_arrays(fg, 200)
    P, Lm = w["particles"][rb.POSE2], w["particles"][rb.POINT2]
    ctx.set_particles(rb.POSE2, P)
    ctx.set_particles(rb.POINT2, Lm)
    for fam in (rb.POSE2POSE2, rb.BEARINGRANGE):
        f = w["families"][fam]
        W.upload_family(ctx, fam, f["i0"], f["i1"], f["a"], f["b"])
    f = w["families"][rb.POSE2POSE2]
    _check(ctx, w, rb.POSE2POSE2, lambda m: O.sweep_pose2pose2(f["i0"], f["i1"], P, m), angle_cols=(2,))
    g = w["families"][rb.BEARINGRANGE]
    _check(ctx, w, rb.BEARINGRANGE, lambda m: O.sweep_bearingrange(g["i0"], g["i1"], P, Lm, m), angle_cols=(0,))


def test_heading_spread_across_the_branch_cut(ctx):
    """N = 100 particles per pose whose headings straddle +-pi (anchor near pi, offsets of both signs) on a
    Manhattan-like chain: the wrapped heading offsets of the particle store and the residual's angle wrap must agree with
    the oracle's atan(sin, cos) convention; default and PRECISE arithmetic"""
    rng = np.random.default_rng(33)
    nv, N = 400, 100
    w = W.manhattan_arrays(nv, seed=7, N=N, particle_seed=8)
    P = w["particles"][rb.POSE2].copy()
    # every third pose: mean heading exactly at the cut, spread 0.4 rad -> about half the particles on either side
    for v in range(0, nv, 3):
        P[v, :, 2] = O.np_wrap(np.pi + rng.normal(size=N) * 0.4)
    for v in range(1, nv, 3):
        P[v, :, 2] = O.np_wrap(-np.pi + 1e-9 + rng.normal(size=N) * 1e-3)
    w["particles"][rb.POSE2] = P
    ctx.set_particles(rb.POSE2, P)
    back = ctx.get_particles(rb.POSE2)
    d = back - P
    d[..., 2] = O.np_wrap(d[..., 2])
    assert np.abs(d).max() < 1e-6
    f = w["families"][rb.POSE2POSE2]
    W.upload_family(ctx, rb.POSE2POSE2, f["i0"], f["i1"], f["a"], f["b"])
    _check(ctx, w, rb.POSE2POSE2, lambda m: O.sweep_pose2pose2(f["i0"], f["i1"], P, m), angle_cols=(2,))
