"""GPU parity of the next-row families through the C ABI: PriorPoint2, Point2Point2, Pose2Point2,
Pose2Point2Range, Point2Point2Range, Pose2Point2Bearing.  Same tolerance definitions as test_gpu_parity_raw.py."""
import json
import os

import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O
from test_gpu_parity_raw import FLOOR_SAME, assert_close, make_pose2_graph, rand_cov, seen, seen_meas

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


def _setup(ctx, rng, N, nvars=40, nl=25):
    poses, _, _ = make_pose2_graph(rng, nvars, 1, N)
    pts = poses.mean(1)[rng.integers(0, nvars, nl), :2][:, None, :] + rng.normal(size=(nl, 1, 2)) * 8 \
        + rng.normal(size=(nl, N, 2)) * 0.3
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_particles(rb.POINT2, pts)
    return poses, pts


@pytest.mark.parametrize("N", [100, 200])
def test_point2_gaussian_families(ctx, N):
    rng = np.random.default_rng(40 + N)
    poses, pts = _setup(ctx, rng, N)
    nF = 77
    i0p = rng.integers(0, len(pts), nF).astype(np.int32)
    i1p = ((i0p + 1 + rng.integers(0, 3, nF)) % len(pts)).astype(np.int32)
    i0s = rng.integers(0, len(poses), nF).astype(np.int32)
    cov = rand_cov(rng, nF, 2, [0.1, 0.2])
    Lc = np.linalg.cholesky(cov)
    cases = [
        (rb.PRIORPOINT2, i0p, None, pts.mean(1)[i0p], lambda m: O.np_priorpoint2(m, seen(ctx, rb.POINT2, N)[i0p])),
        (rb.POINT2POINT2, i0p, i1p, pts.mean(1)[i1p] - pts.mean(1)[i0p],
         lambda m: O.np_point2point2(m, seen(ctx, rb.POINT2, N)[i0p], seen(ctx, rb.POINT2, N)[i1p])),
        (rb.POSE2POINT2, i0s, i1p, rng.normal(size=(nF, 2)) * 5,
         lambda m: O.np_pose2point2(m, seen(ctx, rb.POSE2, N)[i0s], seen(ctx, rb.POINT2, N)[i1p])),
    ]
    for fam, i0, i1, mu, ref_fn in cases:
        ctx.set_factors_point2(fam, i0, i1, mu, cov)
        meas = mu[:, None, :] + np.einsum("fij,fnj->fni", Lc, rng.normal(size=(nF, N, 2)))
        moff = rb.meas_to_offsets(meas, mu)
        flags = rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD
        out = ctx.alloc_host_outputs(fam, flags)
        ctx.eval_host(fam, flags, meas=moff, **out)
        res = rb.rows_to_particle_major(out["res"], N)
        assert_close(res, ref_fn(seen_meas(moff, mu, N)), what=f"family {fam} (A)", floor=FLOOR_SAME)
        assert np.allclose(out["stats"][:, :2], res.sum(1), rtol=1e-3, atol=1e-3)
        # the forward proposal is a root of the residual
        tgt_idx = i0 if i1 is None else i1
        prop = rb.rows_to_particle_major(out["prop_fwd"], N) + ctx.get_anchors(rb.POINT2)[tgt_idx][:, None, :]
        P2, PT = seen(ctx, rb.POSE2, N), seen(ctx, rb.POINT2, N)
        if fam == rb.PRIORPOINT2:
            r0 = O.np_priorpoint2(seen_meas(moff, mu, N), prop)
        elif fam == rb.POINT2POINT2:
            r0 = O.np_point2point2(seen_meas(moff, mu, N), PT[i0], prop)
        else:
            r0 = O.np_pose2point2(seen_meas(moff, mu, N), P2[i0], prop)
        assert np.abs(r0).max() < 2e-5
        # fused sampling == supplied on the written-back samples
        fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
        o1 = ctx.alloc_host_outputs(fam, fl)
        ctx.eval_host(fam, fl, seed=9, **o1)
        o2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
        ctx.eval_host(fam, rb.RESIDUAL, meas=o1["meas_out"], **o2)
        assert np.array_equal(o1["res"][:, :N], o2["res"][:, :N])
        d = rb.rows_to_particle_major(o1["meas_out"], N)
        zw = np.linalg.solve(Lc, np.transpose(d, (0, 2, 1)))
        assert np.abs(np.cov(np.transpose(zw, (1, 0, 2)).reshape(2, -1)) - np.eye(2)).max() < 0.08


@pytest.mark.parametrize("N", [100, 37])
def test_scalar_families(ctx, N):
    rng = np.random.default_rng(50 + N)
    poses, pts = _setup(ctx, rng, N)
    nF = 61
    i0s = rng.integers(0, len(poses), nF).astype(np.int32)
    i0p = rng.integers(0, len(pts), nF).astype(np.int32)
    i1p = ((i0p + 1) % len(pts)).astype(np.int32)
    for fam, i0, vt0, fn in ((rb.POSE2POINT2RANGE, i0s, rb.POSE2, O.np_range2), (rb.POINT2POINT2RANGE, i0p, rb.POINT2, O.np_range2),
                             (rb.POSE2POINT2BEARING, i0s, rb.POSE2, O.np_pose2point2bearing)):
        mean = rng.uniform(5, 30, nF) if fam != rb.POSE2POINT2BEARING else rng.uniform(-3, 3, nF)
        belief = np.column_stack([mean, rng.uniform(0.05, 0.5, nF)])
        ctx.set_factors_scalar(fam, i0, i1p, belief)
        meas = mean[:, None, None] + belief[:, 1][:, None, None] * rng.normal(size=(nF, N, 1))
        moff = rb.meas_to_offsets(meas, mean[:, None])
        out = ctx.alloc_host_outputs(fam, rb.RESIDUAL | rb.STATS)
        ctx.eval_host(fam, rb.RESIDUAL | rb.STATS, meas=moff, **out)
        res = rb.rows_to_particle_major(out["res"], N)
        ref = fn(seen_meas(moff, mean[:, None], N), seen(ctx, vt0, N)[i0], seen(ctx, rb.POINT2, N)[i1p])
        assert_close(res, ref, angle_cols=(0,) if fam == rb.POSE2POINT2BEARING else (), what=f"family {fam} (A)",
                     floor=FLOOR_SAME)
        assert np.allclose(out["stats"][:, 0], res[..., 0].sum(1), rtol=1e-3, atol=1e-3)
        with pytest.raises(rb.RomeB200Error):  # no closed-form proposal for a scalar factor
            ctx.eval_host(fam, rb.PROPOSAL_FWD, meas=moff, prop_fwd=np.zeros(4, np.float32))
        fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
        o1 = ctx.alloc_host_outputs(fam, fl)
        ctx.eval_host(fam, fl, seed=2, **o1)
        o2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
        ctx.eval_host(fam, rb.RESIDUAL, meas=o1["meas_out"], **o2)
        assert np.array_equal(o1["res"][:, :N], o2["res"][:, :N])
        zs = rb.rows_to_particle_major(o1["meas_out"], N)[..., 0] / belief[:, 1][:, None]
        assert abs(zs.mean()) < 0.06 and abs(zs.std() - 1) < 0.06


def test_bearing_known_answers(golden_dir):
    """test/testBearing2D.jl:13-47, :59-66 through calcFactorResidualTemporary"""
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    for c in ka["pose2point2bearing"]:
        f = rb.Pose2Point2Bearing(rb.Normal(c["b"], 0.05))
        r = rb.calcFactorResidualTemporary(f, (rb.Pose2, rb.Point2), [c["b"]], (c["p"], c["l"]))
        d = r - np.array(c["expect"])
        if c.get("modulo_2pi"):
            d = O.np_wrap(d)
        assert np.all(np.abs(d) < c["atol"]), (c, r)


def test_graph_api_with_point_factors():
    """Boxes2D-style fragment: priors on two points, a Point2Point2 between them, range and bearing from a pose"""
    fg = rb.initfg()
    rb.addVariable(fg, "l0", rb.Point2)
    rb.addVariable(fg, "l1", rb.Point2)
    rb.addVariable(fg, "x0", rb.Pose2)
    rb.addFactor(fg, ["l0"], rb.PriorPoint2(rb.MvNormal([1.0, 2.0], 0.01 * np.eye(2))))
    rb.addFactor(fg, ["l0", "l1"], rb.Point2Point2(rb.MvNormal([10.0, 0.0], 0.01 * np.eye(2))))
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal([0.0, 0.0, 0.0], 0.01 * np.eye(3))))
    rb.addFactor(fg, ["x0", "l1"], rb.Pose2Point2Range(rb.Normal(11.2, 0.1)))
    rb.addFactor(fg, ["x0", "l1"], rb.Pose2Point2Bearing(rb.Normal(0.18, 0.01)))
    rb.addFactor(fg, ["x0", "l0"], rb.Pose2Point2(rb.MvNormal([1.0, 2.0], 0.01 * np.eye(2))))
    rb.initAll(fg, seed=1)
    assert np.allclose(rb.getVal(fg, "l0").mean(0), [1, 2], atol=0.1)
    assert np.allclose(rb.getVal(fg, "l1").mean(0), [11, 2], atol=0.15)
    dg = rb.DeviceGraph(fg)
    r = dg.eval(rb.POSE2POINT2RANGE, rb.RESIDUAL | rb.SAMPLE)["res"]
    assert abs(r.mean()) < 0.2  # |l1 - x0| = sqrt(11^2 + 2^2) = 11.18
    b = dg.eval(rb.POSE2POINT2BEARING, rb.RESIDUAL | rb.SAMPLE)["res"]
    assert abs(b.mean()) < 0.05  # atan(2/11) = 0.18
    with pytest.raises(NotImplementedError):
        rb.approxConv(fg, "x0l1f1", "l1")
