"""GPU parity of the 3-D next-row families through the C ABI: PriorPoint3, Point3Point3,
Pose3Pose3XYYaw, Pose3Pose3Rotation, Pose3Pose3UnitTrans.  Tolerances as in test_gpu_parity_raw.py."""
import json
import os

import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O
from test_gpu_parity_raw import FLOOR_SAME, assert_close, make_pose3, rand_cov, seen, seen_meas

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("N", [100, 200])
def test_point3_families(ctx, N):
    rng = np.random.default_rng(70 + N)
    nv, nF = 30, 61
    pts = rng.normal(size=(nv, 1, 3)) * 50 + rng.normal(size=(nv, N, 3)) * 0.3
    ctx.set_particles(rb.POINT3, pts)
    i0 = rng.integers(0, nv, nF).astype(np.int32)
    i1 = ((i0 + 1 + rng.integers(0, 3, nF)) % nv).astype(np.int32)
    cov = rand_cov(rng, nF, 3, [0.1, 0.2, 0.15])
    Lc = np.linalg.cholesky(cov)
    for fam, a, b, mu in ((rb.PRIORPOINT3, i0, None, pts.mean(1)[i0]),
                          (rb.POINT3POINT3, i0, i1, pts.mean(1)[i1] - pts.mean(1)[i0])):
        ctx.set_factors_gaussian(fam, a, b, mu, cov)
        meas = mu[:, None, :] + np.einsum("fij,fnj->fni", Lc, rng.normal(size=(nF, N, 3)))
        moff = rb.meas_to_offsets(meas, mu)
        flags = rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD
        out = ctx.alloc_host_outputs(fam, flags)
        ctx.eval_host(fam, flags, meas=moff, **out)
        res = rb.rows_to_particle_major(out["res"], N)
        S, M = seen(ctx, rb.POINT3, N), seen_meas(moff, mu, N)
        ref = M - S[a] if b is None else M - (S[b] - S[a])
        assert_close(res, ref, what=f"family {fam} (A)", floor=FLOOR_SAME)
        ref_b = meas - pts[a] if b is None else meas - (pts[b] - pts[a])
        assert_close(res, ref_b, what=f"family {fam} (B)")
        assert np.allclose(out["stats"][:, :3], res.sum(1), rtol=1e-3, atol=1e-3)
        tgt = a if b is None else b
        prop_off = rb.rows_to_particle_major(out["prop_fwd"], N)
        prop = prop_off + ctx.get_anchors(rb.POINT3)[tgt][:, None, :]
        r0 = M - prop if b is None else M - (prop - S[a])
        assert np.abs(r0).max() < 2e-5
        assert np.allclose(out["stats"][:, 9:12], prop_off.sum(1), rtol=1e-3, atol=1e-2)
        # fused sampling == supplied on the written-back samples; whitened draws are standard normal
        fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
        o1 = ctx.alloc_host_outputs(fam, fl)
        ctx.eval_host(fam, fl, seed=11, **o1)
        o2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
        ctx.eval_host(fam, rb.RESIDUAL, meas=o1["meas_out"], **o2)
        assert np.array_equal(o1["res"][:, :N], o2["res"][:, :N])
        d = rb.rows_to_particle_major(o1["meas_out"], N)
        zw = np.linalg.solve(Lc, np.transpose(d, (0, 2, 1)))
        assert np.abs(np.cov(np.transpose(zw, (1, 0, 2)).reshape(3, -1)) - np.eye(3)).max() < 0.08


@pytest.mark.parametrize("N", [100, 48])
def test_pose3_partial_families(ctx, N):
    rng = np.random.default_rng(80 + N)
    nvars, nF = 31, 83
    poses = make_pose3(rng, nvars, N)
    ip = rng.integers(0, nvars - 3, nF).astype(np.int32)
    iq = (ip + rng.integers(1, 4, nF)).astype(np.int32)
    ctx.set_particles(rb.POSE3, poses)
    P0, Q0 = poses[ip, 0], poses[iq, 0]
    # measurement means consistent with the first particles so residuals are small
    mu_xyy = np.zeros((nF, 3))
    mu_rot = np.zeros((nF, 3))
    mu_se3 = np.zeros((nF, 6))
    for f in range(nF):
        Rp, Rq = O.so3_exp(P0[f, 3:]), O.so3_exp(Q0[f, 3:])
        yp, yq = np.arctan2(Rp[1, 0], Rp[0, 0]), np.arctan2(Rq[1, 0], Rq[0, 0])
        d = Q0[f, :2] - P0[f, :2]
        mu_xyy[f] = [np.cos(yp) * d[0] + np.sin(yp) * d[1], -np.sin(yp) * d[0] + np.cos(yp) * d[1], O.np_wrap(yq - yp)]
        mu_rot[f] = O.so3_log(Rp.T @ Rq)
        mu_se3[f, :3] = Rp.T @ (Q0[f, :3] - P0[f, :3]) * 0.5  # half the step: the unit translation residual is O(1)
        mu_se3[f, 3:] = mu_rot[f]
    cases = [
        (rb.POSE3POSE3XYYAW, mu_xyy, [0.1, 0.1, 0.02], O.np_pose3pose3xyyaw, (2,)),
        (rb.POSE3POSE3ROTATION, mu_rot, [0.01, 0.01, 0.01], O.np_pose3pose3rotation, ()),
        (rb.POSE3POSE3UNITTRANS, mu_se3, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01], O.np_pose3pose3unittrans, ()),
    ]
    for fam, mu, scale, ref_fn, ang in cases:
        d = mu.shape[1]
        cov = rand_cov(rng, nF, d, scale)
        Lc = np.linalg.cholesky(cov)
        ctx.set_factors_gaussian(fam, ip, iq, mu, cov)
        meas = mu[:, None, :] + np.einsum("fij,fnj->fni", Lc, rng.normal(size=(nF, N, d)))
        moff = rb.meas_to_offsets(meas, mu)
        flags = rb.RESIDUAL | rb.STATS
        out = ctx.alloc_host_outputs(fam, flags)
        ctx.eval_host(fam, flags, meas=moff, **out)
        res = rb.rows_to_particle_major(out["res"], N)
        S = seen(ctx, rb.POSE3, N)
        assert_close(res, ref_fn(seen_meas(moff, mu, N), S[ip], S[iq]), angle_cols=ang, what=f"family {fam} (A)",
                     floor=FLOOR_SAME)
        assert_close(res, ref_fn(meas, poses[ip], poses[iq]), angle_cols=ang, what=f"family {fam} (B)")
        assert np.allclose(out["stats"][:, :d], res.sum(1), rtol=1e-3, atol=1e-3)
        with pytest.raises(rb.RomeB200Error):  # partial constraints have no closed-form proposal
            bad = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
            ctx.eval_host(fam, rb.RESIDUAL | rb.PROPOSAL_FWD, meas=moff, prop_fwd=np.zeros((nF, rb.npad(N), d), np.float32),
                          **bad)
        fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
        o1 = ctx.alloc_host_outputs(fam, fl)
        ctx.eval_host(fam, fl, seed=13, **o1)
        o2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
        ctx.eval_host(fam, rb.RESIDUAL, meas=o1["meas_out"], **o2)
        assert np.array_equal(o1["res"][:, :N], o2["res"][:, :N])
        dm = rb.rows_to_particle_major(o1["meas_out"], N)
        zw = np.linalg.solve(Lc, np.transpose(dm, (0, 2, 1)))
        assert np.abs(np.cov(np.transpose(zw, (1, 0, 2)).reshape(d, -1)) - np.eye(d)).max() < 0.08


def test_known_answers_3d(ctx, golden_dir):
    """test/testPartialPose3.jl known answers through calcFactorResidualTemporary (the reference's own probe)"""
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    I3 = np.diag([0.01, 0.01, 0.001]) ** 2
    for c in ka["pose3pose3xyyaw"]:
        f = rb.Pose3Pose3XYYaw(rb.MvNormal(c["X"], I3))
        r = rb.calcFactorResidualTemporary(f, (rb.Pose3, rb.Pose3), c["X"], (c["p"], c["q"]), ctx=ctx)
        d = r - np.array(c["expect"], dtype=float)
        d[2] = O.np_wrap(d[2])
        assert np.all(np.abs(d) < 1e-5), (c["src"], r)
    for c in ka["pose3pose3rotation"]:
        f = rb.Pose3Pose3Rotation(rb.MvNormal(c["m"], 0.001 * np.eye(3)))
        r = rb.calcFactorResidualTemporary(f, (rb.Pose3, rb.Pose3), c["m"], (c["p"], c["q"]), ctx=ctx)
        assert np.linalg.norm(r) < 1e-7, (c["src"], r)
    # Point3 sign convention mirrors the Point2 check of test/testPartialPose3.jl:264-269
    f = rb.Point3Point3(rb.MvNormal([20.0, 5.0, 1.0], np.diag([0.01, 0.01, 0.01]) ** 2))
    r = rb.calcFactorResidualTemporary(f, (rb.Point3, rb.Point3), [20.0, 5.0, 1.0], ([10, 0, 0], [20, 10, 3]), ctx=ctx)
    assert np.allclose(r, [10, -5, -2], atol=1e-6)
    f = rb.PriorPoint3(rb.MvNormal(np.zeros(3), np.eye(3)))
    r = rb.calcFactorResidualTemporary(f, (rb.Point3,), [1.0, 2.0, 3.0], ([0.5, 0.5, 0.5],), ctx=ctx)
    assert np.allclose(r, [0.5, 1.5, 2.5], atol=1e-6)
    # convolution onto the second point is closed-form; partial constraints refuse
    fg = rb.initfg(rb.SolverParams(N=64))
    rb.addVariable(fg, "a", rb.Point3); rb.addVariable(fg, "b", rb.Point3)
    rb.setVal(fg, "a", np.tile([1.0, 2.0, 3.0], (64, 1)))
    rb.addFactor(fg, ["a", "b"], rb.Point3Point3(rb.MvNormal([1.0, 1.0, 1.0], np.eye(3) * 1e-4)))
    pts = rb.approxConv(fg, "abf1", "b", ctx=ctx)
    assert pts.shape == (64, 3) and np.allclose(pts.mean(0), [2, 3, 4], atol=0.02)


@pytest.mark.parametrize("N", [100, 48])
def test_pose3_ternary_families(ctx, N):
    """Pose3Pose3RotOffset (third variable: Rotation3) and Pose3Pose3Transform (third variable: Pose3),
    src/factors/Pose3Pose3.jl:57-95: residual parity (A) / (B), statistics, forward proposal = the root for the SECOND
    variable, fused sampling == supplied on the written-back samples."""
    rng = np.random.default_rng(90 + N)
    nvars, nF, nrot = 33, 77, 5
    poses = make_pose3(rng, nvars, N)
    rots = rng.normal(size=(nrot, 1, 3)) * 0.4 + rng.normal(size=(nrot, N, 3)) * 0.03
    ip = rng.integers(0, nvars - 6, nF).astype(np.int32)
    iq = (ip + rng.integers(1, 4, nF)).astype(np.int32)
    ir_rot = rng.integers(0, nrot, nF).astype(np.int32)
    ir_pose = (nvars - 3 + rng.integers(0, 3, nF)).astype(np.int32)  # the last three poses serve as extrinsics Delta
    poses[nvars - 3:] = rng.normal(size=(3, 1, 6)) * [0.5, 0.5, 0.5, 0.3, 0.3, 0.3] + rng.normal(size=(3, N, 6)) * 0.02
    ctx.set_particles(rb.POSE3, poses)
    ctx.set_particles(rb.ROTATION3, rots)
    P0, Q0 = poses[ip, 0], poses[iq, 0]
    for fam, i2, third, ref_fn in ((rb.POSE3POSE3ROTOFFSET, ir_rot, rots, O.np_pose3pose3rotoffset),
                                   (rb.POSE3POSE3TRANSFORM, ir_pose, poses, O.np_pose3pose3transform)):
        T0 = third[i2, 0]
        # measurement means consistent with the first particles, so that residuals are small
        mu = np.zeros((nF, 6))
        for f in range(nF):
            Rp, Rq = O.so3_exp(P0[f, 3:]), O.so3_exp(Q0[f, 3:])
            if fam == rb.POSE3POSE3ROTOFFSET:
                B = O.so3_exp(T0[f])
                mu[f, :3] = Rp.T @ (Q0[f, :3] - P0[f, :3])
                mu[f, 3:] = O.so3_log(B.T @ Rp.T @ Rq)
            else:
                RD = O.so3_exp(T0[f, 3:])
                mu[f, :3] = RD.T @ (Rp.T @ (Q0[f, :3] - P0[f, :3]) - T0[f, :3])
                mu[f, 3:] = O.so3_log(RD.T @ Rp.T @ Rq)
        cov = rand_cov(rng, nF, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
        Lc = np.linalg.cholesky(cov)
        with pytest.raises(rb.RomeB200Error):  # the binary entry point refuses a family with a third variable
            ctx.set_factors_gaussian(fam, ip, iq, mu, cov)
        ctx.set_factors_ternary(fam, ip, iq, i2, mu, cov)
        assert ctx.num_factors(fam) == nF
        meas = mu[:, None, :] + np.einsum("fij,fnj->fni", Lc, rng.normal(size=(nF, N, 6)))
        moff = rb.meas_to_offsets(meas, mu)
        flags = rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD
        out = ctx.alloc_host_outputs(fam, flags)
        ctx.eval_host(fam, flags, meas=moff, **out)
        res = rb.rows_to_particle_major(out["res"], N)
        S = seen(ctx, rb.POSE3, N)
        S3 = S if fam == rb.POSE3POSE3TRANSFORM else seen(ctx, rb.ROTATION3, N)
        M = seen_meas(moff, mu, N)
        assert_close(res, ref_fn(M, S[ip], S[iq], S3[i2]), what=f"family {fam} (A)", floor=FLOOR_SAME)
        assert_close(res, ref_fn(meas, poses[ip], poses[iq], third[i2]), what=f"family {fam} (B)")
        assert np.abs(res).max() < 2.0  # consistent factors: the comparison is a relative one on small residuals
        assert np.allclose(out["stats"][:, :6], res.sum(1), rtol=1e-3, atol=1e-3)
        # the forward proposal is the root for the second variable: residual(meas, p, proposal, third) = 0
        prop = rb.rows_to_particle_major(out["prop_fwd"], N) + ctx.get_anchors(rb.POSE3)[iq][:, None, :]
        assert np.abs(ref_fn(M, S[ip], prop, S3[i2])).max() < 2e-5
        # the compile-time RESIDUAL|STATS variant computes the same rows
        o9 = ctx.alloc_host_outputs(fam, rb.RESIDUAL | rb.STATS)
        ctx.eval_host(fam, rb.RESIDUAL | rb.STATS, meas=moff, **o9)
        assert np.array_equal(o9["res"][:, :N], out["res"][:, :N])
        fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
        o1 = ctx.alloc_host_outputs(fam, fl)
        ctx.eval_host(fam, fl, seed=17, **o1)
        o2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
        ctx.eval_host(fam, rb.RESIDUAL, meas=o1["meas_out"], **o2)
        assert np.array_equal(o1["res"][:, :N], o2["res"][:, :N])
        dm = rb.rows_to_particle_major(o1["meas_out"], N)
        zw = np.linalg.solve(Lc, np.transpose(dm, (0, 2, 1)))
        assert np.abs(np.cov(np.transpose(zw, (1, 0, 2)).reshape(6, -1)) - np.eye(6)).max() < 0.08
    # Rotation3 particles round-trip as rotations, reported as principal rotation vectors
    back = ctx.get_particles(rb.ROTATION3)
    assert np.abs(back - rots).max() < 1e-6


def test_rotoffset_known_answer_and_graph_api(ctx):
    """test/testPose3.jl:72-125: x0 = (0, R_z(pi/2)), odometry (1,0,0, 0,0,0.1) measured in a frame rotated by bRa about z:
    the parametric solution the reference asserts -- x1 = (0,1,0), x2 = (0,2,0), bRa = (0,0,-0.1) -- has zero residual."""
    odo = rb.MvNormal([1.0, 0, 0, 0, 0, 0.1], np.diag([0.1, 0.1, 0.1, 0.01, 0.01, 0.01]) ** 2)
    f = rb.Pose3Pose3RotOffset(odo)
    x0, x1, x2, bRa = [0, 0, 0, 0, 0, np.pi / 2], [0, 1.0, 0, 0, 0, np.pi / 2], [0, 2.0, 0, 0, 0, np.pi / 2], [0, 0, -0.1]
    for p, q in ((x0, x1), (x1, x2)):
        r = rb.calcFactorResidualTemporary(f, (rb.Pose3, rb.Pose3, rb.Rotation3), odo.mu, (p, q, bRa), ctx=ctx)
        assert np.abs(r).max() < 1e-6, r
        assert np.abs(O.pose3pose3rotoffset(odo.mu, p, q, bRa)).max() < 1e-12
    r = rb.calcFactorResidualTemporary(f, (rb.Pose3, rb.Pose3, rb.Rotation3), odo.mu, (x0, x1, [0, 0, 0.0]), ctx=ctx)
    assert abs(r[5] - 0.1) < 1e-6 and np.abs(r[:5]).max() < 1e-6   # without the offset the yaw is off by the measured 0.1
    # Transform with Delta = a pure translation along x: the step is Delta.t + R_Delta m.t = 2 m along p's x axis
    g = rb.Pose3Pose3Transform(rb.MvNormal([1.0, 0, 0, 0, 0, 0], np.eye(6) * 1e-4))
    r = rb.calcFactorResidualTemporary(g, (rb.Pose3, rb.Pose3, rb.Pose3), g.Z.mu, (x0, [0, 2.0, 0, 0, 0, np.pi / 2], [1.0, 0, 0, 0, 0, 0]),
                                       ctx=ctx)
    assert np.abs(r).max() < 1e-6, r
    # graph API: the convolution onto the second variable is closed-form; onto the third the reference runs its solver
    fg = rb.initfg(rb.SolverParams(N=64))
    for l, t in (("x0", rb.Pose3), ("x1", rb.Pose3), ("bRa", rb.Rotation3)):
        rb.addVariable(fg, l, t)
    rb.setVal(fg, "x0", np.tile(x0, (64, 1)))
    rb.setVal(fg, "bRa", np.tile(bRa, (64, 1)))
    rb.addFactor(fg, ["x0", "x1", "bRa"], rb.Pose3Pose3RotOffset(rb.MvNormal(odo.mu, np.eye(6) * 1e-6)))
    pts = rb.approxConv(fg, "x0x1bRaf1", "x1", ctx=ctx)
    assert pts.shape == (64, 6) and np.allclose(pts.mean(0)[:3], [0, 1, 0], atol=0.01)
    assert np.allclose(O.so3_exp(pts.mean(0)[3:]), O.so3_exp(np.array([0, 0, np.pi / 2])), atol=0.01)
    with pytest.raises(NotImplementedError):
        rb.approxConv(fg, "x0x1bRaf1", "bRa", ctx=ctx)
