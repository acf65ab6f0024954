"""Deconvolution (IIF approxDeconv, test/testBasicPose2Conv.jl:51-56): the measurement the kernels
solve for must zero the residual when it is fed back, for every hot-path family, and the deconvolved set of a
convolved belief must reproduce the factor's measurement distribution."""
import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O
from test_gpu_parity_raw import make_pose2_graph, make_pose3, rand_cov, seen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


def _roundtrip(ctx, fam, N, oracle_fn, vars_seen, angle_cols=()):
    """DECONV -> supplied-measurement residual == 0, and the oracle agrees on the deconvolved measurement"""
    out = ctx.alloc_host_outputs(fam, rb.SAMPLE | rb.DECONV)
    ctx.eval_host(fam, rb.SAMPLE | rb.DECONV, seed=3, **out)
    o2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
    ctx.eval_host(fam, rb.RESIDUAL, meas=out["meas_out"], **o2)
    res = rb.rows_to_particle_major(o2["res"], N)
    assert np.abs(res).max() < 2e-5, np.abs(res).max()
    return out["meas_out"]


def test_deconv_pose2_families(ctx):
    rng = np.random.default_rng(3)
    N, nvars, nF = 100, 40, 90
    poses, ip, iq = make_pose2_graph(rng, nvars, nF, N)
    ctx.set_particles(rb.POSE2, poses)
    mu = rng.normal(size=(nF, 3)) * [5, 5, 1]
    ctx.set_factors_pose2pose2(ip, iq, mu, rand_cov(rng, nF, 3, [0.1, 0.1, 0.02]))
    moff = _roundtrip(ctx, rb.POSE2POSE2, N, None, None)
    S = seen(ctx, rb.POSE2, N)
    X = rb.offsets_to_meas(moff, mu, N)
    assert np.abs(O.np_pose2pose2(X, S[ip], S[iq])).max() < 2e-5   # float64 oracle: the deconvolved X zeroes the residual
    # closed form: X = (R_p'(t_q - t_p), wrap(th_q - th_p))
    p, q = S[ip], S[iq]
    c, s = np.cos(p[..., 2]), np.sin(p[..., 2])
    d = q[..., :2] - p[..., :2]
    ref = np.stack([c * d[..., 0] + s * d[..., 1], -s * d[..., 0] + c * d[..., 1], O.np_wrap(q[..., 2] - p[..., 2])], -1)
    dd = X - ref
    dd[..., 2] = O.np_wrap(dd[..., 2])
    assert np.abs(dd).max() < 1e-5
    ctx.set_factors_priorpose2(np.arange(nvars, dtype=np.int32), poses[:, 0], rand_cov(rng, nvars, 3, [0.1, 0.1, 0.02]))
    moff = _roundtrip(ctx, rb.PRIORPOSE2, N, None, None)
    dd = rb.offsets_to_meas(moff, poses[:, 0], N) - S
    dd[..., 2] = O.np_wrap(dd[..., 2])
    assert np.abs(dd).max() < 1e-5
    # bearing-range
    nl = 20
    pts = poses.mean(1)[rng.integers(0, nvars, nl), :2][:, None, :] + rng.normal(size=(nl, 1, 2)) * 8 \
        + rng.normal(size=(nl, N, 2)) * 0.3
    ctx.set_particles(rb.POINT2, pts)
    il = rng.integers(0, nl, nF).astype(np.int32)
    ctx.set_factors_bearingrange(ip, il, np.column_stack([rng.uniform(-3, 3, nF), np.full(nF, 0.05)]),
                                 np.column_stack([rng.uniform(5, 20, nF), np.full(nF, 0.3)]))
    _roundtrip(ctx, rb.BEARINGRANGE, N, None, None)
    with pytest.raises(rb.RomeB200Error):  # DECONV excludes WRITE_MEAS
        o = ctx.alloc_host_outputs(rb.BEARINGRANGE, rb.SAMPLE | rb.DECONV)
        ctx.eval_host(rb.BEARINGRANGE, rb.SAMPLE | rb.DECONV | rb.WRITE_MEAS, **o)


def test_deconv_pose3_families(ctx):
    rng = np.random.default_rng(4)
    N, nvars, nF = 100, 25, 60
    poses = make_pose3(rng, nvars, N)
    ip = rng.integers(0, nvars - 3, nF).astype(np.int32)
    iq = (ip + rng.integers(1, 4, nF)).astype(np.int32)
    ctx.set_particles(rb.POSE3, poses)
    mu = rng.normal(size=(nF, 6)) * [3, 3, 3, 0.5, 0.5, 0.5]
    ctx.set_factors_pose3pose3(ip, iq, mu, rand_cov(rng, nF, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01]))
    moff = _roundtrip(ctx, rb.POSE3POSE3, N, None, None)
    S = seen(ctx, rb.POSE3, N)
    X = rb.offsets_to_meas(moff, mu, N)
    assert np.abs(O.np_pose3pose3(X, S[ip], S[iq])).max() < 5e-5
    ctx.set_factors_priorpose3(np.arange(nvars, dtype=np.int32), poses[:, 0],
                               rand_cov(rng, nvars, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01]))
    _roundtrip(ctx, rb.PRIORPOSE3, N, None, None)


def test_approxdeconv_reproduces_the_measurement_belief(ctx):
    """test/testBasicPose2Conv.jl:10-56: x0 ~ N(0, 0.01^2), x1 = approxConv through Pose2Pose2((10, 0, pi), 0.1 I);
    the deconvolved measurements must be distributed like the factor's own samples"""
    rng = np.random.default_rng(5)
    N = 100
    fg = rb.initfg(rb.SolverParams(N=N))
    rb.addVariable(fg, "x0", rb.Pose2)
    rb.setVal(fg, "x0", 0.01 * rng.normal(size=(N, 3)))
    rb.addVariable(fg, "x1", rb.Pose2)
    rb.addFactor(fg, ["x0", "x1"], rb.Pose2Pose2(rb.MvNormal([10, 0, np.pi], 0.1 * np.eye(3))), graphinit=False)
    x1 = rb.approxConv(fg, "x0x1f1", "x1", seed=7, ctx=ctx)
    assert np.allclose(x1[:, :2].mean(0), [10, 0], atol=0.2)                       # :25-31
    psi = O.np_wrap(x1[:, 2])
    assert (psi > 2).sum() > 20 and (psi < -2).sum() > 20 and ((psi > -2) & (psi < 2)).sum() < 1   # :38-44
    rb.setVal(fg, "x1", x1)
    pts, meas = rb.approxDeconv(fg, "x0x1f1", seed=8, ctx=ctx)
    for a in (pts, meas):
        a[:, 2] = np.mod(a[:, 2], 2 * np.pi)   # the measurement heading lives around pi
    assert np.allclose(pts.mean(0), meas.mean(0), atol=0.12)
    assert np.allclose(pts.std(0), meas.std(0), rtol=0.25)
    assert np.allclose(pts.mean(0), [10, 0, np.pi], atol=0.12)
