"""Parametric solve against the reference's deterministic parametric tests: every residual and
finite-difference column comes from the CUDA residual kernels, the host only solves the sparse linear system."""
import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O

pytestmark = pytest.mark.gpu
pi = np.pi


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


def test_square_loop(ctx):
    """test/testParametric.jl:16-57: prior (10, 10, -pi + 1e-5), 4 x Pose2Pose2 (10, 0, pi/2)"""
    fg = rb.initfg()
    rb.addVariable(fg, "x0", rb.Pose2)
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal([10, 10, -pi + 1e-5], [0.1, 0.1, 0.05])))
    for i in range(4):
        rb.addVariable(fg, f"x{i + 1}", rb.Pose2)
        rb.addFactor(fg, [f"x{i}", f"x{i + 1}"], rb.Pose2Pose2(rb.MvNormal([10.0, 0, pi / 2], [0.1, 0.1, 0.1])))
    rb.initAll(fg, seed=1, ctx=ctx)
    labels, x, cost, Sigma = rb.solveGraphParametric(fg, ctx=ctx)
    expect = {"x0": [10, 10, -pi], "x1": [0, 10, -pi / 2], "x2": [0, 0, 0], "x3": [10, 0, pi / 2], "x4": [10, 10, -pi]}
    for l, e in expect.items():
        d = x[l] - np.array(e)
        d[2] = O.np_wrap(d[2])
        assert np.all(np.abs(d) < 1e-3), (l, x[l])
    assert cost < 1e-8 and Sigma.shape == (15, 15)
    # marginal covariance of :x0 = the prior's (nothing else constrains it absolutely)
    assert np.allclose(np.sqrt(np.diag(Sigma)[:3]), [0.1, 0.1, 0.05], rtol=1e-2)


def test_information_weighting(ctx):
    """test/testParametricCovariances.jl:41-52: x1 = (1.05, 0, 0)"""
    fg = rb.initfg()
    rb.addVariable(fg, "x0", rb.Pose2)
    rb.addVariable(fg, "x1", rb.Pose2)
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal([0.0, 0, 0], np.diag([0.1, 0.1, 0.01]) ** 2)))
    rb.addFactor(fg, ["x0", "x1"], rb.Pose2Pose2(rb.MvNormal([1.1, 0, 0], np.diag([0.1, 0.1, 0.01]) ** 2)))
    rb.addFactor(fg, ["x0", "x1"], rb.Pose2Pose2(rb.MvNormal([0.9, 0, 0], np.diag([np.sqrt(0.03), 0.1, 0.01]) ** 2)))
    labels, x, cost, Sigma = rb.solveGraphParametric(fg, ctx=ctx)
    assert np.allclose(x["x0"], [0, 0, 0], atol=1e-4)
    assert np.allclose(x["x1"], [1.05, 0, 0], atol=1e-4)


def test_bearing_range_triangulation(ctx):
    """test/testParametric.jl:155-177: two landmarks with priors, two sightings -> pose ((2, 0), R(pi/2))"""
    fg = rb.initfg()
    rb.addVariable(fg, "x1", rb.Pose2)
    rb.addVariable(fg, "l1", rb.Point2)
    rb.addVariable(fg, "l2", rb.Point2)
    rb.addFactor(fg, ["l1"], rb.PriorPoint2(rb.MvNormal([1.0, 1], [0.01, 0.01])))
    rb.addFactor(fg, ["l2"], rb.PriorPoint2(rb.MvNormal([1.0, -1], [0.01, 0.01])))
    rb.addFactor(fg, ["x1", "l1"], rb.Pose2Point2BearingRange(rb.Normal(pi / 4, 0.01), rb.Normal(np.sqrt(2), 0.1)))
    rb.addFactor(fg, ["x1", "l2"], rb.Pose2Point2BearingRange(rb.Normal(3 * pi / 4, 0.01), rb.Normal(np.sqrt(2), 0.1)))
    fg["x1"].parametric = np.array([1.5, 0.3, 1.2])  # the reference starts from initAll!'s (noisy) estimate
    fg["l1"].parametric = np.array([1.1, 0.9])
    fg["l2"].parametric = np.array([0.9, -1.1])
    labels, x, cost, Sigma = rb.solveGraphParametric(fg, ctx=ctx)
    assert np.allclose(x["x1"][:2], [2, 0], atol=1e-3) and abs(O.np_wrap(x["x1"][2] - pi / 2)) < 1e-3
    assert np.allclose(x["l1"], [1, 1], atol=1e-3) and np.allclose(x["l2"], [1, -1], atol=1e-3)


def test_pose3_loop(ctx):
    """test/testPose3.jl:27-56: prior pitch -pi/4, 4 x Pose3Pose3 (sqrt 2, 0, 0, 0, 0, pi/2): x4 closes onto x0"""
    fg = rb.initfg()
    rb.addVariable(fg, "x0", rb.Pose3)
    rb.addFactor(fg, ["x0"], rb.PriorPose3(rb.MvNormal([0.0, 0, 0, 0, -pi / 4, 0],
                                                       np.diag([0.1, 0.1, 0.1, 0.01, 0.01, 0.01]) ** 2)))
    odo = rb.MvNormal([np.sqrt(2), 0, 0, 0, 0, pi / 2], np.diag([0.1, 0.1, 0.1, 0.01, 0.01, 0.01]) ** 2)
    for i in range(1, 5):
        rb.addVariable(fg, f"x{i}", rb.Pose3)
        rb.addFactor(fg, [f"x{i - 1}", f"x{i}"], rb.Pose3Pose3(odo))
    rb.initAll(fg, seed=2, ctx=ctx)
    labels, x, cost, Sigma = rb.solveGraphParametric(fg, ctx=ctx)
    assert cost < 1e-8
    assert np.allclose(x["x0"][:3], x["x4"][:3], atol=1e-3)
    Rd = O.np_so3_exp(x["x0"][3:]).T @ O.np_so3_exp(x["x4"][3:])
    assert np.abs(O.np_so3_log(Rd)).max() < 1e-3
    assert np.allclose(x["x0"], [0, 0, 0, 0, -pi / 4, 0], atol=1e-3)
    # every odometry residual vanishes at the solution (float64 oracle on the solved coordinates)
    for i in range(1, 5):
        assert np.abs(O.pose3pose3(odo.mu, x[f"x{i - 1}"], x[f"x{i}"])).max() < 1e-4


def test_analytic_and_finite_difference_jacobians_give_the_same_solution(ctx):
    """the kernels' analytic Jacobian blocks (default for the hot families) against the finite-difference path on a graph
    with all five hot families' 2-D members and on the Pose3 loop"""
    fg = rb.generateGraph_Hexagonal()
    rb.initAll(fg, seed=1, ctx=ctx)
    start = {l: (v.val.mean(0) if v.val is not None else np.zeros(v.variableType.dim)) for l, v in fg.variables.items()}

    def solve(analytic):
        for l, v in fg.variables.items():
            v.parametric = start[l].copy()
        return rb.solveGraphParametric(fg, ctx=ctx, analytic=analytic)
    _, xa, ca, Sa = solve(True)
    _, xf, cf, Sf = solve(False)
    for l in xa:
        d = xa[l] - xf[l]
        if fg[l].variableType is rb.Pose2:
            d[2] = O.np_wrap(d[2])
        assert np.abs(d).max() < 2e-4, (l, d)
    assert abs(ca - cf) < 1e-6 * max(1.0, cf)
    assert np.allclose(Sa, Sf, rtol=2e-2, atol=1e-6)
    # known answer of the Hexagonal graph (test/testHexagonal2D_CliqByCliq.jl:37-79 ground truth)
    assert np.allclose(xa["x1"][:2], [10, 0], atol=0.5) and np.allclose(xa["l1"], [20, 0], atol=1.0)
