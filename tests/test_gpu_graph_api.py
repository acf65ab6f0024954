"""GPU tests through the reference-named host API: the reference's own known-answer vectors via
calcFactorResidualTemporary, canonical graphs, golden data fixtures, approxConv / sampleFactor."""
import json
import math
import os

import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O

pytestmark = pytest.mark.gpu
pi = math.pi


@pytest.fixture(scope="module")
def ka(golden_dir):
    with open(os.path.join(golden_dir, "known_answers.json")) as fh:
        return json.load(fh)


def _check(r, case, atol_floor):
    r = np.asarray(r, float)
    atol = max(case.get("atol", 0.0), atol_floor)
    if "expect" in case:
        assert np.allclose(r, case["expect"], rtol=0, atol=atol), (case["src"], r)
    if "expect_abs" in case:
        assert np.allclose(np.abs(r), case["expect_abs"], rtol=0, atol=atol), (case["src"], r)
    if "expect_norm_below" in case:
        assert np.linalg.norm(r) < max(case["expect_norm_below"], atol_floor), (case["src"], r)


def test_known_answers_pose2pose2(ka):
    """test/testParametricSimulated.jl:37-46,99-144.  The reference asserts 1e-14 in Float64; the device stores
    the residual as float32, so exact zeros stay exact and pi is matched to float32 rounding (2.4e-7)."""
    f = rb.Pose2Pose2(rb.MvNormal([0, 0, -pi + 0.01], np.diag([0.03] * 3)))
    for c in ka["pose2pose2"]:
        r = rb.calcFactorResidualTemporary(f, (rb.Pose2, rb.Pose2), c["X"], (c["p"], c["q"]))
        _check(r, c, 3e-7)


def test_known_answers_bearingrange(ka):
    """test/testBearingRange2D.jl:55-250"""
    f = rb.Pose2Point2BearingRange(rb.Normal(0, 0.1), rb.Normal(20.0, 1.0))
    for c in ka["bearingrange"]:
        r = rb.calcFactorResidualTemporary(f, (rb.Pose2, rb.Point2), c["meas"], (c["p"], c["l"]))
        _check(r, c, 1e-7)


def test_known_answers_pose3pose3(ka):
    """test/testPartialPose3.jl:398-436, test/threeDimLinearProductTest.jl:150-167 (incl. [pi,pi,pi])"""
    # the factor mean is 0 here while the tested measurement is (10, .., pi, pi, pi): the measurement offset from
    # the mean is stored as float32, so pi carries float32 rounding (8.7e-8 per component)
    f = rb.Pose3Pose3()
    for c in ka["pose3pose3"]:
        r = rb.calcFactorResidualTemporary(f, (rb.Pose3, rb.Pose3), c["X"], (c["p"], c["q"]))
        _check(r, c, 5e-7)
        # with the factor mean AT the measurement the offset is exactly 0 and the residual is float32-exact
        g = rb.Pose3Pose3(rb.MvNormal(c["X"], np.diag([0.01] * 3 + [1e-4] * 3)))
        r = rb.calcFactorResidualTemporary(g, (rb.Pose3, rb.Pose3), c["X"], (c["p"], c["q"]))
        _check(r, c, 1e-12)


def test_priors_zero_at_measurement():
    r = rb.calcFactorResidualTemporary(rb.PriorPose2(), (rb.Pose2,), [1.0, -2.0, 3.0], ([1.0, -2.0, 3.0],))
    assert np.abs(r).max() == 0
    r = rb.calcFactorResidualTemporary(rb.PriorPose2(), (rb.Pose2,), [1.0, -2.0, 3.0], ([0.5, -1.0, -3.0],))
    assert np.allclose(r, O.priorpose2([1.0, -2.0, 3.0], [0.5, -1.0, -3.0]), atol=1e-6)
    m, p = [1, 2, 3, 0.1, -0.2, 0.3], [0.5, 2.5, 3, -0.3, 0.2, 0.1]
    r = rb.calcFactorResidualTemporary(rb.PriorPose3(), (rb.Pose3,), m, (p,))
    assert np.allclose(r, O.priorpose3(m, p), atol=1e-6)


def test_hexagonal_graph_all_families():
    """config C1/C2: generateGraph_Hexagonal, N=100 -- graph init by propagation, then every family's residual
    against the oracle on the same particles and the same (in-kernel drawn, written back) samples."""
    fg = rb.generateGraph_Hexagonal()
    rb.initAll(fg, seed=3)
    truth = {l: fg[l].simulated for l in rb.ls(fg)}
    for l in rb.ls(fg, rb.Pose2):  # forward propagation stays near the simulated truth
        m = rb.getVal(fg, l).mean(0)
        assert np.hypot(*(m[:2] - truth[l][:2])) < 1.5, (l, m, truth[l])
    assert np.hypot(*(rb.getVal(fg, "l1").mean(0) - [20, 0])) < 1.0
    dg = rb.DeviceGraph(fg)
    N = 100
    poses = np.stack([fg[f"x{i}"].val for i in range(7)])
    points = fg["l1"].val[None]
    flags = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS | rb.STATS
    out = dg.eval(rb.POSE2POSE2, flags, seed=5)
    ref = O.sweep_pose2pose2(np.arange(6), np.arange(1, 7), poses, out["meas"])
    assert np.abs(out["res"] - ref).max() < 1e-6
    assert np.allclose(out["meas"].mean(1), [10, 0, pi / 3], atol=0.05) and np.allclose(out["meas"].std(1), 0.1, atol=0.03)
    out = dg.eval(rb.PRIORPOSE2, flags, seed=5)
    assert np.abs(out["res"] - O.sweep_priorpose2([0], poses, out["meas"])).max() < 1e-6
    out = dg.eval(rb.BEARINGRANGE, flags, seed=5)
    ref = O.sweep_bearingrange([0, 6], [0, 0], poses, points, out["meas"])
    assert np.abs(out["res"] - ref).max() < 1e-6
    assert np.allclose(out["meas"].std(1), [[0.1, 1.0]] * 2, rtol=0.25)
    assert dg.ctx.launch_count >= 3


def test_manhattan500_fixture_parity(golden_dir):
    """Solved reference graph (examples/manhattan-batch-500-fg.tar.gz subset): real reference particles + factors."""
    z = np.load(os.path.join(golden_dir, "manhattan500_fixture.npz"))
    parts, ip, iq, mu, Sg = z["particles"], z["ip"], z["iq"], z["mu"], z["Sigma"]
    ctx = rb.Context(0)
    ctx.set_particles(rb.POSE2, parts)
    ctx.set_factors_pose2pose2(ip, iq, mu, Sg)
    rng = np.random.default_rng(0)
    meas = mu[:, None, :] + np.einsum("fij,fnj->fni", np.linalg.cholesky(Sg), rng.normal(size=(len(ip), 100, 3)))
    out = ctx.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS)
    ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS, meas=rb.meas_to_offsets(meas, mu), **out)
    res = rb.rows_to_particle_major(out["res"], 100)
    ref = O.sweep_pose2pose2(ip, iq, parts, meas)
    d = res - ref
    d[..., 2] = O.np_wrap(d[..., 2])
    rel = np.abs(d) / np.maximum(np.abs(ref), 1e-2)
    assert rel.max() < 1e-5, rel.max()
    assert np.abs(d).max() < 5e-7
    ctx.close()


def test_manhattan_g2o_full_graph(golden_dir):
    """config C3: examples/manhattan.g2o (3500 poses, 5453 EDGE_SE2) + the example's prior, N=100; particles from
    dead-reckoned means + fixture-like spread; fused-sample residuals + stats vs oracle."""
    z = np.load(os.path.join(golden_dir, "manhattan_g2o.npz"))
    fg = rb.graphFromEdgeArrays(z["ids"], z["mu"], z["info"])
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal(np.zeros(3), np.diag([0.1, 0.1, 0.05]) ** 2)))
    assert len(rb.ls(fg)) == 3500 and len(rb.lsf(fg, rb.Pose2Pose2)) == 5453
    # dead reckoning along the odometry chain (edges i -> i+1)
    odo = {(a, b): m for (a, b), m in zip(map(tuple, z["ids"]), z["mu"]) if b == a + 1}
    pose = np.zeros(3)
    fg["x0"].simulated = pose.copy()
    for i in range(3499):
        pose = O.pose2pose2_fwd(odo[(i, i + 1)], pose)
        fg[f"x{i+1}"].simulated = pose.copy()
    rb.seed_particles(fg, seed=1)
    dg = rb.DeviceGraph(fg)
    out = dg.eval(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS | rb.STATS, seed=11)
    facs = dg.by_family[rb.POSE2POSE2]
    ip = np.array([fg[f.variableOrderSymbols[0]].index for f in facs])
    iq = np.array([fg[f.variableOrderSymbols[1]].index for f in facs])
    poses = np.stack([v.val for v in dg.by_type[rb.POSE2]])
    ref = O.sweep_pose2pose2(ip, iq, poses, out["meas"])
    d = out["res"] - ref
    d[..., 2] = O.np_wrap(d[..., 2])
    odo_f = np.array([b == a + 1 for a, b in zip(ip, iq)])
    # odometry factors: residual magnitudes ~0.1-0.3; loop closures of the dead-reckoned graph can be metres off
    assert (np.abs(d[odo_f]) / np.maximum(np.abs(ref[odo_f]), 1e-2)).max() < 1e-5
    assert (np.abs(d) / np.maximum(np.abs(ref), 1e-2)).max() < 1e-5
    assert np.allclose(out["stats"][:, :3], out["res"].sum(1), rtol=1e-3, atol=1e-3)


def test_approxconv_and_samplefactor():
    """test/testBasicPose2Conv.jl:25-34 shape: prior at 0, Pose2Pose2 [10,0,pi/2?]; approxConv onto x1 lands at the
    measurement; backward conv returns to x0; sampleFactor moments (test/testBearingRange2D.jl:12-41)."""
    fg = rb.initfg()
    rb.addVariable(fg, "x0", rb.Pose2)
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal(np.zeros(3), 0.01 * np.eye(3))))
    rb.addVariable(fg, "x1", rb.Pose2)
    rb.addFactor(fg, ["x0", "x1"], rb.Pose2Pose2(rb.MvNormal([10, 0, pi / 2], 0.01 * np.eye(3))))
    rb.initAll(fg)
    pts = rb.approxConv(fg, "x0x1f1", "x1")
    assert pts.shape == (100, 3)
    assert np.allclose(pts.mean(0), [10, 0, pi / 2], atol=0.3)
    back = rb.approxConv(fg, "x0x1f1", "x0")
    assert np.allclose(back.mean(0), [0, 0, 0], atol=0.3)
    br = rb.Pose2Point2BearingRange(rb.Normal(0.3, 0.1), rb.Normal(20.0, 1.0))
    s = rb.sampleFactor(br, N=2000, seed=2)
    assert s.shape == (2000, 2)
    assert abs(s[:, 0].mean() - 0.3) < 0.01 and abs(s[:, 1].mean() - 20) < 0.1
    assert abs(s[:, 0].std() - 0.1) < 0.01 and abs(s[:, 1].std() - 1.0) < 0.08
    rb.addVariable(fg, "l1", rb.Point2)
    rb.addFactor(fg, ["x1", "l1"], br)
    with pytest.raises(NotImplementedError):
        rb.setVal(fg, "l1", np.zeros((100, 2)))
        rb.approxConv(fg, "x1l1f1", "x1")
    lm = rb.approxConv(fg, "x1l1f1", "l1")
    want = np.array([10, 0]) + 20 * np.array([math.cos(pi / 2 + 0.3), math.sin(pi / 2 + 0.3)])
    assert np.allclose(lm.mean(0), want, atol=1.0)


def test_beehive_and_pose3_chain_parity():
    """configs C4 / C5 at parity size: Beehive (Pose2 + BR landmarks, N=200) and SE(3) chain + loops (N=100)."""
    bh = rb.generateGraph_Beehive(60, N=200)
    rb.seed_particles(bh, N=200, seed=3)
    dg = rb.DeviceGraph(bh)
    fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
    poses = np.stack([v.val for v in dg.by_type[rb.POSE2]])
    points = np.stack([v.val for v in dg.by_type[rb.POINT2]])
    facs = dg.by_family[rb.BEARINGRANGE]
    ip = [bh[f.variableOrderSymbols[0]].index for f in facs]
    il = [bh[f.variableOrderSymbols[1]].index for f in facs]
    # default arithmetic (float32 per particle; ranges of 20 m +- 0.5 m): floor 0.1; Float64 chain (PRECISE): floor 0.01
    for flags, floor in ((fl, 1e-1), (fl | rb.PRECISE, 1e-2)):
        out = dg.eval(rb.BEARINGRANGE, flags, seed=1)
        ref = O.sweep_bearingrange(ip, il, poses, points, out["meas"])
        d = out["res"] - ref
        d[..., 0] = O.np_wrap(d[..., 0])
        assert (np.abs(d) / np.maximum(np.abs(ref), floor)).max() < 1e-5
    p3 = rb.generateGraph_Pose3Chain(400, loops=40)
    rb.seed_particles(p3, seed=4)
    dg3 = rb.DeviceGraph(p3)
    out = dg3.eval(rb.POSE3POSE3, fl, seed=1)
    facs = dg3.by_family[rb.POSE3POSE3]
    ip = [p3[f.variableOrderSymbols[0]].index for f in facs]
    iq = [p3[f.variableOrderSymbols[1]].index for f in facs]
    poses3 = np.stack([v.val for v in dg3.by_type[rb.POSE3]])
    ref = O.sweep_pose3pose3(ip, iq, poses3, out["meas"])
    assert (np.abs(out["res"] - ref) / np.maximum(np.abs(ref), 1e-2)).max() < 1e-5
    out = dg3.eval(rb.PRIORPOSE3, fl, seed=1)
    ref = O.sweep_priorpose3([0], poses3, out["meas"])
    assert (np.abs(out["res"] - ref) / np.maximum(np.abs(ref), 1e-2)).max() < 1e-5
