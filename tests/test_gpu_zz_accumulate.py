"""GPU test of the parametric chain helpers (IIF.accumulateFactorMeans / solveFactorParametric) through the
closed-form proposal kernels; known answer of the reference's test/testAccumulateFactors.jl:19-30."""
import math

import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_accumulate_factor_means_reference_case():
    fg = rb.initfg()
    rb.addVariable(fg, "x0", rb.Pose2)
    rb.addFactor(fg, ["x0"], rb.PriorPose2(rb.MvNormal(np.zeros(3), 0.001 * np.eye(3))))
    rb.addVariable(fg, "x1", rb.Pose2)
    rb.addFactor(fg, ["x0", "x1"], rb.Pose2Pose2(rb.MvNormal([10, 0, 0.0], 0.001 * np.eye(3))), graphinit=False)
    val = rb.accumulateFactorMeans(fg, ["x0f1", "x0x1f1"])
    assert np.allclose(val, [10, 0, 0], atol=1e-3)  # the reference's assertion
    assert np.allclose(val, [10, 0, 0], atol=1e-6)


def test_accumulate_along_hexagon_and_back():
    """prior + six odometry means walk the hexagon back to the start (the truth of generateGraph_Hexagonal); solving
    the chain backwards from x6 returns to x0; an SE(3) chain agrees with the oracle's closed form."""
    fg = rb.generateGraph_Hexagonal(graphinit=False)
    chain = ["x0f1"] + [f"x{i}x{i + 1}f1" for i in range(6)]
    for k in range(1, 7):
        val = rb.accumulateFactorMeans(fg, chain[:k + 1])
        truth = fg[f"x{k}"].simulated
        # every link returns float32 offsets from the (zero) anchor of its target: <= 1e-6 m / 1.2e-7 rad of rounding per link
        assert np.allclose(val[:2], truth[:2], atol=5e-5)
        assert abs(math.remainder(val[2] - truth[2], 2 * math.pi)) < 5e-6
    x = fg["x6"].simulated.copy()
    for i in range(5, -1, -1):
        x = rb.solveFactorParametric(fg, f"x{i}x{i + 1}f1", (f"x{i + 1}", x), f"x{i}")
    assert np.allclose(x[:2], 0, atol=5e-5) and abs(math.remainder(x[2], 2 * math.pi)) < 5e-6
    g3 = rb.generateGraph_Pose3Chain(6, loops=0)
    labels = rb.lsf(g3, rb.Pose3Pose3)
    val = rb.accumulateFactorMeans(g3, rb.lsf(g3, rb.PriorPose3) + labels)
    ref = np.asarray(g3[rb.lsf(g3, rb.PriorPose3)[0]].fnc.Z.mu, dtype=np.float64)
    for l in labels:
        ref = O.pose3pose3_fwd(g3[l].fnc.Z.mu, ref)
    assert np.allclose(val[:3], ref[:3], atol=5e-5)
    assert np.linalg.norm(O.so3_log(O.so3_exp(val[3:]).T @ O.so3_exp(ref[3:]))) < 5e-6
    with pytest.raises(NotImplementedError):  # one equation for two unknowns: no closed-form solve
        gr = rb.initfg()
        rb.addVariable(gr, "x0", rb.Pose2)
        rb.addVariable(gr, "l1", rb.Point2)
        rb.addFactor(gr, ["x0", "l1"], rb.Pose2Point2Range(rb.Normal(5.0, 0.1)), graphinit=False)
        rb.solveFactorParametric(gr, "x0l1f1", ("x0", np.zeros(3)), "l1")


@pytest.mark.parametrize("N", [1, 5, 17])
def test_statistics_with_fewer_particles_than_lanes(N, nF=4000):
    """Npad < 32: the dead lanes of the slot group must read a particle that exists -- whatever lies behind the block in
    shared memory (possibly a NaN bit pattern) would otherwise enter the zero-masked statistics as NaN * 0.  Thousands of
    factors so that, were the defect back, some factor would meet such a pattern (found by tests/test_emulated_kernels.py)."""
    rng = np.random.default_rng(N)
    ctx = rb.default_context()
    nv = 300
    poses = rng.normal(size=(nv, 1, 3)) * [30, 30, 1] + rng.normal(size=(nv, N, 3)) * [0.1, 0.1, 0.02]
    p3 = rng.normal(size=(nv, 1, 6)) * [5, 5, 5, 0.5, 0.5, 0.5] + rng.normal(size=(nv, N, 6)) * 0.05
    ip = rng.integers(0, nv, nF).astype(np.int32)
    iq = rng.integers(0, nv, nF).astype(np.int32)
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_particles(rb.POSE3, p3)
    ctx.set_factors_pose2pose2(ip, iq, rng.normal(size=(nF, 3)), np.tile(0.01 * np.eye(3), (nF, 1, 1)))
    ctx.set_factors_pose3pose3(ip, iq, rng.normal(size=(nF, 6)) * 0.3, np.tile(0.01 * np.eye(6), (nF, 1, 1)))
    for fam, dr in ((rb.POSE2POSE2, 3), (rb.POSE3POSE3, 6)):
        for flags in (rb.RESIDUAL | rb.STATS | rb.SAMPLE, rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD | rb.SAMPLE):
            out = ctx.alloc_host_outputs(fam, flags)
            ctx.eval_host(fam, flags, seed=3, **out)
            res = out["res"][:, :N].astype(np.float64)
            assert np.isfinite(out["stats"]).all() and np.isfinite(res).all(), (fam, flags)
            assert np.allclose(out["stats"][:, :dr], res.sum(1), rtol=1e-4, atol=1e-4), (fam, flags)
