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
