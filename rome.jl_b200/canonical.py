"""Canonical graph generators of the hot-path workloads (host side, restated -- not ported).

generateGraph_Hexagonal / generateGraph_Circle   src/canonical/GenerateHexagonal.jl:27-42, GenerateCircular.jl:31-94
generateGraph_ZeroPose                           src/canonical/GenerateCommon.jl:70-102
generateGraph_Beehive (honeycomb walk + BR)      src/canonical/GenerateBeehive.jl:20-72, GenerateHoneycomb.jl:59-170
generateGraph_ManhattanShaped                    synthetic 10k-pose SE(2) grid walk (SURVEY.md 8d "T")
generateGraph_Pose3Chain                         synthetic SE(3) helix chain + loop closures (SURVEY.md 8d "C5")
Each variable also receives a `simulated` ground-truth coordinate vector (the reference stores it as the
:simulated PPE, GenerateCommon.jl:36-50), used here to seed particles without running a solver.
"""
from __future__ import annotations

import math
import re

import numpy as np

from .factors import (MvNormal, Normal, Point2, Pose2, Pose2Point2BearingRange, Pose2Pose2, Pose3, Pose3Pose3,
                      PriorPose2, PriorPose3)
from .graph import FactorGraph, SolverParams, addFactor, addVariable, initfg


# ---- small Float64 SE(2)/SE(3) helpers for ground truth (graph construction only) ----------------------
def _se2_compose(p, m):
    c, s = math.cos(p[2]), math.sin(p[2])
    th = p[2] + m[2]
    return np.array([p[0] + c * m[0] - s * m[1], p[1] + s * m[0] + c * m[1], math.atan2(math.sin(th), math.cos(th))])


def _so3_exp(w):
    t = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])
    if t < 1e-8:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + math.sin(t) / t * K + (1 - math.cos(t)) / t ** 2 * K @ K


def _so3_log(R):
    c = min(1.0, max(-1.0, 0.5 * (np.trace(R) - 1)))
    t = math.acos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return 0.5 * v if t < 1e-8 else t / (2 * math.sin(t)) * v


def _truth(fg: FactorGraph, label) -> np.ndarray:
    return fg.variables[label].simulated


def seed_particles(fg: FactorGraph, sigma_pose2=(0.1, 0.12, 0.02), sigma_point2=(0.3, 0.3),
                   sigma_pose3=(0.1, 0.1, 0.1, 0.02, 0.02, 0.02), seed=0, N=None):
    """particles = simulated truth + Gaussian spread (fixture median std for Pose2)"""
    rng = np.random.default_rng(seed)
    N = N or fg.solverParams.N
    sig = {Pose2: sigma_pose2, Point2: sigma_point2, Pose3: sigma_pose3}
    for v in fg.variables.values():
        v.val = v.simulated[None, :] + rng.normal(size=(N, v.variableType.dim)) * np.asarray(sig[v.variableType])
    return fg


# ---- reference generators -----------------------------------------------------------------------------
def generateGraph_ZeroPose(fg: FactorGraph | None = None, varType=Pose2, mu0=None, cov=None, graphinit=True):
    """:x0 + prior; defaults follow GenerateCommon.jl:70-102 (Sigma0 = 0.01*I)."""
    fg = fg or initfg()
    d = varType.dim
    mu0 = np.zeros(d) if mu0 is None else np.asarray(mu0, dtype=np.float64)
    cov = 0.01 * np.eye(d) if cov is None else np.asarray(cov)
    v = addVariable(fg, "x0", varType)
    v.simulated = mu0.copy()
    prior = PriorPose2 if varType is Pose2 else PriorPose3
    addFactor(fg, ["x0"], prior(MvNormal(mu0, cov)), graphinit=graphinit)
    return fg


def generateGraph_Circle(poses=6, fg: FactorGraph | None = None, graphinit=True, landmark=True, loopClosure=True,
                         biasTurn=0.0, kappaOdo=1.0, cyclePoses=None, offsetPoses=None, stopEarly=9999999):
    """GenerateCircular.jl:31-94 with the same parameters: prior 0.01*I on :x0, `poses` odometry legs
    [10, 0, 2pi/cyclePoses + biasTurn] with sigma kappaOdo*0.1, landmark :l1 sighted (0, 20) from :x0 and :x<poses>."""
    fg = fg or initfg()
    cyclePoses = cyclePoses or poses
    if offsetPoses is None:  # continue an existing graph behind its last pose (GenerateCircular.jl:33)
        offsetPoses = max(sum(1 for l in fg.variables if re.search(r"x\d", l)) - 1, 0)
    if not offsetPoses < poses:
        raise ValueError("`offsetPoses` must be smaller than total number of `poses`")
    if "x0" not in fg.variables:
        v = addVariable(fg, "x0", Pose2)
        v.simulated = np.zeros(3)
        addFactor(fg, ["x0"], PriorPose2(MvNormal(np.zeros(3), 0.01 * np.eye(3))), graphinit=graphinit)
    X = np.array([10.0, 0.0, 2 * math.pi / cyclePoses + biasTurn])
    for i in range(offsetPoses, poses):
        if stopEarly <= i:
            break
        p, n = f"x{i}", f"x{i + 1}"
        v = addVariable(fg, n, Pose2)
        v.simulated = _se2_compose(_truth(fg, p), X)
        addFactor(fg, [p, n], Pose2Pose2(MvNormal(X, np.diag((kappaOdo * np.array([0.1, 0.1, 0.1])) ** 2))),
                  graphinit=graphinit)
    if not landmark:
        return fg
    if "l1" not in fg.variables:
        l = addVariable(fg, "l1", Point2, tags=["LANDMARK"])
        l.simulated = np.array([20.0, 0.0])
        addFactor(fg, ["x0", "l1"], Pose2Point2BearingRange(Normal(0, 0.1), Normal(20.0, 1.0)), graphinit=graphinit)
    if loopClosure and f"x{poses}" in fg.variables:
        addFactor(fg, [f"x{poses}", "l1"], Pose2Point2BearingRange(Normal(0, 0.1), Normal(20.0, 1.0)),
                  graphinit=graphinit)
    return fg


def generateGraph_Hexagonal(fg: FactorGraph | None = None, landmark=True, loopClosure=None, N=100, graphinit=True):
    """GenerateHexagonal.jl:27-42: 7 Pose2, 1 Point2, 1 PriorPose2 + 6 Pose2Pose2 + 2 Pose2Point2BearingRange."""
    fg = fg or initfg()
    fg.solverParams.N = N
    return generateGraph_Circle(6, fg=fg, graphinit=graphinit, landmark=landmark,
                                loopClosure=landmark if loopClosure is None else loopClosure)


class _LandmarkSighter:
    """_addLandmarkBeehive! (GenerateHoneycomb.jl:59-100): every pose sights the lattice landmark 20 m ahead through
    Pose2Point2BearingRange(Normal(0, 0.03), Normal(20, 0.5)); a landmark within `atol` of an existing one is that one
    (the reference's `_checkVariableByReference` data association), otherwise a new Point2 is added."""

    def __init__(self, fg, atol, label, association=None):
        self.fg, self.atol, self.label, self.association = fg, atol, label, association
        self.landmarks: list[tuple[str, np.ndarray]] = []
        self.cells: dict[tuple[int, int], list[int]] = {}

    def __call__(self, pose_label):
        fg = self.fg
        p = _truth(fg, pose_label)
        pos = p[:2] + 20.0 * np.array([math.cos(p[2]), math.sin(p[2])])
        key = (int(math.floor(pos[0] / 4.0)), int(math.floor(pos[1] / 4.0)))
        found = None
        if self.association is not None:
            # `_doHoneycomb` (:73-83): the table decides, not the geometry
            forced = self.association.get(self.label(pose_label, len(self.landmarks)))
            found = None if forced is None else next(k for k, (lab, _) in enumerate(self.landmarks) if lab == forced)
        else:
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for k in self.cells.get((key[0] + dx, key[1] + dy), []):
                        if np.linalg.norm(self.landmarks[k][1] - pos) < self.atol:
                            found = k
        if found is None:
            lab = self.label(pose_label, len(self.landmarks))
            v = addVariable(fg, lab, Point2, tags=["LANDMARK"])
            v.simulated = pos
            self.landmarks.append((lab, pos))
            self.cells.setdefault(key, []).append(len(self.landmarks) - 1)
            found = len(self.landmarks) - 1
        addFactor(fg, [pose_label, self.landmarks[found][0]], Pose2Point2BearingRange(Normal(0, 0.03), Normal(20, 0.5)),
                  graphinit=False)
        return self.landmarks[found][0]


def generateGraph_Honeycomb(poseCountTarget=36, fg: FactorGraph | None = None, graphinit=False, direction="right",
                            addLandmarks=True, atol=1.0, N=100, extraLeftLegs=(41, 63, 78), association=None):
    """generateGraph_Honeycomb! (GenerateHoneycomb.jl:179-231): the predetermined honeycomb -- repeat { six hexagon legs
    [10, 0, +pi/3] (`_driveHex!` :103-132), an extra left leg after poses 41, 63, 78 (the pose entries of
    `_honeycombRecipe` :47-49), one offset leg [10, 0, -+pi/3] in `direction` (`_offsetHexLeg` :134-170) } until
    poseCountTarget; every new pose sights the landmark 20 m ahead, named l<pose number> when new.  The reference forces
    its landmark data association from a hard-coded table (`_honeycombRecipe` :3-46, `_doHoneycomb` :73-83): pass that
    table as `association` ({"l6": "l0", ...}; tests/golden/known_answers.json holds it) to rebuild the reference's graph
    label for label.  By default the association follows from the geometry (same landmark = within `atol`), which agrees
    with the table wherever the table is consistent with the simulated poses (tests/test_host_logic.py)."""
    fg = fg or initfg(SolverParams(N=N, graphinit=graphinit))
    generateGraph_ZeroPose(fg, Pose2, graphinit=graphinit)
    sight = _LandmarkSighter(fg, atol, label=lambda pose_label, count: "l" + pose_label[1:], association=association)
    if addLandmarks:
        sight("x0")
    state = {"n": 0}

    def leg(turn):
        if poseCountTarget <= state["n"]:
            return
        i = state["n"]
        X = np.array([10.0, 0.0, turn])
        v = addVariable(fg, f"x{i + 1}", Pose2)
        v.simulated = _se2_compose(_truth(fg, f"x{i}"), X)
        addFactor(fg, [f"x{i}", f"x{i + 1}"], Pose2Pose2(MvNormal(X, np.diag([0.1, 0.1, 0.1]) ** 2)), graphinit=graphinit)
        state["n"] = i + 1
        if addLandmarks:
            sight(f"x{i + 1}")

    off = {"right": -math.pi / 3, "left": math.pi / 3}
    while state["n"] < poseCountTarget:
        for _ in range(6):
            leg(math.pi / 3)
        if state["n"] in extraLeftLegs:
            leg(off["left"])
        leg(off[direction])
    return fg


def generateGraph_Beehive(poseCountTarget=10, fg: FactorGraph | None = None, graphinit=True, addLandmarks=True,
                          yaw0=None, locality=1.0, atol=1.0, seed=3, N=200):
    """Honeycomb walk of GenerateBeehive.jl:20-72: legs [10, 0, +-pi/3] (GenerateHoneycomb.jl:134-170), direction
    kept/flipped with probability 1/(1+locality); every pose sights the lattice landmark 20 m ahead through
    Pose2Point2BearingRange(Normal(0, 0.03), Normal(20, 0.5)), landmarks within `atol` are the same one
    (GenerateHoneycomb.jl:59-100)."""
    rng = np.random.default_rng(seed)
    fg = fg or initfg(SolverParams(N=N, graphinit=graphinit))
    yaw0 = [0.0, -2 * math.pi / 3, 2 * math.pi / 3][rng.integers(0, 3)] if yaw0 is None else yaw0
    generateGraph_ZeroPose(fg, Pose2, mu0=[0.0, 0.0, yaw0], graphinit=graphinit)
    sight = _LandmarkSighter(fg, atol, label=lambda pose_label, count: f"l{count + 1}")

    if addLandmarks:
        sight("x0")
    direction = 1 if rng.integers(0, 2) == 0 else -1
    flip = 1.0 / (1.0 + locality)
    for i in range(poseCountTarget):
        if rng.random() < flip:
            direction = -direction
        X = np.array([10.0, 0.0, direction * math.pi / 3])
        v = addVariable(fg, f"x{i + 1}", Pose2)
        v.simulated = _se2_compose(_truth(fg, f"x{i}"), X)
        addFactor(fg, [f"x{i}", f"x{i + 1}"], Pose2Pose2(MvNormal(X, np.diag([0.1, 0.1, 0.1]) ** 2)), graphinit=graphinit)
        if addLandmarks:
            sight(f"x{i + 1}")
    return fg


# ---- synthetic throughput workloads ----------------------------------------------------------
def generateGraph_ManhattanShaped(poses=10000, loop_fraction=0.2, seed=2, N=100, pose_offset=0,
                                  fg: FactorGraph | None = None):
    """Manhattan-world style SE(2) graph: unit-step grid walk (turn 0 / +-pi/2 with probability .7/.15/.15),
    poses-1 odometry factors + about loop_fraction*poses loop closures between poses that revisit the same
    neighbourhood (<= 2 m apart, >= 20 steps apart), information like examples/manhattan.g2o's median
    diag(44, 380, 9700) -> Sigma = diag(1/44, 1/380, 1/9700), and the prior of
    examples/ManhattanDatasetBatch.jl:30-32 (sigma 0.1, 0.1, 0.05)."""
    rng = np.random.default_rng(seed)
    fg = fg or initfg(SolverParams(N=N))
    Sigma = np.diag([1 / 44.0, 1 / 380.0, 1 / 9700.0])
    L = np.linalg.cholesky(Sigma)
    lab = lambda i: f"x{pose_offset + i}"
    truth = np.zeros((poses, 3))
    turns = rng.choice([0.0, math.pi / 2, -math.pi / 2], size=poses - 1, p=[0.7, 0.15, 0.15])
    v0 = addVariable(fg, lab(0), Pose2)
    v0.simulated = truth[0].copy()
    addFactor(fg, [lab(0)], PriorPose2(MvNormal(np.zeros(3), np.diag([0.1, 0.1, 0.05]) ** 2)))
    for i in range(poses - 1):
        X = np.array([1.0, 0.0, turns[i]])
        truth[i + 1] = _se2_compose(truth[i], X)
        v = addVariable(fg, lab(i + 1), Pose2)
        v.simulated = truth[i + 1].copy()
        addFactor(fg, [lab(i), lab(i + 1)], Pose2Pose2(MvNormal(X + L @ rng.normal(size=3), Sigma)))
    # loop closures by spatial hashing of the (integer) grid positions
    cells: dict[tuple[int, int], list[int]] = {}
    want = int(loop_fraction * poses)
    cand = []
    for i in range(poses):
        key = (int(round(truth[i, 0])), int(round(truth[i, 1])))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for j in cells.get((key[0] + dx, key[1] + dy), []):
                    if i - j >= 20 and np.hypot(*(truth[i, :2] - truth[j, :2])) <= 2.0:
                        cand.append((j, i))
        cells.setdefault(key, []).append(i)
    if len(cand) > want:
        cand = [cand[k] for k in sorted(rng.choice(len(cand), want, replace=False))]
    for j, i in cand:
        c, s = math.cos(truth[j, 2]), math.sin(truth[j, 2])
        d = truth[i, :2] - truth[j, :2]
        rel = np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1],
                        math.atan2(math.sin(truth[i, 2] - truth[j, 2]), math.cos(truth[i, 2] - truth[j, 2]))])
        addFactor(fg, [lab(j), lab(i)], Pose2Pose2(MvNormal(rel + L @ rng.normal(size=3), Sigma)))
    return fg


def generateGraph_Pose3Chain(poses=10000, loops=1000, seed=4, N=100, fg: FactorGraph | None = None):
    """SE(3) helix chain: odometry mu = [1,0,0, 0,0.02,0.05], Sigma = diag([0.01*1_3; 1e-4*1_3]) (the default of
    src/factors/Pose3Pose3.jl:10), `loops` closures between poses >= 50 apart and <= 3 m apart, PriorPose3 on x0."""
    rng = np.random.default_rng(seed)
    fg = fg or initfg(SolverParams(N=N))
    Sigma = np.diag([0.01] * 3 + [1e-4] * 3)
    sd = np.sqrt(np.diag(Sigma))
    mu = np.array([1.0, 0, 0, 0, 0.02, 0.05])
    t, R = np.zeros(3), np.eye(3)
    Ts, Rs = [t.copy()], [R.copy()]
    v = addVariable(fg, "x0", Pose3)
    v.simulated = np.zeros(6)
    addFactor(fg, ["x0"], PriorPose3(MvNormal(np.zeros(6), Sigma)))
    M = _so3_exp(mu[3:])
    for i in range(poses - 1):
        t = t + R @ mu[:3]
        R = R @ M
        Ts.append(t.copy()); Rs.append(R.copy())
        v = addVariable(fg, f"x{i + 1}", Pose3)
        v.simulated = np.concatenate([t, _so3_log(R)])
        addFactor(fg, [f"x{i}", f"x{i + 1}"], Pose3Pose3(MvNormal(mu + sd * rng.normal(size=6), Sigma)))
    T = np.array(Ts)
    cells: dict[tuple, list[int]] = {}
    cand = []
    for i in range(poses):
        key = tuple(np.floor(T[i] / 3.0).astype(int))
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for j in cells.get((key[0] + dx, key[1] + dy, key[2] + dz), []):
                        if i - j >= 50 and np.linalg.norm(T[i] - T[j]) <= 3.0:
                            cand.append((j, i))
        cells.setdefault(key, []).append(i)
    if len(cand) > loops:
        cand = [cand[k] for k in sorted(rng.choice(len(cand), loops, replace=False))]
    for j, i in cand:
        rel = np.concatenate([Rs[j].T @ (T[i] - T[j]), _so3_log(Rs[j].T @ Rs[i])])
        addFactor(fg, [f"x{j}", f"x{i}"], Pose3Pose3(MvNormal(rel + sd * rng.normal(size=6), Sigma)))
    return fg
