"""g2o reader for the hot-path workloads: EDGE_SE2 -> Pose2Pose2, EDGE_SE3:QUAT -> Pose3Pose3.
Restates src/services/g2oParser.jl:39-49 (importG2o), :91-122 (EDGE_SE2), :123-168 (EDGE_SE3:QUAT);
export and VERTEX_* handling are out of scope."""
from __future__ import annotations

import numpy as np

from .factors import MvNormal, Pose2, Pose2Pose2, Pose3, Pose3Pose3
from .graph import FactorGraph, addFactor, addVariable, initfg


def importG2o(input_file: str):
    """every line split on whitespace (g2oParser.jl:39-49)"""
    with open(input_file) as fh:
        return [ln.split() for ln in fh if ln.strip()]


def _sym_inv(info):
    cov = np.linalg.inv(info)
    return 0.5 * (cov + cov.T)  # g2oParser.jl:106-109 "workaround to ensure cov_mat is Hermitian"


def _quat_to_rotvec(qw, qx, qy, qz):
    n = np.sqrt(qw * qw + qx * qx + qy * qy + qz * qz)
    qw, qx, qy, qz = qw / n, qx / n, qy / n, qz / n
    if qw < 0:
        qw, qx, qy, qz = -qw, -qx, -qy, -qz
    s = np.sqrt(qx * qx + qy * qy + qz * qz)
    k = 2.0 if s < 1e-12 else 2.0 * np.arctan2(s, qw) / s
    return np.array([qx, qy, qz]) * k


def parseG2oInstruction(fg: FactorGraph, instruction: list):
    """parseG2oInstruction!(fg, pieces).  Variables are labelled Symbol("x", id)."""
    kind = instruction[0]
    if kind == "EDGE_SE2":
        a, b = "x" + instruction[1], "x" + instruction[2]
        mu = np.array([float(v) for v in instruction[3:6]])
        i = [float(v) for v in instruction[6:12]]
        info = np.array([[i[0], i[1], i[2]], [i[1], i[3], i[4]], [i[2], i[4], i[5]]])  # g2oParser.jl:103-105
        for l in (a, b):
            if l not in fg.variables:
                addVariable(fg, l, Pose2)
        addFactor(fg, [a, b], Pose2Pose2(MvNormal(mu, _sym_inv(info))))
    elif kind == "EDGE_SE3:QUAT":
        a, b = "x" + instruction[1], "x" + instruction[2]
        t = np.array([float(v) for v in instruction[3:6]])
        qx, qy, qz, qw = (float(v) for v in instruction[6:10])  # file order x y z w, reordered at :136
        mu = np.concatenate([t, _quat_to_rotvec(qw, qx, qy, qz)])
        u = [float(v) for v in instruction[10:31]]
        info = np.zeros((6, 6))
        k = 0
        for r in range(6):
            for c in range(r, 6):
                info[r, c] = info[c, r] = u[k]
                k += 1
        for l in (a, b):
            if l not in fg.variables:
                addVariable(fg, l, Pose3)
        addFactor(fg, [a, b], Pose3Pose3(MvNormal(mu, _sym_inv(info))))
    # VERTEX_* lines only carry initial values in the reference (parametric init); ignored here
    return fg


def loadG2o(input_file: str, fg: FactorGraph | None = None) -> FactorGraph:
    fg = fg or initfg()
    for ins in importG2o(input_file):
        parseG2oInstruction(fg, ins)
    return fg


def graphFromEdgeArrays(ids, mu, info_upper, fg: FactorGraph | None = None) -> FactorGraph:
    """Same as loadG2o for already-tokenised EDGE_SE2 records (tests/golden/manhattan_g2o.npz)."""
    fg = fg or initfg()
    for (a, b), m, i in zip(ids, mu, info_upper):
        parseG2oInstruction(fg, ["EDGE_SE2", str(a), str(b)] + [repr(float(v)) for v in m] + [repr(float(v)) for v in i])
    return fg
