"""g2o reader / writer for the hot-path workloads.
Restates src/services/g2oParser.jl:39-49 (importG2o), :62-90 (VERTEX_SE2, VERTEX_SE3:QUAT), :91-122 (EDGE_SE2),
:123-168 (EDGE_SE3:QUAT), :176-186 (variable numbering), :188-290 (stringG2o!), :295-400 (exportG2o); the exact text
the reference's own test expects for the Hexagonal graph (test/testG2oParser.jl:26-51) is reproduced character for
character, including the signed zeros of LAPACK's potri-based inverse."""
from __future__ import annotations

import numpy as np

import re

from .factors import (MvNormal, Pose2, Pose2Point2BearingRange, Pose2Pose2, Pose3, Pose3Pose3, PriorPoint3)
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


def parseG2oInstruction(fg: FactorGraph, instruction: list, initialize: bool = True):
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
    elif kind == "VERTEX_SE2":  # g2oParser.jl:62-73: the value seeds the :parametric solve key
        l = "x" + instruction[1]
        if l not in fg.variables:
            addVariable(fg, l, Pose2)
        if initialize:
            fg.variables[l].parametric = np.array([float(v) for v in instruction[2:5]])
    elif kind == "VERTEX_SE3:QUAT":  # g2oParser.jl:74-90 (file order qx qy qz qw)
        l = "x" + instruction[1]
        if l not in fg.variables:
            addVariable(fg, l, Pose3)
        if initialize:
            t = [float(v) for v in instruction[2:5]]
            qx, qy, qz, qw = (float(v) for v in instruction[5:9])
            fg.variables[l].parametric = np.concatenate([t, _quat_to_rotvec(qw, qx, qy, qz)])
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


# ---- export (g2oParser.jl:176-400) ---------------------------------------------------------------------------------
def _invcov(Sigma):
    """Distributions.invcov of a full MvNormal = inverse through the upper Cholesky factor (LAPACK potrf/potri); kept
    identical so that the exported text (including signed zeros) matches the reference's"""
    from scipy.linalg import lapack
    c, info = lapack.dpotrf(np.asarray(Sigma, dtype=np.float64), lower=0)
    inv, info2 = lapack.dpotri(c, lower=0)
    iu = np.triu_indices_from(inv, 1)
    inv[(iu[1], iu[0])] = inv[iu]  # potri fills the upper triangle only; mirror it without touching signed zeros
    inv[np.isinf(inv)] = 0.0
    return inv


def _jl(x) -> str:
    """Julia's string interpolation of a Float64 (shortest round-trip representation, like Python's repr)"""
    return repr(float(x))


def _rotvec_to_quat(w):
    th = float(np.linalg.norm(w))
    if th < 1e-12:
        return 1.0, 0.5 * w[0], 0.5 * w[1], 0.5 * w[2]
    k = np.sin(0.5 * th) / th
    return float(np.cos(0.5 * th)), float(k * w[0]), float(k * w[1]), float(k * w[2])


def _natural_key(label):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", label)]


def stringG2o(fnc, varlist) -> str:
    """one line per factor (g2oParser.jl:188-273)"""
    if isinstance(fnc, Pose2Pose2):
        I, m = _invcov(fnc.Z.Sigma), fnc.Z.mu
        return (f"EDGE_SE2 {varlist[0]} {varlist[1]} {_jl(m[0])} {_jl(m[1])} {_jl(m[2])} "
                f"{_jl(I[0, 0])} {_jl(I[0, 1])} {_jl(I[0, 2])} {_jl(I[1, 1])} {_jl(I[1, 2])} {_jl(I[2, 2])}")
    if isinstance(fnc, Pose2Point2BearingRange):
        return (f"LANDMARK {varlist[0]} {varlist[1]} {_jl(fnc.bearing.mu)} {_jl(fnc.range.mu)} "
                f"{_jl(1 / fnc.bearing.sigma ** 2)} {_jl(0.0)} {_jl(1 / fnc.range.sigma ** 2)}")
    if isinstance(fnc, Pose3Pose3):
        I, m = _invcov(fnc.Z.Sigma), fnc.Z.mu
        qw, qx, qy, qz = _rotvec_to_quat(m[3:])
        up = " ".join(_jl(I[r, c]) for r in range(6) for c in range(r, 6))
        return (f"EDGE_SE3:QUAT {varlist[0]} {varlist[1]} {_jl(m[0])} {_jl(m[1])} {_jl(m[2])} "
                f"{_jl(qx)} {_jl(qy)} {_jl(qz)} {_jl(qw)} {up}")
    if isinstance(fnc, PriorPoint3):  # const PriorPose3XYZ = PriorPoint3 (g2oParser.jl:188)
        I, m = _invcov(fnc.Z.Sigma), fnc.Z.mu
        return (f"EDGE_SE3_XYZ_PRIOR {varlist[0]} {_jl(m[0])} {_jl(m[1])} {_jl(m[2])} "
                f"{_jl(I[0, 0])} {_jl(I[0, 1])} {_jl(I[0, 2])} {_jl(I[1, 1])} {_jl(I[1, 2])} {_jl(I[2, 2])}")
    raise TypeError(f"unknown factor type {type(fnc).__name__}")  # g2oParser.jl:275-283


def exportG2o(fg: FactorGraph, poseRegex: str = r"x\d", ignorePriors: bool = True, filename: str = "/tmp/test.txt",
              solveKey=None, varIntLabel: dict | None = None) -> str:
    """exportG2o(dfg; poseRegex, ignorePriors, filename, varIntLabel, solveKey) (g2oParser.jl:373-400): factors in pose
    order, every factor once, variables numbered in order of first appearance (landmarks interleave with poses exactly
    as in the reference), continuing after the numbers `varIntLabel` ({label: int}, ordered) already assigns.  With a
    solveKey the VERTEX_* lines come first: for the entries of `varIntLabel` like the reference (`_writeG2oVertexes`
    :323-339 -- it writes none when the mapping is empty), or, when no mapping is passed, for every numbered variable.
    The estimate is the variable's PPE `suggested` for that key (`setPPE`), else the parametric solution
    ("parametric"), else the particle mean."""
    pat = re.compile(poseRegex)
    poses = sorted((l for l in fg.variables if pat.search(l)), key=_natural_key)  # occursin, like ls(dfg, regex)
    remaining = list(fg.factors)  # insertion order, like DFG's neighbour lists
    ids, lines = {str(k): int(v) for k, v in (varIntLabel or {}).items()}, []
    vertex_labels = list(ids) if varIntLabel is not None else None
    nxt = max(ids.values(), default=-1) + 1  # uniqVarInt: one past the largest number in use
    for vs in poses:
        for fl in [f for f in remaining if vs in fg.factors[f].variableOrderSymbols]:
            f = fg.factors[fl]
            remaining.remove(fl)
            if ignorePriors and f.fnc.is_prior:
                continue
            varlist = []
            for l in f.variableOrderSymbols:
                if l not in ids:
                    ids[l] = nxt
                    nxt += 1
                varlist.append(ids[l])
            lines.append(stringG2o(f.fnc, varlist))
    head = []
    if solveKey is not None:
        for l in (ids if vertex_labels is None else vertex_labels):
            v, i = fg.variables[l], ids[l]
            x = getattr(v, "ppes", {}).get(solveKey, {}).get("suggested")
            if x is None and solveKey == "parametric":
                x = getattr(v, "parametric", None)
            if x is None:
                if v.val is None:
                    raise ValueError(f"variable {l} has no estimate for solve key {solveKey}")
                x = v.val.mean(0)
                if v.variableType is Pose2:
                    x[2] = np.arctan2(np.sin(v.val[:, 2]).mean(), np.cos(v.val[:, 2]).mean())
            if v.variableType is Pose2:
                head.append(f"VERTEX_SE2 {i} {_jl(x[0])} {_jl(x[1])} {_jl(x[2])}")
            elif v.variableType is Pose3:
                qw, qx, qy, qz = _rotvec_to_quat(np.asarray(x[3:]))
                head.append(f"VERTEX_SE3:QUAT {i} {_jl(x[0])} {_jl(x[1])} {_jl(x[2])} {_jl(qx)} {_jl(qy)} {_jl(qz)} {_jl(qw)}")
            # other variable types have no vertex line in the reference either (g2oParser.jl:285-318)
    with open(filename, "w") as fh:
        for ln in head + lines:
            fh.write(ln + "\n")
    return filename
