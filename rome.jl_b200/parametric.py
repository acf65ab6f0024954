"""Parametric (Gaussian, single-point) solve on top of the batched residual kernels.

The reference's `IIF.solveGraphParametric!` minimises  sum_f r_f(x)' Omega_f r_f(x)  with r_f the SAME factor functor
evaluated at the belief mean (measurement = the factor mean, `getMeasurementParametric`, e.g.
src/factors/BearingRange2D.jl:30-37) and differentiates the functor numerically / by ForwardDiff [IIF-knowledge].
For graphs made of the five hot families (Pose2Pose2, PriorPose2, Pose2Point2BearingRange, Pose3Pose3, PriorPose3) the
Jacobian blocks are the kernels' ANALYTIC ones (ROME_B200_JACOBIAN): one particle per
variable, one launch per family per iteration; Pose3 increments are right perturbations (t + dt, R Exp(delta)).
Other families fall back to finite differences, where every residual AND every finite-difference column comes from
ONE launch per factor family per iteration:
the variables are coloured so that no two variables sharing a factor have the same colour, and the "particles" of a
variable are  [mean, mean +- h e_i for every coordinate i in its colour's slot]  -- particle n of all variables
forms one perturbed copy of the graph, so the residual rows of a family evaluated over the N = 1 + 2*D*C particles
hold r and the central differences for both variables of every factor.  The host then assembles the sparse
Levenberg-Marquardt system (SciPy) -- no residual arithmetic happens on the CPU.
Pinned by the reference's deterministic parametric tests (test/testParametric.jl:16-57,155-181,
test/testParametricCovariances.jl:41-52, test/testPose3.jl:27-56) in tests/test_gpu_parametric.py.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import FAMILY, VAR_DIM, Context, npad
from .factors import getMeasurementParametric
from .graph import DeviceGraph, FactorGraph

# residual components that are angles (finite differences are taken modulo 2 pi)
_ANGLE_COMPONENTS = {L.POSE2POSE2: (2,), L.PRIORPOSE2: (2,), L.BEARINGRANGE: (0,), L.POSE2POINT2BEARING: (0,),
                     L.POSE3POSE3XYYAW: (2,)}
_WRAP_COORD = {L.POSE2: 2}


def color_variables(fg: FactorGraph) -> dict:
    """greedy colouring of the variable adjacency graph (two variables are adjacent when a factor joins them)"""
    adj = {l: set() for l in fg.variables}
    for f in fg.factors.values():
        ls = f.variableOrderSymbols
        for a in ls:
            adj[a].update(b for b in ls if b != a)
    color = {}
    for l in sorted(adj, key=lambda l: -len(adj[l])):
        used = {color[n] for n in adj[l] if n in color}
        c = 0
        while c in used:
            c += 1
        color[l] = c
    return color


def _wrap(a):
    return a - 2 * np.pi * np.round(a / (2 * np.pi))


def _initial_values(fg):
    x = {}
    for l, v in fg.variables.items():
        if getattr(v, "parametric", None) is not None:
            x[l] = np.array(v.parametric, dtype=np.float64)
        elif v.val is not None:
            m = v.val.mean(0)
            w = _WRAP_COORD.get(v.variableType.vartype)
            if w is not None:
                m[w] = np.arctan2(np.sin(v.val[:, w]).mean(), np.cos(v.val[:, w]).mean())
            x[l] = m
        else:
            x[l] = np.zeros(v.variableType.dim)
    return x


_ANALYTIC = (L.POSE2POSE2, L.PRIORPOSE2, L.BEARINGRANGE, L.POSE3POSE3, L.PRIORPOSE3)


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-8:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def _so3_log(R):
    c = np.clip((np.trace(R) - 1) / 2, -1.0, 1.0)
    th = np.arccos(c)
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    if th < 1e-8:
        return 0.5 * v
    if np.pi - th < 1e-6:  # near pi: axis from the symmetric part
        A = (R + np.eye(3)) / 2
        ax = np.sqrt(np.maximum(np.diag(A), 0))
        k = int(np.argmax(ax))
        ax = A[:, k] / ax[k]
        return th * ax / np.linalg.norm(ax)
    return th / (2 * np.sin(th)) * v


def _analytic_blocks(fam, jac, res_dim):
    """per-variable Jacobian blocks [(slot, d r / d increment)] of one factor from the kernel's compact output"""
    I3 = np.eye(3)
    if fam == L.POSE2POSE2:      # jac = (-ry, rx, cos, sin):  d r / d (xp, yp, thp), d r / d q = -I
        Jp = np.array([[1, 0, jac[0]], [0, 1, jac[1]], [0, 0, 1.0]])
        return [(0, Jp), (1, -I3)]
    if fam == L.PRIORPOSE2:
        return [(0, -I3)]
    if fam == L.BEARINGRANGE:    # jac = (d r1 / d l (2), d r2 / d l (2));  d r / d t_p = -d r / d l, d r1 / d thp = 1
        Jl = np.array([[jac[0], jac[1]], [jac[2], jac[3]]])
        Jp = np.column_stack([-Jl, [1.0, 0.0]])
        return [(0, Jp), (1, Jl)]
    if fam == L.POSE3POSE3:      # blocks A, B, C, R_p (include/rome_b200.h); increments (dt, delta) per pose
        A, B, Cq = jac[0:9].reshape(3, 3), jac[9:18].reshape(3, 3), jac[18:27].reshape(3, 3)
        Z = np.zeros((3, 3))
        return [(0, np.block([[I3, A], [Z, B]])), (1, np.block([[-I3, Z], [Z, Cq]]))]
    if fam == L.PRIORPOSE3:
        Z = np.zeros((3, 3))
        return [(0, np.block([[-I3, Z], [Z, jac[0:9].reshape(3, 3)]]))]
    raise ValueError(fam)


def solveGraphParametric(fg: FactorGraph, ctx: Context | None = None, iters: int = 100, h: float = 1e-3,
                         tol: float = 1e-10, covariance: bool = True, analytic: bool | None = None):
    """Levenberg-Marquardt on the variable coordinates.  Returns (labels, values {label: coords}, residual cost, Sigma)
    with Sigma the dense inverse of the final normal matrix (None for graphs above 2000 coordinates).
    analytic: use the kernels' analytic Jacobian blocks (default: whenever every factor family of the graph has them)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    labels = list(fg.variables)
    present = {f.fnc.family for f in fg.factors.values()}
    if analytic is None:
        analytic = present <= set(_ANALYTIC)
    if analytic and not present <= set(_ANALYTIC):
        raise ValueError("analytic Jacobians exist for the five hot families only")
    color = color_variables(fg)
    C = max(color.values()) + 1 if color else 1
    D = max(v.variableType.dim for v in fg.variables.values())
    N = 1 if analytic else 1 + 2 * D * C
    x = _initial_values(fg)
    off, tot = {}, 0
    for l in labels:
        off[l] = tot
        tot += fg.variables[l].variableType.dim
    dg = DeviceGraph(fg, ctx=ctx, N=N, upload_particles=False)
    c = dg.ctx
    fams = [f for f, facs in dg.by_family.items() if facs]
    whiten = {}
    for fam in fams:  # W_f with W_f' W_f = Omega_f
        W = []
        for f in dg.by_family[fam]:
            _, info = getMeasurementParametric(f.fnc)
            W.append(np.linalg.cholesky(np.atleast_2d(info)).T)
        whiten[fam] = np.stack(W)
    zero_meas = {fam: np.zeros((len(dg.by_family[fam]), npad(N), FAMILY[fam][2]), np.float32) for fam in fams}

    def linearize_analytic(x):
        """one launch per family with ONE particle per variable: residuals + analytic Jacobian blocks"""
        for t, vs in dg.by_type.items():
            if vs:
                arr = np.zeros((len(vs), 1, VAR_DIM[t]))
                for v in vs:
                    arr[v.index, 0] = x[v.label]
                c.set_particles(t, arr)
        rows, cols, vals, rvec = [], [], [], []
        r0 = 0
        for fam in fams:
            facs = dg.by_family[fam]
            dr, dj = FAMILY[fam][3], FAMILY[fam][5]
            flags = L.RESIDUAL | (L.JACOBIAN if dj else 0)
            out = c.alloc_host_outputs(fam, flags)
            c.eval_host(fam, flags, meas=zero_meas[fam], **out)
            res = np.asarray(out["res"], dtype=np.float64)
            jac = np.asarray(out["jac"], dtype=np.float64) if dj else None
            W = whiten[fam]
            for k, f in enumerate(facs):
                rvec.append(W[k] @ res[k, 0])
                for slot, Jb in _analytic_blocks(fam, jac[k, 0] if dj else None, dr):
                    l = f.variableOrderSymbols[slot]
                    Jw = W[k] @ Jb
                    rr, cc = np.meshgrid(np.arange(dr), np.arange(Jb.shape[1]), indexing="ij")
                    rows.append((r0 + rr).ravel()); cols.append((off[l] + cc).ravel()); vals.append(Jw.ravel())
                r0 += dr
        Jm = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, tot))
        return np.concatenate(rvec), Jm

    def linearize_fd(x):
        """one launch per family: whitened residual vector and sparse Jacobian at x"""
        for t, vs in dg.by_type.items():
            if not vs:
                continue
            d = VAR_DIM[t]
            arr = np.zeros((len(vs), N, d))
            for v in vs:
                arr[v.index] = x[v.label]
                base = 1 + 2 * D * color[v.label]
                for i in range(d):
                    arr[v.index, base + 2 * i, i] += h
                    arr[v.index, base + 2 * i + 1, i] -= h
            c.set_particles(t, arr)
        rows, cols, vals, rvec = [], [], [], []
        r0 = 0
        for fam in fams:
            facs = dg.by_family[fam]
            dr = FAMILY[fam][3]
            out = c.alloc_host_outputs(fam, L.RESIDUAL)
            c.eval_host(fam, L.RESIDUAL, meas=zero_meas[fam], **out)
            res = np.asarray(out["res"], dtype=np.float64)  # [nF][Npad][dr]
            ang = _ANGLE_COMPONENTS.get(fam, ())
            W = whiten[fam]
            for k, f in enumerate(facs):
                r = res[k, 0]
                rvec.append(W[k] @ r)
                for l in f.variableOrderSymbols:
                    v = fg.variables[l]
                    d = v.variableType.dim
                    base = 1 + 2 * D * color[l]
                    J = (res[k, base:base + 2 * d:2] - res[k, base + 1:base + 2 * d:2]).T  # [dr][d]
                    for a in ang:
                        J[a] = _wrap(J[a])
                    J = W[k] @ (J / (2 * h))
                    rr, cc = np.meshgrid(np.arange(dr), np.arange(d), indexing="ij")
                    rows.append((r0 + rr).ravel()); cols.append((off[l] + cc).ravel()); vals.append(J.ravel())
                r0 += dr
        Jm = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(r0, tot))
        return np.concatenate(rvec), Jm

    def retract(x, delta):
        y = {}
        for l in labels:
            v = fg.variables[l]
            d = v.variableType.dim
            z = x[l] + delta[off[l]:off[l] + d]
            w = _WRAP_COORD.get(v.variableType.vartype)
            if w is not None:
                z[w] = _wrap(z[w])
            if v.variableType.vartype == L.POSE3:
                if analytic:  # the analytic blocks are for right perturbations: R <- R Exp(delta)
                    z[3:] = _so3_log(_so3_exp(x[l][3:]) @ _so3_exp(delta[off[l] + 3:off[l] + 6]))
                th = np.linalg.norm(z[3:])  # keep the rotation vector in the principal range
                if th > np.pi:
                    z[3:] *= 1 - 2 * np.pi / th
            y[l] = z
        return y

    linearize = linearize_analytic if analytic else linearize_fd

    lam = 1e-6
    r, J = linearize(x)
    cost = float(r @ r)
    H = None
    for _ in range(iters):
        H = (J.T @ J).tocsc()
        g = J.T @ r
        step = spla.spsolve(H + lam * sp.diags(H.diagonal() + 1e-12), -g)
        xn = retract(x, step)
        rn, Jn = linearize(xn)
        cn = float(rn @ rn)
        if cn <= cost * (1 + 1e-12):
            x, r, J = xn, rn, Jn
            lam = max(lam * 0.1, 1e-12)
            done = abs(cost - cn) <= tol * max(cost, 1e-30) or np.abs(step).max() < tol
            cost = cn
            if done:
                break
        else:
            lam *= 10.0
            if lam > 1e8:
                break
    Sigma = None
    if covariance and tot <= 2000:
        Sigma = np.linalg.inv((J.T @ J).toarray())
    for l in labels:
        fg.variables[l].parametric = x[l].copy()
    return labels, x, cost, Sigma
