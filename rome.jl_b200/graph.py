"""Host-side factor-graph container + the solver-facing entry points of the hot path.

Mirrors the calls RoME's generators, tests and examples make (DFG / IIF names):
    initfg, addVariable (addVariable!), addFactor (addFactor!), ls, lsf, getVal, setVal (initVariable!),
    sampleFactor, approxConv / approxConvBelief, calcFactorResidualTemporary, calcFactorResidual,
    initAll (graphinit), and `DeviceGraph` -- the batched per-family evaluation that replaces IIF's
    per-particle functor loop (SURVEY.md 3.1 HOT LOOP).
All arithmetic runs in librome_b200.so on the GPU; this module only moves parameters and arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib as L
from .engine import FAMILY, VAR_DIM, Context, meas_to_offsets, offsets_to_meas, rows_to_particle_major
from .factors import (PARTIAL_FACTORS, POINT2_FACTORS, SCALAR_FACTORS, TERNARY_FACTORS, Rotation3, AbstractFactor, InferenceVariable, Point2, Point3,
                      Pose2, Pose2Point2BearingRange, Pose3, factor_mean)

_VARCLASS = {L.POSE2: Pose2, L.POINT2: Point2, L.POSE3: Pose3, L.POINT3: Point3, L.ROTATION3: Rotation3}


@dataclass
class SolverParams:
    """IIF SolverParams defaults as serialized in the reference's test/testdata/g2otest.tar.gz:dfg.json"""
    N: int = 100
    gibbsIters: int = 3
    inflateCycles: int = 3
    inflation: float = 5.0
    spreadNH: float = 3.0
    graphinit: bool = True
    multiproc: bool = False


@dataclass
class DFGVariable:
    label: str
    variableType: type
    index: int  # position inside the device particle store of its type
    val: np.ndarray | None = None  # [N][d] coordinates (reference `vecval` layout)
    tags: list = field(default_factory=list)

    @property
    def initialized(self):
        return self.val is not None


@dataclass
class DFGFactor:
    label: str
    variableOrderSymbols: list
    fnc: AbstractFactor
    index: int  # position inside the device factor table of its family
    graphinit: bool = True
    tags: list = field(default_factory=list)


class FactorGraph:
    def __init__(self, solverParams: SolverParams | None = None):
        self.solverParams = solverParams or SolverParams()
        self.variables: dict[str, DFGVariable] = {}
        self.factors: dict[str, DFGFactor] = {}
        self._nvar = {t: 0 for t in VAR_DIM}
        self._nfac = {f: 0 for f in FAMILY}

    def __getitem__(self, label):
        label = str(label)
        return self.variables[label] if label in self.variables else self.factors[label]


def initfg(solverParams: SolverParams | None = None) -> FactorGraph:
    return FactorGraph(solverParams)


def getSolverParams(fg: FactorGraph) -> SolverParams:
    return fg.solverParams


def addVariable(fg: FactorGraph, label, variableType, tags=None, N=None) -> DFGVariable:
    """addVariable!(fg, :x0, Pose2)"""
    label = str(label)
    if label in fg.variables:
        raise KeyError(f"variable {label} already exists")
    if not (isinstance(variableType, type) and issubclass(variableType, InferenceVariable)):
        raise TypeError("variableType must be Pose2, Point2, Pose3 or Point3")
    v = DFGVariable(label, variableType, fg._nvar[variableType.vartype], tags=list(tags or []))
    fg._nvar[variableType.vartype] += 1
    fg.variables[label] = v
    return v


def addFactor(fg: FactorGraph, labels, fnc: AbstractFactor, graphinit=None, tags=None) -> DFGFactor:
    """addFactor!(fg, [:x0; :x1], Pose2Pose2(...)); label follows DFG's default `x0x1f1` scheme."""
    labels = [str(l) for l in labels]
    if len(labels) != len(fnc.variabletypes):
        raise ValueError(f"{type(fnc).__name__} connects {len(fnc.variabletypes)} variable(s), got {len(labels)}")
    for l, t in zip(labels, fnc.variabletypes):
        if l not in fg.variables:
            raise KeyError(f"variable {l} does not exist")
        if fg.variables[l].variableType is not t:
            raise TypeError(f"variable {l} is {fg.variables[l].variableType.__name__}, factor expects {t.__name__}")
    base, k = "".join(labels) + "f", 1
    while base + str(k) in fg.factors:
        k += 1
    f = DFGFactor(base + str(k), labels, fnc, fg._nfac[fnc.family],
                  graphinit=fg.solverParams.graphinit if graphinit is None else graphinit, tags=list(tags or []))
    fg._nfac[fnc.family] += 1
    fg.factors[f.label] = f
    return f


def ls(fg: FactorGraph, variableType=None):
    return [l for l, v in fg.variables.items() if variableType is None or v.variableType is variableType]


def lsf(fg: FactorGraph, factorType=None):
    return [l for l, f in fg.factors.items() if factorType is None or isinstance(f.fnc, factorType)]


def getVal(fg: FactorGraph, label) -> np.ndarray:
    v = fg.variables[str(label)]
    if v.val is None:
        raise ValueError(f"variable {label} is not initialized")
    return v.val


def setVal(fg: FactorGraph, label, val):
    """initVariable!(fg, :x0, pts): val is [N][d] coordinates."""
    v = fg.variables[str(label)]
    val = np.ascontiguousarray(val, dtype=np.float64)
    if val.ndim != 2 or val.shape[1] != v.variableType.dim:
        raise ValueError("val must be [N][d] coordinates of the variable's type")
    v.val = val


# ------------------------------------------------------------------------------------------------------
# device binding
# ------------------------------------------------------------------------------------------------------
_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _factor_arrays(facs):
    """per-family parameter arrays in table order for Context.set_factors_*"""
    fam = facs[0].fnc.family
    if fam == L.BEARINGRANGE:
        return (np.array([[f.fnc.bearing.mu, f.fnc.bearing.sigma] for f in facs]),
                np.array([[f.fnc.range.mu, f.fnc.range.sigma] for f in facs]))
    if isinstance(facs[0].fnc, SCALAR_FACTORS):
        return np.array([[f.fnc.Z.mu, f.fnc.Z.sigma] for f in facs]), None
    return np.stack([f.fnc.Z.mu for f in facs]), np.stack([f.fnc.Z.Sigma for f in facs])


class DeviceGraph:
    """The whole graph resident on one GPU: particle stores per variable type + one factor table per
    family.  `eval(family, flags)` is the batched replacement of IIF's per-factor / per-particle loop."""

    def __init__(self, fg: FactorGraph, ctx: Context | None = None, N: int | None = None, upload_particles: bool = True):
        self.fg = fg
        self.ctx = ctx or Context(0)
        self.N = N or fg.solverParams.N
        self.by_type = {t: sorted((v for v in fg.variables.values() if v.variableType.vartype == t),
                                  key=lambda v: v.index) for t in VAR_DIM}
        self.by_family = {f: sorted((x for x in fg.factors.values() if x.fnc.family == f), key=lambda x: x.index)
                          for f in FAMILY}
        self.upload_factors()
        if upload_particles:
            self.upload_particles()

    def upload_particles(self):
        for t, vs in self.by_type.items():
            if not vs:
                continue
            arr = np.zeros((len(vs), self.N, VAR_DIM[t]))
            for v in vs:
                if v.val is not None:
                    if v.val.shape[0] != self.N:
                        raise ValueError(f"{v.label} holds {v.val.shape[0]} particles, graph N is {self.N}")
                    arr[v.index] = v.val
            self.ctx.set_particles(t, arr)

    def upload_factors(self):
        fg, c = self.fg, self.ctx
        for fam, facs in self.by_family.items():
            if not facs:
                continue
            i0 = [fg.variables[f.variableOrderSymbols[0]].index for f in facs]
            i1 = [fg.variables[f.variableOrderSymbols[1]].index for f in facs] if FAMILY[fam][1] is not None else None
            a, b = _factor_arrays(facs)
            if fam == L.POSE2POSE2:
                c.set_factors_pose2pose2(i0, i1, a, b)
            elif fam == L.PRIORPOSE2:
                c.set_factors_priorpose2(i0, a, b)
            elif fam == L.BEARINGRANGE:
                c.set_factors_bearingrange(i0, i1, a, b)
            elif fam == L.POSE3POSE3:
                c.set_factors_pose3pose3(i0, i1, a, b)
            elif fam == L.PRIORPOSE3:
                c.set_factors_priorpose3(i0, a, b)
            elif isinstance(facs[0].fnc, TERNARY_FACTORS):
                i2 = [fg.variables[f.variableOrderSymbols[2]].index for f in facs]
                c.set_factors_ternary(fam, i0, i1, i2, a, b)
            elif isinstance(facs[0].fnc, POINT2_FACTORS):
                c.set_factors_point2(fam, i0, i1, a, b)
            elif isinstance(facs[0].fnc, SCALAR_FACTORS):
                c.set_factors_scalar(fam, i0, i1, a)
            else:  # every other family holds one MvNormal belief
                c.set_factors_gaussian(fam, i0, i1, a, b)

    def means(self, family) -> np.ndarray:
        return np.stack([factor_mean(f.fnc) for f in self.by_family[family]])

    def download_particles(self):
        """write device particles back into the host graph (updateFromSubgraph analogue)"""
        for t, vs in self.by_type.items():
            if vs:
                arr = self.ctx.get_particles(t)
                for v in vs:
                    v.val = arr[v.index].copy()

    def eval(self, family, flags, meas=None, seed=0, stream_id=0, first=0, count=-1):
        """Host-array convenience around Context.eval_host.  meas: Float64 samples [nF][N][dm] in the
        reference's coordinates (or None with SAMPLE).  Returns dict of Float64 arrays in the reference layout
        ([nF][N][d]); proposals are absolute coordinates."""
        c, N = self.ctx, self.N
        mu = self.means(family)
        kw = c.alloc_host_outputs(family, flags)
        if not (flags & L.SAMPLE):
            kw["meas"] = meas_to_offsets(meas, mu)
        c.eval_host(family, flags, seed=seed, stream_id=stream_id, first=first, count=count, **kw)
        vt0, vt1 = FAMILY[family][0], FAMILY[family][1]
        out = {}
        if "res" in kw:
            out["res"] = rows_to_particle_major(kw["res"], N)
        if "meas_out" in kw:
            out["meas"] = offsets_to_meas(kw["meas_out"], mu, N)
        if "stats" in kw:
            out["stats"] = kw["stats"]
        if "jac" in kw:
            out["jac"] = rows_to_particle_major(kw["jac"], N)
        facs = self.by_family[family]
        if "prop_fwd" in kw:
            tgt = vt1 if vt1 is not None else vt0
            slot = 1 if vt1 is not None else 0
            idx = [self.fg.variables[f.variableOrderSymbols[slot]].index for f in facs]
            out["prop_fwd"] = rows_to_particle_major(kw["prop_fwd"], N) + c.get_anchors(tgt)[idx][:, None, :]
        if "prop_bwd" in kw:
            idx = [self.fg.variables[f.variableOrderSymbols[0]].index for f in facs]
            out["prop_bwd"] = rows_to_particle_major(kw["prop_bwd"], N) + c.get_anchors(vt0)[idx][:, None, :]
        return out


# ------------------------------------------------------------------------------------------------------
# IIF-named entry points used by RoME's tests
# ------------------------------------------------------------------------------------------------------
def _mini_graph(fnc, points_by_slot, N):
    """temporary graph holding one factor (calcFactorResidualTemporary builds one too [IIF-knowledge])"""
    fg = initfg(SolverParams(N=N))
    labels = []
    for k, (t, pts) in enumerate(zip(fnc.variabletypes, points_by_slot)):
        v = addVariable(fg, f"t{k}", t)
        v.val = np.ascontiguousarray(pts, dtype=np.float64).reshape(N, t.dim)
        labels.append(v.label)
    addFactor(fg, labels, fnc)
    return fg


def calcFactorResidualTemporary(fnc, varTypes, meas, points, ctx: Context | None = None) -> np.ndarray:
    """One residual evaluation `cf(meas, points...)` on coordinates.
    meas: measurement coordinates (tangent coordinates for relative factors, point coordinates for priors,
    (bearing, range) for Pose2Point2BearingRange); points: one coordinate vector per variable."""
    if tuple(varTypes) != tuple(fnc.variabletypes):
        raise TypeError("variable types do not match the factor")
    fg = _mini_graph(fnc, [np.asarray(p, dtype=np.float64)[None, :] for p in points], 1)
    dg = DeviceGraph(fg, ctx or default_context(), N=1)
    out = dg.eval(fnc.family, L.RESIDUAL, meas=np.asarray(meas, dtype=np.float64).reshape(1, 1, -1))
    return out["res"][0, 0]


def calcFactorResidual(fg: FactorGraph, flabel, meas, *points, ctx: Context | None = None) -> np.ndarray:
    f = fg.factors[str(flabel)]
    return calcFactorResidualTemporary(f.fnc, f.fnc.variabletypes, meas, points, ctx=ctx)


def sampleFactor(fg_or_fnc, flabel=None, N: int = 1, seed=0, ctx: Context | None = None) -> np.ndarray:
    """N draws of getSample for a factor, as coordinates [N][dm] (tangent coordinates for relative factors,
    point coordinates for priors, (bearing, range) for Pose2Point2BearingRange; BearingRange2D.jl:17-27)."""
    fnc = fg_or_fnc.factors[str(flabel)].fnc if isinstance(fg_or_fnc, FactorGraph) else fg_or_fnc
    pts = [np.zeros((N, t.dim)) for t in fnc.variabletypes]
    fg = _mini_graph(fnc, pts, N)
    dg = DeviceGraph(fg, ctx or default_context(), N=N)
    return dg.eval(fnc.family, L.SAMPLE | L.WRITE_MEAS, seed=seed)["meas"][0]


getSample = sampleFactor


def approxConv(fg: FactorGraph, flabel, target, N: int | None = None, seed=0, ctx: Context | None = None):
    """Convolve the factor's measurement belief with the other variable's particles onto `target`
    (IIF approxConv / approxConvBelief) using the closed-form root of the residual.
    Returns [N][d] coordinates on the target variable."""
    f = fg.factors[str(flabel)]
    target = str(target)
    if target not in f.variableOrderSymbols:
        raise KeyError(f"{target} is not connected to {flabel}")
    fnc = f.fnc
    N = N or fg.solverParams.N
    slot = f.variableOrderSymbols.index(target)
    if isinstance(fnc, TERNARY_FACTORS) and slot != 1:
        # qhat = p o (...) has a closed form onto the SECOND variable; onto p or onto the third variable (bRa / Delta)
        # the reference runs its numeric solver -- not shipped
        raise NotImplementedError(f"{type(fnc).__name__}: only the convolution onto the second variable is closed-form here")
    pts = []
    for k, (l, t) in enumerate(zip(f.variableOrderSymbols, fnc.variabletypes)):
        v = fg.variables[l]
        if k == slot:
            pts.append(np.zeros((N, t.dim)) if v.val is None or v.val.shape[0] != N else v.val)
        else:
            if v.val is None:
                raise ValueError(f"variable {l} must be initialized to convolve through {flabel}")
            val = v.val
            if val.shape[0] != N:  # resample with replacement like IIF does when N differs
                val = val[np.random.default_rng(seed).integers(0, val.shape[0], N)]
            pts.append(val)
    last = len(fnc.variabletypes) - 1
    if isinstance(fnc, PARTIAL_FACTORS):
        raise NotImplementedError(f"{type(fnc).__name__} is a partial constraint: the convolution has no unique root")
    if isinstance(fnc, SCALAR_FACTORS):
        # one equation for a 2- or 3-dimensional target: a 1-parameter family of roots, left to the optimiser's
        # start point in the reference; no closed form is shipped
        raise NotImplementedError(f"{type(fnc).__name__}: the convolution has no unique root")
    if isinstance(fnc, POINT2_FACTORS) and not fnc.is_prior and slot != last:
        raise NotImplementedError(f"{type(fnc).__name__}: only the convolution onto the last variable is closed-form here")
    if isinstance(fnc, TERNARY_FACTORS):
        flag, key = L.PROPOSAL_FWD, "prop_fwd"
    elif fnc.is_prior or slot == last:
        flag, key = L.PROPOSAL_FWD, "prop_fwd"
    elif isinstance(fnc, Pose2Point2BearingRange):
        # pose from landmark is a 1-parameter family: the reference leaves it to the optimiser's start
        # point + inflation noise; no closed form is shipped for it.
        raise NotImplementedError("Pose2Point2BearingRange: convolution onto the pose has no unique root")
    elif FAMILY[fnc.family][7] == 0:
        raise NotImplementedError(f"{type(fnc).__name__}: only the convolution onto the last variable is closed-form here")
    else:
        flag, key = L.PROPOSAL_BWD, "prop_bwd"
    dg = DeviceGraph(_mini_graph(fnc, pts, N), ctx or default_context(), N=N)
    return dg.eval(fnc.family, L.SAMPLE | flag, seed=seed)[key][0]


approxConvBelief = approxConv


def approxDeconv(fg: FactorGraph, flabel, N: int | None = None, seed=0, ctx: Context | None = None):
    """IIF approxDeconv(fg, :x0x1f1) -> (pts, meas): `pts` = the measurement each particle pair implies (the residual
    solved for the measurement, ROME_B200_DECONV), `meas` = N fresh samples of the factor's belief, both as [N][dm]
    measurement coordinates (test/testBasicPose2Conv.jl:51-56 compares the two sets)."""
    f = fg.factors[str(flabel)]
    fnc = f.fnc
    N = N or fg.solverParams.N
    pts = []
    for l in f.variableOrderSymbols:
        v = fg.variables[l]
        if v.val is None or v.val.shape[0] != N:
            raise ValueError(f"variable {l} must hold {N} particles to deconvolve {flabel}")
        pts.append(v.val)
    dg = DeviceGraph(_mini_graph(fnc, pts, N), ctx or default_context(), N=N)
    c, mu = dg.ctx, dg.means(fnc.family)
    out = c.alloc_host_outputs(fnc.family, L.SAMPLE | L.DECONV)
    c.eval_host(fnc.family, L.SAMPLE | L.DECONV, seed=seed, **out)
    dec = offsets_to_meas(out["meas_out"], mu, N)[0]
    meas = dg.eval(fnc.family, L.SAMPLE | L.WRITE_MEAS, seed=seed)["meas"][0]
    return dec, meas


def solveFactorParametric(fg: FactorGraph, flabel, src, target, ctx: Context | None = None) -> np.ndarray:
    """IIF.solveFactorParametric(dfg, fct, [srcsym => val], trgsym): the target variable's coordinates that zero the
    residual at the factor's MEAN measurement, given the other variable's coordinates -- the closed-form proposal
    kernels evaluated on one particle with the measurement supplied (no sampling).  `src` = (label, coordinates)."""
    f = fg.factors[str(flabel)]
    fnc, target = f.fnc, str(target)
    if fnc.is_prior or target not in f.variableOrderSymbols or len(f.variableOrderSymbols) != 2:
        raise ValueError(f"{flabel} is not a two-variable factor connected to {target}")
    slot = f.variableOrderSymbols.index(target)
    if str(src[0]) != f.variableOrderSymbols[1 - slot]:
        raise KeyError(f"{src[0]} is not the other variable of {flabel}")
    if isinstance(fnc, PARTIAL_FACTORS) or isinstance(fnc, SCALAR_FACTORS) or FAMILY[fnc.family][7 if slot == 0 else 6] == 0:
        raise NotImplementedError(f"{type(fnc).__name__}: no closed-form solve onto {'the first' if slot == 0 else 'the last'} variable")
    pts = [None, None]
    pts[1 - slot] = np.asarray(src[1], dtype=np.float64).reshape(1, -1)
    pts[slot] = np.zeros((1, fnc.variabletypes[slot].dim))
    dg = DeviceGraph(_mini_graph(fnc, pts, 1), ctx or default_context(), N=1)
    flag, key = (L.PROPOSAL_FWD, "prop_fwd") if slot == 1 else (L.PROPOSAL_BWD, "prop_bwd")
    return dg.eval(fnc.family, L.RESIDUAL | flag, meas=factor_mean(fnc).reshape(1, 1, -1))[key][0, 0]


def accumulateFactorMeans(fg: FactorGraph, fctsyms, ctx: Context | None = None) -> np.ndarray:
    """IIF.accumulateFactorMeans(dfg, [:x0f1; :x0x1f1; ...]) (test/testAccumulateFactors.jl:19-30): start from the
    first factor's mean when it is a prior (else from the mean of the chain's first variable's particles) and carry
    the value along the chain of relative factors through `solveFactorParametric`."""
    fctsyms = [str(s) for s in fctsyms]
    if not fctsyms:
        raise ValueError("accumulateFactorMeans needs at least one factor")
    f0 = fg.factors[fctsyms[0]]
    if f0.fnc.is_prior:
        val, cur, rest = factor_mean(f0.fnc).astype(np.float64), f0.variableOrderSymbols[0], fctsyms[1:]
    else:
        vs = f0.variableOrderSymbols
        nxt = [v for v in vs if len(fctsyms) > 1 and v in fg.factors[fctsyms[1]].variableOrderSymbols]
        cur = next(v for v in vs if v not in nxt[:1]) if nxt else vs[0]
        v = fg.variables[cur]
        if v.val is None:
            raise ValueError(f"variable {cur} has no estimate to start the chain from")
        val = v.val.mean(0)
        if v.variableType is Pose2:
            val[2] = np.arctan2(np.sin(v.val[:, 2]).mean(), np.cos(v.val[:, 2]).mean())
        rest = fctsyms
    for fl in rest:
        others = [v for v in fg.factors[fl].variableOrderSymbols if v != cur]
        if len(others) != 1:
            raise ValueError(f"{fl} does not continue the chain from {cur}")
        val = solveFactorParametric(fg, fl, (cur, val), others[0], ctx=ctx)
        cur = others[0]
    return val


def initAll(fg: FactorGraph, seed=0, ctx: Context | None = None):
    """graphinit: initialise every variable by propagating priors through the factors in insertion order
    (IIF initAll!/doautoinit! uses the same approxConv path)."""
    progress = True
    k = 0
    while progress:
        progress = False
        for f in fg.factors.values():
            vs = [fg.variables[l] for l in f.variableOrderSymbols]
            if f.fnc.is_prior:
                if not vs[0].initialized:
                    vs[0].val = approxConv(fg, f.label, vs[0].label, seed=seed + k, ctx=ctx)
                    progress, k = True, k + 1
            elif vs[0].initialized and not vs[1].initialized:
                if len(vs) > 2 and not vs[2].initialized:
                    continue  # a third variable (bRa / Delta) nothing has initialised yet: test/testPose3.jl:118-119
                if FAMILY[f.fnc.family][6] == 0:
                    continue  # no closed-form root toward the last variable (ranges, bearing, partial Pose3): defer it
                vs[1].val = approxConv(fg, f.label, vs[1].label, seed=seed + k, ctx=ctx)
                progress, k = True, k + 1
            elif vs[1].initialized and not vs[0].initialized:
                # a backward root that needs no starting point exists for Pose2Pose2 / Pose3Pose3 only (BearingRange's
                # keeps the pose's current heading, the point / scalar families have none): defer the variable otherwise
                if FAMILY[f.fnc.family][7] == 0 or isinstance(f.fnc, Pose2Point2BearingRange):
                    continue
                vs[0].val = approxConv(fg, f.label, vs[0].label, seed=seed + k, ctx=ctx)
                progress, k = True, k + 1
    return fg
