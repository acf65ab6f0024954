"""Array-level workloads (BASELINE.json configs / SURVEY.md 8d) and their rank-local views for owner-sharded runs.

A workload is a dict
    N          particles per variable
    particles  {vartype: Float64 [nvars][N][d]}           (reference layout, `vecval`)
    families   {family: dict(i0, i1 | None, a, b)}        (a, b) = the arrays Context.set_factors_* takes
built either from a host FactorGraph (`graph_arrays`: the canonical generators of canonical.py) or directly as arrays
(`manhattan_arrays`: the bench's 10k-poses-per-GPU Manhattan-shaped graph, vectorised so that the 80k-pose graph of an
8-GPU run is built in a second).  `local_view` cuts a workload down to what one rank of an OwnerSharding holds."""
from __future__ import annotations

import math

import numpy as np

from . import _lib as L
from .engine import FAMILY, VAR_DIM


def upload_family(ctx, fam, i0, i1, a, b):
    """Context.set_factors_* by family id (the arrays of graph._factor_arrays)"""
    if fam == L.POSE2POSE2:
        ctx.set_factors_pose2pose2(i0, i1, a, b)
    elif fam == L.PRIORPOSE2:
        ctx.set_factors_priorpose2(i0, a, b)
    elif fam == L.BEARINGRANGE:
        ctx.set_factors_bearingrange(i0, i1, a, b)
    elif fam == L.POSE3POSE3:
        ctx.set_factors_pose3pose3(i0, i1, a, b)
    elif fam == L.PRIORPOSE3:
        ctx.set_factors_priorpose3(i0, a, b)
    elif fam in (L.PRIORPOINT2, L.POINT2POINT2, L.POSE2POINT2):
        ctx.set_factors_point2(fam, i0, i1, a, b)
    elif fam in (L.POSE2POINT2RANGE, L.POINT2POINT2RANGE, L.POSE2POINT2BEARING):
        ctx.set_factors_scalar(fam, i0, i1, a)
    else:
        ctx.set_factors_gaussian(fam, i0, i1, a, b)


def graph_arrays(fg, N=None):
    """workload arrays of a host FactorGraph whose variables hold particles (`seed_particles` / a solve)"""
    from .graph import _factor_arrays
    N = N or fg.solverParams.N
    particles, families = {}, {}
    for t in VAR_DIM:
        vs = sorted((v for v in fg.variables.values() if v.variableType.vartype == t), key=lambda v: v.index)
        if vs:
            particles[t] = np.stack([v.val for v in vs])
    for fam in FAMILY:
        facs = sorted((x for x in fg.factors.values() if x.fnc.family == fam), key=lambda x: x.index)
        if not facs:
            continue
        i0 = np.array([fg.variables[f.variableOrderSymbols[0]].index for f in facs], np.int32)
        i1 = np.array([fg.variables[f.variableOrderSymbols[1]].index for f in facs], np.int32) \
            if FAMILY[fam][1] is not None else None
        a, b = _factor_arrays(facs)
        families[fam] = dict(i0=i0, i1=i1, a=a, b=b)
    return dict(N=N, particles=particles, families=families)


def manhattan_arrays(poses=10000, loop_fraction=0.2, seed=2, N=100, particle_seed=1, sigma=(0.1, 0.12, 0.02)):
    """canonical.generateGraph_ManhattanShaped + seed_particles as arrays (same random streams, same graph): unit-step
    grid walk with turns 0 / +-pi/2 (p = .7 / .15 / .15), poses-1 odometry Pose2Pose2 + ~loop_fraction*poses loop
    closures between poses <= 2 m and >= 20 steps apart, Sigma = diag(1/44, 1/380, 1/9700) (examples/manhattan.g2o
    median information), PriorPose2 sigma (0.1, 0.1, 0.05) on x0 (examples/ManhattanDatasetBatch.jl:30-32)."""
    rng = np.random.default_rng(seed)
    Sigma = np.diag([1 / 44.0, 1 / 380.0, 1 / 9700.0])
    Lc = np.linalg.cholesky(Sigma)
    turns = rng.choice([0.0, math.pi / 2, -math.pi / 2], size=poses - 1, p=[0.7, 0.15, 0.15])
    # truth: heading = running sum of the turns (wrapped like _se2_compose), position = running sum of unit steps
    k = np.concatenate([[0], np.cumsum(np.rint(turns / (math.pi / 2)).astype(np.int64))])
    heading = np.arctan2(np.sin(k * (math.pi / 2)), np.cos(k * (math.pi / 2)))
    cs = np.array([[1, 0], [0, 1], [-1, 0], [0, -1]], np.float64)[k[:-1] % 4]  # exact unit steps on the grid
    xy = np.concatenate([np.zeros((1, 2)), np.cumsum(cs, 0)])
    truth = np.column_stack([xy, heading])
    odo_noise = rng.normal(size=(poses - 1, 3))
    mu_odo = np.column_stack([np.ones(poses - 1), np.zeros(poses - 1), turns]) + odo_noise @ Lc.T
    # loop closures by spatial hashing of the integer grid positions (as in canonical.py)
    key = np.rint(xy).astype(np.int64)
    cells, cand = {}, []
    for i in range(poses):
        kx, ky = int(key[i, 0]), int(key[i, 1])
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for j in cells.get((kx + dx, ky + dy), ()):
                    if i - j >= 20 and math.hypot(xy[i, 0] - xy[j, 0], xy[i, 1] - xy[j, 1]) <= 2.0:
                        cand.append((j, i))
        cells.setdefault((kx, ky), []).append(i)
    want = int(loop_fraction * poses)
    if len(cand) > want:
        cand = [cand[c] for c in sorted(rng.choice(len(cand), want, replace=False))]
    cand = np.asarray(cand, np.int64).reshape(-1, 2)
    j, i = cand[:, 0], cand[:, 1]
    c, s = np.cos(truth[j, 2]), np.sin(truth[j, 2])
    d = truth[i, :2] - truth[j, :2]
    dth = truth[i, 2] - truth[j, 2]
    rel = np.column_stack([c * d[:, 0] + s * d[:, 1], -s * d[:, 0] + c * d[:, 1], np.arctan2(np.sin(dth), np.cos(dth))])
    mu_loop = rel + rng.normal(size=(len(cand), 3)) @ Lc.T
    ip = np.concatenate([np.arange(poses - 1), j]).astype(np.int32)
    iq = np.concatenate([np.arange(1, poses), i]).astype(np.int32)
    mu = np.concatenate([mu_odo, mu_loop])
    prng = np.random.default_rng(particle_seed)
    particles = truth[:, None, :] + prng.normal(size=(poses, N, 3)) * np.asarray(sigma)
    return dict(N=N, truth=truth, particles={L.POSE2: particles},
                families={L.POSE2POSE2: dict(i0=ip, i1=iq, a=mu, b=np.broadcast_to(Sigma, (len(ip), 3, 3)).copy()),
                          L.PRIORPOSE2: dict(i0=np.zeros(1, np.int32), i1=None, a=np.zeros((1, 3)),
                                             b=(np.diag([0.1, 0.1, 0.05]) ** 2)[None])})


def sharding_of(w, world, balance=True):
    """OwnerSharding of a workload: contiguous variable ranges per type, bounds chosen so that every rank evaluates
    the same number of factors (factors live on the owner of their first variable)"""
    from .sharding import OwnerSharding, balanced_bounds
    nvars = {vt: p.shape[0] for vt, p in w["particles"].items()}
    fams = {fam: (FAMILY[fam][0], FAMILY[fam][1], f["i0"], f["i1"]) for fam, f in w["families"].items()}
    bounds = None
    if balance and world > 1:
        bounds = {}
        for vt in nvars:
            firsts = [f["i0"] for fam, f in w["families"].items() if FAMILY[fam][0] == vt]
            if firsts:
                bounds[vt] = balanced_bounds(firsts, nvars[vt], world)
    # launch geometry of every family's sampled RESIDUAL|STATS launch on a B200 (148 SMs): where the cut block is placed
    from . import _lib as L
    from .engine import plan_query
    geometry = {}
    for fam in fams:
        try:
            p = plan_query(fam, L.RESIDUAL | L.STATS | L.SAMPLE, int(w["N"]))
            geometry[fam] = (p["warps"], 148 * p["ctas_per_sm"])
        except Exception:
            pass
    return OwnerSharding(world, nvars, fams, bounds, geometry)


def local_view(w, sh, rank, fill_halo=False):
    """rank-local arrays of workload `w` under sharding `sh`: particles of the owned variables followed by the halo
    slots (zeros unless fill_halo: the owners push them), every family's factors in local table order (cut factors at
    local indices [cut_first, cut_first + n_cut)) with local variable indices"""
    loc = sh.local(rank)
    particles = {}
    for vt, p in w["particles"].items():
        gid = loc["var_global"][vt]
        arr = p[gid].copy()
        if not fill_halo:
            n_own = loc["own"][vt][1] - loc["own"][vt][0]
            arr[n_own:] = 0.0
        particles[vt] = arr
    families = {}
    for fam, f in w["families"].items():
        lf = loc["fam"][fam]
        o = lf["order"]
        families[fam] = dict(i0=lf["i0"], i1=lf["i1"], a=np.asarray(f["a"])[o],
                             b=None if f["b"] is None else np.asarray(f["b"])[o],
                             n_interior=lf["n_interior"], n_cut=lf["n_cut"], cut_first=lf["cut_first"], order=o,
                             dst_rank=lf["dst_rank"], dst_row=lf["dst_row"], recv=lf["recv"])
    return dict(N=w["N"], particles=particles, families=families, loc=loc)
