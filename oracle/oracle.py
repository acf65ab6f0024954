"""ctypes loader for the float64 C oracle + a NumPy twin of the same arithmetic.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from the product package
(rome.jl_b200/ must not import anything under oracle/).

Parity status: PINNED (tests/test_oracle_golden.py checks every function here against
the reference's own known-answer vectors in tests/golden/known_answers.json).

Reference lines restated (paths relative to /root/reference):
  src/factors/Pose2D.jl:51-67, src/factors/PriorPose2.jl:19-25,37-47,
  src/factors/BearingRange2D.jl:48-64, src/factors/Pose3Pose3.jl:17-29,
  src/factors/Pose3D.jl:15-19.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librome_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/rome_oracle.c with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "rome_oracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        d, i32, u64 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_uint64
        _lib.rome_oracle_sym_rem.restype = C.c_double
        _lib.rome_oracle_sym_rem.argtypes = [C.c_double]
        for name, n in [("pose2pose2", 4), ("priorpose2", 3), ("bearingrange", 4), ("pose3pose3", 4),
                        ("priorpose3", 3), ("pose2pose2_fwd", 3), ("pose2pose2_bwd", 3),
                        ("bearingrange_fwd", 3), ("pose3pose3_fwd", 3), ("pose3pose3_bwd", 3),
                        ("so3_exp", 2), ("so3_log", 2), ("priorpoint2", 3), ("point2point2", 4), ("pose2point2", 4),
                        ("range2", 4), ("pose2point2bearing", 4), ("priorpoint3", 3), ("point3point3", 4),
                        ("pose3pose3xyyaw", 4), ("pose3pose3rotation", 4), ("pose3pose3unittrans", 4),
                        ("pose3_point", 3), ("pose3_coords", 3), ("pose3pose3rotoffset", 5),
                        ("pose3pose3transform", 5)]:
            fn = getattr(_lib, "rome_oracle_" + name)
            fn.restype = None
            fn.argtypes = [d] * n
        _lib.rome_oracle_sweep_pose2pose2.argtypes = [C.c_int, C.c_int, i32, i32, d, d, d, C.c_int]
        _lib.rome_oracle_step.argtypes = [C.c_int, C.c_int, C.c_int, i32, i32, d, d, d, d, u64, d, d, C.c_int]
        _lib.rome_oracle_sweep_pose3pose3.argtypes = [C.c_int, C.c_int, i32, i32, d, d, d, C.c_int]
        _lib.rome_oracle_sweep_priorpose2.argtypes = [C.c_int, C.c_int, i32, d, d, d, C.c_int]
        _lib.rome_oracle_sweep_priorpose3.argtypes = [C.c_int, C.c_int, i32, d, d, d, C.c_int]
        _lib.rome_oracle_sweep_bearingrange.argtypes = [C.c_int, C.c_int, i32, i32, d, d, d, d, C.c_int]
        _lib.rome_oracle_conv_nm_pose2pose2.argtypes = [C.c_int, C.c_int, i32, i32, d, d, C.c_int, C.c_int,
                                                        C.c_double, u64, d, C.POINTER(u64), C.c_int]
        _lib.rome_oracle_product.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(d), C.c_int, C.c_int, u64, d]
        _lib.rome_oracle_product_sweep.argtypes = [C.c_int, i32, i32, d, C.c_int, C.c_int, C.c_int, C.c_int, u64, d, C.c_int]
        _lib.rome_oracle_philox4x32_10.restype = None
        _lib.rome_oracle_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        _lib.rome_oracle_normal4.restype = None
        _lib.rome_oracle_normal4.argtypes = [u64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, d]
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------------
# single evaluations (== calcFactorResidualTemporary on coordinates)
# ----------------------------------------------------------------------------------
def _call(name, nout, *args):
    arrs = [_f64(a) for a in args]
    out = np.zeros(nout)
    getattr(lib(), "rome_oracle_" + name)(*[_dp(a) for a in arrs], _dp(out))
    return out


def pose2pose2(X, p, q):
    return _call("pose2pose2", 3, X, p, q)


def priorpose2(m, p):
    return _call("priorpose2", 3, m, p)


def bearingrange(meas, p, l):
    return _call("bearingrange", 2, meas, p, l)


def pose3pose3(X, p, q):
    return _call("pose3pose3", 6, X, p, q)


def priorpose3(m, p):
    return _call("priorpose3", 6, m, p)


def priorpoint2(m, x):
    return _call("priorpoint2", 2, m, x)


def point2point2(m, xi, xj):
    return _call("point2point2", 2, m, xi, xj)


def pose2point2(m, p, l):
    return _call("pose2point2", 2, m, p, l)


def range2(rho, xi, l):
    return _call("range2", 1, np.atleast_1d(rho), np.asarray(xi, dtype=float)[:2], l)


def pose2point2bearing(b, p, l):
    return _call("pose2point2bearing", 1, np.atleast_1d(b), p, l)


def priorpoint3(m, x):
    return _call("priorpoint3", 3, m, x)


def point3point3(m, xi, xj):
    return _call("point3point3", 3, m, xi, xj)


def pose3pose3xyyaw(X, p, q):
    return _call("pose3pose3xyyaw", 3, X, p, q)


def pose3pose3rotation(m, p, q):
    return _call("pose3pose3rotation", 3, m, p, q)


def pose3pose3rotoffset(X, p, q, w):
    """src/factors/Pose3Pose3.jl:57-78; w = rotation-vector coordinates of the Rotation3 variable bRa"""
    return _call("pose3pose3rotoffset", 6, X, p, q, w)


def pose3pose3transform(X, p, q, D):
    """src/factors/Pose3Pose3.jl:80-95; D = coordinates of the Pose3 variable Delta"""
    return _call("pose3pose3transform", 6, X, p, q, D)


def pose3pose3unittrans(X, p, q):
    return _call("pose3pose3unittrans", 6, X, p, q)


def pose3_coords(t, R):
    """coordinates (t, rotation vector) of the Pose3 point (t, R)"""
    return _call("pose3_coords", 6, t, np.asarray(R, dtype=np.float64).reshape(9))


def np_priorpoint2(m, x):
    return np.asarray(m, dtype=np.float64) - np.asarray(x, dtype=np.float64)


def np_point2point2(m, xi, xj):
    return np.asarray(m, dtype=np.float64) - (np.asarray(xj, dtype=np.float64) - np.asarray(xi, dtype=np.float64))


def np_pose2point2(m, p, l):
    m, p, l = (np.asarray(a, dtype=np.float64) for a in (m, p, l))
    c, s = np.cos(p[..., 2]), np.sin(p[..., 2])
    return np.stack([l[..., 0] - (p[..., 0] + c * m[..., 0] - s * m[..., 1]),
                     l[..., 1] - (p[..., 1] + s * m[..., 0] + c * m[..., 1])], -1)


def np_range2(rho, xi, l):
    rho, xi, l = (np.asarray(a, dtype=np.float64) for a in (rho, xi, l))
    return (rho[..., 0] - np.hypot(l[..., 0] - xi[..., 0], l[..., 1] - xi[..., 1]))[..., None]


def np_pose2point2bearing(b, p, l):
    b, p, l = (np.asarray(a, dtype=np.float64) for a in (b, p, l))
    c, s = np.cos(p[..., 2]), np.sin(p[..., 2])
    dx, dy = l[..., 0] - p[..., 0], l[..., 1] - p[..., 1]
    return np_sym_rem(b[..., 0] - np.arctan2(-s * dx + c * dy, c * dx + s * dy))[..., None]


def pose2pose2_fwd(X, p):
    return _call("pose2pose2_fwd", 3, X, p)


def pose2pose2_bwd(X, q):
    return _call("pose2pose2_bwd", 3, X, q)


def bearingrange_fwd(meas, p):
    return _call("bearingrange_fwd", 2, meas, p)


def pose3pose3_fwd(X, p):
    return _call("pose3pose3_fwd", 6, X, p)


def pose3pose3_bwd(X, q):
    return _call("pose3pose3_bwd", 6, X, q)


def so3_exp(w):
    return _call("so3_exp", 9, w).reshape(3, 3)


def so3_log(R):
    return _call("so3_log", 3, np.asarray(R).reshape(9))


def sym_rem(x):
    return lib().rome_oracle_sym_rem(float(x))


# ----------------------------------------------------------------------------------
# batched sweeps; arrays in the reference's particle-major layout
#   vars [nvars][N][d], meas [nF][N][dm] -> res [nF][N][dr]
# ----------------------------------------------------------------------------------
def step_prepare(family, i0, i1, v0, v1, mu, Lc, nthreads=0):
    """one STEP of the hot path on the CPU (rome_oracle_step: getSample + residual + per-factor statistics for every
    factor x particle of `family`, 0..4): inputs converted and outputs allocated once; returns (call, res, stats) where
    call(seed) -> threads used"""
    i0, v0, mu, Lc = _i32(i0), _f64(v0), _f64(mu), _f64(Lc)
    i1 = _i32(i1) if i1 is not None else i0
    v1 = _f64(v1) if v1 is not None else v0
    nF, N, dr = len(i0), v0.shape[1], mu.shape[1]
    res, stats = np.empty((nF, N, dr)), np.empty((nF, dr + dr * (dr + 1) // 2))

    def call(seed=0):
        return lib().rome_oracle_step(family, nF, N, _ip(i0), _ip(i1), _dp(v0), _dp(v1), _dp(mu), _dp(Lc), seed, _dp(res),
                                      _dp(stats), nthreads)
    return call, res, stats


def sweep_pose2pose2(ip, iq, poses, meas, nthreads=0):
    ip, iq, poses, meas = _i32(ip), _i32(iq), _f64(poses), _f64(meas)
    nF, N = meas.shape[0], meas.shape[1]
    res = np.empty((nF, N, 3))
    lib().rome_oracle_sweep_pose2pose2(nF, N, _ip(ip), _ip(iq), _dp(poses), _dp(meas), _dp(res), nthreads)
    return res


def sweep_priorpose2(ip, poses, meas, nthreads=0):
    ip, poses, meas = _i32(ip), _f64(poses), _f64(meas)
    nF, N = meas.shape[0], meas.shape[1]
    res = np.empty((nF, N, 3))
    lib().rome_oracle_sweep_priorpose2(nF, N, _ip(ip), _dp(poses), _dp(meas), _dp(res), nthreads)
    return res


def sweep_bearingrange(ip, il, poses, points, meas, nthreads=0):
    ip, il, poses, points, meas = _i32(ip), _i32(il), _f64(poses), _f64(points), _f64(meas)
    nF, N = meas.shape[0], meas.shape[1]
    res = np.empty((nF, N, 2))
    lib().rome_oracle_sweep_bearingrange(nF, N, _ip(ip), _ip(il), _dp(poses), _dp(points), _dp(meas),
                                         _dp(res), nthreads)
    return res


def sweep_pose3pose3(ip, iq, poses, meas, nthreads=0):
    ip, iq, poses, meas = _i32(ip), _i32(iq), _f64(poses), _f64(meas)
    nF, N = meas.shape[0], meas.shape[1]
    res = np.empty((nF, N, 6))
    lib().rome_oracle_sweep_pose3pose3(nF, N, _ip(ip), _ip(iq), _dp(poses), _dp(meas), _dp(res), nthreads)
    return res


def sweep_priorpose3(ip, poses, meas, nthreads=0):
    ip, poses, meas = _i32(ip), _f64(poses), _f64(meas)
    nF, N = meas.shape[0], meas.shape[1]
    res = np.empty((nF, N, 6))
    lib().rome_oracle_sweep_priorpose3(nF, N, _ip(ip), _dp(poses), _dp(meas), _dp(res), nthreads)
    return res


def conv_nm_pose2pose2(ip, iq, poses, meas, fwd=True, inflate_cycles=3, inflation=5.0, seed=0, nthreads=0):
    """Reference-shaped convolution (Nelder-Mead per particle). Returns (proposals, n_evals, threads)."""
    ip, iq, poses, meas = _i32(ip), _i32(iq), _f64(poses), _f64(meas)
    nF, N = meas.shape[0], meas.shape[1]
    out = np.empty((nF, N, 3))
    ne = C.c_uint64(0)
    nt = lib().rome_oracle_conv_nm_pose2pose2(nF, N, _ip(ip), _ip(iq), _dp(poses), _dp(meas), int(fwd),
                                              inflate_cycles, inflation, seed, _dp(out), C.byref(ne), nthreads)
    return out, int(ne.value), nt


# ----------------------------------------------------------------------------------
# sampler twin (device Philox4x32-10 + Box-Muller restated on the host)
# ----------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    u = C.POINTER(C.c_uint32)
    lib().rome_oracle_philox4x32_10(c.ctypes.data_as(u), k.ctypes.data_as(u), o.ctypes.data_as(u))
    return o


def normal4(seed, stream, factor, particle, block=0):
    z = np.zeros(4)
    lib().rome_oracle_normal4(seed, stream, factor, particle, block, _dp(z))
    return z


# ----------------------------------------------------------------------------------
# NumPy twin (vectorised over leading axes) -- an independent second statement of the
# same formulas, used to cross-check the C code and for quick array-level checks.
# ----------------------------------------------------------------------------------
def np_wrap(a):
    return np.arctan2(np.sin(a), np.cos(a))


def np_sym_rem(x):
    x = np.asarray(x, dtype=np.float64)
    r = np.remainder(x + np.pi, 2 * np.pi) - np.pi  # [-pi, pi)
    # IEEE remainder keeps +pi for inputs just below an odd multiple... sym_rem maps x~pi to -pi
    return np.where(np.isclose(x, np.pi, rtol=1.4901161193847656e-08, atol=0), -np.pi, r)


def np_pose2pose2(X, p, q):
    X, p, q = (np.asarray(a, dtype=np.float64) for a in (X, p, q))
    c, s = np.cos(p[..., 2]), np.sin(p[..., 2])
    r = np.empty(np.broadcast_shapes(X.shape, p.shape, q.shape))
    r[..., 0] = p[..., 0] + c * X[..., 0] - s * X[..., 1] - q[..., 0]
    r[..., 1] = p[..., 1] + s * X[..., 0] + c * X[..., 1] - q[..., 1]
    r[..., 2] = np_wrap(p[..., 2] + X[..., 2] - q[..., 2])
    return r


def np_priorpose2(m, p):
    m, p = np.asarray(m, dtype=np.float64), np.asarray(p, dtype=np.float64)
    r = np.empty(np.broadcast_shapes(m.shape, p.shape))
    r[..., :2] = m[..., :2] - p[..., :2]
    r[..., 2] = np_wrap(m[..., 2] - p[..., 2])
    return r


def np_bearingrange(meas, p, l):
    meas, p, l = (np.asarray(a, dtype=np.float64) for a in (meas, p, l))
    c, s = np.cos(p[..., 2]), np.sin(p[..., 2])
    dx, dy = l[..., 0] - p[..., 0], l[..., 1] - p[..., 1]
    plx, ply = c * dx + s * dy, -s * dx + c * dy
    return np.stack([np_sym_rem(meas[..., 0] - np.arctan2(ply, plx)), meas[..., 1] - np.hypot(plx, ply)], -1)


def np_so3_exp(w):
    w = np.asarray(w, dtype=np.float64)
    t = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        a = np.where(t < 1e-6, 1 - t * t / 6, np.sin(t) / t)
        b = np.where(t < 1e-6, 0.5 - t * t / 24, (1 - np.cos(t)) / (t * t))
    return np.eye(3) + a * K + b * (K @ K)


def np_so3_log(R):
    """Quaternion-free log via the reference's generic branch; inputs away from theta=pi."""
    R = np.asarray(R, dtype=np.float64)
    c = np.clip(0.5 * (np.trace(R, axis1=-2, axis2=-1) - 1), -1, 1)
    th = np.arccos(c)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1)
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(th < 1e-8, 0.5, th / (2 * np.sin(th)))
    return k[..., None] * v


def np_pose3pose3(X, p, q):
    X, p, q = (np.asarray(a, dtype=np.float64) for a in (X, p, q))
    Rp, Rq, M = np_so3_exp(p[..., 3:]), np_so3_exp(q[..., 3:]), np_so3_exp(X[..., 3:])
    rt = p[..., :3] + np.einsum("...ij,...j->...i", Rp, X[..., :3]) - q[..., :3]
    U = np.swapaxes(Rq, -1, -2) @ Rp @ M
    return np.concatenate([rt, np_so3_log(U)], -1)


def np_priorpose3(m, p):
    m, p = np.asarray(m, dtype=np.float64), np.asarray(p, dtype=np.float64)
    U = np.swapaxes(np_so3_exp(p[..., 3:]), -1, -2) @ np_so3_exp(m[..., 3:])
    return np.concatenate([m[..., :3] - p[..., :3], np_so3_log(U)], -1)


def np_pose3pose3xyyaw(X, p, q):
    """src/factors/PartialPose3.jl:116-134 (vectorised twin)"""
    X, p, q = (np.asarray(a, dtype=np.float64) for a in (X, p, q))
    Rp, Rq = np_so3_exp(p[..., 3:]), np_so3_exp(q[..., 3:])
    yp, yq = np.arctan2(Rp[..., 1, 0], Rp[..., 0, 0]), np.arctan2(Rq[..., 1, 0], Rq[..., 0, 0])
    p2 = np.stack([p[..., 0], p[..., 1], yp], -1)
    q2 = np.stack([q[..., 0], q[..., 1], yq], -1)
    return np_pose2pose2(X, p2, q2)


def np_pose3pose3rotation(m, p, q):
    m, p, q = (np.asarray(a, dtype=np.float64) for a in (m, p, q))
    U = np.swapaxes(np_so3_exp(p[..., 3:]), -1, -2) @ np_so3_exp(q[..., 3:])
    return m - np_so3_log(U)


def np_pose3pose3rotoffset(X, p, q, w):
    """NumPy twin of src/factors/Pose3Pose3.jl:57-78 (batched)"""
    X, p, q, w = (np.asarray(a, dtype=np.float64) for a in (X, p, q, w))
    Rp, Rq = np_so3_exp(p[..., 3:]), np_so3_exp(q[..., 3:])
    rt = p[..., :3] + np.einsum("...ij,...j->...i", Rp, X[..., :3]) - q[..., :3]
    U = np.swapaxes(Rq, -1, -2) @ Rp @ np_so3_exp(w) @ np_so3_exp(X[..., 3:])
    return np.concatenate([rt, np_so3_log(U)], -1)


def np_pose3pose3transform(X, p, q, D):
    """NumPy twin of src/factors/Pose3Pose3.jl:80-95 (batched)"""
    X, p, q, D = (np.asarray(a, dtype=np.float64) for a in (X, p, q, D))
    Rp, Rq, RD = np_so3_exp(p[..., 3:]), np_so3_exp(q[..., 3:]), np_so3_exp(D[..., 3:])
    lever = D[..., :3] + np.einsum("...ij,...j->...i", RD, X[..., :3])
    rt = p[..., :3] + np.einsum("...ij,...j->...i", Rp, lever) - q[..., :3]
    U = np.swapaxes(Rq, -1, -2) @ Rp @ RD @ np_so3_exp(X[..., 3:])
    return np.concatenate([rt, np_so3_log(U)], -1)


def np_pose3pose3unittrans(X, p, q):
    r = np_pose3pose3(X, p, q)
    r[..., :3] /= np.linalg.norm(r[..., :3], axis=-1, keepdims=True)
    return r


# ----------------------------------------------------------------------------------
# product of proposal KDEs.  PARITY UNPINNED: the reference's product sampler lives in
# ApproxManifoldProducts / KernelDensityEstimate (absent from /root/reference, versions Project.toml:51,67) and is
# stochastic; this is a NumPy restatement of the algorithm the CUDA kernel implements (label Gibbs sampling over
# Gaussian-kernel KDEs with rule-of-thumb bandwidths), checked statistically, never bit for bit.
# ----------------------------------------------------------------------------------
def kde_bandwidth(pts, wrap_dim=None):
    """per-dimension rule-of-thumb bandwidth h = std * (4 / ((d + 2) N))^(1 / (d + 4)); circular std for wrap_dim"""
    pts = np.asarray(pts, dtype=np.float64)
    N, d = pts.shape
    x = pts.copy()
    if wrap_dim is not None:
        m = np.arctan2(np.sin(x[:, wrap_dim]).sum(), np.cos(x[:, wrap_dim]).sum())
        x[:, wrap_dim] = np_wrap(x[:, wrap_dim] - m)
    std = np.sqrt(((x - x.mean(0) * (np.arange(d) != (wrap_dim if wrap_dim is not None else -1))) ** 2).sum(0) / max(N - 1, 1))
    return np.maximum(std * (4.0 / ((d + 2.0) * N)) ** (1.0 / (d + 4.0)), 1e-6)


def product_gibbs(props, n_out, iters=3, wrap_dim=None, seed=0):
    """props: list of [N][d] particle sets (k >= 2).  Returns [n_out][d] samples of the product of their KDEs:
    sources 0 and 1 are sampled exactly from the N^2-component pair mixture, further sources enter conditioned on the
    components already chosen, then `iters` Gibbs sweeps over all labels (k > 2 only) -- the CUDA kernel's algorithm."""
    rng = np.random.default_rng(seed)
    props = [np.asarray(p, dtype=np.float64) for p in props]
    k, (N, d) = len(props), props[0].shape
    w = [1.0 / kde_bandwidth(p, wrap_dim) ** 2 for p in props]  # precisions per source and dimension
    out = np.zeros((n_out, d))

    def wrapdiff(dlt):
        if wrap_dim is not None:
            dlt[..., wrap_dim] = np_wrap(dlt[..., wrap_dim])
        return dlt

    def fuse(sel, upto, skip):
        lam, s, ref = np.zeros(d), np.zeros(d), None
        for j in range(upto):
            if j == skip:
                continue
            x = props[j][sel[j]].copy()
            if wrap_dim is not None:
                if ref is None:
                    ref = x[wrap_dim]
                x[wrap_dim] = ref + np_wrap(x[wrap_dim] - ref)
            lam += w[j]
            s += w[j] * x
        return s / lam, 1.0 / lam

    def draw(j, mu, var):
        q = -0.5 * (wrapdiff(props[j] - mu) ** 2 / (1.0 / w[j] + var)).sum(1)
        return int(np.argmax(q + rng.gumbel(size=N)))

    q01 = -0.5 * (wrapdiff(props[0][:, None, :] - props[1][None, :, :]) ** 2 / (1.0 / w[0] + 1.0 / w[1])).sum(-1)
    lw = q01.max(1) + np.log(np.exp(q01 - q01.max(1, keepdims=True)).sum(1))
    cdf = np.cumsum(np.exp(lw - lw.max()))
    for c in range(n_out):
        sel = np.zeros(k, dtype=int)
        sel[0] = min(int(np.searchsorted(cdf, rng.random() * cdf[-1])), N - 1)
        sel[1] = draw(1, *fuse(sel, 1, -1))
        for j in range(2, k):
            sel[j] = draw(j, *fuse(sel, j, -1))
        if k > 2:
            for _ in range(iters):
                for j in range(k):
                    sel[j] = draw(j, *fuse(sel, k, j))
        mu, var = fuse(sel, k, -1)
        x = mu + np.sqrt(var) * rng.normal(size=d)
        if wrap_dim is not None:
            x[wrap_dim] = np_wrap(x[wrap_dim])
        out[c] = x
    return out


def product_c(props, n_out, iters=2, wrap_dim=None, seed=0):
    """C twin of product_gibbs (oracle/rome_oracle.c rome_oracle_product): same algorithm, own RNG"""
    props = [_f64(p) for p in props]
    k, (N, d) = len(props), props[0].shape
    ptrs = (C.POINTER(C.c_double) * k)(*[_dp(p) for p in props])
    out = np.zeros((n_out, d))
    rc = lib().rome_oracle_product(k, N, d, -1 if wrap_dim is None else wrap_dim, ptrs, n_out, iters, seed, _dp(out))
    if rc != 0:
        raise ValueError("rome_oracle_product: bad arguments")
    return out


def product_sweep_c(var_off, src_row, rows, wrap_dim=None, iters=2, seed=0, nthreads=0):
    """all variables of one type at once (OpenMP over variables): rows [nrows][N][d] -> [nvars][N][d]"""
    var_off, src_row, rows = _i32(var_off), _i32(src_row), _f64(rows)
    _, N, d = rows.shape
    out = np.zeros((len(var_off) - 1, N, d))
    rc = lib().rome_oracle_product_sweep(len(var_off) - 1, _ip(var_off), _ip(src_row), _dp(rows), N, d,
                                         -1 if wrap_dim is None else wrap_dim, iters, seed, _dp(out), nthreads)
    if rc < 0:
        raise ValueError("rome_oracle_product_sweep: bad arguments")
    return out, rc
