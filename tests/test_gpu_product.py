"""The step after the path: product of proposal KDEs on the device and device-resident sweeps.
The product sampler is stochastic (and the reference's lives in absent packages), so the checks are statistical:
analytic Gaussian products, circular headings, multi-modal selection against the NumPy twin of the same algorithm, and
the reference's own acceptance boxes for the Hexagonal graph (test/testHexagonal2D_CliqByCliq.jl:37-79)."""
import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


def _run_product(ctx, vartype, particles, rows, off, sb, sr, seed=1, iters=3, reanchor=False, shift=0, manifold=True):
    """rows: [nrows][N][d] Float64 proposal offsets (from the target's anchor); returns new particles [nvars][N][d]"""
    nv, N, d = particles.shape
    Np = rb.npad(N)
    ctx.set_particles(vartype, particles)
    buf = np.zeros((len(rows), Np, d), np.float32)
    buf[:, :N] = rows
    base = ctx.malloc_device(buf.nbytes + 16)
    ptr = base + shift  # shift = 4: rows that are not 16-byte aligned (scalar staging loads)
    ctx.memcpy_h2d(ptr, buf)
    ctx.set_product_plan(vartype, off, sb, sr)
    bw = ctx.malloc_device(max(1, len(sb)) * d * 4)
    ctx.product(vartype, [ptr], seed=seed, gibbs_iters=iters, reanchor=reanchor, bw_out=bw, manifold=manifold)
    out = ctx.get_particles(vartype)
    h = np.zeros((max(1, len(sb)), d), np.float32)
    ctx.memcpy_d2h(h, bw)
    ctx.free_device(base); ctx.free_device(bw)
    return out, h


def test_product_of_gaussians_point2(ctx):
    rng = np.random.default_rng(0)
    N, nv = 200, 40
    m1, m2, s1, s2 = np.array([1.0, -0.5]), np.array([-0.4, 0.7]), 0.5, 0.8
    parts = np.zeros((nv + 2, N, 2))  # anchors at 0: proposal offsets are absolute coordinates
    parts[nv] = 5.0 + rng.normal(size=(N, 2))      # one source -> adopted
    parts[nv + 1] = -3.0 + rng.normal(size=(N, 2))  # no source -> unchanged
    parts[nv, 0] = 0.0                              # anchor (= first particle) at the origin
    rows = np.concatenate([m1 + s1 * rng.normal(size=(nv, N, 2)), m2 + s2 * rng.normal(size=(nv, N, 2)),
                           7.0 + rng.normal(size=(1, N, 2))])
    off = np.concatenate([2 * np.arange(nv + 1), [2 * nv + 1, 2 * nv + 1]]).astype(np.int32)
    sb = np.zeros(2 * nv + 1, np.int32)
    sr = np.concatenate([np.stack([np.arange(nv), nv + np.arange(nv)], 1).reshape(-1), [2 * nv]]).astype(np.int32)
    before = parts.copy()
    out, h = _run_product(ctx, rb.POINT2, parts, rows, off, sb, sr)
    # bandwidths: rule of thumb from each proposal's own spread
    scale = (4.0 / (4.0 * N)) ** (1.0 / 6.0)
    assert np.allclose(h[0], rows[0].std(0, ddof=1) * scale, rtol=1e-3)
    assert np.allclose(h[1], rows[nv].std(0, ddof=1) * scale, rtol=1e-3)
    v1, v2 = s1 ** 2 * (1 + scale ** 2), s2 ** 2 * (1 + scale ** 2)
    mean = (m1 / v1 + m2 / v2) / (1 / v1 + 1 / v2)
    var = 1.0 / (1 / v1 + 1 / v2)
    got = out[:nv].reshape(-1, 2)
    assert np.allclose(got.mean(0), mean, atol=0.03), (got.mean(0), mean)
    assert np.allclose(got.var(0), var, rtol=0.1), (got.var(0), var)
    # per-variable means scatter like the finite-sample product does, not more
    assert np.abs(out[:nv].mean(1) - mean).max() < 0.45
    assert np.allclose(out[nv], rows[2 * nv], atol=1e-6)          # single proposal adopted
    assert np.allclose(out[nv + 1], before[nv + 1], atol=1e-5)    # no proposal: untouched


def test_product_heading_wraps(ctx):
    """Pose2 headings straddling the +-pi cut relative to the anchor: the circular mean must come out near pi"""
    rng = np.random.default_rng(1)
    N, nv = 128, 16
    parts = np.zeros((nv, N, 3))
    a = np.concatenate([np.pi - 0.05 + 0.1 * rng.normal(size=(nv, N, 1))], 2)
    b = np.concatenate([-np.pi + 0.08 + 0.1 * rng.normal(size=(nv, N, 1))], 2)
    xy1, xy2 = 0.3 * rng.normal(size=(nv, N, 2)) + 1.0, 0.3 * rng.normal(size=(nv, N, 2)) + 1.2
    rows = np.concatenate([np.concatenate([xy1, O.np_wrap(a)], 2), np.concatenate([xy2, O.np_wrap(b)], 2)])
    off = (2 * np.arange(nv + 1)).astype(np.int32)
    sr = np.stack([np.arange(nv), nv + np.arange(nv)], 1).reshape(-1).astype(np.int32)
    out, h = _run_product(ctx, rb.POSE2, parts, rows, off, np.zeros(2 * nv, np.int32), sr)
    th = out[..., 2].reshape(-1)
    circ = np.arctan2(np.sin(th).mean(), np.cos(th).mean())
    assert abs(O.np_wrap(circ - (np.pi + 0.015))) < 0.03, circ
    assert np.abs(O.np_wrap(th - np.pi)).max() < 0.6        # nothing lands on the far side of the circle
    assert np.allclose(out[..., :2].reshape(-1, 2).mean(0), [1.1, 1.1], atol=0.05)
    assert h[:, 2].max() < 0.1                               # circular spread, not the +-pi jump


def test_product_multimodal_matches_numpy_twin(ctx):
    """a bimodal proposal times a unimodal one keeps the shared mode; same behaviour as the NumPy twin"""
    rng = np.random.default_rng(2)
    N, nv = 100, 24
    bim = np.where(rng.random((nv, N, 1)) < 0.5, -2.0, 2.0) + 0.3 * rng.normal(size=(nv, N, 2))
    uni = np.array([1.8, 1.9]) + 0.4 * rng.normal(size=(nv, N, 2))
    rows = np.concatenate([bim, uni])
    off = (2 * np.arange(nv + 1)).astype(np.int32)
    sr = np.stack([np.arange(nv), nv + np.arange(nv)], 1).reshape(-1).astype(np.int32)
    out, _ = _run_product(ctx, rb.POINT2, np.zeros((nv, N, 2)), rows, off, np.zeros(2 * nv, np.int32), sr)
    twin = np.stack([O.product_gibbs([bim[v], uni[v]], N, iters=3, seed=v) for v in range(4)])
    frac_gpu = (out[..., 0] > 0).mean()
    frac_twin = (twin[..., 0] > 0).mean()
    assert frac_gpu > 0.97 and frac_twin > 0.97, (frac_gpu, frac_twin)
    g, t = out.reshape(-1, 2), twin.reshape(-1, 2)
    assert np.allclose(g.mean(0), t.mean(0), atol=0.06), (g.mean(0), t.mean(0))
    assert np.allclose(g.std(0), t.std(0), rtol=0.2), (g.std(0), t.std(0))


def _mixture_moments(rows, h):
    """exact mean / variance per dimension of the product of k KDEs (rows[j]: [N][d] components, h[j]: [d] bandwidths):
    a mixture over all N^k index tuples, component = precision-weighted fusion, weight = the Gaussian overlap"""
    import itertools
    k, (N, d) = len(rows), rows[0].shape
    prec = np.array([1.0 / np.square(h[j].astype(np.float64)) for j in range(k)])     # [k][d]
    lam = prec.sum(0)
    idx = np.array(list(itertools.product(range(N), repeat=k)))                       # [N^k][k]
    x = np.stack([rows[j][idx[:, j]] for j in range(k)], 1)                           # [N^k][k][d]
    m = (x * prec[None]).sum(1) / lam
    logw = (-0.5 * ((x * x * prec[None]).sum(1) - lam * m * m)).sum(1)
    w = np.exp(logw - logw.max()); w /= w.sum()
    mean = (w[:, None] * m).sum(0)
    var = (w[:, None] * (1.0 / lam + m * m)).sum(0) - mean * mean
    return mean, var


@pytest.mark.parametrize("case", ["pair_point2_N10", "unaligned_pair_point2_N11", "pair_pose2_N13", "triple_point2_N6", "far_apart_pose2_N10",
                                  "wide_headings_pose2_N10"])
def test_product_matches_exact_mixture(ctx, case):
    """few components, many variables with the SAME proposals (every variable = fresh chains): the pooled samples have
    the mean and variance of the exact N^k-component product mixture.  Particle counts that are not multiples of four
    exercise the padding of the staged rows; `far_apart` (densities 60 bandwidths apart: every direct weight underflows)
    and `wide_headings` (offsets beyond 1.5 rad) take the log-domain code of the general path."""
    rng = np.random.default_rng(5)
    k = 3 if case.startswith("triple") else 2
    N = int(case.rsplit("N", 1)[1])
    pose = "pose2" in case
    vt, d = (rb.POSE2, 3) if pose else (rb.POINT2, 2)
    nv = 1500
    base = [rng.normal(size=(N, d)) * 0.5 + 0.4 * j for j in range(k)]
    if pose:
        for b in base:
            b[:, 2] *= 0.3
    if case.startswith("far_apart"):
        base[1][:, 0] += 60.0
    if case.startswith("wide_headings"):
        for b in base:
            b[:, 2] += 2.0
    rows = np.concatenate([np.repeat(b[None], nv, 0) for b in base])       # row j * nv + v
    off = (k * np.arange(nv + 1)).astype(np.int32)
    sr = np.stack([j * nv + np.arange(nv) for j in range(k)], 1).reshape(-1).astype(np.int32)
    out, h = _run_product(ctx, vt, np.zeros((nv, N, d)), rows, off, np.zeros(k * nv, np.int32), sr, iters=4,
                          shift=4 if case.startswith("unaligned") else 0)
    assert np.isfinite(out).all()
    mean, var = _mixture_moments(base, [h[j] for j in range(k)])
    got = out.reshape(-1, d)
    se = np.sqrt(var / got.shape[0])
    # pooled samples: 5 standard errors (+ a little slack for the finite Gibbs sweeps of the triple product)
    slack = 0.02 if k > 2 else 0.0
    assert np.all(np.abs(got.mean(0) - mean) < 5 * se + slack * np.sqrt(var)), (got.mean(0), mean, se)
    assert np.allclose(got.var(0), var, rtol=0.08 + 2 * slack), (got.var(0), var)


def test_product_pose3_on_manifold(ctx):
    """ROME_B200_PRODUCT_MANIFOLD: two wide rotation clouds (0.30 / 0.45 rad) around an anchor next to the |w| = pi
    sphere, Gaussian in the tangent space at the anchor rotation.  The product sampled in that tangent space has the
    analytic Gaussian-product moments there (bandwidths are those of the tangent coordinates); multiplying the
    rotation-vector offsets as Euclidean coordinates (flag off) is measurably further from them."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(11)
    N, nv = 100, 60
    w0 = np.array([2.9, 0.3, -0.2])
    Ra = R.from_rotvec(w0)
    m, sg = [np.array([0.25, -0.1, 0.15]), np.array([-0.15, 0.2, 0.05])], [0.30, 0.45]
    tm, ts = [np.array([1.0, -0.5, 0.2]), np.array([0.6, 0.1, -0.3])], [0.5, 0.8]
    parts = np.zeros((nv, N, 6))
    parts[..., 3:] = w0                     # anchor (= first particle) at the rotation w0, translation 0
    rows, xis = [], []
    for j in range(2):
        xi = m[j] + sg[j] * rng.normal(size=(nv, N, 3))
        w = (Ra * R.from_rotvec(xi.reshape(-1, 3))).as_rotvec().reshape(nv, N, 3)
        th = np.linalg.norm(w, axis=-1, keepdims=True)
        alt = w * (1.0 - 2.0 * np.pi / np.maximum(th, 1e-12))   # the other representative of the same rotation
        near = np.linalg.norm(alt - w0, axis=-1, keepdims=True) < np.linalg.norm(w - w0, axis=-1, keepdims=True)
        w = np.where(near, alt, w)
        rows.append(np.concatenate([tm[j] + ts[j] * rng.normal(size=(nv, N, 3)), w - w0], -1))
        xis.append(xi)
    rows = np.concatenate(rows)
    off = (2 * np.arange(nv + 1)).astype(np.int32)
    sr = np.stack([np.arange(nv), nv + np.arange(nv)], 1).reshape(-1).astype(np.int32)
    scale = (4.0 / (8.0 * N)) ** (1.0 / 10.0)

    def tangent_moments(manifold):
        out, h = _run_product(ctx, rb.POSE3, parts.copy(), rows, off, np.zeros(2 * nv, np.int32), sr, manifold=manifold)
        assert np.isfinite(out).all()
        xi = (Ra.inv() * R.from_rotvec(out[..., 3:].reshape(-1, 3))).as_rotvec()
        return out, h, xi.mean(0), xi.var(0)

    out, h, mean_w, var_w = tangent_moments(True)
    v = [np.square(s_) * (1 + scale ** 2) for s_ in sg]
    want_mean = (m[0] / v[0] + m[1] / v[1]) / (1 / v[0] + 1 / v[1])
    want_var = 1.0 / (1 / v[0] + 1 / v[1])
    assert np.allclose(h[0, 3:], xis[0][0].std(0, ddof=1) * scale, rtol=5e-3), (h[0], xis[0][0].std(0, ddof=1) * scale)
    assert np.allclose(mean_w, want_mean, atol=0.03), (mean_w, want_mean)
    assert np.allclose(var_w, want_var, rtol=0.15), (var_w, want_var)
    vt = [np.square(s_) * (1 + scale ** 2) for s_ in ts]
    t = out[..., :3].reshape(-1, 3)
    assert np.allclose(t.mean(0), (tm[0] / vt[0] + tm[1] / vt[1]) / (1 / vt[0] + 1 / vt[1]), atol=0.05)
    assert np.allclose(t.var(0), 1.0 / (1 / vt[0] + 1 / vt[1]), rtol=0.15)
    _, _, mean_e, var_e = tangent_moments(False)
    err_m = np.abs(mean_w - want_mean).max() + np.abs(var_w / want_var - 1).max()
    err_e = np.abs(mean_e - want_mean).max() + np.abs(var_e / want_var - 1).max()
    assert err_m < err_e, (err_m, err_e)


def test_plan_errors(ctx):
    ctx.set_particles(rb.POINT2, np.zeros((2, 16, 2)))
    with pytest.raises(rb.RomeB200Error):  # more sources than ROME_B200_MAX_PRODUCT_SOURCES
        ctx.set_product_plan(rb.POINT2, [0, 40, 40], np.zeros(40, np.int32), np.zeros(40, np.int32))
    ctx.set_product_plan(rb.POINT2, [0, 1, 2], [0, 3], [0, 0])
    p = ctx.malloc_device(4096)
    with pytest.raises(rb.RomeB200Error):  # the plan indexes buffer 3, only one is passed
        ctx.product(rb.POINT2, [p])
    ctx.free_device(p)


def test_hexagonal_solve_reference_boxes(ctx):
    """generateGraph_Hexagonal + device-resident sweeps: the acceptance boxes of the reference's own solve test
    (test/testHexagonal2D_CliqByCliq.jl:37-79: more than 35 of 100 particles inside each box)"""
    fg = rb.generateGraph_Hexagonal()
    rb.initAll(fg, seed=3, ctx=ctx)
    rb.solveGraphGibbs(fg, sweeps=3, seed=5, ctx=ctx)
    boxes = {
        "x0": [(-3, 3), (-3, 3), (-0.3, 0.3)], "x1": [(7, 13), (-3, 3), (0.7, 1.3)],
        "x2": [(12, 18), (6, 11), (1.8, 2.4)], "x4": [(-5, 5), (13, 22), (-2.8, -1.5)],
        "x5": [(-8, -2), (6, 11), (-1.3, -0.7)], "x6": [(-3, 3), (-3, 3), (-0.3, 0.3)],
        "l1": [(17, 23), (-5, 5)],
    }
    for label, bx in boxes.items():
        v = rb.getVal(fg, label)
        assert v.shape[0] == 100
        for c, (lo, hi) in enumerate(bx):
            x = O.np_wrap(v[:, c]) if c == 2 else v[:, c]
            assert ((x > lo) & (x < hi)).sum() > 35, (label, c, x.mean())
    v3 = rb.getVal(fg, "x3")  # :59-64: mean near (11, 17.5), heading near pi, bounded covariance
    assert np.allclose(v3[:, :2].mean(0), [11, 17.5], atol=3.0)
    assert abs(O.np_wrap(np.arctan2(np.sin(v3[:, 2]).mean(), np.cos(v3[:, 2]).mean()) - np.pi)) < 0.5
    assert np.all(v3[:, :2].var(0) < 25)
    # the loop closure through :l1 ties :x6 back to :x0
    assert np.linalg.norm(rb.getVal(fg, "x6")[:, :2].mean(0)) < 1.5


def test_sweeps_on_pose3_chain_and_beehive(ctx):
    """device-resident sweeps over the other BASELINE graph shapes: SE(3) chain with loop closures (rotation-vector
    coordinates treated as Euclidean inside the product) and Beehive (Pose2Pose2 + bearing-range + landmarks):
    beliefs stay centred on the simulated truth and contract instead of diffusing"""
    p3 = rb.generateGraph_Pose3Chain(60, loops=6)
    rb.seed_particles(p3, seed=4, N=100)
    truth = {l: v.simulated.copy() for l, v in p3.variables.items()}
    spread0 = np.mean([v.val[:, :3].std(0).mean() for v in p3.variables.values()])
    rb.solveGraphGibbs(p3, sweeps=3, seed=2, ctx=ctx)
    err = np.array([np.abs(v.val.mean(0)[:3] - truth[l][:3]).max() for l, v in p3.variables.items()])
    # mean rotation error in the tangent space at the truth (rotation vectors do not average across |w| = pi)
    rot = np.array([np.abs(O.np_so3_log(O.np_so3_exp(truth[l][3:]).T @ O.np_so3_exp(v.val[:, 3:])).mean(0)).max()
                    for l, v in p3.variables.items()])
    spread1 = np.mean([v.val[:, :3].std(0).mean() for v in p3.variables.values()])
    assert err.max() < 0.5 and rot.max() < 0.1, (err.max(), rot.max())
    assert spread1 < 1.5 * spread0
    bh = rb.generateGraph_Beehive(30, N=100)
    rb.seed_particles(bh, seed=3, N=100)
    truth = {l: v.simulated.copy() for l, v in bh.variables.items()}
    rb.solveGraphGibbs(bh, sweeps=3, seed=4, ctx=ctx)
    for l, v in bh.variables.items():
        m = v.val.mean(0)
        assert np.abs(m[:2] - truth[l][:2]).max() < 1.5, (l, m, truth[l])
        if v.variableType is rb.Pose2:
            th = np.arctan2(np.sin(v.val[:, 2]).mean(), np.cos(v.val[:, 2]).mean())
            assert abs(O.np_wrap(th - truth[l][2])) < 0.3, (l, th, truth[l][2])
