"""GPU parity of every factor family through the C ABI against the float64 oracle, on random
graphs (seeded), in parity mode (measurement supplied) and fused-sample mode.

Tolerance (north_star: 1e-5 relative on residuals), per residual component:
  (A) same inputs: the oracle is evaluated in Float64 on exactly the values the kernel sees
      (anchor + float32 offset, read back from the device):  |gpu - ref| <= 1e-5 * max(|ref|, 1e-7),
      i.e. PURE relative 1e-5 down to residuals of 1e-7;
  (B) original inputs: the oracle is evaluated on the caller's Float64 arrays before the anchored
      float32 quantisation:  |gpu - ref| <= 1e-5 * max(|ref|, 1e-2)  (relative 1e-5 for residuals above
      1 cm / 0.01 rad, absolute 1e-7 below -- the storage quantisation of a 0.1..1 m offset is 1e-8).
(A) and (B) are the bars of the Float64 chain (ROME_B200_PRECISE; Jacobian / deconvolution outputs always use it).
  (F) the DEFAULT arithmetic -- Float64 per factor, float32 per particle on the small offsets -- is held to
      |gpu - ref| <= 1e-5 * max(|ref|, 1e-1) against BOTH references (same inputs and original inputs): relative 1e-5
      for residuals above 0.1, absolute 1e-6 below.  Its error is 2^-24 x the magnitude of the per-particle terms (the
      offsets and the rotation-induced displacement |heading offset| x |lever arm|, here up to 30 m x 0.08 rad), i.e.
      1e-8 .. 2e-7 on these graphs; SURVEY.md section 7 (hard part 2) defines the parity metric with a floor of 1.
Angular components are compared modulo 2 pi (the +-pi branch cut is a sign choice)."""
import numpy as np
import pytest

import rome_b200 as rb
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL, FLOOR_SAME, FLOOR_ORIG, FLOOR_F32 = 1e-5, 1e-7, 1e-2, 1e-1


def assert_close(gpu, ref, angle_cols=(), what="", floor=FLOOR_ORIG):
    gpu, ref = np.asarray(gpu, float), np.asarray(ref, float)
    d = gpu - ref
    for c in angle_cols:
        d[..., c] = np.abs(O.np_wrap(d[..., c]))
    tol = RTOL * np.maximum(np.abs(ref), floor)
    bad = np.abs(d) > tol
    assert not bad.any(), (f"{what}: {bad.sum()} of {bad.size} off; worst abs {np.abs(d).max():.3e}, "
                           f"worst rel {(np.abs(d) / np.maximum(np.abs(ref), floor)).max():.3e}")


def seen(ctx, vartype, N):
    """the exact Float64 particle values the kernels see (anchor + float32 offset)"""
    return rb.dequantized_particles(ctx.get_anchors(vartype), ctx.get_offsets(vartype), N)


def seen_meas(moff, mu, N):
    return rb.offsets_to_meas(moff, mu, N)


@pytest.fixture(scope="module")
def ctx():
    c = rb.Context(0)
    yield c
    c.close()


def make_pose2_graph(rng, nvars, nF, N, spread=(0.1, 0.12, 0.02), step=10.0):
    """SLAM-shaped: variable means on a random walk (10 m legs, arbitrary headings, |coords| up to a few
    hundred metres), factors between variables at most 3 steps apart (lever arm <= 30 m)."""
    heading = rng.uniform(-np.pi, np.pi, nvars)
    xy = np.cumsum(step * np.column_stack([np.cos(heading), np.sin(heading)]), 0) + rng.uniform(-200, 200, 2)
    mean = np.column_stack([xy, rng.uniform(-np.pi, np.pi, nvars)])
    poses = mean[:, None, :] + rng.normal(size=(nvars, N, 3)) * spread
    ip = rng.integers(0, nvars, nF)
    iq = (ip + rng.integers(1, 4, nF)) % nvars
    far = np.abs(ip - iq) > 3  # wrapped around the end of the walk: re-link to a neighbour
    iq[far] = ip[far] - 1
    return poses, ip.astype(np.int32), iq.astype(np.int32)


def rand_cov(rng, nF, d, scale):
    A = rng.normal(size=(nF, d, d)) * 0.3 + np.eye(d)
    S = A @ np.transpose(A, (0, 2, 1)) * (np.asarray(scale)[:, None] * np.asarray(scale)[None, :])
    return S


@pytest.mark.parametrize("N", [100, 37, 200])
def test_pose2pose2_parity(ctx, N):
    rng = np.random.default_rng(10 + N)
    nvars, nF = 53, 211
    poses, ip, iq = make_pose2_graph(rng, nvars, nF, N)
    # measurement consistent with the graph so residuals are small (the hard case for relative error)
    mu = np.zeros((nF, 3))
    for f in range(nF):
        p, q = poses[ip[f]].mean(0), poses[iq[f]].mean(0)
        c, s = np.cos(p[2]), np.sin(p[2])
        d = q[:2] - p[:2]
        mu[f] = [c * d[0] + s * d[1], -s * d[0] + c * d[1], O.np_wrap(q[2] - p[2])]
    cov = rand_cov(rng, nF, 3, [0.1, 0.1, 0.02])
    meas = mu[:, None, :] + np.einsum("fij,fnj->fni", np.linalg.cholesky(cov), rng.normal(size=(nF, N, 3)))
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2(ip, iq, mu, cov)
    flags = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.PROPOSAL_BWD | rb.STATS | rb.JACOBIAN | rb.PRECISE
    out = ctx.alloc_host_outputs(rb.POSE2POSE2, flags)
    moff = rb.meas_to_offsets(meas, mu)
    ctx.eval_host(rb.POSE2POSE2, flags, meas=moff, **out)
    ref = O.sweep_pose2pose2(ip, iq, poses, meas)
    res = rb.rows_to_particle_major(out["res"], N)
    assert_close(res, ref, angle_cols=(2,), what="pose2pose2 residual (B)")
    ref_same = O.sweep_pose2pose2(ip, iq, seen(ctx, rb.POSE2, N), seen_meas(moff, mu, N))
    assert_close(res, ref_same, angle_cols=(2,), what="pose2pose2 residual (A)", floor=FLOOR_SAME)
    assert np.abs(ref).max() < 5.0  # residuals are small: the comparison is a relative one
    # default arithmetic (float32 per particle): residuals and both proposals against the same references
    f32 = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.PROPOSAL_BWD | rb.STATS
    o32 = ctx.alloc_host_outputs(rb.POSE2POSE2, f32)
    ctx.eval_host(rb.POSE2POSE2, f32, meas=moff, **o32)
    res32 = rb.rows_to_particle_major(o32["res"], N)
    assert_close(res32, ref, angle_cols=(2,), what="pose2pose2 residual (F, original inputs)", floor=FLOOR_F32)
    assert_close(res32, ref_same, angle_cols=(2,), what="pose2pose2 residual (F, same inputs)", floor=FLOOR_F32)
    for key in ("prop_fwd", "prop_bwd"):
        d = rb.rows_to_particle_major(o32[key], N) - rb.rows_to_particle_major(out[key], N)
        d[..., 2] = O.np_wrap(d[..., 2])
        assert np.abs(d).max() < 2e-6, (key, np.abs(d).max())
    assert np.allclose(o32["stats"], out["stats"], rtol=1e-4, atol=1e-4)
    # proposals are roots of the residual
    anchors = ctx.get_anchors(rb.POSE2)
    fwd = rb.rows_to_particle_major(out["prop_fwd"], N) + anchors[iq][:, None, :]
    bwd = rb.rows_to_particle_major(out["prop_bwd"], N) + anchors[ip][:, None, :]
    r_f = O.np_pose2pose2(meas, poses[ip], fwd)
    r_b = O.np_pose2pose2(meas, bwd, poses[iq])
    assert np.abs(r_f).max() < 2e-5 and np.abs(r_b).max() < 2e-5, (np.abs(r_f).max(), np.abs(r_b).max())
    # statistics
    st = out["stats"]
    assert np.allclose(st[:, 0:3], res.sum(1), rtol=1e-4, atol=1e-4)
    rr = np.einsum("fni,fnj->fij", res, res)
    assert np.allclose(st[:, 3:9], np.stack([rr[:, 0, 0], rr[:, 0, 1], rr[:, 0, 2], rr[:, 1, 1], rr[:, 1, 2],
                                             rr[:, 2, 2]], 1), rtol=1e-4, atol=1e-4)
    po = rb.rows_to_particle_major(out["prop_fwd"], N)
    assert np.allclose(st[:, 9:11], po[:, :, :2].sum(1), rtol=1e-4, atol=1e-3)
    assert np.allclose(st[:, 11], np.cos(po[:, :, 2]).sum(1), rtol=1e-4, atol=1e-3)
    assert np.allclose(st[:, 12], np.sin(po[:, :, 2]).sum(1), rtol=1e-4, atol=1e-3)
    # jacobian entries vs finite differences of the oracle
    jac = rb.rows_to_particle_major(out["jac"], N)
    h = 1e-6
    pp = poses[ip].copy(); pp[..., 2] += h
    pm = poses[ip].copy(); pm[..., 2] -= h
    fd = (O.np_pose2pose2(meas, pp, poses[iq]) - O.np_pose2pose2(meas, pm, poses[iq])) / (2 * h)
    assert np.allclose(jac[..., 0], fd[..., 0], atol=1e-4) and np.allclose(jac[..., 1], fd[..., 1], atol=1e-4)
    assert np.allclose(jac[..., 2], np.cos(poses[ip][..., 2]), atol=1e-6)
    assert np.allclose(jac[..., 3], np.sin(poses[ip][..., 2]), atol=1e-6)


def test_priorpose2_parity(ctx):
    rng = np.random.default_rng(3)
    nvars, N = 40, 100
    poses, _, _ = make_pose2_graph(rng, nvars, 1, N)
    ip = np.arange(nvars, dtype=np.int32)
    mu = poses.mean(1)
    cov = rand_cov(rng, nvars, 3, [0.1, 0.1, 0.05])
    meas = mu[:, None, :] + np.einsum("fij,fnj->fni", np.linalg.cholesky(cov), rng.normal(size=(nvars, N, 3)))
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_priorpose2(ip, mu, cov)
    flags = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.STATS | rb.PRECISE
    out = ctx.alloc_host_outputs(rb.PRIORPOSE2, flags)
    moff = rb.meas_to_offsets(meas, mu)
    ctx.eval_host(rb.PRIORPOSE2, flags, meas=moff, **out)
    ref = O.sweep_priorpose2(ip, poses, meas)
    res = rb.rows_to_particle_major(out["res"], N)
    assert_close(res, ref, angle_cols=(2,), what="priorpose2 residual (B)")
    ref_same = O.sweep_priorpose2(ip, seen(ctx, rb.POSE2, N), seen_meas(moff, mu, N))
    assert_close(res, ref_same, angle_cols=(2,), what="priorpose2 residual (A)", floor=FLOOR_SAME)
    o32 = ctx.alloc_host_outputs(rb.PRIORPOSE2, flags & ~rb.PRECISE)
    ctx.eval_host(rb.PRIORPOSE2, flags & ~rb.PRECISE, meas=moff, **o32)
    res32 = rb.rows_to_particle_major(o32["res"], N)
    assert_close(res32, ref, angle_cols=(2,), what="priorpose2 residual (F, original inputs)", floor=FLOOR_F32)
    assert_close(res32, ref_same, angle_cols=(2,), what="priorpose2 residual (F, same inputs)", floor=FLOOR_F32)
    d = rb.rows_to_particle_major(o32["prop_fwd"], N) - rb.rows_to_particle_major(out["prop_fwd"], N)
    assert np.abs(d[..., :2]).max() < 1e-6 and np.abs(O.np_wrap(d[..., 2])).max() < 1e-6
    prop = rb.rows_to_particle_major(out["prop_fwd"], N) + ctx.get_anchors(rb.POSE2)[ip][:, None, :]
    d = prop - meas
    d[..., 2] = O.np_wrap(d[..., 2])
    assert np.abs(d).max() < 1e-5


def test_bearingrange_parity(ctx):
    rng = np.random.default_rng(4)
    nvars, nl, nF, N = 60, 17, 150, 200
    poses, _, _ = make_pose2_graph(rng, nvars, 1, N)
    ip = rng.integers(0, nvars, nF).astype(np.int32)
    il = rng.integers(0, nl, nF).astype(np.int32)
    lm_mean = np.zeros((nl, 2))
    for k in range(nl):  # landmark about 20 m from the first pose that sights it
        src = poses.mean(1)[ip[np.argmax(il == k)] if (il == k).any() else 0]
        lm_mean[k] = src[:2] + rng.normal(size=2) * 14
    points = lm_mean[:, None, :] + rng.normal(size=(nl, N, 2)) * 0.3
    near = np.hypot(*(lm_mean[il] - poses.mean(1)[ip][:, :2]).T) < 60  # keep sightings within 60 m
    ip, il, nF = ip[near], il[near], int(near.sum())
    pm, lmn = poses.mean(1)[ip], lm_mean[il]
    d = lmn - pm[:, :2]
    mu_b = O.np_wrap(np.arctan2(d[:, 1], d[:, 0]) - pm[:, 2])
    mu_r = np.hypot(d[:, 0], d[:, 1])
    bearing = np.column_stack([mu_b, np.full(nF, 0.03)])
    rng_ = np.column_stack([mu_r, np.full(nF, 0.5)])
    meas = np.stack([mu_b[:, None] + 0.03 * rng.normal(size=(nF, N)), mu_r[:, None] + 0.5 * rng.normal(size=(nF, N))], -1)
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_particles(rb.POINT2, points)
    ctx.set_factors_bearingrange(ip, il, bearing, rng_)
    flags = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.STATS | rb.JACOBIAN | rb.PRECISE
    out = ctx.alloc_host_outputs(rb.BEARINGRANGE, flags)
    mu = np.column_stack([mu_b, mu_r])
    moff = rb.meas_to_offsets(meas, mu)
    ctx.eval_host(rb.BEARINGRANGE, flags, meas=moff, **out)
    ref = O.sweep_bearingrange(ip, il, poses, points, meas)
    res = rb.rows_to_particle_major(out["res"], N)
    assert_close(res, ref, angle_cols=(0,), what="bearingrange residual (B)")
    ref_same = O.sweep_bearingrange(ip, il, seen(ctx, rb.POSE2, N), seen(ctx, rb.POINT2, N), seen_meas(moff, mu, N))
    assert_close(res, ref_same, angle_cols=(0,), what="bearingrange residual (A)", floor=FLOOR_SAME)
    f32 = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.PROPOSAL_BWD | rb.STATS
    o32 = ctx.alloc_host_outputs(rb.BEARINGRANGE, f32)
    ctx.eval_host(rb.BEARINGRANGE, f32, meas=moff, **o32)
    o64 = ctx.alloc_host_outputs(rb.BEARINGRANGE, f32 | rb.PRECISE)
    ctx.eval_host(rb.BEARINGRANGE, f32 | rb.PRECISE, meas=moff, **o64)
    res32 = rb.rows_to_particle_major(o32["res"], N)
    assert_close(res32, ref, angle_cols=(0,), what="bearingrange residual (F, original inputs)", floor=FLOOR_F32)
    assert_close(res32, ref_same, angle_cols=(0,), what="bearingrange residual (F, same inputs)", floor=FLOOR_F32)
    for key in ("prop_fwd", "prop_bwd"):
        d = rb.rows_to_particle_major(o32[key], N) - rb.rows_to_particle_major(o64[key], N)
        assert np.abs(d).max() < 5e-6, (key, np.abs(d).max())
    prop = rb.rows_to_particle_major(out["prop_fwd"], N) + ctx.get_anchors(rb.POINT2)[il][:, None, :]
    r_f = O.np_bearingrange(meas, poses[ip], prop)
    assert np.abs(r_f).max() < 2e-5
    jac = rb.rows_to_particle_major(out["jac"], N)
    h = 1e-6
    for k in range(2):
        lp = points[il].copy(); lp[..., k] += h
        lm = points[il].copy(); lm[..., k] -= h
        fd = (O.np_bearingrange(meas, poses[ip], lp) - O.np_bearingrange(meas, poses[ip], lm)) / (2 * h)
        assert np.allclose(jac[..., k], fd[..., 0], atol=1e-5)
        assert np.allclose(jac[..., 2 + k], fd[..., 1], atol=1e-5)


def make_pose3(rng, nvars, N):
    """means on a 3-D random walk with 5 m legs (neighbouring variables are <= 15 m apart)"""
    xyz = np.cumsum(rng.normal(size=(nvars, 3)) * 3.0, 0) + rng.uniform(-100, 100, 3)
    mean = np.column_stack([xyz, rng.normal(size=(nvars, 3)) * 0.9])
    return mean[:, None, :] + rng.normal(size=(nvars, N, 6)) * [0.1, 0.1, 0.1, 0.02, 0.02, 0.02]


def rot_close(w_gpu, w_ref, tol):
    """compare rotation vectors as rotations"""
    d = O.np_so3_log(np.swapaxes(O.np_so3_exp(w_ref), -1, -2) @ O.np_so3_exp(w_gpu))
    return np.abs(d).max() < tol


@pytest.mark.parametrize("N", [100, 64])
def test_pose3pose3_parity(ctx, N):
    rng = np.random.default_rng(5)
    nvars, nF = 31, 97
    poses = make_pose3(rng, nvars, N)
    ip = rng.integers(0, nvars - 3, nF).astype(np.int32)
    iq = (ip + rng.integers(1, 4, nF)).astype(np.int32)
    mu = np.zeros((nF, 6))
    for f in range(nF):  # mean measurement = relative pose of the variable means (so residuals are small)
        p, q = poses[ip[f], 0], poses[iq[f], 0]
        Rp, Rq = O.so3_exp(p[3:]), O.so3_exp(q[3:])
        mu[f, :3] = Rp.T @ (q[:3] - p[:3])
        mu[f, 3:] = O.so3_log(Rp.T @ Rq)
    cov = rand_cov(rng, nF, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
    meas = mu[:, None, :] + np.einsum("fij,fnj->fni", np.linalg.cholesky(cov), rng.normal(size=(nF, N, 6)))
    ctx.set_particles(rb.POSE3, poses)
    ctx.set_factors_pose3pose3(ip, iq, mu, cov)
    flags = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.PROPOSAL_BWD | rb.STATS
    out = ctx.alloc_host_outputs(rb.POSE3POSE3, flags)
    moff = rb.meas_to_offsets(meas, mu)
    ctx.eval_host(rb.POSE3POSE3, flags, meas=moff, **out)
    ref = O.sweep_pose3pose3(ip, iq, poses, meas)
    res = rb.rows_to_particle_major(out["res"], N)
    assert_close(res, ref, what="pose3pose3 residual (B)")
    ref_same = O.sweep_pose3pose3(ip, iq, seen(ctx, rb.POSE3, N), seen_meas(moff, mu, N))
    assert_close(res, ref_same, what="pose3pose3 residual (A)", floor=FLOOR_SAME)
    anchors = ctx.get_anchors(rb.POSE3)
    fwd = rb.rows_to_particle_major(out["prop_fwd"], N) + anchors[iq][:, None, :]
    bwd = rb.rows_to_particle_major(out["prop_bwd"], N) + anchors[ip][:, None, :]
    assert np.abs(O.np_pose3pose3(meas, poses[ip], fwd)).max() < 5e-5
    assert np.abs(O.np_pose3pose3(meas, bwd, poses[iq])).max() < 5e-5
    st = out["stats"]
    assert np.allclose(st[:, :6], res.sum(1), rtol=1e-4, atol=1e-4)
    assert np.allclose(st[:, 31], (res ** 2).sum((1, 2)), rtol=1e-4, atol=1e-4)


def test_priorpose3_parity(ctx):
    rng = np.random.default_rng(6)
    nvars, N = 20, 100
    poses = make_pose3(rng, nvars, N)
    ip = np.arange(nvars, dtype=np.int32)
    mu = poses[:, 0, :].copy()
    cov = rand_cov(rng, nvars, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
    meas = mu[:, None, :] + np.einsum("fij,fnj->fni", np.linalg.cholesky(cov), rng.normal(size=(nvars, N, 6)))
    ctx.set_particles(rb.POSE3, poses)
    ctx.set_factors_priorpose3(ip, mu, cov)
    flags = rb.RESIDUAL | rb.PROPOSAL_FWD | rb.STATS
    out = ctx.alloc_host_outputs(rb.PRIORPOSE3, flags)
    moff = rb.meas_to_offsets(meas, mu)
    ctx.eval_host(rb.PRIORPOSE3, flags, meas=moff, **out)
    ref = O.sweep_priorpose3(ip, poses, meas)
    res = rb.rows_to_particle_major(out["res"], N)
    assert_close(res, ref, what="priorpose3 residual (B)")
    ref_same = O.sweep_priorpose3(ip, seen(ctx, rb.POSE3, N), seen_meas(moff, mu, N))
    assert_close(res, ref_same, what="priorpose3 residual (A)", floor=FLOOR_SAME)
    prop = rb.rows_to_particle_major(out["prop_fwd"], N) + ctx.get_anchors(rb.POSE3)[ip][:, None, :]
    assert np.abs(prop - meas).max() < 1e-5


def twin_normals(seed, sid, f, n, d):
    """host twin of the device sampler's addressing (csrc/factor_kernels.cu): particle n is slot n>>5 of lane n&31.
    SE(2) families (d = 2, 3): one batch of d Philox blocks serves the 4 slots of a group; Pose3Pose3 / PriorPose3
    (d = 6): three blocks serve the lane's pair of slots (2 it, 2 it + 1) of iteration it -- 12 normals, none wasted."""
    lane, slot = n & 31, n >> 5
    if d == 6:
        it, j = slot >> 1, slot & 1
        z = np.concatenate([O.normal4(seed, sid, f, lane, 3 * it + b) for b in range(3)])
        return z[6 * j:6 * j + 6]
    g, k = slot >> 2, slot & 3
    z = np.concatenate([O.normal4(seed, sid, f, lane, g * d + b) for b in range(d)])
    return z[d * k:d * k + d]


def test_particle_roundtrip(ctx):
    rng = np.random.default_rng(7)
    poses, _, _ = make_pose2_graph(rng, 25, 1, 100)
    poses[..., 2] = O.np_wrap(poses[..., 2])
    ctx.set_particles(rb.POSE2, poses)
    back = ctx.get_particles(rb.POSE2)
    d = back - poses
    d[..., 2] = O.np_wrap(d[..., 2])
    assert np.abs(d).max() < 1e-6
    p3 = make_pose3(rng, 9, 100)
    ctx.set_particles(rb.POSE3, p3)
    back3 = ctx.get_particles(rb.POSE3)
    assert np.abs(back3[..., :3] - p3[..., :3]).max() < 1e-6
    assert rot_close(back3[..., 3:], p3[..., 3:], 1e-6)                 # the same rotations ...
    assert np.linalg.norm(back3[..., 3:], axis=-1).max() <= np.pi + 1e-9  # ... reported as principal rotation vectors


@pytest.mark.parametrize("family", ["pose2pose2", "priorpose2", "bearingrange", "pose3pose3", "priorpose3"])
def test_fused_sampling_matches_supplied_and_host_twin(ctx, family):
    """SAMPLE|WRITE_MEAS: (1) residuals are bit-identical to a second call that is handed the written
    samples; (2) the samples reproduce the host Philox/Box-Muller twin; (3) moments match (mu, Sigma)."""
    rng = np.random.default_rng(8)
    N, seed, sid = 100, 0x1234567811, 3
    if family in ("pose2pose2", "priorpose2", "bearingrange"):
        poses, ip, iq = make_pose2_graph(rng, 30, 64, N)
        ctx.set_particles(rb.POSE2, poses)
    if family == "pose2pose2":
        fam, d = rb.POSE2POSE2, 3
        mu = rng.normal(size=(64, 3)) * [5, 5, 1]
        cov = rand_cov(rng, 64, 3, [0.1, 0.2, 0.05])
        ctx.set_factors_pose2pose2(ip, iq, mu, cov)
        Lc = np.linalg.cholesky(cov)
    elif family == "priorpose2":
        fam, d = rb.PRIORPOSE2, 3
        mu = rng.normal(size=(64, 3)) * [5, 5, 1]
        cov = rand_cov(rng, 64, 3, [0.1, 0.2, 0.05])
        ctx.set_factors_priorpose2(ip, mu, cov)
        Lc = np.linalg.cholesky(cov)
    elif family == "bearingrange":
        fam, d = rb.BEARINGRANGE, 2
        points = rng.uniform(-50, 50, (8, 1, 2)) + rng.normal(size=(8, N, 2)) * 0.2
        ctx.set_particles(rb.POINT2, points)
        il = rng.integers(0, 8, 64).astype(np.int32)
        mu = np.column_stack([rng.uniform(-3, 3, 64), rng.uniform(5, 30, 64)])
        sig = np.column_stack([rng.uniform(0.01, 0.1, 64), rng.uniform(0.1, 1, 64)])
        ctx.set_factors_bearingrange(ip, il, np.column_stack([mu[:, 0], sig[:, 0]]), np.column_stack([mu[:, 1], sig[:, 1]]))
        Lc = np.stack([np.diag(s) for s in sig])
    else:
        poses = make_pose3(rng, 30, N)
        ctx.set_particles(rb.POSE3, poses)
        ip = rng.integers(0, 30, 64).astype(np.int32)
        iq = ((ip + 1) % 30).astype(np.int32)
        d = 6
        mu = rng.normal(size=(64, 6)) * [1, 1, 1, .2, .2, .2]
        cov = rand_cov(rng, 64, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
        Lc = np.linalg.cholesky(cov)
        if family == "pose3pose3":
            fam = rb.POSE3POSE3
            ctx.set_factors_pose3pose3(ip, iq, mu, cov)
        else:
            fam = rb.PRIORPOSE3
            ctx.set_factors_priorpose3(ip, mu, cov)
    flags = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
    out = ctx.alloc_host_outputs(fam, flags)
    ctx.eval_host(fam, flags, seed=seed, stream_id=sid, **out)
    out2 = ctx.alloc_host_outputs(fam, rb.RESIDUAL)
    ctx.eval_host(fam, rb.RESIDUAL, meas=out["meas_out"], **out2)
    assert np.array_equal(out["res"][:, :N], out2["res"][:, :N])
    delta = rb.rows_to_particle_major(out["meas_out"], N)  # [nF][N][d] offsets from mu
    # host twin for a few (factor, particle) pairs
    for f in (0, 7, 63):
        for n in (0, 1, 50, 99):
            z = twin_normals(seed, sid, f, n, d)
            want = Lc[f].astype(np.float32).astype(np.float64) @ z
            # the device draws on the special-function unit (MUFU log/sin/cos): ~1e-4 sigma from the Float64 twin
            assert np.allclose(delta[f, n], want, rtol=1e-3, atol=3e-4 * np.abs(Lc[f]).max()), (f, n, delta[f, n], want)
    # moments: whiten and check identity covariance / zero mean over all factors x particles
    zw = np.linalg.solve(Lc, np.transpose(delta, (0, 2, 1)))  # [nF][d][N]
    zs = np.transpose(zw, (1, 0, 2)).reshape(d, -1)
    assert np.abs(zs.mean(1)).max() < 0.05
    assert np.abs(np.cov(zs) - np.eye(d)).max() < 0.06
    # a different stream id gives different draws
    out3 = ctx.alloc_host_outputs(fam, flags)
    ctx.eval_host(fam, flags, seed=seed, stream_id=sid + 1, **out3)
    assert not np.array_equal(out3["meas_out"], out["meas_out"])
    # the compile-time flag variants the timed loops run (SAMPLE|RESIDUAL|STATS [|PROPOSAL_FWD]) draw the SAME samples and
    # compute the SAME rows as the run-time-flag variant checked above: bit-identical residuals, equal statistics
    for hot in (rb.RESIDUAL | rb.STATS | rb.SAMPLE, rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD | rb.SAMPLE):
        oh = ctx.alloc_host_outputs(fam, hot)
        ctx.eval_host(fam, hot, seed=seed, stream_id=sid, **oh)
        assert np.array_equal(oh["res"][:, :N], out["res"][:, :N]), hex(hot)
        assert np.allclose(oh["stats"][:, :out["res"].shape[2]], rb.rows_to_particle_major(out["res"], N).sum(1), rtol=1e-3, atol=1e-3)


def test_device_pointer_path_and_graph(ctx):
    """eval() with device buffers (torch tensors) == eval_host(); factor sub-ranges write disjoint slices;
    a captured CUDA graph replays the same result."""
    import torch
    rng = np.random.default_rng(9)
    N, nF = 100, 1000
    poses, ip, iq = make_pose2_graph(rng, 200, nF, N)
    mu = rng.normal(size=(nF, 3))
    cov = rand_cov(rng, nF, 3, [0.1, 0.1, 0.02])
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2(ip, iq, mu, cov)
    Np = rb.npad(N)
    moff = rb.meas_to_offsets(mu[:, None, :] + rng.normal(size=(nF, N, 3)) * 0.1, mu)
    out = ctx.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS)
    ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS, meas=moff, **out)
    ctx.use_torch_stream()
    try:
        dm = torch.from_numpy(moff).cuda()
        res = torch.zeros((nF, Np, 3), device="cuda")
        st = torch.zeros((nF, 16), device="cuda")
        ctx.eval(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS, first=0, count=300, meas=dm, res=res, stats=st)
        ctx.eval(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS, first=300, count=-1, meas=dm, res=res, stats=st)
        torch.cuda.synchronize()
        assert np.array_equal(res.cpu().numpy(), out["res"])
        assert np.array_equal(st.cpu().numpy(), out["stats"])
        res.zero_()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ctx.use_torch_stream()
            ctx.graph_begin()
            ctx.eval(rb.POSE2POSE2, rb.RESIDUAL | rb.STATS, meas=dm, res=res, stats=st)
            g = ctx.graph_end()
            before = ctx.launch_count
            ctx.graph_launch(g)
            s.synchronize()
            assert ctx.launch_count == before + 1
        assert np.array_equal(res.cpu().numpy(), out["res"])
    finally:
        ctx.set_stream(None)


def test_error_behaviour(ctx):
    with pytest.raises(rb.RomeB200Error):
        ctx.set_factors_pose2pose2([0], [1], [[0, 0, 0]], [np.zeros((3, 3))])  # not positive definite
    c2 = rb.Context(0)
    try:
        with pytest.raises(rb.RomeB200Error) as ei:
            c2.eval_host(rb.POSE2POSE2, rb.RESIDUAL, meas=np.zeros(1, np.float32), res=np.zeros(1, np.float32))
        assert ei.value.code == -4  # NOT_SET
        c2.set_particles(rb.POSE2, np.zeros((2, 10, 3)))
        c2.set_factors_pose2pose2([0], [5], [[0, 0, 0]], [np.eye(3)])
        with pytest.raises(rb.RomeB200Error) as ei:
            c2.eval_host(rb.POSE2POSE2, rb.RESIDUAL, meas=np.zeros(3 * 16, np.float32), res=np.zeros(3 * 16, np.float32))
        assert ei.value.code == -3  # SHAPE_MISMATCH: variable index beyond the store
        with pytest.raises(rb.RomeB200Error):
            c2.eval_host(rb.PRIORPOSE2, rb.PROPOSAL_BWD, prop_bwd=np.zeros(1, np.float32))
    finally:
        c2.close()


def test_nan_propagates(ctx):
    poses = np.zeros((2, 16, 3))
    poses[1, 3, 0] = np.nan
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2([0], [1], [[1, 0, 0]], [np.eye(3) * 0.01])
    out = ctx.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL)
    ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL, meas=np.zeros((1, 16, 3), np.float32), **out)
    assert np.isnan(out["res"][0, 3, 0]) and np.isfinite(out["res"][0, 2, 0])


def test_wide_heading_spread_takes_general_sincos_path(ctx):
    """heading offsets from the anchor beyond the small-angle range (|delta| > 0.78 rad, up to +-pi): the kernel
    must fall back to the general sin/cos; uniform headings on the whole circle"""
    rng = np.random.default_rng(21)
    N, nvars, nF = 100, 12, 40
    poses = np.column_stack([rng.uniform(-5, 5, (nvars * N, 2)), rng.uniform(-np.pi, np.pi, nvars * N)]).reshape(nvars, N, 3)
    ip = rng.integers(0, nvars, nF).astype(np.int32)
    iq = ((ip + 1) % nvars).astype(np.int32)
    mu = rng.normal(size=(nF, 3)) * [3, 3, 1]
    cov = rand_cov(rng, nF, 3, [0.1, 0.1, 0.05])
    meas = mu[:, None, :] + rng.normal(size=(nF, N, 3)) * 0.1
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2(ip, iq, mu, cov)
    for flags in (rb.RESIDUAL | rb.STATS, rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD | rb.PROPOSAL_BWD):
        out = ctx.alloc_host_outputs(rb.POSE2POSE2, flags)
        moff = rb.meas_to_offsets(meas, mu)
        ctx.eval_host(rb.POSE2POSE2, flags, meas=moff, **out)
        ref_same = O.sweep_pose2pose2(ip, iq, seen(ctx, rb.POSE2, N), seen_meas(moff, mu, N))
        assert_close(rb.rows_to_particle_major(out["res"], N), ref_same, angle_cols=(2,), what="wide headings (A)",
                     floor=FLOOR_SAME)


@pytest.mark.parametrize("N", [1, 8, 33, 129, 500, 2000])
def test_particle_count_edge_cases(ctx, N):
    """N = 1 (calcFactorResidualTemporary), exact multiples of 8/32, partial trailing groups, and N large enough to
    switch the kernels to the 2- and 1-factor tiles (shared-memory budget)"""
    rng = np.random.default_rng(100 + N)
    nvars, nF = 9, 21
    poses, ip, iq = make_pose2_graph(rng, nvars, nF, N)
    mu = rng.normal(size=(nF, 3))
    cov = rand_cov(rng, nF, 3, [0.1, 0.1, 0.02])
    meas = mu[:, None, :] + rng.normal(size=(nF, N, 3)) * 0.1
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2(ip, iq, mu, cov)
    moff = rb.meas_to_offsets(meas, mu)
    for flags in (rb.RESIDUAL | rb.STATS, rb.RESIDUAL | rb.STATS | rb.PROPOSAL_FWD, rb.RESIDUAL | rb.JACOBIAN):
        out = ctx.alloc_host_outputs(rb.POSE2POSE2, flags)
        ctx.eval_host(rb.POSE2POSE2, flags, meas=moff, **out)
        res = rb.rows_to_particle_major(out["res"], N)
        ref_same = O.sweep_pose2pose2(ip, iq, seen(ctx, rb.POSE2, N), seen_meas(moff, mu, N))
        f64_only = bool(flags & rb.JACOBIAN)
        assert_close(res, ref_same, angle_cols=(2,), what=f"N={N} flags={flags}", floor=FLOOR_SAME if f64_only else FLOOR_F32)
        if not f64_only:
            o64 = ctx.alloc_host_outputs(rb.POSE2POSE2, flags | rb.PRECISE)
            ctx.eval_host(rb.POSE2POSE2, flags | rb.PRECISE, meas=moff, **o64)
            assert_close(rb.rows_to_particle_major(o64["res"], N), ref_same, angle_cols=(2,), what=f"N={N} flags={flags} PRECISE",
                         floor=FLOOR_SAME)
        if flags & rb.STATS:
            assert np.allclose(out["stats"][:, 0:3], res.sum(1), rtol=1e-3, atol=1e-3)
    # sampled run == supplied run on the written-back samples, for every tile variant
    fl = rb.RESIDUAL | rb.SAMPLE | rb.WRITE_MEAS
    o1 = ctx.alloc_host_outputs(rb.POSE2POSE2, fl)
    ctx.eval_host(rb.POSE2POSE2, fl, seed=3, **o1)
    o2 = ctx.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL)
    ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL, meas=o1["meas_out"], **o2)
    assert np.array_equal(o1["res"][:, :N], o2["res"][:, :N])


def test_too_many_particles_is_an_error(ctx):
    ctx.set_particles(rb.POSE2, np.zeros((2, 6000, 3)))
    ctx.set_factors_pose2pose2([0], [1], [[0, 0, 0]], [np.eye(3)])
    out = ctx.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE)
    with pytest.raises(rb.RomeB200Error) as ei:
        ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE, **out)
    assert ei.value.code == -3


def test_empty_range_and_zero_factors(ctx):
    rng = np.random.default_rng(5)
    poses, ip, iq = make_pose2_graph(rng, 5, 7, 16)
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2(ip, iq, np.zeros((7, 3)), np.tile(np.eye(3), (7, 1, 1)))
    out = ctx.alloc_host_outputs(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE)
    before = ctx.launch_count
    ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE, first=3, count=0, **out)  # nothing to do, no launch
    assert ctx.launch_count == before and not out["res"].any()
    ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE, first=6, count=1, **out)
    assert out["res"][6].any() and not out["res"][:6].any()
    with pytest.raises(rb.RomeB200Error):
        ctx.eval_host(rb.POSE2POSE2, rb.RESIDUAL | rb.SAMPLE, first=5, count=3, **out)


def test_independent_flag_gives_identical_results(ctx):
    """ROME_B200_INDEPENDENT only changes scheduling (programmatic dependent launch)"""
    import torch
    rng = np.random.default_rng(31)
    N, nF = 100, 3000
    poses, ip, iq = make_pose2_graph(rng, 400, nF, N)
    mu = rng.normal(size=(nF, 3))
    ctx.set_particles(rb.POSE2, poses)
    ctx.set_factors_pose2pose2(ip, iq, mu, rand_cov(rng, nF, 3, [0.1, 0.1, 0.02]))
    Np = rb.npad(N)
    ctx.use_torch_stream()
    try:
        bufs = [dict(res=torch.zeros((nF, Np, 3), device="cuda"), stats=torch.zeros((nF, 16), device="cuda"))
                for _ in range(4)]
        fl = rb.RESIDUAL | rb.STATS | rb.SAMPLE
        ctx.eval(rb.POSE2POSE2, fl, seed=4, stream_id=0, **bufs[0])
        for b in bufs[1:]:
            ctx.eval(rb.POSE2POSE2, fl | rb.INDEPENDENT, seed=4, stream_id=0, **b)
        torch.cuda.synchronize()
        for b in bufs[1:]:
            assert torch.equal(b["res"], bufs[0]["res"]) and torch.equal(b["stats"], bufs[0]["stats"])
    finally:
        ctx.set_stream(None)


def _right_perturb(poses, delta):
    """coordinates of (t, R Exp(delta)) for Pose3 coordinates (t, rotation vector)"""
    R = O.np_so3_exp(poses[..., 3:]) @ O.np_so3_exp(np.broadcast_to(delta, poses[..., 3:].shape))
    return np.concatenate([poses[..., :3], O.np_so3_log(R)], -1)


def test_se3_jacobians_match_finite_differences_of_the_oracle(ctx):
    """analytic Jacobian blocks of Pose3Pose3 / PriorPose3 (right perturbations R <- R Exp(delta))
    against central finite differences of the float64 oracle"""
    rng = np.random.default_rng(15)
    nvars, nF, N = 12, 30, 16
    poses = make_pose3(rng, nvars, N)
    ip = rng.integers(0, nvars - 2, nF).astype(np.int32)
    iq = (ip + rng.integers(1, 3, nF)).astype(np.int32)
    mu = np.zeros((nF, 6))
    for f in range(nF):
        p, q = poses[ip[f], 0], poses[iq[f], 0]
        Rp = O.so3_exp(p[3:])
        mu[f, :3] = Rp.T @ (q[:3] - p[:3])
        mu[f, 3:] = O.so3_log(Rp.T @ O.so3_exp(q[3:])) + rng.normal(size=3) * 0.3   # residual rotations up to ~1 rad
    cov = rand_cov(rng, nF, 6, [0.1, 0.1, 0.1, 0.01, 0.01, 0.01])
    meas = mu[:, None, :] + np.einsum("fij,fnj->fni", np.linalg.cholesky(cov), rng.normal(size=(nF, N, 6)))
    ctx.set_particles(rb.POSE3, poses)
    ctx.set_factors_pose3pose3(ip, iq, mu, cov)
    flags = rb.RESIDUAL | rb.JACOBIAN
    out = ctx.alloc_host_outputs(rb.POSE3POSE3, flags)
    assert out["jac"].shape[-1] == 36
    moff = rb.meas_to_offsets(meas, mu)
    ctx.eval_host(rb.POSE3POSE3, flags, meas=moff, **out)
    P, Q = seen(ctx, rb.POSE3, N)[ip], seen(ctx, rb.POSE3, N)[iq]
    M = seen_meas(moff, mu, N)
    J = out["jac"][:, :N].reshape(nF, N, 4, 3, 3).astype(np.float64)
    h = 1e-6
    for k in range(3):
        d = np.zeros(3)
        d[k] = h
        fd_p = (O.np_pose3pose3(M, _right_perturb(P, d), Q) - O.np_pose3pose3(M, _right_perturb(P, -d), Q)) / (2 * h)
        fd_q = (O.np_pose3pose3(M, P, _right_perturb(Q, d)) - O.np_pose3pose3(M, P, _right_perturb(Q, -d))) / (2 * h)
        assert np.abs(fd_p[..., :3] - J[:, :, 0, :, k]).max() < 2e-5      # A = d r_t / d delta_p
        assert np.abs(fd_p[..., 3:] - J[:, :, 1, :, k]).max() < 2e-5      # B = d r_w / d delta_p
        assert np.abs(fd_q[..., 3:] - J[:, :, 2, :, k]).max() < 2e-5      # C = d r_w / d delta_q
        assert np.abs(fd_q[..., :3]).max() < 1e-6                          # d r_t / d delta_q = 0
        Mp, Mm_ = M.copy(), M.copy()
        Mp[..., k] += h
        Mm_[..., k] -= h
        fd_m = (O.np_pose3pose3(Mp, P, Q) - O.np_pose3pose3(Mm_, P, Q)) / (2 * h)
        assert np.abs(fd_m[..., :3] - J[:, :, 3, :, k]).max() < 2e-5      # R_p = d r_t / d m_t
    # PriorPose3
    ctx.set_factors_priorpose3(ip, poses[ip, 0] + rng.normal(size=(nF, 6)) * [0.1, 0.1, 0.1, 0.3, 0.3, 0.3], cov)
    mu_p = poses[ip, 0]
    outp = ctx.alloc_host_outputs(rb.PRIORPOSE3, flags)
    assert outp["jac"].shape[-1] == 9
    measp = mu_p[:, None, :] + rng.normal(size=(nF, N, 6)) * [0.1, 0.1, 0.1, 0.3, 0.3, 0.3]
    ctx.set_factors_priorpose3(ip, mu_p, cov)
    moffp = rb.meas_to_offsets(measp, mu_p)
    ctx.eval_host(rb.PRIORPOSE3, flags, meas=moffp, **outp)
    Mp_ = seen_meas(moffp, mu_p, N)
    Jp = outp["jac"][:, :N].reshape(nF, N, 3, 3).astype(np.float64)
    for k in range(3):
        d = np.zeros(3)
        d[k] = h
        fd = (O.np_priorpose3(Mp_, _right_perturb(P, d)) - O.np_priorpose3(Mp_, _right_perturb(P, -d))) / (2 * h)
        assert np.abs(fd[..., 3:] - Jp[:, :, :, k]).max() < 2e-5
