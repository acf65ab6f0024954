/*
 * rome_oracle.c -- float64 CPU restatement of the RoME.jl factor-residual hot path.
 * TEST INFRASTRUCTURE ONLY (see rome_oracle.h).  Parity status: PINNED against the
 * reference's own known-answer tests (tests/golden/known_answers.json).
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  Arithmetic that lives in unvendored dependencies is restated
 * from their published algorithm and named where used:
 *   Manifolds.jl 0.10  (Project.toml:69)  exp/log/compose/hat/vee/sym_rem/usinc_from_cos
 *   Optim.jl 0.22/1    (Project.toml:71)  NelderMead (adaptive parameters, affine simplexer)
 */
#include "rome_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ================================================================================== */
/* scalar helpers                                                                     */
/* ================================================================================== */

/* Manifolds.sym_rem(x, T=pi) [Manifolds.jl utils]: (x ~ T ? -T : rem(x, 2T, RoundNearest)).
 * Used by src/factors/BearingRange2D.jl:61. isapprox default rtol = sqrt(eps). */
double rome_oracle_sym_rem(double x) {
    const double T = M_PI;
    if (fabs(x - T) <= 1.4901161193847656e-08 * fmax(fabs(x), T)) return -T;
    return remainder(x, 2.0 * T); /* IEEE remainder == rem(.., RoundNearest) */
}

/* so(2) log as the reference evaluates it: atan(U21, U11) of the relative rotation
 * (Manifolds log on SpecialOrthogonal(2); see src/factors/Pose2D.jl:64). */
double rome_oracle_wrap_atan(double a) { return atan2(sin(a), cos(a)); }

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void mat3_tmul(const double A[9], const double B[9], double C[9]) { /* A' * B */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
static void mat3_vec(const double A[9], const double v[3], double o[3]) {
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

/* Exp on SO(3) (Rodrigues), Manifolds.jl exp for Rotations(3):
 *   R = I + sin(t)/t K + (1-cos t)/t^2 K^2,  K = hat(w).  */
void rome_oracle_so3_exp(const double w[3], double R[9]) {
    const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const double t = sqrt(t2);
    double a, b; /* sin t / t, (1 - cos t)/t^2 */
    if (t < 1e-6) {
        a = 1.0 - t2 / 6.0 + t2 * t2 / 120.0;
        b = 0.5 - t2 / 24.0 + t2 * t2 / 720.0;
    } else {
        a = sin(t) / t;
        const double sh = sin(0.5 * t);
        b = 2.0 * sh * sh / t2;
    }
    const double x = w[0], y = w[1], z = w[2];
    R[0] = 1.0 - b * (y * y + z * z); R[1] = -a * z + b * x * y;       R[2] = a * y + b * x * z;
    R[3] = a * z + b * x * y;         R[4] = 1.0 - b * (x * x + z * z); R[5] = -a * x + b * y * z;
    R[6] = -a * y + b * x * z;        R[7] = a * x + b * y * z;         R[8] = 1.0 - b * (x * x + y * y);
}

/* Log on SO(3), restating Manifolds.jl log for Rotations(3):
 *   cos t = (tr R - 1)/2;  generic: X = (R - R')/(2 usinc_from_cos(cos t));
 *   cos t ~ -1 (isapprox, rtol sqrt(eps)): pi * (eigenvector of eigenvalue 1).
 * The pi-branch sign is ambiguous in the reference too (compare as rotations there);
 * here the axis is taken from the dominant column of (R + I)/2 and its sign fixed so
 * it agrees with vee(R - R') whenever that is non-zero. */
void rome_oracle_so3_log(const double R[9], double w[3]) {
    double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    const double v[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]}; /* (X32, X13, X21) * 2 sin t */
    if (fabs(c + 1.0) <= 1.4901161193847656e-08 * fmax(fabs(c), 1.0)) {
        double B[9];
        for (int i = 0; i < 9; ++i) B[i] = 0.5 * R[i];
        B[0] += 0.5; B[4] += 0.5; B[8] += 0.5; /* (R+I)/2 = a a' at t = pi */
        int k = 0;
        if (B[4] > B[0]) k = 1;
        if (B[8] > B[4 * k]) k = 2;
        double ax[3] = {B[k], B[3 + k], B[6 + k]};
        const double n = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        for (int i = 0; i < 3; ++i) ax[i] /= n;
        if (ax[0] * v[0] + ax[1] * v[1] + ax[2] * v[2] < 0.0)
            for (int i = 0; i < 3; ++i) ax[i] = -ax[i];
        for (int i = 0; i < 3; ++i) w[i] = M_PI * ax[i];
        return;
    }
    double us; /* usinc_from_cos */
    if (c >= 1.0) us = 1.0;
    else if (c <= -1.0) us = 0.0;
    else us = sqrt(1.0 - c * c) / acos(c);
    for (int i = 0; i < 3; ++i) w[i] = v[i] / (2.0 * us);
}

/* getPoint(Pose3, c) = exp(M, e, hat(M, e, c)) under Hybrid semantics:
 * (t = c[1:3], R = Exp(c[4:6])); test/testPose3.jl:9-23, src/services/ManifoldUtils.jl:8-23 */
void rome_oracle_pose3_point(const double c[6], double t[3], double R[9]) {
    t[0] = c[0]; t[1] = c[1]; t[2] = c[2];
    rome_oracle_so3_exp(c + 3, R);
}
void rome_oracle_pose3_coords(const double t[3], const double R[9], double c[6]) {
    c[0] = t[0]; c[1] = t[1]; c[2] = t[2];
    rome_oracle_so3_log(R, c + 3);
}

/* ================================================================================== */
/* residuals                                                                          */
/* ================================================================================== */

/* src/factors/Pose2D.jl:51-67 with _compose/_vee of src/factors/PriorPose2.jl:19-25:
 *   eX = exp(M, e0, X) = (X.t, R(X.theta))
 *   qhat = (p.t + p.R eX.t, p.R eX.R)
 *   Xhat = log(M, q, qhat) = (qhat.t - q.t, log(q.R' qhat.R))
 *   return (Xhat.t[1], Xhat.t[2], Xhat.R[2,1]) */
void rome_oracle_pose2pose2(const double X[3], const double p[3], const double q[3], double r[3]) {
    const double cp = cos(p[2]), sp = sin(p[2]);
    const double cm = cos(X[2]), sm = sin(X[2]);
    const double cq = cos(q[2]), sq = sin(q[2]);
    /* qhat */
    const double tx = p[0] + cp * X[0] - sp * X[1];
    const double ty = p[1] + sp * X[0] + cp * X[1];
    const double h11 = cp * cm - sp * sm, h21 = sp * cm + cp * sm; /* first column of p.R eX.R */
    /* U = q.R' qhat.R, first column */
    const double u11 = cq * h11 + sq * h21;
    const double u21 = -sq * h11 + cq * h21;
    r[0] = tx - q[0];
    r[1] = ty - q[1];
    r[2] = atan2(u21, u11);
}

/* src/factors/PriorPose2.jl:37-47: _vee(log(M, p, m)) = (m.t - p.t, log(p.R' m.R)[2,1]) */
void rome_oracle_priorpose2(const double m[3], const double p[3], double r[3]) {
    const double cp = cos(p[2]), sp = sin(p[2]);
    const double cm = cos(m[2]), sm = sin(m[2]);
    r[0] = m[0] - p[0];
    r[1] = m[1] - p[1];
    r[2] = atan2(-sp * cm + cp * sm, cp * cm + sp * sm);
}

/* src/factors/BearingRange2D.jl:48-64:
 *   pl = p.R' (l - p.t);  dtheta = sym_rem(b - atan(pl[2], pl[1]));  dr = rho - norm(pl) */
void rome_oracle_bearingrange(const double meas[2], const double p[3], const double l[2], double r[2]) {
    const double cp = cos(p[2]), sp = sin(p[2]);
    const double dx = l[0] - p[0], dy = l[1] - p[1];
    const double plx = cp * dx + sp * dy;
    const double ply = -sp * dx + cp * dy;
    r[0] = rome_oracle_sym_rem(meas[0] - atan2(ply, plx));
    r[1] = meas[1] - sqrt(plx * plx + ply * ply); /* LinearAlgebra.norm of a 2-vector */
}

/* src/factors/Pose3Pose3.jl:17-29:
 *   qhat = compose(M, p, exp(M, e, X)) = (p.t + p.R X.t, p.R Exp(X.w))
 *   Xc = get_coordinates(M, q, log(M, q, qhat)) = (qhat.t - q.t, vee(Log(q.R' qhat.R))) */
void rome_oracle_pose3pose3(const double X[6], const double p[6], const double q[6], double r[6]) {
    double tp[3], Rp[9], tq[3], Rq[9], M[9], Rh[9], U[9], v[3];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_pose3_point(q, tq, Rq);
    rome_oracle_so3_exp(X + 3, M);
    mat3_vec(Rp, X, v);
    mat3_mul(Rp, M, Rh);
    mat3_tmul(Rq, Rh, U);
    for (int i = 0; i < 3; ++i) r[i] = tp[i] + v[i] - tq[i];
    rome_oracle_so3_log(U, r + 3);
}

/* src/factors/Pose3Pose3.jl:57-78 (Pose3Pose3RotOffset): the measurement is taken in frame a, p and q live in frame b,
 * bRa (a Rotation3 variable, rotation-vector coordinates w) rotates a into b:
 *   a_m = exp(M, e, aX) = (X.t, Exp(X.w));  b_m = (a_m.t, bRa a_m.R);  qhat = compose(M, p, b_m)
 *   return vee(M, q, log(M, q, qhat)) = (qhat.t - q.t, vee(Log(q.R' qhat.R))) */
void rome_oracle_pose3pose3rotoffset(const double X[6], const double p[6], const double q[6], const double w[3],
                                     double r[6]) {
    double tp[3], Rp[9], tq[3], Rq[9], M[9], B[9], BM[9], Rh[9], U[9], v[3];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_pose3_point(q, tq, Rq);
    rome_oracle_so3_exp(X + 3, M);
    rome_oracle_so3_exp(w, B);
    mat3_vec(Rp, X, v);
    mat3_mul(B, M, BM);
    mat3_mul(Rp, BM, Rh);
    mat3_tmul(Rq, Rh, U);
    for (int i = 0; i < 3; ++i) r[i] = tp[i] + v[i] - tq[i];
    rome_oracle_so3_log(U, r + 3);
}
/* src/factors/Pose3Pose3.jl:80-95 (Pose3Pose3Transform): Delta is a Pose3 variable:
 *   Dn = compose(M, Delta, exp(M, e, X)) = (D.t + D.R X.t, D.R Exp(X.w));  qhat = compose(M, p, Dn)
 *   return get_coordinates(M, q, log(M, q, qhat)) */
void rome_oracle_pose3pose3transform(const double X[6], const double p[6], const double q[6], const double D[6],
                                     double r[6]) {
    double tp[3], Rp[9], tq[3], Rq[9], tD[3], RD[9], M[9], DM[9], Rh[9], U[9], v[3], l[3];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_pose3_point(q, tq, Rq);
    rome_oracle_pose3_point(D, tD, RD);
    rome_oracle_so3_exp(X + 3, M);
    mat3_vec(RD, X, v);
    for (int i = 0; i < 3; ++i) l[i] = tD[i] + v[i];
    mat3_vec(Rp, l, v);
    mat3_mul(RD, M, DM);
    mat3_mul(Rp, DM, Rh);
    mat3_tmul(Rq, Rh, U);
    for (int i = 0; i < 3; ++i) r[i] = tp[i] + v[i] - tq[i];
    rome_oracle_so3_log(U, r + 3);
}

/* src/factors/Pose3D.jl:15-19: vee(M, p, log(M, p, m)) = (m.t - p.t, vee(Log(p.R' m.R))) */
void rome_oracle_priorpose3(const double m[6], const double p[6], double r[6]) {
    double tp[3], Rp[9], tm[3], Rm[9], U[9];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_pose3_point(m, tm, Rm);
    mat3_tmul(Rp, Rm, U);
    for (int i = 0; i < 3; ++i) r[i] = tm[i] - tp[i];
    rome_oracle_so3_log(U, r + 3);
}

/* ---- next-row families ---- */
/* src/factors/Point2D.jl:14-18 */
void rome_oracle_priorpoint2(const double m[2], const double x[2], double r[2]) {
    r[0] = m[0] - x[0];
    r[1] = m[1] - x[1];
}
/* src/factors/Point2D.jl:30-35 */
void rome_oracle_point2point2(const double m[2], const double xi[2], const double xj[2], double r[2]) {
    r[0] = m[0] - (xj[0] - xi[0]);
    r[1] = m[1] - (xj[1] - xi[1]);
}
/* src/factors/Pose2Point2.jl:23-40: w_H_qhat = affine(w_T_p) * affine((m, I)); return l - w_H_qhat[1:2, end] */
void rome_oracle_pose2point2(const double m[2], const double p[3], const double l[2], double r[2]) {
    const double c = cos(p[2]), s = sin(p[2]);
    r[0] = l[0] - (p[0] + c * m[0] - s * m[1]);
    r[1] = l[1] - (p[1] + s * m[0] + c * m[1]);
}
/* src/factors/Range2D.jl:14-18, 51-54 */
void rome_oracle_range2(const double rho[1], const double xi[2], const double l[2], double r[1]) {
    const double dx = l[0] - xi[0], dy = l[1] - xi[1];
    r[0] = rho[0] - sqrt(dx * dx + dy * dy);
}
/* src/factors/Bearing2D.jl:23-32 */
void rome_oracle_pose2point2bearing(const double b[1], const double p[3], const double l[2], double r[1]) {
    const double c = cos(p[2]), s = sin(p[2]);
    const double dx = l[0] - p[0], dy = l[1] - p[1];
    r[0] = rome_oracle_sym_rem(b[0] - atan2(-s * dx + c * dy, c * dx + s * dy));
}

/* ---- next-row 3-D families ---- */
/* src/factors/Point3D.jl:13-20 */
void rome_oracle_priorpoint3(const double m[3], const double x[3], double r[3]) {
    for (int i = 0; i < 3; ++i) r[i] = m[i] - x[i];
}
/* src/factors/Point3Point3.jl:11-15 */
void rome_oracle_point3point3(const double m[3], const double xi[3], const double xj[3], double r[3]) {
    for (int i = 0; i < 3; ++i) r[i] = m[i] - (xj[i] - xi[i]);
}
/* src/factors/PartialPose3.jl:116-134: p2 = (t_p[1:2], R from normalize(R_p[1:2,1])), q2 likewise;
 *   qhat = compose(M2, p2, exp(M2, e, X)); Xc = vee(M2, q2, log(M2, q2, qhat))   (Hybrid SE(2): same reading as Pose2Pose2) */
void rome_oracle_pose3pose3xyyaw(const double X[3], const double p[6], const double q[6], double r[3]) {
    double tp[3], Rp[9], tq[3], Rq[9];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_pose3_point(q, tq, Rq);
    const double np_ = hypot(Rp[0], Rp[3]), nq = hypot(Rq[0], Rq[3]);
    const double cp = Rp[0] / np_, sp = Rp[3] / np_, cq = Rq[0] / nq, sq = Rq[3] / nq;
    const double cm = cos(X[2]), sm = sin(X[2]);
    r[0] = tp[0] + cp * X[0] - sp * X[1] - tq[0];
    r[1] = tp[1] + sp * X[0] + cp * X[1] - tq[1];
    /* U = R_q2' R_p2 R(X.theta); angle = atan(U21, U11) */
    const double ch = cp * cm - sp * sm, sh = sp * cm + cp * sm;
    r[2] = atan2(sh * cq - ch * sq, ch * cq + sh * sq);
}
/* src/factors/PartialPose3.jl:212-226: Xc = vee(log(SO3, R_p, R_q)) = Log(R_p' R_q); return Xc_m - Xc */
void rome_oracle_pose3pose3rotation(const double m[3], const double p[6], const double q[6], double r[3]) {
    double tp[3], Rp[9], tq[3], Rq[9], U[9], w[3];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_pose3_point(q, tq, Rq);
    mat3_tmul(Rp, Rq, U);
    rome_oracle_so3_log(U, w);
    for (int i = 0; i < 3; ++i) r[i] = m[i] - w[i];
}
/* src/factors/Pose3Pose3.jl:107-116: Pose3Pose3 coordinates with normalize(Xc[1:3]) */
void rome_oracle_pose3pose3unittrans(const double X[6], const double p[6], const double q[6], double r[6]) {
    rome_oracle_pose3pose3(X, p, q, r);
    const double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    for (int i = 0; i < 3; ++i) r[i] /= n;
}

/* ================================================================================== */
/* closed-form roots                                                                  */
/* ================================================================================== */

/* q = p o Exp(X); in-tree twin: addPose2Pose2 = se2vee(SE2(x)*SE2(dx)),
 * src/services/OdometryUtils.jl:132-142 */
void rome_oracle_pose2pose2_fwd(const double X[3], const double p[3], double q[3]) {
    const double cp = cos(p[2]), sp = sin(p[2]);
    q[0] = p[0] + cp * X[0] - sp * X[1];
    q[1] = p[1] + sp * X[0] + cp * X[1];
    q[2] = rome_oracle_wrap_atan(p[2] + X[2]);
}
/* p with residual(X, p, q) == 0: theta_p = theta_q - m_theta, t_p = t_q - R(theta_p) m_t */
void rome_oracle_pose2pose2_bwd(const double X[3], const double q[3], double p[3]) {
    const double th = rome_oracle_wrap_atan(q[2] - X[2]);
    const double c = cos(th), s = sin(th);
    p[0] = q[0] - (c * X[0] - s * X[1]);
    p[1] = q[1] - (s * X[0] + c * X[1]);
    p[2] = th;
}
/* l = t_p + rho R(theta_p) (cos b, sin b); inverse of calcPosePointBearingRange,
 * src/services/SimulationUtils.jl:47-62 */
void rome_oracle_bearingrange_fwd(const double meas[2], const double p[3], double l[2]) {
    const double a = p[2] + meas[0];
    l[0] = p[0] + meas[1] * cos(a);
    l[1] = p[1] + meas[1] * sin(a);
}
void rome_oracle_pose3pose3_fwd(const double X[6], const double p[6], double q[6]) {
    double tp[3], Rp[9], M[9], Rh[9], v[3], t[3];
    rome_oracle_pose3_point(p, tp, Rp);
    rome_oracle_so3_exp(X + 3, M);
    mat3_vec(Rp, X, v);
    mat3_mul(Rp, M, Rh);
    for (int i = 0; i < 3; ++i) t[i] = tp[i] + v[i];
    rome_oracle_pose3_coords(t, Rh, q);
}
void rome_oracle_pose3pose3_bwd(const double X[6], const double q[6], double p[6]) {
    double tq[3], Rq[9], M[9], Mt[9], Rp[9], v[3], t[3];
    rome_oracle_pose3_point(q, tq, Rq);
    rome_oracle_so3_exp(X + 3, M);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Mt[3 * i + j] = M[3 * j + i];
    mat3_mul(Rq, Mt, Rp);
    mat3_vec(Rp, X, v);
    for (int i = 0; i < 3; ++i) t[i] = tq[i] - v[i];
    rome_oracle_pose3_coords(t, Rp, p);
}

/* ================================================================================== */
/* batched sweeps -- the loop IIF runs per factor over n = 1:N (SURVEY.md 3.1 HOT LOOP)*/
/* ================================================================================== */

static int set_threads(int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    return nthreads;
#else
    (void)nthreads;
    return 1;
#endif
}

/* ---- one STEP of the hot path on the CPU, as bench.py's GPU step defines it: for every factor x particle getSample
 * (rand(MvNormal) = mu + L z, IIF's default sampler on the `.Z` field; BearingRange2D.jl:17-27 for the two scalar
 * beliefs), the residual functor, and the per-factor statistics (sum r, sum r r').  The reference draws from Julia's
 * Xoshiro256++ stream with a ziggurat randn; this restatement uses Xoshiro256++ with the Marsaglia polar method, one
 * generator per factor seeded from (seed, factor) so the result does not depend on the thread count.
 * family: 0 Pose2Pose2, 1 PriorPose2, 2 BearingRange, 3 Pose3Pose3, 4 PriorPose3 (include/rome_b200.h numbering).
 * mu [nF][dm], Lc [nF][dm][dm] lower Cholesky factor row-major (BearingRange: diag(sig_b, sig_r)),
 * res [nF][N][dr], stats [nF][dr + dr(dr+1)/2]. */
typedef struct { uint64_t s[4]; int have; double spare; } xo_t;
static inline uint64_t xo_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xo_next(xo_t* g) {
    uint64_t* s = g->s;
    const uint64_t r = xo_rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = xo_rotl(s[3], 45);
    return r;
}
static inline void xo_seed(xo_t* g, uint64_t seed, uint64_t stream) {
    uint64_t z = seed ^ (stream * 0x9E3779B97F4A7C15ull);
    for (int i = 0; i < 4; ++i) {  /* splitmix64 */
        z += 0x9E3779B97F4A7C15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        g->s[i] = x ^ (x >> 31);
    }
    g->have = 0;
}
static inline double xo_normal(xo_t* g) {
    if (g->have) { g->have = 0; return g->spare; }
    double u, v, q;
    do {
        u = (double)(xo_next(g) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
        v = (double)(xo_next(g) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
        q = u * u + v * v;
    } while (q >= 1.0 || q == 0.0);
    const double m = sqrt(-2.0 * log(q) / q);
    g->spare = v * m; g->have = 1;
    return u * m;
}
int rome_oracle_step(int family, int nF, int N, const int32_t* i0, const int32_t* i1, const double* v0, const double* v1,
                     const double* mu, const double* Lc, uint64_t seed, double* res, double* stats, int nthreads) {
    static const int DM[5] = {3, 3, 2, 6, 6}, D0[5] = {3, 3, 3, 6, 6}, D1[5] = {3, 0, 2, 6, 0};
    if (family < 0 || family > 4) return -1;
    const int dm = DM[family], dr = dm, d0 = D0[family], d1 = D1[family], ns = dr + dr * (dr + 1) / 2;
    const int nt = set_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int f = 0; f < nF; ++f) {
        xo_t g;
        xo_seed(&g, seed, (uint64_t)f);
        const double* P = v0 + (size_t)i0[f] * N * d0;
        const double* Q = d1 ? v1 + (size_t)i1[f] * N * d1 : 0;
        const double *m0 = mu + (size_t)f * dm, *L = Lc + (size_t)f * dm * dm;
        double* R = res + (size_t)f * N * dr;
        double st[27];
        for (int i = 0; i < ns; ++i) st[i] = 0.0;
        for (int n = 0; n < N; ++n) {
            double z[6], X[6], r[6];
            for (int i = 0; i < dm; ++i) z[i] = xo_normal(&g);
            for (int i = 0; i < dm; ++i) {
                double a = m0[i];
                for (int j = 0; j <= i; ++j) a += L[i * dm + j] * z[j];
                X[i] = a;
            }
            switch (family) {
                case 0: rome_oracle_pose2pose2(X, P + 3 * n, Q + 3 * n, r); break;
                case 1: rome_oracle_priorpose2(X, P + 3 * n, r); break;
                case 2: rome_oracle_bearingrange(X, P + 3 * n, Q + 2 * n, r); break;
                case 3: rome_oracle_pose3pose3(X, P + 6 * n, Q + 6 * n, r); break;
                default: rome_oracle_priorpose3(X, P + 6 * n, r); break;
            }
            int k = dr;
            for (int i = 0; i < dr; ++i) {
                R[(size_t)n * dr + i] = r[i];
                st[i] += r[i];
                for (int j = i; j < dr; ++j) st[k++] += r[i] * r[j];
            }
        }
        for (int i = 0; i < ns; ++i) stats[(size_t)f * ns + i] = st[i];
    }
    return nt;
}

int rome_oracle_sweep_pose2pose2(int nF, int N, const int32_t* ip, const int32_t* iq,
                                 const double* poses, const double* meas, double* res, int nthreads) {
    const int nt = set_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int f = 0; f < nF; ++f) {
        const double* P = poses + (size_t)ip[f] * N * 3;
        const double* Q = poses + (size_t)iq[f] * N * 3;
        const double* X = meas + (size_t)f * N * 3;
        double* R = res + (size_t)f * N * 3;
        for (int n = 0; n < N; ++n) rome_oracle_pose2pose2(X + 3 * n, P + 3 * n, Q + 3 * n, R + 3 * n);
    }
    return nt;
}

int rome_oracle_sweep_priorpose2(int nF, int N, const int32_t* ip, const double* poses,
                                 const double* meas, double* res, int nthreads) {
    const int nt = set_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int f = 0; f < nF; ++f) {
        const double* P = poses + (size_t)ip[f] * N * 3;
        const double* X = meas + (size_t)f * N * 3;
        double* R = res + (size_t)f * N * 3;
        for (int n = 0; n < N; ++n) rome_oracle_priorpose2(X + 3 * n, P + 3 * n, R + 3 * n);
    }
    return nt;
}

int rome_oracle_sweep_bearingrange(int nF, int N, const int32_t* ip, const int32_t* il,
                                   const double* poses, const double* points, const double* meas,
                                   double* res, int nthreads) {
    const int nt = set_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int f = 0; f < nF; ++f) {
        const double* P = poses + (size_t)ip[f] * N * 3;
        const double* L = points + (size_t)il[f] * N * 2;
        const double* X = meas + (size_t)f * N * 2;
        double* R = res + (size_t)f * N * 2;
        for (int n = 0; n < N; ++n) rome_oracle_bearingrange(X + 2 * n, P + 3 * n, L + 2 * n, R + 2 * n);
    }
    return nt;
}

int rome_oracle_sweep_pose3pose3(int nF, int N, const int32_t* ip, const int32_t* iq,
                                 const double* poses, const double* meas, double* res, int nthreads) {
    const int nt = set_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int f = 0; f < nF; ++f) {
        const double* P = poses + (size_t)ip[f] * N * 6;
        const double* Q = poses + (size_t)iq[f] * N * 6;
        const double* X = meas + (size_t)f * N * 6;
        double* R = res + (size_t)f * N * 6;
        for (int n = 0; n < N; ++n) rome_oracle_pose3pose3(X + 6 * n, P + 6 * n, Q + 6 * n, R + 6 * n);
    }
    return nt;
}

int rome_oracle_sweep_priorpose3(int nF, int N, const int32_t* ip, const double* poses,
                                 const double* meas, double* res, int nthreads) {
    const int nt = set_threads(nthreads);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int f = 0; f < nF; ++f) {
        const double* P = poses + (size_t)ip[f] * N * 6;
        const double* X = meas + (size_t)f * N * 6;
        double* R = res + (size_t)f * N * 6;
        for (int n = 0; n < N; ++n) rome_oracle_priorpose3(X + 6 * n, P + 6 * n, R + 6 * n);
    }
    return nt;
}

/* ================================================================================== */
/* sampler twin: Philox4x32-10 + Box-Muller, identical to csrc/philox.cuh             */
/* ================================================================================== */

void rome_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void rome_oracle_normal4(uint64_t seed, uint32_t stream, uint32_t factor, uint32_t particle,
                         uint32_t block, double z[4]) {
    const uint32_t ctr[4] = {particle, factor, stream, block};
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t x[4];
    rome_oracle_philox4x32_10(ctr, key, x);
    for (int h = 0; h < 2; ++h) {
        const double u1 = ((double)(x[2 * h] >> 9) + 0.5) * (1.0 / 8388608.0);
        const double u2 = ((double)(x[2 * h + 1] >> 9) + 0.5) * (1.0 / 8388608.0);
        const double rad = sqrt(-2.0 * log(u1));
        z[2 * h] = rad * cos(2.0 * M_PI * u2);
        z[2 * h + 1] = rad * sin(2.0 * M_PI * u2);
    }
}

/* ================================================================================== */
/* reference-shaped convolution: Nelder-Mead per particle                             */
/* ================================================================================== */

/* cost(x) = || cf(meas, p, q(x)) ||^2 with the solve-for variable replaced by the
 * candidate coordinates x (_solveCCWNumeric! [IIF-knowledge]). */
typedef struct {
    const double* X; /* measurement */
    const double* other; /* the fixed variable's particle */
    int fwd;
    uint64_t evals;
} nm_ctx;

static double nm_cost(nm_ctx* c, const double x[3]) {
    double r[3];
    if (c->fwd) rome_oracle_pose2pose2(c->X, c->other, x, r);
    else rome_oracle_pose2pose2(c->X, x, c->other, r);
    c->evals++;
    return r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
}

/* Optim.jl NelderMead() defaults restated: AffineSimplexer(a=0.025, b=0.5) initial simplex
 * x_i = x0 + (b*x0[i] + a) e_i; AdaptiveParameters alpha=1, beta=1+2/n, gamma=0.75-1/(2n),
 * delta=1-1/n; stop when sqrt(var(f_simplex) * n/(n+1)) < g_tol (1e-8) or iterations = 1000. */
static void nelder_mead3(nm_ctx* c, double x0[3]) {
    enum { n = 3 };
    const double alpha = 1.0, beta = 1.0 + 2.0 / n, gamma = 0.75 - 1.0 / (2.0 * n), delta = 1.0 - 1.0 / n;
    double S[n + 1][n], f[n + 1];
    for (int i = 0; i <= n; ++i) {
        for (int j = 0; j < n; ++j) S[i][j] = x0[j];
        if (i > 0) S[i][i - 1] += 0.5 * x0[i - 1] + 0.025;
        f[i] = nm_cost(c, S[i]);
    }
    for (int it = 0; it < 1000; ++it) {
        /* order */
        int idx[n + 1] = {0, 1, 2, 3};
        for (int a = 0; a <= n; ++a)
            for (int b = a + 1; b <= n; ++b)
                if (f[idx[b]] < f[idx[a]]) { int t = idx[a]; idx[a] = idx[b]; idx[b] = t; }
        const int lo = idx[0], hi = idx[n], nh = idx[n - 1];
        double mean = 0.0, var = 0.0;
        for (int i = 0; i <= n; ++i) mean += f[i];
        mean /= (n + 1);
        for (int i = 0; i <= n; ++i) var += (f[i] - mean) * (f[i] - mean);
        var /= n; /* Julia var: unbiased */
        if (sqrt(var * n / (n + 1.0)) < 1e-8) break;
        double cen[n] = {0, 0, 0};
        for (int i = 0; i <= n; ++i)
            if (i != hi) for (int j = 0; j < n; ++j) cen[j] += S[i][j] / n;
        double xr[n], xe[n], xc[n];
        for (int j = 0; j < n; ++j) xr[j] = cen[j] + alpha * (cen[j] - S[hi][j]);
        const double fr = nm_cost(c, xr);
        if (fr < f[lo]) {
            for (int j = 0; j < n; ++j) xe[j] = cen[j] + beta * (xr[j] - cen[j]);
            const double fe = nm_cost(c, xe);
            if (fe < fr) { memcpy(S[hi], xe, sizeof xe); f[hi] = fe; }
            else { memcpy(S[hi], xr, sizeof xr); f[hi] = fr; }
        } else if (fr < f[nh]) {
            memcpy(S[hi], xr, sizeof xr); f[hi] = fr;
        } else {
            int shrink = 0;
            if (fr < f[hi]) { /* outside contraction */
                for (int j = 0; j < n; ++j) xc[j] = cen[j] + gamma * (xr[j] - cen[j]);
                const double fc = nm_cost(c, xc);
                if (fc <= fr) { memcpy(S[hi], xc, sizeof xc); f[hi] = fc; } else shrink = 1;
            } else { /* inside contraction */
                for (int j = 0; j < n; ++j) xc[j] = cen[j] - gamma * (xr[j] - cen[j]);
                const double fc = nm_cost(c, xc);
                if (fc < f[hi]) { memcpy(S[hi], xc, sizeof xc); f[hi] = fc; } else shrink = 1;
            }
            if (shrink)
                for (int i = 0; i <= n; ++i)
                    if (i != lo) {
                        for (int j = 0; j < n; ++j) S[i][j] = S[lo][j] + delta * (S[i][j] - S[lo][j]);
                        f[i] = nm_cost(c, S[i]);
                    }
        }
    }
    int best = 0;
    for (int i = 1; i <= n; ++i) if (f[i] < f[best]) best = i;
    memcpy(x0, S[best], sizeof(double) * n);
}

int rome_oracle_conv_nm_pose2pose2(int nF, int N, const int32_t* ip, const int32_t* iq,
                                   const double* poses, const double* meas, int fwd,
                                   int inflate_cycles, double inflation, uint64_t seed,
                                   double* out, uint64_t* n_evals, int nthreads) {
    const int nt = set_threads(nthreads);
    uint64_t total = 0;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(+ : total)
    for (int f = 0; f < nF; ++f) {
        const double* P = poses + (size_t)ip[f] * N * 3;
        const double* Q = poses + (size_t)iq[f] * N * 3;
        const double* tgt = fwd ? Q : P;
        /* spread of the target belief per coordinate (IIF perturbs the start point by
         * inflation * spread * randn each cycle [IIF-knowledge]) */
        double mu[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
        for (int n = 0; n < N; ++n) for (int j = 0; j < 3; ++j) mu[j] += tgt[3 * n + j] / N;
        for (int n = 0; n < N; ++n)
            for (int j = 0; j < 3; ++j) sd[j] += (tgt[3 * n + j] - mu[j]) * (tgt[3 * n + j] - mu[j]) / N;
        for (int j = 0; j < 3; ++j) sd[j] = sqrt(sd[j]);
        for (int n = 0; n < N; ++n) {
            nm_ctx c = {meas + ((size_t)f * N + n) * 3, (fwd ? P : Q) + 3 * n, fwd, 0};
            double x[3] = {tgt[3 * n], tgt[3 * n + 1], tgt[3 * n + 2]};
            for (int cyc = 0; cyc < inflate_cycles; ++cyc) {
                double z[4];
                rome_oracle_normal4(seed, 0x4e4du + (uint32_t)cyc, (uint32_t)f, (uint32_t)n, 0, z);
                for (int j = 0; j < 3; ++j) x[j] += inflation * sd[j] * z[j];
                nelder_mead3(&c, x);
            }
            memcpy(out + ((size_t)f * N + n) * 3, x, sizeof x);
            total += c.evals;
        }
    }
    if (n_evals) *n_evals = total;
    return nt;
}


/* ================================================================================== */
/* product of proposal KDEs -- PARITY UNPINNED: the reference's     */
/* sampler (ApproxManifoldProducts / KernelDensityEstimate) is absent and stochastic; */
/* this restates the algorithm of csrc/product_kernels.cu in Float64 with its own RNG */
/* and is compared statistically (tests/test_oracle_golden.py, tests/test_gpu_product) */
/* ================================================================================== */
static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline double unif01(uint64_t* s) { return ((double)(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double wrapd(double a) { return a - 6.283185307179586476925 * nearbyint(a * 0.15915494309189533577); }

/* rule-of-thumb bandwidth per dimension (circular spread for wrap_dim); out: 1/h^2 */
static void kde_precision(const double* x, int N, int d, int wrap_dim, double* w) {
    const double scale = pow(4.0 / ((d + 2.0) * N), 1.0 / (d + 4.0));
    for (int c = 0; c < d; ++c) {
        double mean = 0.0;
        if (c == wrap_dim) {
            double sc = 0.0, ss = 0.0;
            for (int i = 0; i < N; ++i) { sc += cos(x[i * d + c]); ss += sin(x[i * d + c]); }
            mean = atan2(ss, sc);
        } else {
            for (int i = 0; i < N; ++i) mean += x[i * d + c];
            mean /= N;
        }
        double v = 0.0;
        for (int i = 0; i < N; ++i) {
            double dl = x[i * d + c] - mean;
            if (c == wrap_dim) dl = wrapd(dl);
            v += dl * dl;
        }
        double h = sqrt(v / (N > 1 ? N - 1 : 1)) * scale;
        if (h < 1e-6) h = 1e-6;
        w[c] = 1.0 / (h * h);
    }
}
#define ROME_ORACLE_MAXD 6
#define ROME_ORACLE_MAXK 32
static void fuse_sel(int k, int d, int wrap_dim, const double* const* props, const double* w, const int* sel, int upto,
                     int skip, double* mu, double* var) {
    double lam[ROME_ORACLE_MAXD] = {0}, s[ROME_ORACLE_MAXD] = {0}, ref = 0.0;
    int have_ref = 0;
    (void)k;
    for (int j = 0; j < upto; ++j) {
        if (j == skip) continue;
        const double* x = props[j] + (size_t)sel[j] * d;
        for (int c = 0; c < d; ++c) {
            double v = x[c];
            if (c == wrap_dim) {
                if (!have_ref) { ref = v; have_ref = 1; }
                v = ref + wrapd(v - ref);
            }
            lam[c] += w[j * d + c];
            s[c] += w[j * d + c] * v;
        }
    }
    for (int c = 0; c < d; ++c) { var[c] = 1.0 / lam[c]; mu[c] = s[c] * var[c]; }
}
static int draw_label(int N, int d, int wrap_dim, const double* x, const double* wj, const double* mu, const double* var,
                      uint64_t* rng) {
    double best = -1e300;
    int arg = 0;
    for (int i = 0; i < N; ++i) {
        double q = 0.0;
        for (int c = 0; c < d; ++c) {
            double dl = x[i * d + c] - mu[c];
            if (c == wrap_dim) dl = wrapd(dl);
            q += -0.5 * dl * dl / (1.0 / wj[c] + var[c]);
        }
        const double key = q - log(-log(unif01(rng)));
        if (key > best) { best = key; arg = i; }
    }
    return arg;
}
/* props: k pointers to [N][d] particle sets; out: [n_out][d] samples of the product of their KDEs */
int rome_oracle_product(int k, int N, int d, int wrap_dim, const double* const* props, int n_out, int iters,
                        uint64_t seed, double* out) {
    if (k < 2 || k > ROME_ORACLE_MAXK || d > ROME_ORACLE_MAXD || N < 1) return -1;
    double w[ROME_ORACLE_MAXK * ROME_ORACLE_MAXD];
    for (int j = 0; j < k; ++j) kde_precision(props[j], N, d, wrap_dim, w + j * d);
    /* exact pair stage: marginal weights of source-0 components over source 1 (log-sum-exp), CDF */
    double* cdf = (double*)malloc(sizeof(double) * N);
    double wmax = -1e300;
    for (int a = 0; a < N; ++a) {
        double m = -1e300, sum = 0.0;
        for (int b = 0; b < N; ++b) {
            double q = 0.0;
            for (int c = 0; c < d; ++c) {
                double dl = props[1][b * d + c] - props[0][a * d + c];
                if (c == wrap_dim) dl = wrapd(dl);
                q += -0.5 * dl * dl / (1.0 / w[c] + 1.0 / w[d + c]);
            }
            const double m2 = q > m ? q : m;
            sum = sum * exp(m - m2) + exp(q - m2);
            m = m2;
        }
        cdf[a] = m + log(sum);
        if (cdf[a] > wmax) wmax = cdf[a];
    }
    double acc = 0.0;
    for (int a = 0; a < N; ++a) { acc += exp(cdf[a] - wmax); cdf[a] = acc; }
    uint64_t rng = seed * 0x9e3779b97f4a7c15ULL + 12345;
    int sel[ROME_ORACLE_MAXK];
    double mu[ROME_ORACLE_MAXD], var[ROME_ORACLE_MAXD];
    for (int n = 0; n < n_out; ++n) {
        const double target = unif01(&rng) * cdf[N - 1];
        int lo = 0, hi = N - 1;
        while (lo < hi) { const int mid = (lo + hi) / 2; if (cdf[mid] < target) lo = mid + 1; else hi = mid; }
        sel[0] = lo;
        fuse_sel(k, d, wrap_dim, props, w, sel, 1, -1, mu, var);
        sel[1] = draw_label(N, d, wrap_dim, props[1], w + d, mu, var, &rng);
        for (int j = 2; j < k; ++j) {
            fuse_sel(k, d, wrap_dim, props, w, sel, j, -1, mu, var);
            sel[j] = draw_label(N, d, wrap_dim, props[j], w + j * d, mu, var, &rng);
        }
        if (k > 2)
            for (int t = 0; t < iters; ++t)
                for (int j = 0; j < k; ++j) {
                    fuse_sel(k, d, wrap_dim, props, w, sel, k, j, mu, var);
                    sel[j] = draw_label(N, d, wrap_dim, props[j], w + j * d, mu, var, &rng);
                }
        fuse_sel(k, d, wrap_dim, props, w, sel, k, -1, mu, var);
        for (int c = 0; c < d; c += 2) {  /* Box-Muller */
            const double r = sqrt(-2.0 * log(unif01(&rng))), t = 6.283185307179586476925 * unif01(&rng);
            double x0 = mu[c] + sqrt(var[c]) * r * cos(t);
            if (c == wrap_dim) x0 = wrapd(x0);
            out[(size_t)n * d + c] = x0;
            if (c + 1 < d) {
                double x1 = mu[c + 1] + sqrt(var[c + 1]) * r * sin(t);
                if (c + 1 == wrap_dim) x1 = wrapd(x1);
                out[(size_t)n * d + c + 1] = x1;
            }
        }
    }
    free(cdf);
    return 0;
}
/* one sweep's worth of products: variable v multiplies rows src_row[var_off[v] .. var_off[v+1]) of `rows`
 * ([nrows][N][d]); out [nvars][N][d]; variables with fewer than two sources are copied / left zero */
int rome_oracle_product_sweep(int nvars, const int32_t* var_off, const int32_t* src_row, const double* rows, int N, int d,
                              int wrap_dim, int iters, uint64_t seed, double* out, int nthreads) {
    const int nt = set_threads(nthreads);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int v = 0; v < nvars; ++v) {
        const int k = var_off[v + 1] - var_off[v];
        double* o = out + (size_t)v * N * d;
        if (k == 0) continue;
        if (k == 1) { memcpy(o, rows + (size_t)src_row[var_off[v]] * N * d, sizeof(double) * N * d); continue; }
        const double* props[ROME_ORACLE_MAXK];
        if (k > ROME_ORACLE_MAXK) { bad = 1; continue; }
        for (int j = 0; j < k; ++j) props[j] = rows + (size_t)src_row[var_off[v] + j] * N * d;
        if (rome_oracle_product(k, N, d, wrap_dim, props, N, iters, seed + (uint64_t)v, o)) bad = 1;
    }
    return bad ? -1 : nt;
}
