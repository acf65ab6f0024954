/*
 * rome_oracle.h -- float64 CPU restatement of the RoME.jl factor-residual hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the CPU baseline.
 *
 * Parity status: PINNED.  The restatement reproduces every known-answer vector the
 * reference's own tests hold for this path (tests/golden/known_answers.json, each
 * with its reference file:line) -- see tests/test_oracle_golden.py.  The reference
 * itself (Julia + Manifolds.jl 0.10 + IncrementalInference 0.35, not vendored under
 * /root/reference, no julia binary in the image) cannot be executed here, so there
 * is no oracle/_ref build; DESIGN.md says so.
 *
 * Conventions (all citations relative to /root/reference):
 *   Pose2 coordinates (x, y, theta)  <-> point (t in R^2, R(theta) in SO(2))
 *       src/variables/VariableTypes.jl:35
 *   Pose3 coordinates (x, y, z, wx, wy, wz): translation first, rotation vector second
 *       src/variables/VariableTypes.jl:47, src/services/ManifoldUtils.jl:14
 *   Point2 coordinates (x, y)          src/variables/VariableTypes.jl:13
 *   SpecialEuclidean(n; vectors=HybridTangentRepresentation()):
 *       exp(M, e, X) = (X.t, Exp_SO(n)(X.w));  log(M, q, s) = (s.t - q.t, Log_SO(n)(q.R' s.R))
 *       [Manifolds.jl 0.10 semantics; pinned by test/testParametric.jl:22-53 and the
 *        optimizer trace at test/testParametricSimulated.jl:118-119]
 */
#ifndef ROME_ORACLE_H
#define ROME_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar helpers -------------------------------------------------------------- */
double rome_oracle_sym_rem(double x);              /* Manifolds.sym_rem: wrap to [-pi,pi], +pi -> -pi */
double rome_oracle_wrap_atan(double a);            /* atan(sin a, cos a) in (-pi, pi] */
void   rome_oracle_so3_exp(const double w[3], double R[9]);      /* row-major 3x3 */
void   rome_oracle_so3_log(const double R[9], double w[3]);
void   rome_oracle_pose3_point(const double c[6], double t[3], double R[9]);   /* getPoint(Pose3, c) */
void   rome_oracle_pose3_coords(const double t[3], const double R[9], double c[6]);

/* ---- one residual evaluation (== calcFactorResidualTemporary) -------------------- */
/* src/factors/Pose2D.jl:51-67 ; X = tangent coordinates (mx,my,mtheta) */
void rome_oracle_pose2pose2(const double X[3], const double p[3], const double q[3], double r[3]);
/* src/factors/PriorPose2.jl:37-47 ; m = sampled POINT coordinates */
void rome_oracle_priorpose2(const double m[3], const double p[3], double r[3]);
/* src/factors/BearingRange2D.jl:48-64 ; meas = (bearing, range) */
void rome_oracle_bearingrange(const double meas[2], const double p[3], const double l[2], double r[2]);
/* src/factors/Pose3Pose3.jl:17-29 ; all arguments as coordinates */
void rome_oracle_pose3pose3(const double X[6], const double p[6], const double q[6], double r[6]);
/* src/factors/Pose3D.jl:15-19 */
void rome_oracle_priorpose3(const double m[6], const double p[6], double r[6]);

/* ---- next-row families ---------------------------------------------------------- */
/* src/factors/Point2D.jl:14-18  PriorPoint2: meas - x */
void rome_oracle_priorpoint2(const double m[2], const double x[2], double r[2]);
/* src/factors/Point2D.jl:30-35  Point2Point2: meas - (xj - xi) */
void rome_oracle_point2point2(const double m[2], const double xi[2], const double xj[2], double r[2]);
/* src/factors/Pose2Point2.jl:23-40  Pose2Point2: l - (p.t + R_p m) */
void rome_oracle_pose2point2(const double m[2], const double p[3], const double l[2], double r[2]);
/* src/factors/Range2D.jl:51-54 Pose2Point2Range and :14-18 Point2Point2Range: rho - |l - x| (xi: first two coords) */
void rome_oracle_range2(const double rho[1], const double xi[2], const double l[2], double r[1]);
/* src/factors/Bearing2D.jl:23-32 Pose2Point2Bearing: sym_rem(b - atan(R_p'(l - p.t))) */
void rome_oracle_pose2point2bearing(const double b[1], const double p[3], const double l[2], double r[1]);

/* next-row 3-D families */
/* src/factors/Point3D.jl:13-20 PriorPoint3: m - x ; src/factors/Point3Point3.jl:11-15: m - (xj - xi) */
void rome_oracle_priorpoint3(const double m[3], const double x[3], double r[3]);
void rome_oracle_point3point3(const double m[3], const double xi[3], const double xj[3], double r[3]);
/* src/factors/PartialPose3.jl:116-134 Pose3Pose3XYYaw: SE(2) residual of the (x, y, yaw) projections */
void rome_oracle_pose3pose3xyyaw(const double X[3], const double p[6], const double q[6], double r[3]);
/* src/factors/PartialPose3.jl:212-226 Pose3Pose3Rotation: m - Log(R_p' R_q) */
void rome_oracle_pose3pose3rotation(const double m[3], const double p[6], const double q[6], double r[3]);
/* src/factors/Pose3Pose3.jl:107-116 Pose3Pose3UnitTrans */
/* families with a third variable: src/factors/Pose3Pose3.jl:57-78 (w: Rotation3 coordinates), :80-95 (D: Pose3) */
void rome_oracle_pose3pose3rotoffset(const double X[6], const double p[6], const double q[6], const double w[3],
                                     double r[6]);
void rome_oracle_pose3pose3transform(const double X[6], const double p[6], const double q[6], const double D[6],
                                     double r[6]);
void rome_oracle_pose3pose3unittrans(const double X[6], const double p[6], const double q[6], double r[6]);

/* ---- closed-form roots of the residual (what the per-particle solve converges to) */
/* cf. src/services/OdometryUtils.jl:132-158 (addPose2Pose2 / odomKDE) */
void rome_oracle_pose2pose2_fwd(const double X[3], const double p[3], double q[3]);
void rome_oracle_pose2pose2_bwd(const double X[3], const double q[3], double p[3]);
/* cf. src/services/SimulationUtils.jl:47-62 (inverse of calcPosePointBearingRange) */
void rome_oracle_bearingrange_fwd(const double meas[2], const double p[3], double l[2]);
void rome_oracle_pose3pose3_fwd(const double X[6], const double p[6], double q[6]);
void rome_oracle_pose3pose3_bwd(const double X[6], const double q[6], double p[6]);

/* ---- batched sweeps (reference layout: particle-major AoS float64 "vecval") ------
 * vars  : [nvars][N][d] coordinates      meas : [nF][N][dm]      res : [nF][N][dr]
 * nthreads <= 0 -> all OpenMP threads.  Return the thread count used.               */
int rome_oracle_sweep_pose2pose2(int nF, int N, const int32_t* ip, const int32_t* iq,
                                 const double* poses, const double* meas, double* res, int nthreads);
int rome_oracle_sweep_priorpose2(int nF, int N, const int32_t* ip,
                                 const double* poses, const double* meas, double* res, int nthreads);
int rome_oracle_sweep_bearingrange(int nF, int N, const int32_t* ip, const int32_t* il,
                                   const double* poses, const double* points, const double* meas,
                                   double* res, int nthreads);
int rome_oracle_sweep_pose3pose3(int nF, int N, const int32_t* ip, const int32_t* iq,
                                 const double* poses, const double* meas, double* res, int nthreads);
int rome_oracle_sweep_priorpose3(int nF, int N, const int32_t* ip,
                                 const double* poses, const double* meas, double* res, int nthreads);

/* One STEP of the hot path as bench.py's GPU step defines it (getSample + residual + per-factor statistics for every
 * factor x particle of one family); see rome_oracle.c.  Returns the thread count used, -1 for an unknown family. */
int rome_oracle_step(int family, int nF, int N, const int32_t* i0, const int32_t* i1, const double* v0, const double* v1,
                     const double* mu, const double* Lc, uint64_t seed, double* res, double* stats, int nthreads);

/* ---- "reference-shaped" convolution: per particle, Nelder-Mead over the target's
 * tangent coordinates wrapping the residual, inflateCycles restarts with inflation
 * noise (IIF 0.35 defaults N=100, inflateCycles=3, inflation=5.0 as serialized in
 * test/testdata/g2otest.tar.gz:dfg.json).  Solves for q (fwd=1) or p (fwd=0).
 * out: [nF][N][3] proposals; *n_evals receives the number of residual calls made.   */
int rome_oracle_conv_nm_pose2pose2(int nF, int N, const int32_t* ip, const int32_t* iq,
                                   const double* poses, const double* meas, int fwd,
                                   int inflate_cycles, double inflation, uint64_t seed,
                                   double* out, uint64_t* n_evals, int nthreads);

/* ---- sampler twin (restates the DEVICE sampler of rome.jl_b200/csrc, so that tests can
 * reproduce fused-getSample draws on the host; the reference's own sampler is
 * rand(MvNormal) on Julia's Xoshiro stream, which is not reproducible outside Julia) */
void rome_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* four N(0,1) draws for (seed, stream, factor, particle, block) */
void rome_oracle_normal4(uint64_t seed, uint32_t stream, uint32_t factor, uint32_t particle,
                         uint32_t block, double z[4]);

/* ---- product of proposal KDEs (PARITY UNPINNED, statistical checks only) ---- */
int rome_oracle_product(int k, int N, int d, int wrap_dim, const double* const* props, int n_out, int iters,
                        uint64_t seed, double* out);
int rome_oracle_product_sweep(int nvars, const int32_t* var_off, const int32_t* src_row, const double* rows, int N, int d,
                              int wrap_dim, int iters, uint64_t seed, double* out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
