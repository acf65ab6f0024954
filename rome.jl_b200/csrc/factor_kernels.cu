// factor_kernels.cu -- hand-written sm_100a kernels of the RoME factor-residual hot path.
//
// One fused kernel per factor family.  For every factor of the family in [first, first+count) and
// every particle n < N it (optionally) draws the measurement (getSample), composes the SE(2)/SE(3)
// group operation, applies the measurement and writes the residual coordinates, the closed-form
// proposal(s), compact Jacobian entries and per-factor statistics.
//
//   persistent CTAs = FT consumer warps (one factor each per tile of FT factors) + 1 producer warp;
//   the producer stages, per tile and S tiles ahead, everything the consumers read into shared memory
//   with 1-D TMA bulk copies (cp.async.bulk, completion on an mbarrier per stage): the tile's factor-table
//   rows {var ids, mu (f64), chol(Sigma) (f32)}, its measurement block, and -- gathered by variable id --
//   one contiguous particle block {anchor (f64), d rows x Npad float32 offsets} per factor slot;
//   consumers never issue a global load: they wait on the stage's "full" barrier, compute in Float64 on
//   the anchored float32 data (DESIGN.md "precision"), write residual / proposal rows with coalesced
//   16-B streaming stores and release the stage through its "empty" barrier;
//   statistics are reduced with a halving-butterfly of __shfl_xor_sync (device_utils.cuh).
//
// Reference arithmetic (paths relative to /root/reference):
//   Pose2Pose2    src/factors/Pose2D.jl:51-67, _compose/_vee src/factors/PriorPose2.jl:19-25
//   PriorPose2    src/factors/PriorPose2.jl:37-47
//   BearingRange  src/factors/BearingRange2D.jl:48-64 (getSample :17-27)
//   Pose3Pose3    src/factors/Pose3Pose3.jl:17-29
//   PriorPose3    src/factors/Pose3D.jl:15-19
#include <cuda_runtime.h>

#include "../../include/rome_b200.h"
#include "device_utils.cuh"
#include "tables.h"

namespace rome {

// =============================================================================================
// Float64 quaternion helpers for SO(3)
// =============================================================================================
struct Quat {
    double w, x, y, z;
};
// general path: any rotation vector
__device__ __noinline__ Quat quat_exp_general(double wx, double wy, double wz) {
    const double t2 = wx * wx + wy * wy + wz * wz;
    double k, c;
    if (t2 < 1e-8) {
        k = 0.5 - t2 * (1.0 / 48.0);
        c = 1.0 - t2 * 0.125 + t2 * t2 * (1.0 / 384.0);
    } else {
        const double t = sqrt(t2);
        double s;
        sincos(0.5 * t, &s, &c);
        k = s / t;
    }
    return {c, k * wx, k * wy, k * wz};
}
// Exp of a rotation vector with |w| <= pi (the principal range) without sqrt, division or range reduction:
// with y = theta/4 <= pi/4 the fdlibm kernels give cos y and sin(y)/y as polynomials in y^2 = |w|^2/16, and
//   cos(theta/2) = 2 cos^2 y - 1,   sin(theta/2)/theta = (sin(y)/y) cos(y) / 2.
__device__ __forceinline__ Quat quat_exp(double wx, double wy, double wz) {
    const double t2 = wx * wx + wy * wy + wz * wz;
    if (t2 > 9.8696) return quat_exp_general(wx, wy, wz);  // |w| > pi: non-principal rotation vector
    const double z = t2 * 0.0625;
    double ps = fma(z, kSinC[5], kSinC[4]);
    double pc = fma(z, kCosC[5], kCosC[4]);
    ps = fma(z, ps, kSinC[3]); pc = fma(z, pc, kCosC[3]);
    ps = fma(z, ps, kSinC[2]); pc = fma(z, pc, kCosC[2]);
    ps = fma(z, ps, kSinC[1]); pc = fma(z, pc, kCosC[1]);
    ps = fma(z, ps, kSinC[0]); pc = fma(z, pc, kCosC[0]);
    const double sy = fma(z, ps, 1.0);                       // sin(y)/y
    const double cy = fma(z * z, pc, fma(z, -0.5, 1.0));     // cos(y)
    const double k = 0.5 * sy * cy;
    return {fma(2.0 * cy, cy, -1.0), k * wx, k * wy, k * wz};
}
__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Quat qconj(const Quat& a) { return {a.w, -a.x, -a.y, -a.z}; }
// rotation vector (angle in [0, pi]) of a unit quaternion: general path
__device__ __noinline__ void quat_log_general(Quat q, double& x, double& y, double& z) {
    const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    double k;
    if (n2 < 1e-16) {
        k = 2.0 / q.w;
    } else {
        const double n = sqrt(n2);
        k = 2.0 * atan2(n, q.w) / n;
    }
    x = k * q.x; y = k * q.y; z = k * q.z;
}
// Small rotations (the residual of a consistent factor): with u = |v|/w <= 0.1 (angle <= 0.2 rad),
//   2 atan2(|v|, w)/|v| = (2/w) * atan(u)/u,  atan(u)/u = sum (-u^2)^k/(2k+1) (8 terms reach 1e-17),
// 1/w by three Newton steps from 2 - w (w >= 0.995).  No sqrt, no division, no atan2.
__device__ __forceinline__ void quat_log(Quat q, double& x, double& y, double& z) {
    if (q.w < 0.0) {
        q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z;
    }
    const double n2 = q.x * q.x + q.y * q.y + q.z * q.z;
    if (n2 > 0.0099 * q.w * q.w) {
        quat_log_general(q, x, y, z);
        return;
    }
    double r = 2.0 - q.w;
    r = r * fma(-q.w, r, 2.0);
    r = r * fma(-q.w, r, 2.0);
    r = r * fma(-q.w, r, 2.0);
    const double u2 = -n2 * r * r;
    double p = fma(u2, 1.0 / 17.0, 1.0 / 15.0);
    p = fma(u2, p, 1.0 / 13.0);
    p = fma(u2, p, 1.0 / 11.0);
    p = fma(u2, p, 1.0 / 9.0);
    p = fma(u2, p, 1.0 / 7.0);
    p = fma(u2, p, 1.0 / 5.0);
    p = fma(u2, p, 1.0 / 3.0);
    p = fma(u2, p, 1.0);
    const double k = 2.0 * r * p;
    x = k * q.x; y = k * q.y; z = k * q.z;
}
__device__ __forceinline__ void quat_rotate(const Quat& q, double vx, double vy, double vz, double& ox, double& oy,
                                            double& oz) {
    const double tx = 2.0 * (q.y * vz - q.z * vy);
    const double ty = 2.0 * (q.z * vx - q.x * vz);
    const double tz = 2.0 * (q.x * vy - q.y * vx);
    ox = vx + q.w * tx + (q.y * tz - q.z * ty);
    oy = vy + q.w * ty + (q.z * tx - q.x * tz);
    oz = vz + q.w * tz + (q.x * ty - q.y * tx);
}

// =============================================================================================
// SE(2) statistics accumulator: 16 additive values per factor
//   0..2 sum r | 3..8 sum r r' (11 12 13 22 23 33) | 9,10 sum proposal (dx,dy) | 11,12 sum cos/sin of the
//   proposal heading offset | 13..15 sum dx^2, dx dy, dy^2     (offsets from the target anchor)
// =============================================================================================
__device__ __forceinline__ void acc_res3(float (&st)[16], float m, float r1, float r2, float r3) {
    r1 *= m; r2 *= m; r3 *= m;
    st[0] += r1; st[1] += r2; st[2] += r3;
    st[3] = fmaf(r1, r1, st[3]); st[4] = fmaf(r1, r2, st[4]); st[5] = fmaf(r1, r3, st[5]);
    st[6] = fmaf(r2, r2, st[6]); st[7] = fmaf(r2, r3, st[7]); st[8] = fmaf(r3, r3, st[8]);
}
__device__ __forceinline__ void acc_prop2(float (&st)[16], float m, float dx, float dy) {
    dx *= m; dy *= m;
    st[9] += dx; st[10] += dy;
    st[13] = fmaf(dx, dx, st[13]); st[14] = fmaf(dx, dy, st[14]); st[15] = fmaf(dy, dy, st[15]);
}
__device__ __forceinline__ void acc_heading(float (&st)[16], float m, float dth) {
    float s, c;
    sincosf(dth, &s, &c);
    st[11] = fmaf(m, c, st[11]);
    st[12] = fmaf(m, s, st[12]);
}
__device__ __forceinline__ void write_stats16(float (&st)[16], float* stats, int f, int lane) {
    const float tot = warp_reduce_scatter16(st, lane);
    if ((lane & 1) == 0) stats[(size_t)f * 16 + (lane >> 1)] = tot;
}
// What a consumer warp sees of its factor: inputs already in shared memory, plus its private output slice.
struct FactorView {
    const unsigned char* b0;  // particle block of the first variable  {anchor, rows}
    const unsigned char* b1;  // particle block of the second variable (nullptr for priors)
    const float* meas;        // [dm][Npad] measurement offsets (nullptr with SAMPLE)
    float* out_res;           // [dr][Npad] residual rows, flushed by a warp-local TMA bulk store
    float* out_fwd;           // [dfwd][Npad] forward-proposal rows, same
};
constexpr uint32_t kHot1 = ROME_B200_RESIDUAL | ROME_B200_STATS;
constexpr uint32_t kHot2 = ROME_B200_RESIDUAL | ROME_B200_STATS | ROME_B200_PROPOSAL_FWD;

// =============================================================================================
// SE(2) families.  Rows are particle-major ([Npad][d], the reference's own `vecval` order): lane l owns
// particles l, l+32, l+64, ...; consecutive lanes read consecutive 12-B (8-B) records -> bank-conflict free.
// The slot loop is unrolled by four: one Philox/Box-Muller batch serves four particles of a lane and the
// Float64 chains of the four slots interleave.  kStatic != 0 fixes the output flags at compile time.
// =============================================================================================
// normals for the lane's slots [4g, 4g+4) of factor f: D normals per particle, 4 per Philox call
template <int D>
__device__ __forceinline__ void normals_for_group(const EvalParams& P, int f, int lane, int g, float (&z)[4 * D]) {
#pragma unroll
    for (int b = 0; b < D; ++b)
        normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(g * D + b), &z[4 * b]);
}

// Slot loop.  A group = the lane's 4 slots {n0, n0+32, n0+64, n0+96}.  A FULL group (its 4th slot still has live
// particles) is evaluated as branch-free straight-line code: first the four slot BODIES (loads + arithmetic into
// registers), then the four slot STORES -- no shared-memory store sits between the loads of different slots, so
// the Float64 chains of the four particles may interleave.  Slots 0-2 are live for every lane; a lane whose 4th
// particle is beyond Npad reads particle `lane` instead and has its stores/statistics masked.  FASTCOND
// (warp-uniform) selects the variant whose body may assume kFast (e.g. small heading offsets -> polynomial
// sin/cos without a fallback branch).  The trailing partial group is evaluated with warp-uniform guards.
// The family defines ROME_SLOT_DECL (per-group register arrays) and ROME_SLOT_STORE (uses k, n, live).
#define ROME_SLOT_LOOP(FASTCOND, ...)                                                          \
    for (int g = 0, n0 = lane; n0 < Npad; ++g, n0 += 128) {                                    \
        float z[4 * DZ];                                                                       \
        if (kSample) normals_for_group<DZ>(P, f, lane, g, z);                                  \
        ROME_SLOT_DECL                                                                         \
        if (n0 - lane + 96 < Npad) {                                                           \
            const int n3 = (n0 + 96 < Npad) ? n0 + 96 : lane;                                  \
            if (FASTCOND) {                                                                    \
                constexpr bool kFast = true; (void)kFast;                                      \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                \
                    const int nn = n0 + 32 * k;                                                \
                    const bool live = (k < 3) || nn < Npad;                                    \
                    const int n = (k < 3) ? nn : n3;                                           \
                    __VA_ARGS__                                                                \
                }                                                                              \
            } else {                                                                           \
                constexpr bool kFast = false; (void)kFast;                                     \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                \
                    const int nn = n0 + 32 * k;                                                \
                    const bool live = (k < 3) || nn < Npad;                                    \
                    const int n = (k < 3) ? nn : n3;                                           \
                    __VA_ARGS__                                                                \
                }                                                                              \
            }                                                                                  \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                    \
                const int nn = n0 + 32 * k;                                                    \
                const bool live = (k < 3) || nn < Npad;                                        \
                const int n = (k < 3) ? nn : n3;                                               \
                ROME_SLOT_STORE                                                                \
            }                                                                                  \
        } else {                                                                               \
            constexpr bool kFast = false; (void)kFast;                                         \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                    \
                const int nn = n0 + 32 * k;                                                    \
                if (n0 - lane + 32 * k < Npad) {                                               \
                    const bool live = nn < Npad;                                               \
                    const int n = live ? nn : lane;                                            \
                    __VA_ARGS__                                                                \
                    ROME_SLOT_STORE                                                            \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }

// SE(2) pose-valued families: residual and forward-proposal rows are 3 floats per particle
#define ROME_SLOT_DECL float o_res[4][3], o_fwd[4][3]; (void)o_res; (void)o_fwd;
#define ROME_SLOT_STORE                                                                        \
    if ((flags & ROME_B200_RESIDUAL) && live) {                                                \
        V.out_res[3 * n] = o_res[k][0]; V.out_res[3 * n + 1] = o_res[k][1]; V.out_res[3 * n + 2] = o_res[k][2]; \
    }                                                                                          \
    if ((flags & ROME_B200_PROPOSAL_FWD) && live) {                                            \
        V.out_fwd[3 * n] = o_fwd[k][0]; V.out_fwd[3 * n + 1] = o_fwd[k][1]; V.out_fwd[3 * n + 2] = o_fwd[k][2]; \
    }

struct FamPose2Pose2 {
    using Row = RowSE2;
    static constexpr int D0 = 3, D1 = 3, DM = 3, DR = 3, DFWD = 3, kMinCtas = 2;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 3;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* ap = reinterpret_cast<const double*>(V.b0);  // {x, y, theta, cos, sin}
        const double* aq = reinterpret_cast<const double*>(V.b1);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(3));
        const float* Qp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(3));
        const double apt = ap[2], ca = ap[3], sa = ap[4];
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1], dat = apt - aq[2];  // anchor deltas (exact Float64)
        const double mu0 = row.mu[0], mu1 = row.mu[1], mu2 = row.mu[2];
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;

        // warp-uniform: every heading offset of this group is small enough for the polynomial sin/cos
#define ROME_P2P2_FAST                                                                                         \
    (!__any_sync(0xffffffffu, fmaxf(fmaxf(fabsf(Pp[3 * n0 + 2]), fabsf(Pp[3 * (n0 + 32) + 2])),               \
                                    fmaxf(fabsf(Pp[3 * (n0 + 64) + 2]), fabsf(Pp[3 * n3 + 2]))) > (float)kSmallAngle))
        ROME_SLOT_LOOP(ROME_P2P2_FAST, {
            const double dpx = Pp[3 * n], dpy = Pp[3 * n + 1], dpt = Pp[3 * n + 2];
            const double dqx = Qp[3 * n], dqy = Qp[3 * n + 1], dqt = Qp[3 * n + 2];
            float mx, my, mt;
            if (!kSample) {
                mx = V.meas[3 * n]; my = V.meas[3 * n + 1]; mt = V.meas[3 * n + 2];
            } else {
                mx = row.L[0] * z[3 * k];
                my = fmaf(row.L[2], z[3 * k + 1], row.L[1] * z[3 * k]);
                mt = fmaf(row.L[5], z[3 * k + 2], fmaf(row.L[4], z[3 * k + 1], row.L[3] * z[3 * k]));
                if ((flags & ROME_B200_WRITE_MEAS) && live) {
                    float* M = P.meas_out + fo + 3 * n;
                    __stcs(M, mx); __stcs(M + 1, my); __stcs(M + 2, mt);
                }
            }
            const double Xx = mu0 + (double)mx, Xy = mu1 + (double)my, Xt = mu2 + (double)mt;
            double s, c;
            if (kFast) {  // sin/cos(anchor + small offset) by angle addition, no fallback branch
                double sx, cx;
                sincos_small(dpt, sx, cx);
                s = fma(sa, cx, ca * sx);
                c = fma(ca, cx, -sa * sx);
            } else {
                sincos_anchored(apt, ca, sa, dpt, s, c);
            }
            const double rx = c * Xx - s * Xy;  // R(theta_p) X.t
            const double ry = s * Xx + c * Xy;
            // qhat - q, Pose2D.jl:62-65 ; qhat offsets are relative to q's anchor
            const double hx = (dax + dpx) + rx, hy = (day + dpy) + ry;
            const double ht = (dat + dpt) + Xt;
            const float e1 = (float)(hx - dqx), e2 = (float)(hy - dqy), e3 = (float)wrap_pi(ht - dqt);
            const float msk = (nn < N) ? 1.f : 0.f;
            o_res[k][0] = e1; o_res[k][1] = e2; o_res[k][2] = e3;
            if (want_stats) acc_res3(st, msk, e1, e2, e3);
            if (flags & ROME_B200_PROPOSAL_FWD) {
                const float ox = (float)hx, oy = (float)hy, ot = (float)wrap_pi(ht);
                o_fwd[k][0] = ox; o_fwd[k][1] = oy; o_fwd[k][2] = ot;
                if (want_stats) { acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot); }
            }
            if (flags & ROME_B200_PROPOSAL_BWD) {
                // theta_p = theta_q - m_theta ; t_p = t_q - R(theta_p) m_t   (offsets from p's anchor)
                const double tb = (dqt - dat) - Xt;  // offset from apt
                double sb, cb;
                sincos(apt + tb, &sb, &cb);
                const float ox = (float)((dqx - dax) - (cb * Xx - sb * Xy));
                const float oy = (float)((dqy - day) - (sb * Xx + cb * Xy));
                const float ot = (float)wrap_pi(tb);
                if (live) {
                    float* B = P.prop_bwd + fo + 3 * n;
                    __stcs(B, ox); __stcs(B + 1, oy); __stcs(B + 2, ot);
                }
                if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) {
                    acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot);
                }
            }
            if ((flags & ROME_B200_JACOBIAN) && live) {  // d r/d theta_p = (-ry, rx, 1); d r/d m = R(theta_p) (+) 1
                float4* J = reinterpret_cast<float4*>(P.jac + ((size_t)f * Npad + n) * 4);
                __stcs(J, make_float4((float)(-ry), (float)rx, (float)c, (float)s));
            }
        })
#undef ROME_P2P2_FAST
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

// PriorPose2: r = (m.t - p.t, wrap(m.theta - p.theta)); proposal = the sampled point m
struct FamPriorPose2 {
    using Row = RowSE2;
    static constexpr int D0 = 3, D1 = 0, DM = 3, DR = 3, DFWD = 3, kMinCtas = 2;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 3;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(3));
        // mean relative to the variable's anchor
        const double mx0 = row.mu[0] - ap[0], my0 = row.mu[1] - ap[1], mt0 = row.mu[2] - ap[2];
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        ROME_SLOT_LOOP(true, {
            const double dpx = Pp[3 * n], dpy = Pp[3 * n + 1], dpt = Pp[3 * n + 2];
            float mx, my, mt;
            if (!kSample) {
                mx = V.meas[3 * n]; my = V.meas[3 * n + 1]; mt = V.meas[3 * n + 2];
            } else {
                mx = row.L[0] * z[3 * k];
                my = fmaf(row.L[2], z[3 * k + 1], row.L[1] * z[3 * k]);
                mt = fmaf(row.L[5], z[3 * k + 2], fmaf(row.L[4], z[3 * k + 1], row.L[3] * z[3 * k]));
                if ((flags & ROME_B200_WRITE_MEAS) && live) {
                    float* M = P.meas_out + fo + 3 * n;
                    __stcs(M, mx); __stcs(M + 1, my); __stcs(M + 2, mt);
                }
            }
            const double hx = mx0 + (double)mx, hy = my0 + (double)my, ht = mt0 + (double)mt;  // m - anchor
            const float e1 = (float)(hx - dpx), e2 = (float)(hy - dpy), e3 = (float)wrap_pi(ht - dpt);
            const float msk = (nn < N) ? 1.f : 0.f;
            o_res[k][0] = e1; o_res[k][1] = e2; o_res[k][2] = e3;
            if (want_stats) acc_res3(st, msk, e1, e2, e3);
            if (flags & ROME_B200_PROPOSAL_FWD) {
                const float ox = (float)hx, oy = (float)hy, ot = (float)wrap_pi(ht);
                o_fwd[k][0] = ox; o_fwd[k][1] = oy; o_fwd[k][2] = ot;
                if (want_stats) { acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot); }
            }
        })
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

#undef ROME_SLOT_DECL
#undef ROME_SLOT_STORE
// point-valued rows: 2 floats per particle
#define ROME_SLOT_DECL float2 o_res[4], o_fwd[4]; (void)o_res; (void)o_fwd;
#define ROME_SLOT_STORE                                                                              \
    if ((flags & ROME_B200_RESIDUAL) && live) *reinterpret_cast<float2*>(V.out_res + 2 * n) = o_res[k]; \
    if ((flags & ROME_B200_PROPOSAL_FWD) && live) *reinterpret_cast<float2*>(V.out_fwd + 2 * n) = o_fwd[k];

// Pose2Point2BearingRange: pl = R_p'(l - t_p); r = (sym_rem(b - atan(pl)), rho - |pl|)
// evaluated as atan(pl) = atan(l - t_p) - theta_p and |pl| = |l - t_p| (same values, no rotation)
struct FamBearingRange {
    using Row = RowBR;
    static constexpr int D0 = 3, D1 = 2, DM = 2, DR = 2, DFWD = 2, kMinCtas = 2;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 2;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const double* al = reinterpret_cast<const double*>(V.b1);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(3));
        const float* Lp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(2));
        const double apt = ap[2];
        const double dax = al[0] - ap[0], day = al[1] - ap[1];  // D = anchor(l) - anchor(p)
        // per factor: direction and inverse squared length of D; per particle the bearing is then
        // atan(d) = atan(D) + atan(cross(D, delta) / (|D|^2 + D.delta)) with a small second term
        const double D2 = dax * dax + day * day;
        const double iD2 = 1.0 / D2;
        const double phi0 = atan2(day, dax);
        const size_t fo = (size_t)f * 2 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        ROME_SLOT_LOOP(true, {
            const double dpx = Pp[3 * n], dpy = Pp[3 * n + 1], dpt = Pp[3 * n + 2];
            const float2 lxy = *reinterpret_cast<const float2*>(Lp + 2 * n);
            const double dlx = lxy.x, dly = lxy.y;
            float mb, mr;
            if (!kSample) {
                const float2 m2 = *reinterpret_cast<const float2*>(V.meas + 2 * n);
                mb = m2.x; mr = m2.y;
            } else {
                mb = row.sig_b * z[2 * k];
                mr = row.sig_r * z[2 * k + 1];
                if ((flags & ROME_B200_WRITE_MEAS) && live)
                    __stcs(reinterpret_cast<float2*>(P.meas_out + fo + 2 * n), make_float2(mb, mr));
            }
            const double b = row.mu_b + (double)mb, rho = row.mu_r + (double)mr;
            const double ex = dlx - dpx, ey = dly - dpy;  // delta: particle offsets (small against D)
            const double dx = dax + ex, dy = day + ey;
            const double th = apt + dpt;
            const double d2 = dx * dx + dy * dy;
            const double rng = sqrt(d2);
            const double cr = dax * ey - day * ex;        // cross(D, delta)
            const double dt = fma(dax, ex, fma(day, ey, D2));  // D.d = |D|^2 + D.delta
            double phi;
            if (fabs(dt * iD2 - 1.0) <= 0.1 && fabs(cr) <= 0.1 * dt) {
                double r = iD2;  // 1/dt by Newton from 1/|D|^2 (relative start error <= 0.1 -> 1e-16 after 4 steps)
                r = r * fma(-dt, r, 2.0);
                r = r * fma(-dt, r, 2.0);
                r = r * fma(-dt, r, 2.0);
                r = r * fma(-dt, r, 2.0);
                const double u = cr * r, u2 = -u * u;     // |u| <= 0.1: atan(u) = u * sum (-u^2)^k / (2k+1)
                double p = fma(u2, 1.0 / 17.0, 1.0 / 15.0);
                p = fma(u2, p, 1.0 / 13.0);
                p = fma(u2, p, 1.0 / 11.0);
                p = fma(u2, p, 1.0 / 9.0);
                p = fma(u2, p, 1.0 / 7.0);
                p = fma(u2, p, 1.0 / 5.0);
                p = fma(u2, p, 1.0 / 3.0);
                p = fma(u2, p, 1.0);
                phi = fma(u, p, phi0);
            } else {
                phi = atan2(dy, dx);
            }
            double e1d = wrap_pi(b + th - phi);
            if (fabs(e1d - kPi) <= 1.4901161193847656e-08 * kPi) e1d = -kPi;  // sym_rem: +pi -> -pi
            const float e1 = (float)e1d, e2 = (float)(rho - rng);
            const float msk = (nn < N) ? 1.f : 0.f;
            o_res[k] = make_float2(e1, e2);
            if (want_stats) acc_res3(st, msk, e1, e2, 0.f);
            if (flags & ROME_B200_PROPOSAL_FWD) {  // l = t_p + rho R(theta_p)(cos b, sin b) - anchor(l)
                double s, c;
                sincos(th + b, &s, &c);
                const float ox = (float)((dpx - dax) + rho * c), oy = (float)((dpy - day) + rho * s);
                o_fwd[k] = make_float2(ox, oy);
                if (want_stats) acc_prop2(st, msk, ox, oy);
            }
            if ((flags & ROME_B200_JACOBIAN) && live) {  // d r1/d l = (dy,-dx)/rho^2 ; d r2/d l = -d/rho
                const double i2 = 1.0 / d2, i1 = 1.0 / rng;
                float4* J = reinterpret_cast<float4*>(P.jac + ((size_t)f * Npad + n) * 4);
                __stcs(J, make_float4((float)(dy * i2), (float)(-dx * i2), (float)(-dx * i1), (float)(-dy * i1)));
            }
        })
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};


// ---------------------------------------------------------------------------------------------
// next-row families: 2-D Gaussian point factors.  Slot 0 is a Point2 (PriorPoint2,
// Point2Point2) or a Pose2 (Pose2Point2); slot 1 a Point2 (absent for the prior).
//   PriorPoint2   r = m - x                          src/factors/Point2D.jl:14-18
//   Point2Point2  r = m - (xj - xi)                  src/factors/Point2D.jl:30-35
//   Pose2Point2   r = l - (p.t + R_p m)              src/factors/Pose2Point2.jl:23-40
// ---------------------------------------------------------------------------------------------
template <int KIND>  // 0 prior, 1 point-point, 2 pose-point
struct FamPoint2Gauss {
    using Row = RowPT2;
    static constexpr int D0 = KIND == 2 ? 3 : 2, D1 = KIND == 0 ? 0 : 2, DM = 2, DR = 2, DFWD = 2, kMinCtas = 2;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 2;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* a0 = reinterpret_cast<const double*>(V.b0);
        const double* a1 = reinterpret_cast<const double*>(KIND == 0 ? V.b0 : V.b1);
        const float* X0 = reinterpret_cast<const float*>(V.b0 + var_header_bytes(D0));
        const float* X1 = reinterpret_cast<const float*>((KIND == 0 ? V.b0 : V.b1) + var_header_bytes(2));
        // anchor(slot 1) - anchor(slot 0) (translation part); prior: mean relative to the anchor
        const double dax = KIND == 0 ? row.mu[0] - a0[0] : a1[0] - a0[0];
        const double day = KIND == 0 ? row.mu[1] - a0[1] : a1[1] - a0[1];
        const double ca = KIND == 2 ? a0[3] : 1.0, sa = KIND == 2 ? a0[4] : 0.0, apt = KIND == 2 ? a0[2] : 0.0;
        const size_t fo = (size_t)f * 2 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        ROME_SLOT_LOOP(true, {
            float2 m2;
            if (!kSample) {
                m2 = *reinterpret_cast<const float2*>(V.meas + 2 * n);
            } else {
                m2.x = row.L[0] * z[2 * k];
                m2.y = fmaf(row.L[2], z[2 * k + 1], row.L[1] * z[2 * k]);
                if ((flags & ROME_B200_WRITE_MEAS) && live)
                    __stcs(reinterpret_cast<float2*>(P.meas_out + fo + 2 * n), m2);
            }
            float e1, e2, ox, oy;
            if (KIND == 0) {  // r = m - x ; proposal = m   (offsets from x's anchor)
                const float2 x = *reinterpret_cast<const float2*>(X0 + 2 * n);
                const double hx = dax + (double)m2.x, hy = day + (double)m2.y;
                e1 = (float)(hx - (double)x.x); e2 = (float)(hy - (double)x.y);
                ox = (float)hx; oy = (float)hy;
            } else if (KIND == 1) {  // r = m - (xj - xi) ; proposal xj = xi + m
                const float2 xi = *reinterpret_cast<const float2*>(X0 + 2 * n);
                const float2 xj = *reinterpret_cast<const float2*>(X1 + 2 * n);
                const double mx = row.mu[0] + (double)m2.x, my = row.mu[1] + (double)m2.y;
                const double hx = ((double)xi.x - dax) + mx, hy = ((double)xi.y - day) + my;  // xi + m - anchor(xj)
                e1 = (float)(hx - (double)xj.x); e2 = (float)(hy - (double)xj.y);
                ox = (float)hx; oy = (float)hy;
            } else {  // r = l - (p.t + R_p m) ; proposal l = p.t + R_p m
                const double dpx = X0[3 * n], dpy = X0[3 * n + 1], dpt = X0[3 * n + 2];
                const float2 l = *reinterpret_cast<const float2*>(X1 + 2 * n);
                const double mx = row.mu[0] + (double)m2.x, my = row.mu[1] + (double)m2.y;
                double s, c;
                sincos_anchored(apt, ca, sa, dpt, s, c);
                const double hx = (dpx - dax) + (c * mx - s * my), hy = (dpy - day) + (s * mx + c * my);
                e1 = (float)((double)l.x - hx); e2 = (float)((double)l.y - hy);
                ox = (float)hx; oy = (float)hy;
            }
            const float msk = (nn < N) ? 1.f : 0.f;
            o_res[k] = make_float2(e1, e2);
            if (want_stats) acc_res3(st, msk, e1, e2, 0.f);
            if (flags & ROME_B200_PROPOSAL_FWD) {
                o_fwd[k] = make_float2(ox, oy);
                if (want_stats) acc_prop2(st, msk, ox, oy);
            }
        })
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

#undef ROME_SLOT_DECL
#undef ROME_SLOT_STORE
// scalar rows: 1 float per particle, no closed-form proposal
#define ROME_SLOT_DECL float o_res[4]; (void)o_res;
#define ROME_SLOT_STORE \
    if ((flags & ROME_B200_RESIDUAL) && live) V.out_res[n] = o_res[k];

// scalar Normal factors.  KIND 0: Pose2Point2Range, 1: Point2Point2Range (rho - |l - x|, src/factors/Range2D.jl:14-18,
// 51-54); 2: Pose2Point2Bearing (sym_rem(b - atan(R_p'(l - p.t))), src/factors/Bearing2D.jl:23-32)
template <int KIND>
struct FamScalar {
    using Row = RowS1;
    static constexpr int D0 = KIND == 1 ? 2 : 3, D1 = 2, DM = 1, DR = 1, DFWD = 0, kMinCtas = 2;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 1;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* a0 = reinterpret_cast<const double*>(V.b0);
        const double* a1 = reinterpret_cast<const double*>(V.b1);
        const float* X0 = reinterpret_cast<const float*>(V.b0 + var_header_bytes(D0));
        const float* X1 = reinterpret_cast<const float*>(V.b1 + var_header_bytes(2));
        const double dax = a1[0] - a0[0], day = a1[1] - a0[1];
        const double apt = KIND == 2 ? a0[2] : 0.0;
        const size_t fo = (size_t)f * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        ROME_SLOT_LOOP(true, {
            float m1;
            if (!kSample) {
                m1 = V.meas[n];
            } else {
                m1 = row.sigma * z[k];
                if ((flags & ROME_B200_WRITE_MEAS) && live) __stcs(P.meas_out + fo + n, m1);
            }
            const float2 l = *reinterpret_cast<const float2*>(X1 + 2 * n);
            const double x0 = X0[D0 * n], y0 = X0[D0 * n + 1];
            const double dx = dax + ((double)l.x - x0), dy = day + ((double)l.y - y0);
            float e1;
            if (KIND == 2) {
                const double th = apt + (double)X0[3 * n + 2];
                double e = wrap_pi((row.mu + (double)m1) + th - atan2(dy, dx));
                if (fabs(e - kPi) <= 1.4901161193847656e-08 * kPi) e = -kPi;  // sym_rem: +pi -> -pi
                e1 = (float)e;
            } else {
                e1 = (float)((row.mu + (double)m1) - sqrt(dx * dx + dy * dy));
            }
            const float msk = (nn < N) ? 1.f : 0.f;
            o_res[k] = e1;
            if (want_stats) acc_res3(st, msk, e1, 0.f, 0.f);
        })
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

#undef ROME_SLOT_DECL
#undef ROME_SLOT_STORE
#undef ROME_SLOT_LOOP

// =============================================================================================
// SE(3) families: one particle per lane per iteration
//   stats[32]: 0..5 sum r | 6..26 sum r r' upper triangle (row-major) | 27..29 sum proposal dt |
//              30 sum |dt|^2 | 31 sum |r|^2
// =============================================================================================
__device__ __forceinline__ void acc_res6(float (&st)[32], float m, const float (&r)[6]) {
    float q[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { q[i] = r[i] * m; st[i] += q[i]; }
    int k = 6;
    float n2 = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = i; j < 6; ++j) { st[k] = fmaf(q[i], q[j], st[k]); ++k; }
        n2 = fmaf(q[i], q[i], n2);
    }
    st[31] += n2;
}
// 24-B particle-major records: three 8-B accesses (conflict-free per half-warp)
__device__ __forceinline__ void load6(const float* p, float (&v)[6]) {
    const float2 a = reinterpret_cast<const float2*>(p)[0], b = reinterpret_cast<const float2*>(p)[1],
                 c = reinterpret_cast<const float2*>(p)[2];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}
__device__ __forceinline__ void store6(float* p, const float (&v)[6]) {
    reinterpret_cast<float2*>(p)[0] = make_float2(v[0], v[1]);
    reinterpret_cast<float2*>(p)[1] = make_float2(v[2], v[3]);
    reinterpret_cast<float2*>(p)[2] = make_float2(v[4], v[5]);
}
__device__ __forceinline__ void store6_global(float* p, const float (&v)[6]) {
    __stcs(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
    __stcs(reinterpret_cast<float2*>(p) + 1, make_float2(v[2], v[3]));
    __stcs(reinterpret_cast<float2*>(p) + 2, make_float2(v[4], v[5]));
}
__device__ __forceinline__ void acc_prop3(float (&st)[32], float m, float x, float y, float z) {
    x *= m; y *= m; z *= m;
    st[27] += x; st[28] += y; st[29] += z;
    st[30] += x * x + y * y + z * z;
}
// measurement offsets L z of slot k of the lane's group (z: 24 normals of the group)
__device__ __forceinline__ void sample6(const RowSE3& row, const float* z, float (&d)[6]) {
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j <= i; ++j) a = fmaf(row.L[k++], z[j], a);
        d[i] = a;
    }
}

struct FamPose3Pose3 {
    using Row = RowSE3;
    static constexpr int D0 = 6, D1 = 6, DM = 6, DR = 6, DFWD = 6, kMinCtas = 1;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const double* aq = reinterpret_cast<const double*>(V.b1);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(6));
        const float* Qp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(6));
        const size_t fo = (size_t)f * 6 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        {
          for (int n = lane, slot = 0; n < Npad; n += 32, ++slot) {
            float p[6], q[6], m[6];
            load6(Pp + 6 * n, p);
            load6(Qp + 6 * n, q);
            if (!kSample) {
                load6(V.meas + 6 * n, m);
            } else {  // two Philox blocks per particle (8 normals, 6 used)
                float z[8];
                normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(2 * slot), z);
                normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(2 * slot + 1), z + 4);
                sample6(row, z, m);
            }
            double X[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) X[i] = row.mu[i] + (double)m[i];
            const Quat Rp = quat_exp(ap[3] + (double)p[3], ap[4] + (double)p[4], ap[5] + (double)p[5]);
            const Quat Rq = quat_exp(aq[3] + (double)q[3], aq[4] + (double)q[4], aq[5] + (double)q[5]);
            const Quat M = quat_exp(X[3], X[4], X[5]);
            double vx, vy, vz;
            quat_rotate(Rp, X[0], X[1], X[2], vx, vy, vz);
            // qhat.t as offset from q's anchor
            const double hx = ((ap[0] - aq[0]) + (double)p[0]) + vx;
            const double hy = ((ap[1] - aq[1]) + (double)p[1]) + vy;
            const double hz = ((ap[2] - aq[2]) + (double)p[2]) + vz;
            const Quat Rh = qmul(Rp, M);
            double wx, wy, wz;
            quat_log(qmul(qconj(Rq), Rh), wx, wy, wz);
            float r[6] = {(float)(hx - (double)q[0]), (float)(hy - (double)q[1]), (float)(hz - (double)q[2]),
                          (float)wx, (float)wy, (float)wz};
            const float msk = (n < N) ? 1.f : 0.f;
            if (want_stats) acc_res6(st, msk, r);
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) {
                store6_global(P.meas_out + fo + 6 * n, m);
            }
            if (flags & ROME_B200_RESIDUAL) {
                store6(V.out_res + 6 * n, r);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {  // q = p o Exp(X): coordinates as offsets from q's anchor
                double ox, oy, oz;
                quat_log(Rh, ox, oy, oz);
                const float o[6] = {(float)hx, (float)hy, (float)hz, (float)(ox - aq[3]), (float)(oy - aq[4]),
                                    (float)(oz - aq[5])};
                store6(V.out_fwd + 6 * n, o);
                if (want_stats) acc_prop3(st, msk, o[0], o[1], o[2]);
            }
            if (flags & ROME_B200_PROPOSAL_BWD) {  // R_p = R_q Exp(X.w)' ; t_p = t_q - R_p X.t
                const Quat Rb = qmul(Rq, qconj(M));
                double bx, by, bz, ox, oy, oz;
                quat_rotate(Rb, X[0], X[1], X[2], bx, by, bz);
                quat_log(Rb, ox, oy, oz);
                const float o[6] = {(float)(((aq[0] - ap[0]) + (double)q[0]) - bx),
                                    (float)(((aq[1] - ap[1]) + (double)q[1]) - by),
                                    (float)(((aq[2] - ap[2]) + (double)q[2]) - bz),
                                    (float)(ox - ap[3]), (float)(oy - ap[4]), (float)(oz - ap[5])};
                store6_global(P.prop_bwd + fo + 6 * n, o);
                if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) acc_prop3(st, msk, o[0], o[1], o[2]);
            }
          }
        }
        if (want_stats) {
            const float tot = warp_reduce_scatter32(st, lane);
            P.stats[(size_t)f * 32 + lane] = tot;
        }
    }
};

struct FamPriorPose3 {
    using Row = RowSE3;
    static constexpr int D0 = 6, D1 = 0, DM = 6, DR = 6, DFWD = 6, kMinCtas = 1;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = kStatic ? kStatic : P.flags;
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(6));
        const size_t fo = (size_t)f * 6 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        {
          for (int n = lane, slot = 0; n < Npad; n += 32, ++slot) {
            float p[6], m[6];
            load6(Pp + 6 * n, p);
            if (!kSample) {
                load6(V.meas + 6 * n, m);
            } else {
                float z[8];
                normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(2 * slot), z);
                normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(2 * slot + 1), z + 4);
                sample6(row, z, m);
            }
            double X[6];  // sampled point coordinates: exp(e, hat(mu + L z)) = (t, Exp(w))
#pragma unroll
            for (int i = 0; i < 6; ++i) X[i] = row.mu[i] + (double)m[i];
            const Quat Rp = quat_exp(ap[3] + (double)p[3], ap[4] + (double)p[4], ap[5] + (double)p[5]);
            const Quat Rm = quat_exp(X[3], X[4], X[5]);
            double wx, wy, wz;
            quat_log(qmul(qconj(Rp), Rm), wx, wy, wz);
            const double hx = X[0] - ap[0], hy = X[1] - ap[1], hz = X[2] - ap[2];
            float r[6] = {(float)(hx - (double)p[0]), (float)(hy - (double)p[1]), (float)(hz - (double)p[2]),
                          (float)wx, (float)wy, (float)wz};
            const float msk = (n < N) ? 1.f : 0.f;
            if (want_stats) acc_res6(st, msk, r);
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) {
                store6_global(P.meas_out + fo + 6 * n, m);
            }
            if (flags & ROME_B200_RESIDUAL) {
                store6(V.out_res + 6 * n, r);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {  // proposal = the sampled point, offsets from the anchor
                const float o[6] = {(float)hx, (float)hy, (float)hz, (float)(X[3] - ap[3]), (float)(X[4] - ap[4]),
                                    (float)(X[5] - ap[5])};
                store6(V.out_fwd + 6 * n, o);
                if (want_stats) acc_prop3(st, msk, o[0], o[1], o[2]);
            }
          }
        }
        if (want_stats) {
            const float tot = warp_reduce_scatter32(st, lane);
            P.stats[(size_t)f * 32 + lane] = tot;
        }
    }
};

// =============================================================================================
// stage layout (shared by host planning and the kernel)
// =============================================================================================
struct StageLayout {
    int rows_off, v0_off, v1_off, meas_off, bytes;
    int b0, b1, mb;  // bytes of one slot-0 block, slot-1 block, one factor's measurement block
};
__host__ __device__ inline StageLayout stage_layout(int ft, int row_bytes, int d0, int d1, int dm, bool sample,
                                                    int Npad) {
    StageLayout L;
    L.b0 = var_block_bytes(d0, Npad);
    L.b1 = d1 ? var_block_bytes(d1, Npad) : 0;
    L.mb = sample ? 0 : dm * Npad * 4;
    L.rows_off = 0;
    L.v0_off = (ft * row_bytes + 127) / 128 * 128;
    L.v1_off = L.v0_off + ft * L.b0;
    L.meas_off = L.v1_off + ft * L.b1;
    L.bytes = (L.meas_off + ft * L.mb + 127) / 128 * 128;
    return L;
}
constexpr int kBarrierBytes = 128;
constexpr int kMaxStages = 6;

// =============================================================================================
// persistent producer/consumer pipeline
//   smem: [full[], empty[] mbarriers | S input stages | FT per-warp output slices]
// =============================================================================================
template <class Fam, uint32_t kStatic, bool kSample, int FT>
__global__ void __launch_bounds__((FT + 1) * 32, Fam::kMinCtas) eval_kernel(const __grid_constant__ EvalParams P) {
    using Row = typename Fam::Row;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kMaxStages;
    unsigned char* stage0 = smem + kBarrierBytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = P.stages;
    const int nTiles = (P.count + FT - 1) / FT;
    const StageLayout L = stage_layout(FT, (int)sizeof(Row), Fam::D0, Fam::D1, Fam::DM, kSample, P.Npad);
    const Row* __restrict__ table = reinterpret_cast<const Row*>(P.rows) + P.first;
    const uint32_t flags = kStatic ? kStatic : P.flags;

    // The producer warp holds the variable ids of a CHUNK of 32/FT tiles at once (lane l <-> tile l/FT of the
    // chunk, factor l%FT): one global-load latency per chunk instead of one per tile; the first chunk is
    // requested before the barrier initialisation is published.
    constexpr int TPC = 32 / FT;  // tiles per chunk
    const int jl = lane / FT, fl_in_tile = lane % FT;
    int2 ids_cur = make_int2(0, 0);
    auto fetch_chunk = [&](int base_tile) {
        const int t = base_tile + jl * (int)gridDim.x;
        const int fl = t * FT + fl_in_tile;
        int2 ids = make_int2(0, 0);
        if (t < nTiles && fl < P.count) ids = __ldg(reinterpret_cast<const int2*>(table + fl));
        return ids;
    };
    if (warp == FT) ids_cur = fetch_chunk(blockIdx.x);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], FT);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // programmatic dependent launch: a following launch flagged ROME_B200_INDEPENDENT may begin as SMs free up
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == FT) {
        // ---------------- producer warp ---------------------------------------------------------------------
        int s = 0;
        uint32_t phase = 1;  // parity of the previous round; the first pass over the ring does not wait
        bool first_round = true;
        for (int base = blockIdx.x; base < nTiles; base += TPC * gridDim.x) {
            const int2 ids_next = fetch_chunk(base + TPC * gridDim.x);  // in flight while this chunk is issued
#pragma unroll 1
            for (int j = 0; j < TPC; ++j) {
                const int tile = base + j * gridDim.x;
                if (tile >= nTiles) break;
                if (!first_round) mbar_wait(&empty[s], phase);
                unsigned char* st = stage0 + (size_t)s * L.bytes;
                const int nf = min(FT, P.count - tile * FT);
                if (lane == j * FT) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(nf * ((int)sizeof(Row) + L.b0 + L.b1 + L.mb)));
                    tma_load_1d(st + L.rows_off, table + (size_t)tile * FT, (uint32_t)(nf * sizeof(Row)), &full[s]);
                    if (!kSample)
                        tma_load_1d(st + L.meas_off, P.meas + (size_t)(P.first + tile * FT) * Fam::DM * P.Npad,
                                    (uint32_t)(nf * L.mb), &full[s]);
                }
                __syncwarp();
                if (jl == j && fl_in_tile < nf) {
                    tma_load_1d(st + L.v0_off + fl_in_tile * L.b0, P.v0 + (size_t)ids_cur.x * L.b0, (uint32_t)L.b0,
                                &full[s]);
                    if (Fam::D1)
                        tma_load_1d(st + L.v1_off + fl_in_tile * L.b1, P.v1 + (size_t)ids_cur.y * L.b1,
                                    (uint32_t)L.b1, &full[s]);
                }
                if (++s == S) { s = 0; phase ^= 1u; first_round = false; }
            }
            ids_cur = ids_next;
        }
    } else {
        // ---------------- consumer warps: warp w owns the tile's w-th factor ----------------------------------
        float* out = reinterpret_cast<float*>(stage0 + (size_t)S * L.bytes + (size_t)warp * P.out_warp_bytes);
        const int res_floats = Fam::DR * P.Npad;
        int s = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
            mbar_wait(&full[s], phase);
            const unsigned char* st = stage0 + (size_t)s * L.bytes;
            const int fl = tile * FT + warp;
            if (fl < P.count) {
                const int f = P.first + fl;
                const Row row = reinterpret_cast<const Row*>(st + L.rows_off)[warp];
                FactorView V;
                V.b0 = st + L.v0_off + warp * L.b0;
                V.b1 = Fam::D1 ? st + L.v1_off + warp * L.b1 : nullptr;
                V.meas = kSample ? nullptr : reinterpret_cast<const float*>(st + L.meas_off + (size_t)warp * L.mb);
                V.out_res = out;
                V.out_fwd = out + res_floats;
                if (flags & (ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD)) {
                    if (lane == 0) tma_store_wait_read();  // the previous tile's rows have left the slice
                    __syncwarp();
                }
                Fam::template factor<kStatic, kSample>(row, P, V, f, lane);
                if (flags & (ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD)) {
                    fence_proxy_async();  // generic-proxy writes of the slice -> visible to the bulk-copy engine
                    __syncwarp();
                    if (lane == 0) {
                        if (flags & ROME_B200_RESIDUAL)
                            tma_store_1d(P.res + (size_t)f * res_floats, V.out_res, (uint32_t)(res_floats * 4));
                        if (flags & ROME_B200_PROPOSAL_FWD) {
                            const size_t off = (size_t)f * Fam::DFWD * P.Npad;
                            const uint32_t bytes = (uint32_t)(Fam::DFWD * P.Npad * 4);
                            tma_store_1d(P.prop_fwd + off, V.out_fwd, bytes);
                            // fused all-gather: the same slice goes to every peer GPU over NVLink
                            for (int r = 0; r < P.n_peers; ++r) tma_store_1d(P.peer_fwd[r] + off, V.out_fwd, bytes);
                        }
                        tma_store_commit();
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) { s = 0; phase ^= 1u; }
        }
        if (lane == 0) tma_store_wait_all();
    }
    // a launch that overlapped its predecessor must not be seen as complete before the predecessor is
    if (P.flags & ROME_B200_INDEPENDENT) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// =============================================================================================
// host side: launch planning + dispatch
// =============================================================================================
struct FamDims {
    int row_bytes, d0, d1, dm, dr, dfwd;
};
static FamDims fam_dims(int family) {
    switch (family) {
        case ROME_B200_POSE2POSE2: return {(int)sizeof(RowSE2), 3, 3, 3, 3, 3};
        case ROME_B200_PRIORPOSE2: return {(int)sizeof(RowSE2), 3, 0, 3, 3, 3};
        case ROME_B200_BEARINGRANGE: return {(int)sizeof(RowBR), 3, 2, 2, 2, 2};
        case ROME_B200_POSE3POSE3: return {(int)sizeof(RowSE3), 6, 6, 6, 6, 6};
        case ROME_B200_PRIORPOINT2: return {(int)sizeof(RowPT2), 2, 0, 2, 2, 2};
        case ROME_B200_POINT2POINT2: return {(int)sizeof(RowPT2), 2, 2, 2, 2, 2};
        case ROME_B200_POSE2POINT2: return {(int)sizeof(RowPT2), 3, 2, 2, 2, 2};
        case ROME_B200_POSE2POINT2RANGE: return {(int)sizeof(RowS1), 3, 2, 1, 1, 0};
        case ROME_B200_POINT2POINT2RANGE: return {(int)sizeof(RowS1), 2, 2, 1, 1, 0};
        case ROME_B200_POSE2POINT2BEARING: return {(int)sizeof(RowS1), 3, 2, 1, 1, 0};
        default: return {(int)sizeof(RowSE3), 6, 0, 6, 6, 6};
    }
}

int plan_launch(int family, uint32_t flags, int Npad, int smem_per_sm, int smem_per_cta_max, LaunchPlan* plan) {
    const FamDims fd = fam_dims(family);
    const bool sample = (flags & ROME_B200_SAMPLE) != 0;
    const bool se3 = family == ROME_B200_POSE3POSE3 || family == ROME_B200_PRIORPOSE3;
    if (fd.dfwd == 0) flags &= ~ROME_B200_PROPOSAL_FWD;
    const uint32_t out_flags = flags & ~(ROME_B200_SAMPLE | ROME_B200_INDEPENDENT);
    const int hot = out_flags == kHot1 ? 1 : out_flags == kHot2 ? 2 : 0;
    static const int fts[3] = {8, 2, 1};
    for (int k = 0; k < 3; ++k) {
        const int ft = fts[k];
        const int variant = ft == 8 ? hot : 0;  // compile-time flag variants exist for the 8-factor tile only
        // per-warp output slice: residual rows, then forward-proposal rows (the generic variant reserves both)
        const int out_warp = (fd.dr + (variant == 1 ? 0 : fd.dfwd)) * Npad * 4;
        const StageLayout L = stage_layout(ft, fd.row_bytes, fd.d0, fd.d1, fd.dm, sample, Npad);
        // prefer 2 CTAs/SM (register cap of the SE(2) kernels allows it) with >= 3 stages each, else 1 CTA/SM
        for (int ctas = (se3 ? 1 : 2); ctas >= 1; --ctas) {
            const int budget = (smem_per_sm / ctas) - 1024;  // 1 KB per CTA is reserved by the system
            const int cap = budget < smem_per_cta_max ? budget : smem_per_cta_max;
            int stages = (cap - kBarrierBytes - ft * out_warp) / L.bytes;
            if (stages > kMaxStages) stages = kMaxStages;
            const int need = ctas == 2 ? 3 : 2;
            if (stages >= need) {
                if (ctas == 2 && stages > 4) stages = 4;
                plan->ft = ft; plan->variant = variant; plan->stages = stages;
                plan->stage_bytes = L.bytes;
                plan->out_warp_bytes = out_warp;
                plan->smem_bytes = kBarrierBytes + stages * L.bytes + ft * out_warp;
                plan->ctas_per_sm = ctas;
                return 0;
            }
        }
    }
    return (int)cudaErrorInvalidConfiguration;  // N too large for the shared-memory pipeline
}

template <class Fam, uint32_t kStatic, bool kSample, int FT>
static int launch_ft(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    auto k = eval_kernel<Fam, kStatic, kSample, FT>;
    static int configured[64] = {0};  // per-instantiation, per-device cache of the opt-in shared memory size
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || plan.smem_bytes > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 64) configured[dev] = plan.smem_bytes;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((FT + 1) * 32);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (p.flags & ROME_B200_INDEPENDENT) ? 1 : 0;
    return (int)cudaLaunchKernelEx(&cfg, k, p);
}
template <class Fam, bool kSample>
static int launch_sample(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    constexpr uint32_t smp = kSample ? ROME_B200_SAMPLE : 0u;
    if (plan.ft == 8) {
        if (plan.variant == 1) return launch_ft<Fam, kHot1 | smp, kSample, 8>(p, plan, grid, s);
        if (plan.variant == 2) return launch_ft<Fam, kHot2 | smp, kSample, 8>(p, plan, grid, s);
        return launch_ft<Fam, 0u, kSample, 8>(p, plan, grid, s);
    }
    if (plan.ft == 2) return launch_ft<Fam, 0u, kSample, 2>(p, plan, grid, s);
    if (plan.ft == 1) return launch_ft<Fam, 0u, kSample, 1>(p, plan, grid, s);
    return (int)cudaErrorInvalidValue;
}
template <class Fam>
static int launch_family(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return (p.flags & ROME_B200_SAMPLE) ? launch_sample<Fam, true>(p, plan, grid, s)
                                        : launch_sample<Fam, false>(p, plan, grid, s);
}

int launch_eval(int family, const EvalParams& p, const LaunchPlan& plan, int grid, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (family) {
        case ROME_B200_POSE2POSE2: return launch_family<FamPose2Pose2>(p, plan, grid, s);
        case ROME_B200_PRIORPOSE2: return launch_family<FamPriorPose2>(p, plan, grid, s);
        case ROME_B200_BEARINGRANGE: return launch_family<FamBearingRange>(p, plan, grid, s);
        case ROME_B200_POSE3POSE3: return launch_family<FamPose3Pose3>(p, plan, grid, s);
        case ROME_B200_PRIORPOSE3: return launch_family<FamPriorPose3>(p, plan, grid, s);
        case ROME_B200_PRIORPOINT2: return launch_family<FamPoint2Gauss<0>>(p, plan, grid, s);
        case ROME_B200_POINT2POINT2: return launch_family<FamPoint2Gauss<1>>(p, plan, grid, s);
        case ROME_B200_POSE2POINT2: return launch_family<FamPoint2Gauss<2>>(p, plan, grid, s);
        case ROME_B200_POSE2POINT2RANGE: return launch_family<FamScalar<0>>(p, plan, grid, s);
        case ROME_B200_POINT2POINT2RANGE: return launch_family<FamScalar<1>>(p, plan, grid, s);
        case ROME_B200_POSE2POINT2BEARING: return launch_family<FamScalar<2>>(p, plan, grid, s);
    }
    return (int)cudaErrorInvalidValue;
}

// =============================================================================================
// layout conversion: reference layout (Float64 particle-major [var][N][d]) <-> particle store blocks
// one warp per variable; anchor = first particle; heading offsets (wrap_dim) wrapped to [-pi, pi]
// =============================================================================================
template <int D>
__global__ void pack_kernel(int nvars, int N, int Npad, int wrap_dim, const double* __restrict__ coords,
                            unsigned char* __restrict__ store) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const double* src = coords + (size_t)v * N * D;
    unsigned char* blk = store + (size_t)v * var_block_bytes(D, Npad);
    double a[D];
#pragma unroll
    for (int c = 0; c < D; ++c) a[c] = src[c];
    double* hdr = reinterpret_cast<double*>(blk);
    if (lane < var_header_bytes(D) / 8) {
        double h = lane < D ? src[lane] : 0.0;
        if (wrap_dim >= 0 && lane == D) h = cos(src[wrap_dim]);      // Pose2 header: {x, y, theta, cos, sin, 0}
        if (wrap_dim >= 0 && lane == D + 1) h = sin(src[wrap_dim]);
        hdr[lane] = h;
    }
    float* dst = reinterpret_cast<float*>(blk + var_header_bytes(D));
    for (int n = lane; n < Npad; n += 32) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double o = 0.0;
            if (n < N) {
                o = src[(size_t)n * D + c] - a[c];
                if (c == wrap_dim) o = wrap_pi(o);
            }
            dst[(size_t)n * D + c] = (float)o;
        }
    }
}
template <int D>
__global__ void unpack_kernel(int nvars, int N, int Npad, int wrap_dim, const unsigned char* __restrict__ store,
                              double* __restrict__ coords) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const unsigned char* blk = store + (size_t)v * var_block_bytes(D, Npad);
    const double* hdr = reinterpret_cast<const double*>(blk);
    const float* src = reinterpret_cast<const float*>(blk + var_header_bytes(D));
    double* dst = coords + (size_t)v * N * D;
    for (int n = lane; n < N; n += 32) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double x = hdr[c] + (double)src[(size_t)n * D + c];
            if (c == wrap_dim) x = wrap_pi(x);
            dst[(size_t)n * D + c] = x;
        }
    }
}
__global__ void adopt_kernel(int d, int Npad, unsigned char* __restrict__ store, int var,
                             const float* __restrict__ prop, int factor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float* dst = reinterpret_cast<float*>(store + (size_t)var * var_block_bytes(d, Npad) + var_header_bytes(d));
    if (i < d * Npad) dst[i] = prop[(size_t)factor * d * Npad + i];
}

int launch_pack(int d, int wrap_dim, int nvars, int N, int Npad, const double* coords, unsigned char* store,
                void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) pack_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, store);
    else if (d == 2) pack_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, store);
    else if (d == 6) pack_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, store);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_unpack(int d, int wrap_dim, int nvars, int N, int Npad, const unsigned char* store, double* coords,
                  void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) unpack_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, store, coords);
    else if (d == 2) unpack_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, store, coords);
    else if (d == 6) unpack_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, store, coords);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_adopt(int d, int Npad, unsigned char* store, int var, const float* prop, int factor, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = d * Npad;
    adopt_kernel<<<(n + 255) / 256, 256, 0, s>>>(d, Npad, store, var, prop, factor);
    return (int)cudaGetLastError();
}

}  // namespace rome
