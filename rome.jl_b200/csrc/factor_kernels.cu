// factor_kernels.cu -- hand-written sm_100a kernels of the RoME factor-residual hot path.
//
// One fused kernel per factor family.  For every factor of the family in [first, first+count) and
// every particle n < N it (optionally) draws the measurement (getSample), composes the SE(2)/SE(3)
// group operation, applies the measurement and writes the residual coordinates, the closed-form
// proposal(s), compact Jacobian entries and per-factor statistics.
//
//   persistent CTAs (8 warps) loop over tiles of 8 factors, one warp per factor;
//   the tile's factor-table rows {var ids, mu (f64), chol(Sigma) (f32)} are staged into shared
//   memory by a 1-D TMA bulk copy (cp.async.bulk + mbarrier), double buffered two tiles ahead;
//   particles / measurements / residuals are SoA float32 rows [.][comp][Npad]: a warp reads a row
//   as coalesced 16-B (SE(2)) or 4-B (SE(3)) lane accesses; residual rows are written with
//   streaming stores; arithmetic is Float64 on anchored float32 storage (DESIGN.md "precision");
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
__device__ __forceinline__ Quat quat_exp(double wx, double wy, double wz) {
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
__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
__device__ __forceinline__ Quat qconj(const Quat& a) { return {a.w, -a.x, -a.y, -a.z}; }
// rotation vector (angle in [0, pi]) of a unit quaternion
__device__ __forceinline__ void quat_log(Quat q, double& x, double& y, double& z) {
    if (q.w < 0.0) {
        q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z;
    }
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

// =============================================================================================
// Pose2Pose2
// =============================================================================================
struct FamPose2Pose2 {
    using Row = RowSE2;
    template <bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, int f, int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = P.flags;
        const float* __restrict__ Pp = P.v0 + (size_t)row.ip * 3 * Npad;
        const float* __restrict__ Qp = P.v1 + (size_t)row.iq * 3 * Npad;
        const double apx = P.a0[row.ip * 3], apy = P.a0[row.ip * 3 + 1], apt = P.a0[row.ip * 3 + 2];
        const double aqx = P.a1[row.iq * 3], aqy = P.a1[row.iq * 3 + 1], aqt = P.a1[row.iq * 3 + 2];
        const double dax = apx - aqx, day = apy - aqy, dat = apt - aqt;  // anchor deltas (exact Float64)
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;

        for (int base = lane * 4; base < Npad; base += 128) {
            const float4 px = ld_reuse4(Pp + base), py = ld_reuse4(Pp + Npad + base),
                         pt = ld_reuse4(Pp + 2 * Npad + base);
            const float4 qx = ld_reuse4(Qp + base), qy = ld_reuse4(Qp + Npad + base),
                         qt = ld_reuse4(Qp + 2 * Npad + base);
            float4 mx, my, mt;
            if (!kSample) {
                mx = ld_stream4(P.meas + fo + base);
                my = ld_stream4(P.meas + fo + Npad + base);
                mt = ld_stream4(P.meas + fo + 2 * Npad + base);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float z[4];
                    normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)(base + j), 0u, z);
                    f4(mx, j) = row.L[0] * z[0];
                    f4(my, j) = fmaf(row.L[2], z[1], row.L[1] * z[0]);
                    f4(mt, j) = fmaf(row.L[5], z[2], fmaf(row.L[4], z[1], row.L[3] * z[0]));
                }
            }
            float4 r1, r2, r3, fx, fy, ft, bx, by, bt, j0, j1, j2, j3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double dpx = f4c(px, j), dpy = f4c(py, j), dpt = f4c(pt, j);
                const double dqx = f4c(qx, j), dqy = f4c(qy, j), dqt = f4c(qt, j);
                const double Xx = row.mu[0] + (double)f4c(mx, j);
                const double Xy = row.mu[1] + (double)f4c(my, j);
                const double Xt = row.mu[2] + (double)f4c(mt, j);
                double s, c;
                sincos(apt + dpt, &s, &c);
                const double rx = c * Xx - s * Xy;  // R(theta_p) X.t
                const double ry = s * Xx + c * Xy;
                // qhat - q, Pose2D.jl:62-65 ; qhat offsets are relative to q's anchor
                const double hx = (dax + dpx) + rx, hy = (day + dpy) + ry;
                const double ht = (dat + dpt) + Xt;
                const float e1 = (float)(hx - dqx), e2 = (float)(hy - dqy), e3 = (float)wrap_pi(ht - dqt);
                f4(r1, j) = e1; f4(r2, j) = e2; f4(r3, j) = e3;
                const float msk = (base + j < N) ? 1.f : 0.f;
                if (want_stats) acc_res3(st, msk, e1, e2, e3);
                if (flags & ROME_B200_PROPOSAL_FWD) {
                    const float ox = (float)hx, oy = (float)hy, ot = (float)wrap_pi(ht);
                    f4(fx, j) = ox; f4(fy, j) = oy; f4(ft, j) = ot;
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
                    f4(bx, j) = ox; f4(by, j) = oy; f4(bt, j) = ot;
                    if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) {
                        acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot);
                    }
                }
                if (flags & ROME_B200_JACOBIAN) {  // d r/d theta_p = (-ry, rx, 1); d r/d m = R(theta_p) (+) 1
                    f4(j0, j) = (float)(-ry); f4(j1, j) = (float)rx; f4(j2, j) = (float)c; f4(j3, j) = (float)s;
                }
            }
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) {
                st_stream4(P.meas_out + fo + base, mx);
                st_stream4(P.meas_out + fo + Npad + base, my);
                st_stream4(P.meas_out + fo + 2 * Npad + base, mt);
            }
            if (flags & ROME_B200_RESIDUAL) {
                st_stream4(P.res + fo + base, r1);
                st_stream4(P.res + fo + Npad + base, r2);
                st_stream4(P.res + fo + 2 * Npad + base, r3);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {
                st_stream4(P.prop_fwd + fo + base, fx);
                st_stream4(P.prop_fwd + fo + Npad + base, fy);
                st_stream4(P.prop_fwd + fo + 2 * Npad + base, ft);
            }
            if (flags & ROME_B200_PROPOSAL_BWD) {
                st_stream4(P.prop_bwd + fo + base, bx);
                st_stream4(P.prop_bwd + fo + Npad + base, by);
                st_stream4(P.prop_bwd + fo + 2 * Npad + base, bt);
            }
            if (flags & ROME_B200_JACOBIAN) {
                float* J = P.jac + (size_t)f * 4 * Npad + base;
                st_stream4(J, j0); st_stream4(J + Npad, j1); st_stream4(J + 2 * Npad, j2); st_stream4(J + 3 * Npad, j3);
            }
        }
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

// =============================================================================================
// PriorPose2: r = (m.t - p.t, wrap(m.theta - p.theta)); proposal = the sampled point m
// =============================================================================================
struct FamPriorPose2 {
    using Row = RowSE2;
    template <bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, int f, int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = P.flags;
        const float* __restrict__ Pp = P.v0 + (size_t)row.ip * 3 * Npad;
        // mean relative to the variable's anchor
        const double mx0 = row.mu[0] - P.a0[row.ip * 3], my0 = row.mu[1] - P.a0[row.ip * 3 + 1],
                     mt0 = row.mu[2] - P.a0[row.ip * 3 + 2];
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        for (int base = lane * 4; base < Npad; base += 128) {
            const float4 px = ld_reuse4(Pp + base), py = ld_reuse4(Pp + Npad + base),
                         pt = ld_reuse4(Pp + 2 * Npad + base);
            float4 mx, my, mt;
            if (!kSample) {
                mx = ld_stream4(P.meas + fo + base);
                my = ld_stream4(P.meas + fo + Npad + base);
                mt = ld_stream4(P.meas + fo + 2 * Npad + base);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float z[4];
                    normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)(base + j), 0u, z);
                    f4(mx, j) = row.L[0] * z[0];
                    f4(my, j) = fmaf(row.L[2], z[1], row.L[1] * z[0]);
                    f4(mt, j) = fmaf(row.L[5], z[2], fmaf(row.L[4], z[1], row.L[3] * z[0]));
                }
            }
            float4 r1, r2, r3, fx, fy, ft;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double hx = mx0 + (double)f4c(mx, j), hy = my0 + (double)f4c(my, j),
                             ht = mt0 + (double)f4c(mt, j);  // m as offset from the anchor
                const float e1 = (float)(hx - (double)f4c(px, j)), e2 = (float)(hy - (double)f4c(py, j)),
                            e3 = (float)wrap_pi(ht - (double)f4c(pt, j));
                f4(r1, j) = e1; f4(r2, j) = e2; f4(r3, j) = e3;
                const float msk = (base + j < N) ? 1.f : 0.f;
                if (want_stats) acc_res3(st, msk, e1, e2, e3);
                if (flags & ROME_B200_PROPOSAL_FWD) {
                    const float ox = (float)hx, oy = (float)hy, ot = (float)wrap_pi(ht);
                    f4(fx, j) = ox; f4(fy, j) = oy; f4(ft, j) = ot;
                    if (want_stats) { acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot); }
                }
            }
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) {
                st_stream4(P.meas_out + fo + base, mx);
                st_stream4(P.meas_out + fo + Npad + base, my);
                st_stream4(P.meas_out + fo + 2 * Npad + base, mt);
            }
            if (flags & ROME_B200_RESIDUAL) {
                st_stream4(P.res + fo + base, r1);
                st_stream4(P.res + fo + Npad + base, r2);
                st_stream4(P.res + fo + 2 * Npad + base, r3);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {
                st_stream4(P.prop_fwd + fo + base, fx);
                st_stream4(P.prop_fwd + fo + Npad + base, fy);
                st_stream4(P.prop_fwd + fo + 2 * Npad + base, ft);
            }
        }
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

// =============================================================================================
// Pose2Point2BearingRange: pl = R_p'(l - t_p); r = (sym_rem(b - atan(pl)), rho - |pl|)
// evaluated as atan(pl) = atan(l - t_p) - theta_p and |pl| = |l - t_p| (same values, no rotation)
// =============================================================================================
struct FamBearingRange {
    using Row = RowBR;
    template <bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, int f, int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = P.flags;
        const float* __restrict__ Pp = P.v0 + (size_t)row.ip * 3 * Npad;
        const float* __restrict__ Lp = P.v1 + (size_t)row.il * 2 * Npad;
        const double apt = P.a0[row.ip * 3 + 2];
        const double dax = P.a1[row.il * 2] - P.a0[row.ip * 3];  // anchor(l) - anchor(p)
        const double day = P.a1[row.il * 2 + 1] - P.a0[row.ip * 3 + 1];
        const size_t fo = (size_t)f * 2 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        for (int base = lane * 4; base < Npad; base += 128) {
            const float4 px = ld_reuse4(Pp + base), py = ld_reuse4(Pp + Npad + base),
                         pt = ld_reuse4(Pp + 2 * Npad + base);
            const float4 lx = ld_reuse4(Lp + base), ly = ld_reuse4(Lp + Npad + base);
            float4 mb, mr;
            if (!kSample) {
                mb = ld_stream4(P.meas + fo + base);
                mr = ld_stream4(P.meas + fo + Npad + base);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // two independent scalar draws, BearingRange2D.jl:23
                    float z[4];
                    normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)(base + j), 0u, z);
                    f4(mb, j) = row.sig_b * z[0];
                    f4(mr, j) = row.sig_r * z[1];
                }
            }
            float4 r1, r2, fx, fy, j0, j1, j2, j3;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double b = row.mu_b + (double)f4c(mb, j), rho = row.mu_r + (double)f4c(mr, j);
                const double dx = dax + ((double)f4c(lx, j) - (double)f4c(px, j));
                const double dy = day + ((double)f4c(ly, j) - (double)f4c(py, j));
                const double th = apt + (double)f4c(pt, j);
                const double d2 = dx * dx + dy * dy;
                const double rng = sqrt(d2);
                double e1d = wrap_pi(b + th - atan2(dy, dx));
                if (fabs(e1d - kPi) <= 1.4901161193847656e-08 * kPi) e1d = -kPi;  // sym_rem: +pi -> -pi
                const float e1 = (float)e1d, e2 = (float)(rho - rng);
                f4(r1, j) = e1; f4(r2, j) = e2;
                const float msk = (base + j < N) ? 1.f : 0.f;
                if (want_stats) acc_res3(st, msk, e1, e2, 0.f);
                if (flags & ROME_B200_PROPOSAL_FWD) {  // l = t_p + rho R(theta_p)(cos b, sin b), offset from l's anchor
                    double s, c;
                    sincos(th + b, &s, &c);
                    const float ox = (float)(((double)f4c(px, j) - dax) + rho * c);
                    const float oy = (float)(((double)f4c(py, j) - day) + rho * s);
                    f4(fx, j) = ox; f4(fy, j) = oy;
                    if (want_stats) acc_prop2(st, msk, ox, oy);
                }
                if (flags & ROME_B200_JACOBIAN) {  // d r1/d l = (dy,-dx)/rho^2 ; d r2/d l = -d/rho
                    const double i2 = 1.0 / d2, i1 = 1.0 / rng;
                    f4(j0, j) = (float)(dy * i2); f4(j1, j) = (float)(-dx * i2);
                    f4(j2, j) = (float)(-dx * i1); f4(j3, j) = (float)(-dy * i1);
                }
            }
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) {
                st_stream4(P.meas_out + fo + base, mb);
                st_stream4(P.meas_out + fo + Npad + base, mr);
            }
            if (flags & ROME_B200_RESIDUAL) {
                st_stream4(P.res + fo + base, r1);
                st_stream4(P.res + fo + Npad + base, r2);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {
                st_stream4(P.prop_fwd + fo + base, fx);
                st_stream4(P.prop_fwd + fo + Npad + base, fy);
            }
            if (flags & ROME_B200_JACOBIAN) {
                float* J = P.jac + (size_t)f * 4 * Npad + base;
                st_stream4(J, j0); st_stream4(J + Npad, j1); st_stream4(J + 2 * Npad, j2); st_stream4(J + 3 * Npad, j3);
            }
        }
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

// =============================================================================================
// SE(3) families: one particle per lane per iteration (coalesced 4-B lane accesses)
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
__device__ __forceinline__ void acc_prop3(float (&st)[32], float m, float x, float y, float z) {
    x *= m; y *= m; z *= m;
    st[27] += x; st[28] += y; st[29] += z;
    st[30] += x * x + y * y + z * z;
}
__device__ __forceinline__ void sample6(const RowSE3& row, const EvalParams& P, int f, int n, float (&d)[6]) {
    float z[8];
    normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)n, 0u, z);
    normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)n, 1u, z + 4);
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
    template <bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, int f, int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = P.flags;
        const float* __restrict__ Pp = P.v0 + (size_t)row.ip * 6 * Npad;
        const float* __restrict__ Qp = P.v1 + (size_t)row.iq * 6 * Npad;
        double ap[6], aq[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { ap[i] = P.a0[row.ip * 6 + i]; aq[i] = P.a1[row.iq * 6 + i]; }
        const size_t fo = (size_t)f * 6 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int n = lane; n < Npad; n += 32) {
            float p[6], q[6], m[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { p[i] = __ldg(Pp + i * Npad + n); q[i] = __ldg(Qp + i * Npad + n); }
            if (!kSample) {
#pragma unroll
                for (int i = 0; i < 6; ++i) m[i] = __ldcs(P.meas + fo + i * Npad + n);
            } else {
                sample6(row, P, f, n, m);
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
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.meas_out + fo + i * Npad + n, m[i]);
            }
            if (flags & ROME_B200_RESIDUAL) {
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.res + fo + i * Npad + n, r[i]);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {  // q = p o Exp(X): coordinates as offsets from q's anchor
                double ox, oy, oz;
                quat_log(Rh, ox, oy, oz);
                const float o[6] = {(float)hx, (float)hy, (float)hz, (float)(ox - aq[3]), (float)(oy - aq[4]),
                                    (float)(oz - aq[5])};
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.prop_fwd + fo + i * Npad + n, o[i]);
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
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.prop_bwd + fo + i * Npad + n, o[i]);
                if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) acc_prop3(st, msk, o[0], o[1], o[2]);
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
    template <bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, int f, int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = P.flags;
        const float* __restrict__ Pp = P.v0 + (size_t)row.ip * 6 * Npad;
        double ap[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) ap[i] = P.a0[row.ip * 6 + i];
        const size_t fo = (size_t)f * 6 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int n = lane; n < Npad; n += 32) {
            float p[6], m[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) p[i] = __ldg(Pp + i * Npad + n);
            if (!kSample) {
#pragma unroll
                for (int i = 0; i < 6; ++i) m[i] = __ldcs(P.meas + fo + i * Npad + n);
            } else {
                sample6(row, P, f, n, m);
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
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.meas_out + fo + i * Npad + n, m[i]);
            }
            if (flags & ROME_B200_RESIDUAL) {
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.res + fo + i * Npad + n, r[i]);
            }
            if (flags & ROME_B200_PROPOSAL_FWD) {  // proposal = the sampled point, offsets from the anchor
                const float o[6] = {(float)hx, (float)hy, (float)hz, (float)(X[3] - ap[3]), (float)(X[4] - ap[4]),
                                    (float)(X[5] - ap[5])};
#pragma unroll
                for (int i = 0; i < 6; ++i) __stcs(P.prop_fwd + fo + i * Npad + n, o[i]);
                if (want_stats) acc_prop3(st, msk, o[0], o[1], o[2]);
            }
        }
        if (want_stats) {
            const float tot = warp_reduce_scatter32(st, lane);
            P.stats[(size_t)f * 32 + lane] = tot;
        }
    }
};

// =============================================================================================
// persistent tile loop with TMA-staged factor rows
// =============================================================================================
template <class Fam, bool kSample>
__global__ void __launch_bounds__(kThreads) eval_kernel(const __grid_constant__ EvalParams P) {
    using Row = typename Fam::Row;
    __shared__ alignas(128) Row rows[2][kWarpsPerCta];
    __shared__ alignas(8) uint64_t full[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nTiles = (P.count + kWarpsPerCta - 1) / kWarpsPerCta;
    const Row* __restrict__ table = reinterpret_cast<const Row*>(P.rows) + P.first;

    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int tile, int buf) {
        const int nrows = min(kWarpsPerCta, P.count - tile * kWarpsPerCta);
        const uint32_t bytes = (uint32_t)(nrows * sizeof(Row));
        mbar_arrive_expect_tx(&full[buf], bytes);
        tma_load_1d(&rows[buf][0], table + (size_t)tile * kWarpsPerCta, bytes, &full[buf]);
    };
    if (threadIdx.x == 0) {
        if ((int)blockIdx.x < nTiles) issue(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < nTiles) issue(blockIdx.x + gridDim.x, 1);
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&full[buf], (uint32_t)((it >> 1) & 1));
        const int fl = tile * kWarpsPerCta + warp;
        const bool valid = fl < P.count;
        Row row;
        if (valid) row = rows[buf][warp];
        __syncthreads();  // every warp holds its row in registers: the buffer may be refilled
        if (threadIdx.x == 0) {
            const int nt = tile + 2 * gridDim.x;
            if (nt < nTiles) {
                fence_proxy_async();
                issue(nt, buf);
            }
        }
        if (valid) Fam::template factor<kSample>(row, P, P.first + fl, lane);
    }
}

template <class Fam>
static int launch_family(const EvalParams& p, int grid, cudaStream_t s) {
    if (p.flags & ROME_B200_SAMPLE)
        eval_kernel<Fam, true><<<grid, kThreads, 0, s>>>(p);
    else
        eval_kernel<Fam, false><<<grid, kThreads, 0, s>>>(p);
    return (int)cudaGetLastError();
}

int launch_eval(int family, const EvalParams& p, int grid, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (family) {
        case ROME_B200_POSE2POSE2: return launch_family<FamPose2Pose2>(p, grid, s);
        case ROME_B200_PRIORPOSE2: return launch_family<FamPriorPose2>(p, grid, s);
        case ROME_B200_BEARINGRANGE: return launch_family<FamBearingRange>(p, grid, s);
        case ROME_B200_POSE3POSE3: return launch_family<FamPose3Pose3>(p, grid, s);
        case ROME_B200_PRIORPOSE3: return launch_family<FamPriorPose3>(p, grid, s);
    }
    return (int)cudaErrorInvalidValue;
}

template <class Fam>
static int occ_family(bool sample) {
    int n = 0;
    cudaError_t e = sample ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, eval_kernel<Fam, true>, kThreads, 0)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, eval_kernel<Fam, false>, kThreads, 0);
    return e == cudaSuccess ? n : 1;
}
int max_resident_ctas(int family, bool sample) {
    switch (family) {
        case ROME_B200_POSE2POSE2: return occ_family<FamPose2Pose2>(sample);
        case ROME_B200_PRIORPOSE2: return occ_family<FamPriorPose2>(sample);
        case ROME_B200_BEARINGRANGE: return occ_family<FamBearingRange>(sample);
        case ROME_B200_POSE3POSE3: return occ_family<FamPose3Pose3>(sample);
        case ROME_B200_PRIORPOSE3: return occ_family<FamPriorPose3>(sample);
    }
    return 1;
}

// =============================================================================================
// layout conversion: reference layout (Float64 particle-major [var][N][d]) <-> anchored float32 SoA
// one warp per variable; anchor = first particle; heading offsets (wrap_dim) wrapped to (-pi, pi]
// =============================================================================================
template <int D>
__global__ void pack_kernel(int nvars, int N, int Npad, int wrap_dim, const double* __restrict__ coords,
                            float* __restrict__ offsets, double* __restrict__ anchors) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const double* src = coords + (size_t)v * N * D;
    double a[D];
#pragma unroll
    for (int c = 0; c < D; ++c) a[c] = src[c];
    if (lane < D) anchors[(size_t)v * D + lane] = src[lane];
    float* dst = offsets + (size_t)v * D * Npad;
    for (int n = lane; n < Npad; n += 32) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double o = 0.0;
            if (n < N) {
                o = src[(size_t)n * D + c] - a[c];
                if (c == wrap_dim) o = wrap_pi(o);
            }
            dst[(size_t)c * Npad + n] = (float)o;
        }
    }
}
template <int D>
__global__ void unpack_kernel(int nvars, int N, int Npad, int wrap_dim, const float* __restrict__ offsets,
                              const double* __restrict__ anchors, double* __restrict__ coords) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const float* src = offsets + (size_t)v * D * Npad;
    double* dst = coords + (size_t)v * N * D;
    for (int n = lane; n < N; n += 32) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double x = anchors[(size_t)v * D + c] + (double)src[(size_t)c * Npad + n];
            if (c == wrap_dim) x = wrap_pi(x);
            dst[(size_t)n * D + c] = x;
        }
    }
}
__global__ void adopt_kernel(int d, int Npad, float* __restrict__ offsets, int var, const float* __restrict__ prop,
                             int factor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d * Npad) offsets[(size_t)var * d * Npad + i] = prop[(size_t)factor * d * Npad + i];
}

int launch_pack(int d, int wrap_dim, int nvars, int N, int Npad, const double* coords, float* offsets, double* anchors,
                void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) pack_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, offsets, anchors);
    else if (d == 2) pack_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, offsets, anchors);
    else if (d == 6) pack_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, offsets, anchors);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_unpack(int d, int wrap_dim, int nvars, int N, int Npad, const float* offsets, const double* anchors,
                  double* coords, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) unpack_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, offsets, anchors, coords);
    else if (d == 2) unpack_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, offsets, anchors, coords);
    else if (d == 6) unpack_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, offsets, anchors, coords);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_adopt(int d, int Npad, float* offsets, int var, const float* prop, int factor, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = d * Npad;
    adopt_kernel<<<(n + 255) / 256, 256, 0, s>>>(d, Npad, offsets, var, prop, factor);
    return (int)cudaGetLastError();
}

}  // namespace rome
