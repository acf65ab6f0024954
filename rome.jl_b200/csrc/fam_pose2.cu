// fam_pose2.cu -- SE(2) pose-valued families (sm_100a).  Reference arithmetic (paths relative to /root/reference):
//   Pose2Pose2    src/factors/Pose2D.jl:51-67, _compose/_vee src/factors/PriorPose2.jl:19-25
//   PriorPose2    src/factors/PriorPose2.jl:37-47
#include "eval_pipeline.cuh"

namespace rome {

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
    static constexpr int D0 = 3, D1 = 3, DM = 3, DR = 3, DFWD = 3, kMinCtas = 2, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 3;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* ap = reinterpret_cast<const double*>(V.b0);  // {x, y, theta, cos, sin}
        const double* aq = reinterpret_cast<const double*>(V.b1);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(3));
        const float* Qp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(3));
        const double apt = ap[2], ca = ap[3], sa = ap[4];
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1], dat = apt - aq[2];  // anchor deltas (exact Float64)
        const double mu0 = row.mu[0], mu1 = row.mu[1], mu2 = row.mu[2];
        const size_t fo = (size_t)f * 3 * Npad;
        float* const bwd = (flags & ROME_B200_PROPOSAL_BWD) ? bwd_row(P, f, (size_t)3 * Npad) : nullptr;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;

        // ---- per-factor Float64 part of the float32 path (warp-uniform): everything LARGE is folded here -----------
        //   A   = R(anchor heading of p) mu_t                      (the lever arm; float32 copy for the small rotation)
        //   c0  = (anchor_p - anchor_q).t + A                      (translation residual of the anchors, hi + lo)
        //   th0 = wrap(anchor_p.theta + mu_theta - anchor_q.theta)
        // per particle everything left is a small offset: r_t = c0 + (dp - dq) + (R(d) - I) A + R(theta_p) m
        const bool f32ok = !(flags & (ROME_B200_PRECISE | ROME_B200_JACOBIAN | ROME_B200_DECONV));
        const double Axd = ca * mu0 - sa * mu1, Ayd = sa * mu0 + ca * mu1;
        const float Ax = (float)Axd, Ay = (float)Ayd;
        float c0x, c0xl, c0y, c0yl;
        split_f64(dax + Axd, c0x, c0xl);
        split_f64(day + Ayd, c0y, c0yl);
        const float th0 = (float)wrap_pi(dat + mu2);
        const float2 csf = *reinterpret_cast<const float2*>(ap + 5);  // float32 copies of (cos, sin) of the anchor heading
        const float caf = csf.x, saf = csf.y;

        // warp-uniform: every heading offset of this group is small enough for the polynomial sin/cos
#define ROME_P2P2_FAST                                                                                         \
    (f32ok && !__any_sync(0xffffffffu, fmaxf(fmaxf(fabsf(Pp[3 * n0 + 2]), fabsf(Pp[3 * (n0 + 32) + 2])),      \
                                             fmaxf(fabsf(Pp[3 * n2 + 2]), fabsf(Pp[3 * n3 + 2]))) > (float)kSmallAngle))
        ROME_SLOT_LOOP(ROME_P2P2_FAST, {
            float mx, my, mt;
            if (!kSample) {
                mx = V.meas[3 * n]; my = V.meas[3 * n + 1]; mt = V.meas[3 * n + 2];
            } else {
                mx = __fmul_rn(row.L[0], z[3 * k]);
                my = fmaf(row.L[2], z[3 * k + 1], row.L[1] * z[3 * k]);
                mt = fmaf(row.L[5], z[3 * k + 2], fmaf(row.L[4], z[3 * k + 1], row.L[3] * z[3 * k]));
                if ((flags & ROME_B200_WRITE_MEAS) && live) {
                    float* M = P.meas_out + fo + 3 * n;
                    __stcs(M, mx); __stcs(M + 1, my); __stcs(M + 2, mt);
                }
            }
            const float msk = (nn < N) ? 1.f : 0.f;
            if (kFast) {
                // ---------------- float32 per-particle arithmetic (default) -----------------------------------------
                const float px = Pp[3 * n], py = Pp[3 * n + 1], pt = Pp[3 * n + 2];
                const float qx = Qp[3 * n], qy = Qp[3 * n + 1], qt = Qp[3 * n + 2];
                float sd, cm1;
                sincosm1_small_f(pt, sd, cm1);
                const float lx = fmaf(cm1, Ax, -sd * Ay), ly = fmaf(cm1, Ay, sd * Ax);       // (R(d) - I) A
                const float c = fmaf(caf, cm1, fmaf(-saf, sd, caf)), s = fmaf(saf, cm1, fmaf(caf, sd, saf));
                const float rmx = fmaf(c, mx, -s * my), rmy = fmaf(s, mx, c * my);           // R(theta_p) m
                const float smx = (c0xl + lx) + rmx, smy = (c0yl + ly) + rmy;                // all the small terms
                const float ht = (pt + th0) + mt;
                const float e1 = ((px - qx) + c0x) + smx, e2 = ((py - qy) + c0y) + smy, e3 = wrap_pi_f(ht - qt);
                o_res[k][0] = e1; o_res[k][1] = e2; o_res[k][2] = e3;
                if (want_stats) acc_res3(st, msk, e1, e2, e3);
                if (flags & ROME_B200_PROPOSAL_FWD) {  // qhat as offsets from q's anchor
                    const float ox = (px + c0x) + smx, oy = (py + c0y) + smy, ot = wrap_pi_f(ht);
                    o_fwd[k][0] = ox; o_fwd[k][1] = oy; o_fwd[k][2] = ot;
                    if (want_stats) { acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot); }
                }
                if (flags & ROME_B200_PROPOSAL_BWD) {
                    // theta_p = theta_q - X_theta = anchor_p.theta + tb ; t_p = t_q - R(tb) (A + R(anchor) m)
                    // R(theta_p) X_t = R(tb) B, B = A + R(anchor) m;  t_p - anchor_p = (dq - c0) - R(anchor) m - (R(tb) - I) B
                    const float tb = (qt - th0) - mt;
                    float sb, cb1;
                    if (fabsf(tb) <= 0.78f) {
                        sincosm1_small_f(tb, sb, cb1);
                    } else {
                        float cb;
                        sincosf(tb, &sb, &cb);
                        cb1 = cb - 1.f;
                    }
                    const float r0x = fmaf(caf, mx, -saf * my), r0y = fmaf(saf, mx, caf * my);
                    const float bx = Ax + r0x, by = Ay + r0y;
                    const float ox = ((qx - c0x) - (c0xl + r0x)) - fmaf(cb1, bx, -sb * by);
                    const float oy = ((qy - c0y) - (c0yl + r0y)) - fmaf(cb1, by, sb * bx);
                    const float ot = wrap_pi_f(tb);
                    if (live) {
                        float* B = bwd + 3 * n;
                        __stcs(B, ox); __stcs(B + 1, oy); __stcs(B + 2, ot);
                    }
                    if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) {
                        acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot);
                    }
                }
            } else {
            // ---------------- Float64 per-particle arithmetic (ROME_B200_PRECISE, wide heading spreads, Jacobian,
            // deconvolution, trailing partial groups) -------------------------------------------------------------
            const double dpx = Pp[3 * n], dpy = Pp[3 * n + 1], dpt = Pp[3 * n + 2];
            const double dqx = Qp[3 * n], dqy = Qp[3 * n + 1], dqt = Qp[3 * n + 2];
            const double Xx = mu0 + (double)mx, Xy = mu1 + (double)my, Xt = mu2 + (double)mt;
            double s, c;
            sincos_anchored(apt, ca, sa, dpt, s, c);
            const double rx = c * Xx - s * Xy;  // R(theta_p) X.t
            const double ry = s * Xx + c * Xy;
            // qhat - q, Pose2D.jl:62-65 ; qhat offsets are relative to q's anchor
            const double hx = (dax + dpx) + rx, hy = (day + dpy) + ry;
            const double ht = (dat + dpt) + Xt;
            const float e1 = (float)(hx - dqx), e2 = (float)(hy - dqy), e3 = (float)wrap_pi(ht - dqt);
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
                    float* B = bwd + 3 * n;
                    __stcs(B, ox); __stcs(B + 1, oy); __stcs(B + 2, ot);
                }
                if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) {
                    acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot);
                }
            }
            if ((flags & ROME_B200_DECONV) && live) {  // X = log(p^-1 q): (R_p'(t_q - t_p), wrap(th_q - th_p)) - mu
                const double vx = (dqx - dpx) - dax, vy = (dqy - dpy) - day;
                float* M = P.meas_out + fo + 3 * n;
                __stcs(M, (float)((c * vx + s * vy) - mu0));
                __stcs(M + 1, (float)((c * vy - s * vx) - mu1));
                __stcs(M + 2, (float)wrap_pi(((dqt - dpt) - dat) - mu2));
            }
            if ((flags & ROME_B200_JACOBIAN) && live) {  // d r/d theta_p = (-ry, rx, 1); d r/d m = R(theta_p) (+) 1
                float4* J = reinterpret_cast<float4*>(P.jac + ((size_t)f * Npad + n) * 4);
                __stcs(J, make_float4((float)(-ry), (float)rx, (float)c, (float)s));
            }
            }
        })
#undef ROME_P2P2_FAST
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

// PriorPose2: r = (m.t - p.t, wrap(m.theta - p.theta)); proposal = the sampled point m
struct FamPriorPose2 {
    using Row = RowSE2;
    static constexpr int D0 = 3, D1 = 0, DM = 3, DR = 3, DFWD = 3, kMinCtas = 2, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 3;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(3));
        // mean relative to the variable's anchor
        const double mx0 = row.mu[0] - ap[0], my0 = row.mu[1] - ap[1], mt0 = row.mu[2] - ap[2];
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        // float32 path: the mean relative to the anchor is the only large quantity (per factor, Float64 -> hi + lo)
        const bool f32ok = !(flags & (ROME_B200_PRECISE | ROME_B200_DECONV));
        float m0x, m0xl, m0y, m0yl;
        split_f64(mx0, m0x, m0xl);
        split_f64(my0, m0y, m0yl);
        const float m0t = (float)wrap_pi(mt0);
        ROME_SLOT_LOOP(f32ok, {
            float mx, my, mt;
            if (!kSample) {
                mx = V.meas[3 * n]; my = V.meas[3 * n + 1]; mt = V.meas[3 * n + 2];
            } else {
                mx = __fmul_rn(row.L[0], z[3 * k]);
                my = fmaf(row.L[2], z[3 * k + 1], row.L[1] * z[3 * k]);
                mt = fmaf(row.L[5], z[3 * k + 2], fmaf(row.L[4], z[3 * k + 1], row.L[3] * z[3 * k]));
                if ((flags & ROME_B200_WRITE_MEAS) && live) {
                    float* M = P.meas_out + fo + 3 * n;
                    __stcs(M, mx); __stcs(M + 1, my); __stcs(M + 2, mt);
                }
            }
            const float msk = (nn < N) ? 1.f : 0.f;
            if (kFast) {
                const float px = Pp[3 * n], py = Pp[3 * n + 1], pt = Pp[3 * n + 2];
                const float hxs = m0xl + mx, hys = m0yl + my, ht = m0t + mt;  // m - anchor = m0 (hi) + small
                const float e1 = (m0x - px) + hxs, e2 = (m0y - py) + hys, e3 = wrap_pi_f(ht - pt);
                o_res[k][0] = e1; o_res[k][1] = e2; o_res[k][2] = e3;
                if (want_stats) acc_res3(st, msk, e1, e2, e3);
                if (flags & ROME_B200_PROPOSAL_FWD) {
                    const float ox = m0x + hxs, oy = m0y + hys, ot = wrap_pi_f(ht);
                    o_fwd[k][0] = ox; o_fwd[k][1] = oy; o_fwd[k][2] = ot;
                    if (want_stats) { acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot); }
                }
            } else {
            const double dpx = Pp[3 * n], dpy = Pp[3 * n + 1], dpt = Pp[3 * n + 2];
            const double hx = mx0 + (double)mx, hy = my0 + (double)my, ht = mt0 + (double)mt;  // m - anchor
            const float e1 = (float)(hx - dpx), e2 = (float)(hy - dpy), e3 = (float)wrap_pi(ht - dpt);
            o_res[k][0] = e1; o_res[k][1] = e2; o_res[k][2] = e3;
            if (want_stats) acc_res3(st, msk, e1, e2, e3);
            if (flags & ROME_B200_PROPOSAL_FWD) {
                const float ox = (float)hx, oy = (float)hy, ot = (float)wrap_pi(ht);
                o_fwd[k][0] = ox; o_fwd[k][1] = oy; o_fwd[k][2] = ot;
                if (want_stats) { acc_prop2(st, msk, ox, oy); acc_heading(st, msk, ot); }
            }
            if ((flags & ROME_B200_DECONV) && live) {  // the measurement that explains the particle is the particle
                float* M = P.meas_out + fo + 3 * n;
                __stcs(M, (float)(dpx - mx0)); __stcs(M + 1, (float)(dpy - my0)); __stcs(M + 2, (float)wrap_pi(dpt - mt0));
            }
            }
        })
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

#undef ROME_SLOT_DECL
#undef ROME_SLOT_STORE

int launch_pose2pose2(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return launch_family<FamPose2Pose2>(p, plan, grid, s);
}
int launch_priorpose2(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return launch_family<FamPriorPose2>(p, plan, grid, s);
}

}  // namespace rome
