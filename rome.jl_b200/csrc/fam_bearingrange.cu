// fam_bearingrange.cu -- Pose2Point2BearingRange (sm_100a).
// Reference: src/factors/BearingRange2D.jl:48-64 (getSample :17-27)
#include "eval_pipeline.cuh"

namespace rome {

// point-valued rows: 2 floats per particle
#define ROME_SLOT_DECL float2 o_res[4], o_fwd[4]; (void)o_res; (void)o_fwd;
#define ROME_SLOT_STORE                                                                              \
    if ((flags & ROME_B200_RESIDUAL) && live) *reinterpret_cast<float2*>(V.out_res + 2 * n) = o_res[k]; \
    if ((flags & ROME_B200_PROPOSAL_FWD) && live) *reinterpret_cast<float2*>(V.out_fwd + 2 * n) = o_fwd[k];

// Pose2Point2BearingRange: pl = R_p'(l - t_p); r = (sym_rem(b - atan(pl)), rho - |pl|)
// evaluated as atan(pl) = atan(l - t_p) - theta_p and |pl| = |l - t_p| (same values, no rotation)
struct FamBearingRange {
    using Row = RowBR;
    static constexpr int D0 = 3, D1 = 2, DM = 2, DR = 2, DFWD = 2, kMinCtas = 2, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 2;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
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
        float* const bwd = (flags & ROME_B200_PROPOSAL_BWD) ? bwd_row(P, f, (size_t)3 * Npad) : nullptr;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        // proposal heading: sin/cos(theta_p + b) = angle addition from the per-factor sin/cos(anchor heading + mu_b)
        double sA = 0.0, cA = 1.0;
        if (flags & (ROME_B200_PROPOSAL_FWD | ROME_B200_PROPOSAL_BWD)) sincos(apt + row.mu_b, &sA, &cA);
        // Warp-uniform fast path: every particle pair of the group has |delta| <= 0.09 |D| (screened in float32), so
        //   |u| = |cross(D,delta)| / (D.d) <= 0.099 (series) and D.d / |D|^2 in [0.91, 1.09] (Newton start) hold for all
        //   lanes and the body is straight-line code: no atan2, no sqrt, no division, no per-particle branch.
        const float fast_lim = (D2 > 1e-20 && D2 < 1e30) ? (float)(0.0081 * D2) : -1.f;
        auto delta2 = [&](int n_) {
            const float ex = Lp[2 * n_] - Pp[3 * n_], ey = Lp[2 * n_ + 1] - Pp[3 * n_ + 1];
            return fmaf(ex, ex, ey * ey);
        };
        // ---- per-factor Float64 part of the float32 path (warp-uniform) ----------------------------------------------
        //   u = D/|D|, beta0 = wrap(mu_b + anchor heading - atan(D)), rho0 = mu_rho - |D|; per particle, with
        //   e = dl - dp (small), (along, cross) = (u.e, u x e):  atan(d) - atan(D) = atan(cross / (|D| + along)) and
        //   |d| - |D| = along + cross * x * (1/2 - x^2/8 + ...), x = cross / (|D| + along)  -- all small quantities
        const bool f32ok = !(flags & (ROME_B200_PRECISE | ROME_B200_JACOBIAN | ROME_B200_DECONV));
        const double Dn = sqrt(D2), iDn = 1.0 / Dn;
        const float ux = (float)(dax * iDn), uy = (float)(day * iDn), Dnf = (float)Dn;
        const float beta0 = (float)wrap_pi((row.mu_b + apt) - phi0);
        const float rho0 = (float)(row.mu_r - Dn), murf = (float)row.mu_r;
        // proposals: rho dir - D = (mu_rho dir0 - D) + m_rho dir0 + rho (R(zeta) - I) dir0, dir0 = (cA, sA)
        const float cAf = (float)cA, sAf = (float)sA;
        float g0x, g0xl, g0y, g0yl;
        split_f64(row.mu_r * cA - dax, g0x, g0xl);
        split_f64(row.mu_r * sA - day, g0y, g0yl);
#define ROME_BR_FAST                                                                                       \
    (f32ok && !__any_sync(0xffffffffu, fmaxf(fmaxf(delta2(n0), delta2(n0 + 32)), fmaxf(delta2(n2), delta2(n3))) > fast_lim))
        ROME_SLOT_LOOP(ROME_BR_FAST, {
            const float2 lxy = *reinterpret_cast<const float2*>(Lp + 2 * n);
            float mb, mr;
            if (!kSample) {
                const float2 m2 = *reinterpret_cast<const float2*>(V.meas + 2 * n);
                mb = m2.x; mr = m2.y;
            } else {
                mb = __fmul_rn(row.sig_b, z[2 * k]);
                mr = __fmul_rn(row.sig_r, z[2 * k + 1]);
                if ((flags & ROME_B200_WRITE_MEAS) && live)
                    __stcs(reinterpret_cast<float2*>(P.meas_out + fo + 2 * n), make_float2(mb, mr));
            }
            const float msk = (nn < N) ? 1.f : 0.f;
            if (kFast) {
                // ---------------- float32 per-particle arithmetic (default) -----------------------------------------
                const float px = Pp[3 * n], py = Pp[3 * n + 1], pt = Pp[3 * n + 2];
                const float exf = lxy.x - px, eyf = lxy.y - py;
                const float al = fmaf(ux, exf, uy * eyf), cr = fmaf(ux, eyf, -uy * exf);
                const float x = __fdividef(cr, Dnf + al), q = x * x;  // |x| <= 0.1 under the fast condition
                float pa = fmaf(q, 1.f / 9.f, -1.f / 7.f);
                pa = fmaf(q, pa, 1.f / 5.f);
                pa = fmaf(q, pa, -1.f / 3.f);
                const float dphi = fmaf(x * q, pa, x);                                        // atan(x)
                const float ps = fmaf(q, fmaf(q, fmaf(q, -5.f / 128.f, 1.f / 16.f), -0.125f), 0.5f);
                const float dr = fmaf(cr * x, ps, al);                                        // |d| - |D|
                float e1 = wrap_pi_f(((beta0 + pt) + mb) - dphi);
                if (fabsf(e1 - 3.14159274f) <= 2.4e-7f) e1 = -3.14159274f;                    // sym_rem: +pi -> -pi
                const float e2 = (rho0 + mr) - dr;
                o_res[k] = make_float2(e1, e2);
                if (want_stats) acc_res3(st, msk, e1, e2, 0.f);
                if (flags & (ROME_B200_PROPOSAL_FWD | ROME_B200_PROPOSAL_BWD)) {
                    float sz, cz1;
                    const float zeta = pt + mb;
                    if (fabsf(zeta) <= 0.78f) {
                        sincosm1_small_f(zeta, sz, cz1);
                    } else {
                        float cz;
                        sincosf(zeta, &sz, &cz);
                        cz1 = cz - 1.f;
                    }
                    const float rho = murf + mr;
                    const float gx = (g0xl + mr * cAf) + rho * fmaf(cz1, cAf, -sz * sAf);   // small part of rho dir - D
                    const float gy = (g0yl + mr * sAf) + rho * fmaf(cz1, sAf, sz * cAf);
                    if (flags & ROME_B200_PROPOSAL_FWD) {
                        const float ox = (px + g0x) + gx, oy = (py + g0y) + gy;
                        o_fwd[k] = make_float2(ox, oy);
                        if (want_stats) acc_prop2(st, msk, ox, oy);
                    }
                    if (flags & ROME_B200_PROPOSAL_BWD) {
                        const float ox = (lxy.x - g0x) - gx, oy = (lxy.y - g0y) - gy;
                        if (live) {
                            float* B = bwd + 3 * n;
                            __stcs(B, ox); __stcs(B + 1, oy); __stcs(B + 2, pt);
                        }
                        if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) acc_prop2(st, msk, ox, oy);
                    }
                }
            } else {
            // ---------------- Float64 per-particle arithmetic (ROME_B200_PRECISE, landmark close to the pose,
            // Jacobian, deconvolution, trailing partial groups) ------------------------------------------------------
            const double dpx = Pp[3 * n], dpy = Pp[3 * n + 1], dpt = Pp[3 * n + 2];
            const double dlx = lxy.x, dly = lxy.y;
            const double b = row.mu_b + (double)mb, rho = row.mu_r + (double)mr;
            const double ex = dlx - dpx, ey = dly - dpy;  // delta: particle offsets (small against D)
            const double dx = dax + ex, dy = day + ey;
            const double th = apt + dpt;
            const double d2 = fma(dx, dx, dy * dy);
            const double cr = dax * ey - day * ex;             // cross(D, delta)
            const double dt = fma(dax, ex, fma(day, ey, D2));  // D.d = |D|^2 + D.delta
            double phi, rng;
            if (fabs(dt * iD2 - 1.0) <= 0.1 && fabs(cr) <= 0.1 * dt) {
                double r = iD2;  // 1/dt by Newton from 1/|D|^2 (relative start error <= 0.1 -> 1e-16 after 4 steps)
                r = r * fma(-dt, r, 2.0);
                r = r * fma(-dt, r, 2.0);
                r = r * fma(-dt, r, 2.0);
                r = r * fma(-dt, r, 2.0);
                const double u = cr * r, u2 = -u * u;  // |u| <= 0.1: atan(u) = u * sum (-u^2)^k / (2k+1)
                double p = fma(u2, kOddInv[8], kOddInv[7]);
                p = fma(u2, p, kOddInv[6]);
                p = fma(u2, p, kOddInv[5]);
                p = fma(u2, p, kOddInv[4]);
                p = fma(u2, p, kOddInv[3]);
                p = fma(u2, p, kOddInv[2]);
                p = fma(u2, p, kOddInv[1]);
                p = fma(u2, p, 1.0);
                phi = fma(u, p, phi0);
            } else {
                phi = atan2(dy, dx);
            }
            rng = sqrt(d2);
            double e1d = wrap_pi(b + th - phi);
            if (fabs(e1d - kPi) <= 1.4901161193847656e-08 * kPi) e1d = -kPi;  // sym_rem: +pi -> -pi
            const float e1 = (float)e1d, e2 = (float)(rho - rng);
            o_res[k] = make_float2(e1, e2);
            if (want_stats) acc_res3(st, msk, e1, e2, 0.f);
            if (flags & (ROME_B200_PROPOSAL_FWD | ROME_B200_PROPOSAL_BWD)) {
                double s, c;
                sincos_anchored(apt + row.mu_b, cA, sA, dpt + (double)mb, s, c);
                if (flags & ROME_B200_PROPOSAL_FWD) {  // l = t_p + rho R(theta_p)(cos b, sin b) - anchor(l)
                    const float ox = (float)((dpx - dax) + rho * c), oy = (float)((dpy - day) + rho * s);
                    o_fwd[k] = make_float2(ox, oy);
                    if (want_stats) acc_prop2(st, msk, ox, oy);
                }
                if (flags & ROME_B200_PROPOSAL_BWD) {
                    // pose from landmark: the residual has a 1-parameter family of roots; the member that keeps the
                    // particle's current heading is t_p = l - rho R(theta_p)(cos b, sin b)   (offsets from anchor(p))
                    const float ox = (float)((dlx + dax) - rho * c), oy = (float)((dly + day) - rho * s);
                    if (live) {
                        float* B = bwd + 3 * n;
                        __stcs(B, ox); __stcs(B + 1, oy); __stcs(B + 2, (float)dpt);
                    }
                    if (want_stats && !(flags & ROME_B200_PROPOSAL_FWD)) acc_prop2(st, msk, ox, oy);
                }
            }
            if ((flags & ROME_B200_DECONV) && live)  // (bearing, range) of the landmark seen from the pose, minus the means
                __stcs(reinterpret_cast<float2*>(P.meas_out + fo + 2 * n),
                       make_float2((float)wrap_pi((phi - th) - row.mu_b), (float)(rng - row.mu_r)));
            if ((flags & ROME_B200_JACOBIAN) && live) {  // d r1/d l = (dy,-dx)/rho^2 ; d r2/d l = -d/rho
                const double i2 = 1.0 / d2, i1 = 1.0 / rng;
                float4* J = reinterpret_cast<float4*>(P.jac + ((size_t)f * Npad + n) * 4);
                __stcs(J, make_float4((float)(dy * i2), (float)(-dx * i2), (float)(-dx * i1), (float)(-dy * i1)));
            }
            }
        })
#undef ROME_BR_FAST
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

#undef ROME_SLOT_DECL
#undef ROME_SLOT_STORE

int launch_bearingrange(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return launch_family<FamBearingRange>(p, plan, grid, s);
}

}  // namespace rome
