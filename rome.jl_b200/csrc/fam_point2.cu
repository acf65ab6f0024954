// fam_point2.cu -- 2-D point / scalar families, sm_100a.
#include "eval_pipeline.cuh"

namespace rome {

// point-valued rows: 2 floats per particle
#define ROME_SLOT_DECL float2 o_res[4], o_fwd[4]; (void)o_res; (void)o_fwd;
#define ROME_SLOT_STORE                                                                              \
    if ((flags & ROME_B200_RESIDUAL) && live) *reinterpret_cast<float2*>(V.out_res + 2 * n) = o_res[k]; \
    if ((flags & ROME_B200_PROPOSAL_FWD) && live) *reinterpret_cast<float2*>(V.out_fwd + 2 * n) = o_fwd[k];


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
    static constexpr int D0 = KIND == 2 ? 3 : 2, D1 = KIND == 0 ? 0 : 2, DM = 2, DR = 2, DFWD = 2, kMinCtas = 2, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 2;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
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
                m2.x = __fmul_rn(row.L[0], z[2 * k]);
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
    static constexpr int D0 = KIND == 1 ? 2 : 3, D1 = 2, DM = 1, DR = 1, DFWD = 0, kMinCtas = 2, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 1;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
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
                m1 = __fmul_rn(row.sigma, z[k]);
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

int launch_point2(int family, const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    switch (family) {
        case ROME_B200_PRIORPOINT2: return launch_family<FamPoint2Gauss<0>>(p, plan, grid, s);
        case ROME_B200_POINT2POINT2: return launch_family<FamPoint2Gauss<1>>(p, plan, grid, s);
        case ROME_B200_POSE2POINT2: return launch_family<FamPoint2Gauss<2>>(p, plan, grid, s);
        case ROME_B200_POSE2POINT2RANGE: return launch_family<FamScalar<0>>(p, plan, grid, s);
        case ROME_B200_POINT2POINT2RANGE: return launch_family<FamScalar<1>>(p, plan, grid, s);
        case ROME_B200_POSE2POINT2BEARING: return launch_family<FamScalar<2>>(p, plan, grid, s);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace rome
