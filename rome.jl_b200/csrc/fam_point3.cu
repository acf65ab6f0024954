// fam_point3.cu -- 3-D point families, sm_100a.
//   PriorPoint3   r = m - x              src/factors/Point3D.jl:13-20
//   Point3Point3  r = m - (xj - xi)      src/factors/Point3Point3.jl:11-15
// Point3 blocks use the 48-B header {x, y, z, 0, 0, 0} followed by [Npad][3] float32 offsets.
#include "eval_pipeline.cuh"

namespace rome {

#define ROME_SLOT_DECL float o_res[4][3], o_fwd[4][3]; (void)o_res; (void)o_fwd;
#define ROME_SLOT_STORE                                                                        \
    if ((flags & ROME_B200_RESIDUAL) && live) {                                                \
        V.out_res[3 * n] = o_res[k][0]; V.out_res[3 * n + 1] = o_res[k][1]; V.out_res[3 * n + 2] = o_res[k][2]; \
    }                                                                                          \
    if ((flags & ROME_B200_PROPOSAL_FWD) && live) {                                            \
        V.out_fwd[3 * n] = o_fwd[k][0]; V.out_fwd[3 * n + 1] = o_fwd[k][1]; V.out_fwd[3 * n + 2] = o_fwd[k][2]; \
    }

// stats[16]: 0..2 sum r | 3..8 sum r r' | 9..11 sum proposal offsets (x, y, z) | 12 sum dz^2 | 13..15 sum dx^2, dx dy, dy^2
template <int KIND>  // 0 prior, 1 point-point
struct FamPoint3Gauss {
    using Row = RowSE2;
    static constexpr int D0 = 3, D1 = KIND == 0 ? 0 : 3, DM = 3, DR = 3, DFWD = 3, kMinCtas = 2, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        constexpr int DZ = 3;
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* a0 = reinterpret_cast<const double*>(V.b0);
        const double* a1 = reinterpret_cast<const double*>(KIND == 0 ? V.b0 : V.b1);
        const float* X0 = reinterpret_cast<const float*>(V.b0 + var_header_bytes(3));
        const float* X1 = reinterpret_cast<const float*>((KIND == 0 ? V.b0 : V.b1) + var_header_bytes(3));
        // prior: mean relative to the anchor; point-point: mu - (anchor(xj) - anchor(xi))
        double c0[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) c0[i] = KIND == 0 ? row.mu[i] - a0[i] : row.mu[i] - (a1[i] - a0[i]);
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        ROME_SLOT_LOOP(true, {
            float m[3];
            if (!kSample) {
                m[0] = V.meas[3 * n]; m[1] = V.meas[3 * n + 1]; m[2] = V.meas[3 * n + 2];
            } else {
                m[0] = __fmul_rn(row.L[0], z[3 * k]);
                m[1] = fmaf(row.L[2], z[3 * k + 1], row.L[1] * z[3 * k]);
                m[2] = fmaf(row.L[5], z[3 * k + 2], fmaf(row.L[4], z[3 * k + 1], row.L[3] * z[3 * k]));
                if ((flags & ROME_B200_WRITE_MEAS) && live) {
                    float* M = P.meas_out + fo + 3 * n;
                    __stcs(M, m[0]); __stcs(M + 1, m[1]); __stcs(M + 2, m[2]);
                }
            }
            float e[3], o[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                // prior: h = m - anchor(x);  point-point: h = xi + m - anchor(xj);  residual = h - (target offset)
                const double h = KIND == 0 ? c0[i] + (double)m[i] : (c0[i] + (double)m[i]) + (double)X0[3 * n + i];
                e[i] = (float)(h - (double)X1[3 * n + i]);
                o[i] = (float)h;
            }
            const float msk = (nn < N) ? 1.f : 0.f;
            o_res[k][0] = e[0]; o_res[k][1] = e[1]; o_res[k][2] = e[2];
            if (want_stats) acc_res3(st, msk, e[0], e[1], e[2]);
            if (flags & ROME_B200_PROPOSAL_FWD) {
                o_fwd[k][0] = o[0]; o_fwd[k][1] = o[1]; o_fwd[k][2] = o[2];
                if (want_stats) {
                    acc_prop2(st, msk, o[0], o[1]);
                    st[11] = fmaf(msk, o[2], st[11]);
                    st[12] = fmaf(msk * o[2], o[2], st[12]);
                }
            }
        })
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

#undef ROME_SLOT_DECL
#undef ROME_SLOT_STORE

int launch_point3(int family, const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    switch (family) {
        case ROME_B200_PRIORPOINT3: return launch_family<FamPoint3Gauss<0>>(p, plan, grid, s);
        case ROME_B200_POINT3POINT3: return launch_family<FamPoint3Gauss<1>>(p, plan, grid, s);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace rome
