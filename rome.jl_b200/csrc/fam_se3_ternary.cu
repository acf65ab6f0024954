// fam_se3_ternary.cu -- SE(3) families with a THIRD variable (sm_100a).  Reference arithmetic (paths relative to
// /root/reference):
//   Pose3Pose3RotOffset   src/factors/Pose3Pose3.jl:57-78   qhat = p o (m.t, bRa Exp(m.w)),   bRa in SO(3) (Rotation3)
//   Pose3Pose3Transform   src/factors/Pose3Pose3.jl:80-95   qhat = p o (Delta o exp(m)),      Delta in SE(3) (Pose3)
// residual (both) = vee(log_q(qhat)) = [qhat.t - q.t ; Log(R_q' R_qhat)], the coordinates Pose3Pose3 uses
// (src/factors/Pose3Pose3.jl:17-29).  The third particle block arrives in the stage like the other two (FactorView::b2).
#include "se3_common.cuh"

namespace rome {

template <int KIND>  // 0: RotOffset (third variable = Rotation3, d = 3), 1: Transform (third variable = Pose3, d = 6)
struct FamPose3Ternary {
    using Row = RowSE3;
    static constexpr int D0 = 6, D1 = 6, D2 = KIND == 0 ? 3 : 6, DM = 6, DR = 6, DFWD = 6, kMinCtas = 1, kWarpFT = 0;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const double* aq = reinterpret_cast<const double*>(V.b1);
        const double* ad = reinterpret_cast<const double*>(V.b2);  // Rotation3: {wx, wy, wz, 0, 0, 0}; Pose3: {t, w}
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(6));
        const float* Qp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(6));
        const float* Dp = reinterpret_cast<const float*>(V.b2 + var_header_bytes(D2));
        const size_t fo = (size_t)f * 6 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1], daz = ap[2] - aq[2];  // anchor delta (exact Float64)
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int n0 = lane; n0 - lane < Npad; n0 += 32) {
            const bool live = n0 < Npad;
            const int n = live ? n0 : 0;  // dead lanes re-read particle 0 (always in range); outputs masked
            float p[6], q[6], m[6];
            load6(Pp + 6 * n, p);
            load6(Qp + 6 * n, q);
            meas6<kSample>(row, P, V, f, lane, n0 >> 5, n, m);
            double X[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) X[i] = row.mu[i] + (double)m[i];
            const Quat Rp = quat_exp<false>(ap[3] + (double)p[3], ap[4] + (double)p[4], ap[5] + (double)p[5]);
            const Quat Rq = quat_exp<false>(aq[3] + (double)q[3], aq[4] + (double)q[4], aq[5] + (double)q[5]);
            const Quat M = quat_exp<false>(X[3], X[4], X[5]);
            Quat Rd;
            double lx = X[0], ly = X[1], lz = X[2];  // the lever arm in p's frame
            if (KIND == 0) {
                Rd = quat_exp<false>(ad[0] + (double)Dp[3 * n], ad[1] + (double)Dp[3 * n + 1], ad[2] + (double)Dp[3 * n + 2]);
            } else {
                float d[6];
                load6(Dp + 6 * n, d);
                Rd = quat_exp<false>(ad[3] + (double)d[3], ad[4] + (double)d[4], ad[5] + (double)d[5]);
                double rx, ry, rz;
                quat_rotate(Rd, X[0], X[1], X[2], rx, ry, rz);  // Delta o exp(m): t = t_D + R_D m.t
                lx = (ad[0] + (double)d[0]) + rx;
                ly = (ad[1] + (double)d[1]) + ry;
                lz = (ad[2] + (double)d[2]) + rz;
            }
            double vx, vy, vz;
            quat_rotate(Rp, lx, ly, lz, vx, vy, vz);
            const double hx = (dax + (double)p[0]) + vx, hy = (day + (double)p[1]) + vy, hz = (daz + (double)p[2]) + vz;
            const Quat Rh = qmul(qmul(Rp, Rd), M);  // R_p bRa Exp(m.w)  resp.  R_p R_D Exp(m.w)
            double wx, wy, wz;
            quat_log_any(qmul(qconj(Rq), Rh), wx, wy, wz);
            const float r[6] = {(float)(hx - (double)q[0]), (float)(hy - (double)q[1]), (float)(hz - (double)q[2]),
                                (float)wx, (float)wy, (float)wz};
            const float msk = (live && n < N) ? 1.f : 0.f;
            if (want_stats) acc_res6(st, msk, r);
            if (kSample && (flags & ROME_B200_WRITE_MEAS) && live) store6_global(P.meas_out + fo + 6 * n, m);
            if ((flags & ROME_B200_RESIDUAL) && live) store6(V.out_res + 6 * n, r);
            if (flags & ROME_B200_PROPOSAL_FWD) {  // qhat as offsets from q's anchor
                double ox, oy, oz;
                quat_log_any(Rh, ox, oy, oz);
                closest_rotvec(ox, oy, oz, aq[3], aq[4], aq[5]);
                const float o[6] = {(float)hx, (float)hy, (float)hz, (float)(ox - aq[3]), (float)(oy - aq[4]), (float)(oz - aq[5])};
                if (live) store6(V.out_fwd + 6 * n, o);
                if (want_stats) acc_prop3(st, msk, o[0], o[1], o[2]);
            }
        }
        if (want_stats) {
            const float tot = warp_reduce_scatter32(st, lane);
            P.stats[(size_t)f * 32 + lane] = tot;
        }
    }
};

int launch_pose3_ternary(int family, const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    if (family == ROME_B200_POSE3POSE3ROTOFFSET) return launch_family<FamPose3Ternary<0>>(p, plan, grid, s);
    return launch_family<FamPose3Ternary<1>>(p, plan, grid, s);
}

}  // namespace rome
