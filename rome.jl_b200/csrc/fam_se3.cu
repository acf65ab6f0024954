// fam_se3.cu -- SE(3) families (sm_100a).  Reference arithmetic (paths relative to /root/reference):
//   Pose3Pose3    src/factors/Pose3Pose3.jl:17-29
//   PriorPose3    src/factors/Pose3D.jl:15-19
#include "se3_common.cuh"

namespace rome {

struct FamPose3Pose3 {
    using Row = RowSE3;
    static constexpr int D0 = 6, D1 = 6, DM = 6, DR = 6, DFWD = 6, kMinCtas = 1, kWarpFT = 12;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const double* aq = reinterpret_cast<const double*>(V.b1);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(6));
        const float* Qp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(6));
        const size_t fo = (size_t)f * 6 * Npad;
        float* const bwd = (flags & ROME_B200_PROPOSAL_BWD) ? bwd_row(P, f, (size_t)6 * Npad) : nullptr;
        const bool want_stats = flags & ROME_B200_STATS;
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1], daz = ap[2] - aq[2];  // anchor delta (exact Float64)
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int it = 0; 64 * it < Npad; ++it) {
            const Pair pr = pair_of(it, lane, Npad);
            float p[2][6], q[2][6], m[2][6];
            double X[2][6], wp[2][3], wq[2][3], t2p[2], t2q[2], t2m[2];
            bool ok = true;
            meas6_pair<kSample>(row, P, V, f, lane, it, pr, m);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                load6(Pp + 6 * pr.n[j], p[j]);
                load6(Qp + 6 * pr.n[j], q[j]);
#pragma unroll
                for (int i = 0; i < 6; ++i) X[j][i] = row.mu[i] + (double)m[j][i];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    wp[j][i] = ap[3 + i] + (double)p[j][3 + i];
                    wq[j][i] = aq[3 + i] + (double)q[j][3 + i];
                }
                t2p[j] = fma(wp[j][0], wp[j][0], fma(wp[j][1], wp[j][1], wp[j][2] * wp[j][2]));
                t2q[j] = fma(wq[j][0], wq[j][0], fma(wq[j][1], wq[j][1], wq[j][2] * wq[j][2]));
                t2m[j] = fma(X[j][3], X[j][3], fma(X[j][4], X[j][4], X[j][5] * X[j][5]));
                ok = ok && fmax(fmax(t2p[j], t2q[j]), t2m[j]) <= kPi2;
            }
            Quat Rp[2], Rq[2], M[2];
            if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    Rp[j] = quat_exp<true>(wp[j][0], wp[j][1], wp[j][2], t2p[j]);
                    Rq[j] = quat_exp<true>(wq[j][0], wq[j][1], wq[j][2], t2q[j]);
                    M[j] = quat_exp<true>(X[j][3], X[j][4], X[j][5], t2m[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    Rp[j] = quat_exp<false>(wp[j][0], wp[j][1], wp[j][2], t2p[j]);
                    Rq[j] = quat_exp<false>(wq[j][0], wq[j][1], wq[j][2], t2q[j]);
                    M[j] = quat_exp<false>(X[j][3], X[j][4], X[j][5], t2m[j]);
                }
            }
            double h[2][3], n2[2];
            Quat Rh[2], E[2];
            bool small = true;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                double vx, vy, vz;
                quat_rotate(Rp[j], X[j][0], X[j][1], X[j][2], vx, vy, vz);
                // qhat.t as offset from q's anchor
                h[j][0] = (dax + (double)p[j][0]) + vx;
                h[j][1] = (day + (double)p[j][1]) + vy;
                h[j][2] = (daz + (double)p[j][2]) + vz;
                Rh[j] = qmul(Rp[j], M[j]);
                E[j] = quat_pos(qmul(qconj(Rq[j]), Rh[j]), n2[j]);
                small = small && n2[j] <= kSmallLog * E[j].w * E[j].w;
            }
            double kl[2];
            if (__all_sync(0xffffffffu, small)) {
#pragma unroll
                for (int j = 0; j < 2; ++j) kl[j] = log_scale_small(n2[j], E[j].w);
            } else {
#pragma unroll
                for (int j = 0; j < 2; ++j)
                    kl[j] = (n2[j] <= kSmallLog * E[j].w * E[j].w) ? log_scale_small(n2[j], E[j].w)
                                                                   : log_scale_general(n2[j], E[j].w);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = pr.n[j];
                const bool live = pr.live[j];
                const float r[6] = {(float)(h[j][0] - (double)q[j][0]), (float)(h[j][1] - (double)q[j][1]),
                                    (float)(h[j][2] - (double)q[j][2]), (float)(kl[j] * E[j].x),
                                    (float)(kl[j] * E[j].y),            (float)(kl[j] * E[j].z)};
                const float msk = (live && n < N) ? 1.f : 0.f;
                if (want_stats) acc_res6(st, msk, r);
                if (kSample && (flags & ROME_B200_WRITE_MEAS) && live) store6_global(P.meas_out + fo + 6 * n, m[j]);
                if ((flags & ROME_B200_RESIDUAL) && live) store6(V.out_res + 6 * n, r);
                if ((flags & ROME_B200_DECONV) && live) {  // X = log(p^-1 q) = (R_p'(t_q - t_p), Log(R_p' R_q)) - mu
                    const Quat Rpc = qconj(Rp[j]);
                    double vx, vy, vz, ox, oy, oz;
                    quat_rotate(Rpc, ((double)q[j][0] - (double)p[j][0]) - dax, ((double)q[j][1] - (double)p[j][1]) - day,
                                ((double)q[j][2] - (double)p[j][2]) - daz, vx, vy, vz);
                    quat_log_any(qmul(Rpc, Rq[j]), ox, oy, oz);
                    const float o[6] = {(float)(vx - row.mu[0]), (float)(vy - row.mu[1]), (float)(vz - row.mu[2]),
                                        (float)(ox - row.mu[3]), (float)(oy - row.mu[4]), (float)(oz - row.mu[5])};
                    store6_global(P.meas_out + fo + 6 * n, o);
                }
                if ((flags & ROME_B200_JACOBIAN) && live) {
                    // 36 floats per particle, four row-major 3x3 blocks (perturbations t <- t + dt, R <- R Exp(delta)):
                    //   A = d r_t / d delta_p = -R_p [m_t]x      B = d r_w / d delta_p = Jr^-1(r_w) M'
                    //   C = d r_w / d delta_q = -Jl^-1(r_w)      R_p = d r_t / d m_t
                    // (d r_t / d t_p = I, d r_t / d t_q = -I, d r_t / d delta_q = 0, d r_w / d m_w = B M)
                    double Rm[9], Mm[9], Jr[9], Jl[9], A[9], B[9];
                    quat_to_rot(Rp[j], Rm);
                    quat_to_rot(M[j], Mm);
                    const double rx = kl[j] * E[j].x, ry = kl[j] * E[j].y, rz = kl[j] * E[j].z;
                    so3_jinv(rx, ry, rz, 1.0, Jr);
                    so3_jinv(rx, ry, rz, -1.0, Jl);
                    const double mx = X[j][0], my = X[j][1], mz = X[j][2];
                    const double K[9] = {0.0, -mz, my, mz, 0.0, -mx, -my, mx, 0.0};  // [m_t]x
#pragma unroll
                    for (int a = 0; a < 3; ++a)
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            A[3 * a + b] = -(Rm[3 * a] * K[b] + Rm[3 * a + 1] * K[3 + b] + Rm[3 * a + 2] * K[6 + b]);
                            B[3 * a + b] = Jr[3 * a] * Mm[3 * b] + Jr[3 * a + 1] * Mm[3 * b + 1] + Jr[3 * a + 2] * Mm[3 * b + 2];
                        }
                    float* J = P.jac + ((size_t)f * Npad + n) * 36;
                    store9_global(J, A, 1.0);
                    store9_global(J + 9, B, 1.0);
                    store9_global(J + 18, Jl, -1.0);
                    store9_global(J + 27, Rm, 1.0);
                }
                if (flags & ROME_B200_PROPOSAL_FWD) {  // q = p o Exp(X): coordinates as offsets from q's anchor
                    double ox, oy, oz;
                    quat_log_any(Rh[j], ox, oy, oz);
                    closest_rotvec(ox, oy, oz, aq[3], aq[4], aq[5]);
                    const float o[6] = {(float)h[j][0],      (float)h[j][1],      (float)h[j][2],
                                        (float)(ox - aq[3]), (float)(oy - aq[4]), (float)(oz - aq[5])};
                    if (live) store6(V.out_fwd + 6 * n, o);
                    if (want_stats) acc_prop3(st, msk, o[0], o[1], o[2]);
                }
                if (flags & ROME_B200_PROPOSAL_BWD) {  // R_p = R_q Exp(X.w)' ; t_p = t_q - R_p X.t
                    const Quat Rb = qmul(Rq[j], qconj(M[j]));
                    double bx, by, bz, ox, oy, oz;
                    quat_rotate(Rb, X[j][0], X[j][1], X[j][2], bx, by, bz);
                    quat_log_any(Rb, ox, oy, oz);
                    closest_rotvec(ox, oy, oz, ap[3], ap[4], ap[5]);
                    const float o[6] = {(float)(((double)q[j][0] - dax) - bx), (float)(((double)q[j][1] - day) - by),
                                        (float)(((double)q[j][2] - daz) - bz), (float)(ox - ap[3]),
                                        (float)(oy - ap[4]),                   (float)(oz - ap[5])};
                    if (live) store6_global(bwd + 6 * n, o);
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
    static constexpr int D0 = 6, D1 = 0, DM = 6, DR = 6, DFWD = 6, kMinCtas = 1, kWarpFT = 12;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(6));
        const size_t fo = (size_t)f * 6 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int it = 0; 64 * it < Npad; ++it) {
            const Pair pr = pair_of(it, lane, Npad);
            float mm[2][6];
            meas6_pair<kSample>(row, P, V, f, lane, it, pr, mm);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = pr.n[j];
                const bool live = pr.live[j];
                float p[6];
                const float (&m)[6] = mm[j];
                load6(Pp + 6 * n, p);
                double X[6];  // sampled point coordinates: exp(e, hat(mu + L z)) = (t, Exp(w))
#pragma unroll
                for (int i = 0; i < 6; ++i) X[i] = row.mu[i] + (double)m[i];
                const Quat Rp = quat_exp<false>(ap[3] + (double)p[3], ap[4] + (double)p[4], ap[5] + (double)p[5]);
                const Quat Rm = quat_exp<false>(X[3], X[4], X[5]);
                double wx, wy, wz;
                quat_log_any(qmul(qconj(Rp), Rm), wx, wy, wz);
                const double hx = X[0] - ap[0], hy = X[1] - ap[1], hz = X[2] - ap[2];
                const float r[6] = {(float)(hx - (double)p[0]), (float)(hy - (double)p[1]), (float)(hz - (double)p[2]),
                                    (float)wx, (float)wy, (float)wz};
                const float msk = (live && n < N) ? 1.f : 0.f;
                if (want_stats) acc_res6(st, msk, r);
                if (kSample && (flags & ROME_B200_WRITE_MEAS) && live) store6_global(P.meas_out + fo + 6 * n, m);
                if ((flags & ROME_B200_RESIDUAL) && live) store6(V.out_res + 6 * n, r);
                if ((flags & ROME_B200_DECONV) && live) {  // the measurement that explains the particle is the particle
                    float o[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) o[i] = (float)((ap[i] + (double)p[i]) - row.mu[i]);
                    store6_global(P.meas_out + fo + 6 * n, o);
                }
                if ((flags & ROME_B200_JACOBIAN) && live) {  // 9 floats: d r_w / d delta_p = -Jl^-1(r_w); d r_t / d t_p = -I
                    double Jl[9];
                    so3_jinv(wx, wy, wz, -1.0, Jl);
                    store9_global(P.jac + ((size_t)f * Npad + n) * 9, Jl, -1.0);
                }
                if (flags & ROME_B200_PROPOSAL_FWD) {  // proposal = the sampled point, offsets from the anchor
                    const float o[6] = {(float)hx,           (float)hy,           (float)hz,
                                        (float)(X[3] - ap[3]), (float)(X[4] - ap[4]), (float)(X[5] - ap[5])};
                    if (live) store6(V.out_fwd + 6 * n, o);
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

int launch_pose3pose3(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return launch_family<FamPose3Pose3>(p, plan, grid, s);
}
int launch_priorpose3(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return launch_family<FamPriorPose3>(p, plan, grid, s);
}

}  // namespace rome
