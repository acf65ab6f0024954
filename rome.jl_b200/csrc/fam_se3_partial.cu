// fam_se3_partial.cu -- partial / derived Pose3-Pose3 families, sm_100a.
//   Pose3Pose3XYYaw      src/factors/PartialPose3.jl:116-134  SE(2) residual of the (x, y, yaw) projections:
//                        p2 = (t_p[1:2], normalize(R_p[1:2,1])), q2 likewise, r = vee(log(q2, p2 o exp(X)))
//   Pose3Pose3Rotation   src/factors/PartialPose3.jl:212-226  r = m - Log(R_p' R_q)
//   Pose3Pose3UnitTrans  src/factors/Pose3Pose3.jl:107-116    Pose3Pose3 residual with normalised translation part
// These are partial constraints: there is no closed-form full proposal (DFWD = 0).
#include "se3_common.cuh"

namespace rome {

__device__ __forceinline__ void load3(const float* p, float (&v)[3]) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }

// measurement offsets L z for a 3-D belief (RowSE2), one particle
template <bool kSample>
__device__ __forceinline__ void meas3(const RowSE2& row, const EvalParams& P, const FactorView& V, int f, int lane,
                                      int slot, int n, float (&m)[3]) {
    if (!kSample) {
        load3(V.meas + 3 * n, m);
    } else {
        float z[4];
        normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)slot, z);
        m[0] = __fmul_rn(row.L[0], z[0]);
        m[1] = fmaf(row.L[2], z[1], row.L[1] * z[0]);
        m[2] = fmaf(row.L[5], z[2], fmaf(row.L[4], z[1], row.L[3] * z[0]));
    }
}
// normalised first-column xy of the rotation matrix of a unit quaternion: (cos yaw, sin yaw)
__device__ __forceinline__ void yaw_dir(const Quat& q, double& c, double& s) {
    const double r00 = 1.0 - 2.0 * (q.y * q.y + q.z * q.z), r10 = 2.0 * (q.x * q.y + q.w * q.z);
    const double inv = 1.0 / sqrt(r00 * r00 + r10 * r10);
    c = r00 * inv; s = r10 * inv;
}

template <int KIND>  // 0 XYYaw, 1 Rotation
struct FamPose3Partial {
    using Row = RowSE2;
    static constexpr int D0 = 6, D1 = 6, DM = 3, DR = 3, DFWD = 0, kMinCtas = 1, kWarpFT = 12;
    template <uint32_t kStatic, bool kSample>
    static __device__ __forceinline__ void factor(const Row& row, const EvalParams& P, const FactorView& V, int f,
                                                  int lane) {
        const int Npad = P.Npad, N = P.N;
        const uint32_t flags = (kStatic ? kStatic : P.flags) & (V.fwd_on ? ~0u : ~ROME_B200_PROPOSAL_FWD);
        const double* ap = reinterpret_cast<const double*>(V.b0);
        const double* aq = reinterpret_cast<const double*>(V.b1);
        const float* Pp = reinterpret_cast<const float*>(V.b0 + var_header_bytes(6));
        const float* Qp = reinterpret_cast<const float*>(V.b1 + var_header_bytes(6));
        const size_t fo = (size_t)f * 3 * Npad;
        const bool want_stats = flags & ROME_B200_STATS;
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1];
        float st[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) st[i] = 0.f;
        for (int n = lane, slot = 0; n < Npad; n += 32, ++slot) {
            float p[6], q[6], m[3];
            load6(Pp + 6 * n, p);
            load6(Qp + 6 * n, q);
            meas3<kSample>(row, P, V, f, lane, slot, n, m);
            const double X0 = row.mu[0] + (double)m[0], X1 = row.mu[1] + (double)m[1], X2 = row.mu[2] + (double)m[2];
            const Quat Rp = quat_exp<false>(ap[3] + (double)p[3], ap[4] + (double)p[4], ap[5] + (double)p[5]);
            const Quat Rq = quat_exp<false>(aq[3] + (double)q[3], aq[4] + (double)q[4], aq[5] + (double)q[5]);
            float e1, e2, e3;
            if (KIND == 0) {
                double cp, sp, cq, sq, sm, cm;
                yaw_dir(Rp, cp, sp);
                yaw_dir(Rq, cq, sq);
                sincos(X2, &sm, &cm);
                const double hx = (dax + (double)p[0]) + (cp * X0 - sp * X1);
                const double hy = (day + (double)p[1]) + (sp * X0 + cp * X1);
                // heading of q2^-1 (p2 o exp(X)): (cp + i sp)(cm + i sm)(cq - i sq)
                const double ch = cp * cm - sp * sm, sh = sp * cm + cp * sm;
                e1 = (float)(hx - (double)q[0]);
                e2 = (float)(hy - (double)q[1]);
                e3 = (float)atan2(sh * cq - ch * sq, ch * cq + sh * sq);
            } else {
                double wx, wy, wz;
                quat_log_any(qmul(qconj(Rp), Rq), wx, wy, wz);
                e1 = (float)(X0 - wx); e2 = (float)(X1 - wy); e3 = (float)(X2 - wz);
            }
            const float msk = (n < N) ? 1.f : 0.f;
            if (want_stats) acc_res3(st, msk, e1, e2, e3);
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) {
                float* M = P.meas_out + fo + 3 * n;
                __stcs(M, m[0]); __stcs(M + 1, m[1]); __stcs(M + 2, m[2]);
            }
            if (flags & ROME_B200_RESIDUAL) {
                V.out_res[3 * n] = e1; V.out_res[3 * n + 1] = e2; V.out_res[3 * n + 2] = e3;
            }
        }
        if (want_stats) write_stats16(st, P.stats, f, lane);
    }
};

struct FamPose3Pose3UnitTrans {
    using Row = RowSE3;
    static constexpr int D0 = 6, D1 = 6, DM = 6, DR = 6, DFWD = 0, kMinCtas = 1, kWarpFT = 12;
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
        const bool want_stats = flags & ROME_B200_STATS;
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1], daz = ap[2] - aq[2];
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int n = lane, slot = 0; n < Npad; n += 32, ++slot) {
            float p[6], q[6], m[6];
            load6(Pp + 6 * n, p);
            load6(Qp + 6 * n, q);
            meas6<kSample>(row, P, V, f, lane, slot, n, m);
            double X[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) X[i] = row.mu[i] + (double)m[i];
            const Quat Rp = quat_exp<false>(ap[3] + (double)p[3], ap[4] + (double)p[4], ap[5] + (double)p[5]);
            const Quat Rq = quat_exp<false>(aq[3] + (double)q[3], aq[4] + (double)q[4], aq[5] + (double)q[5]);
            const Quat M = quat_exp<false>(X[3], X[4], X[5]);
            double vx, vy, vz, wx, wy, wz;
            quat_rotate(Rp, X[0], X[1], X[2], vx, vy, vz);
            const double tx = ((dax + (double)p[0]) + vx) - (double)q[0];
            const double ty = ((day + (double)p[1]) + vy) - (double)q[1];
            const double tz = ((daz + (double)p[2]) + vz) - (double)q[2];
            quat_log_any(qmul(qconj(Rq), qmul(Rp, M)), wx, wy, wz);
            const double inv = 1.0 / sqrt(tx * tx + ty * ty + tz * tz);  // normalize(Xc[1:3]); 0/0 -> NaN as in the reference
            const float r[6] = {(float)(tx * inv), (float)(ty * inv), (float)(tz * inv), (float)wx, (float)wy, (float)wz};
            const float msk = (n < N) ? 1.f : 0.f;
            if (want_stats) acc_res6(st, msk, r);
            if (kSample && (flags & ROME_B200_WRITE_MEAS)) store6_global(P.meas_out + fo + 6 * n, m);
            if (flags & ROME_B200_RESIDUAL) store6(V.out_res + 6 * n, r);
        }
        if (want_stats) {
            const float tot = warp_reduce_scatter32(st, lane);
            P.stats[(size_t)f * 32 + lane] = tot;
        }
    }
};

int launch_pose3_partial(int family, const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    switch (family) {
        case ROME_B200_POSE3POSE3XYYAW: return launch_family<FamPose3Partial<0>>(p, plan, grid, s);
        case ROME_B200_POSE3POSE3ROTATION: return launch_family<FamPose3Partial<1>>(p, plan, grid, s);
        case ROME_B200_POSE3POSE3UNITTRANS: return launch_family<FamPose3Pose3UnitTrans>(p, plan, grid, s);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace rome
