// fam_se3.cu -- SE(3) families (sm_100a).  Reference arithmetic (paths relative to /root/reference):
//   Pose3Pose3    src/factors/Pose3Pose3.jl:17-29
//   PriorPose3    src/factors/Pose3D.jl:15-19
#include "eval_pipeline.cuh"

namespace rome {

// =============================================================================================
// Float64 quaternion helpers for SO(3).  Exp and Log are split into a SCALAR part (the only place where the
// fast polynomial path and the general sqrt/sincos/atan2 path differ) and the common vector part, so the
// general paths are out-of-line functions returning registers (no local-memory traffic) and the fast paths
// are straight-line code that two particles of a lane can interleave.
// =============================================================================================
struct Quat {
    double w, x, y, z;
};
constexpr double kPi2 = 9.8696;  // rotation vectors with |w|^2 <= kPi2 take the polynomial Exp

// Exp(w) = (c, k w) with c = cos(|w|/2), k = sin(|w|/2)/|w|.
// General path: any |w|.
__device__ __noinline__ double2 exp_scale_general(double t2) {
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
    return make_double2(k, c);
}
// |w| <= pi (the principal range) without sqrt, division or range reduction: with y = theta/4 <= pi/4 the fdlibm
// kernels give cos y and sin(y)/y as polynomials in y^2 = |w|^2/16, and
//   cos(theta/2) = 2 cos^2 y - 1,   sin(theta/2)/theta = (sin(y)/y) cos(y) / 2.
__device__ __forceinline__ double2 exp_scale_poly(double t2) {
    const double z = t2 * 0.0625;
    double ps = fma(z, kSinC[5], kSinC[4]);
    double pc = fma(z, kCosC[5], kCosC[4]);
    ps = fma(z, ps, kSinC[3]); pc = fma(z, pc, kCosC[3]);
    ps = fma(z, ps, kSinC[2]); pc = fma(z, pc, kCosC[2]);
    ps = fma(z, ps, kSinC[1]); pc = fma(z, pc, kCosC[1]);
    ps = fma(z, ps, kSinC[0]); pc = fma(z, pc, kCosC[0]);
    const double sy = fma(z, ps, 1.0);                    // sin(y)/y
    const double cy = fma(z * z, pc, fma(z, -0.5, 1.0));  // cos(y)
    return make_double2(0.5 * sy * cy, fma(2.0 * cy, cy, -1.0));
}
template <bool kFast>
__device__ __forceinline__ Quat quat_exp(double wx, double wy, double wz, double t2) {
    const double2 kc = kFast ? exp_scale_poly(t2) : (t2 <= kPi2 ? exp_scale_poly(t2) : exp_scale_general(t2));
    return {kc.y, kc.x * wx, kc.x * wy, kc.x * wz};
}
__device__ __forceinline__ Quat qmul(const Quat& a, const Quat& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
template <bool kFast>
__device__ __forceinline__ Quat quat_exp(double wx, double wy, double wz) {
    return quat_exp<kFast>(wx, wy, wz, fma(wx, wx, fma(wy, wy, wz * wz)));
}
__device__ __forceinline__ Quat qconj(const Quat& a) { return {a.w, -a.x, -a.y, -a.z}; }
// Log(q) = k v for a unit quaternion (w >= 0 after the sign flip), k = 2 atan2(|v|, w)/|v|, angle in [0, pi].
__device__ __noinline__ double log_scale_general(double n2, double w) {
    if (n2 < 1e-16) return 2.0 / w;
    const double n = sqrt(n2);
    return 2.0 * atan2(n, w) / n;
}
// Small rotations (the residual of a consistent factor): with u = |v|/w <= 0.1 (angle <= 0.2 rad),
//   k = (2/w) * atan(u)/u,  atan(u)/u by 9 series terms (1e-17),
// 1/w by three Newton steps from 2 - w (w >= 0.995).  No sqrt, no division, no atan2.
constexpr double kSmallLog = 0.0099;
__device__ __forceinline__ double log_scale_small(double n2, double w) {
    double r = 2.0 - w;
    r = r * fma(-w, r, 2.0);
    r = r * fma(-w, r, 2.0);
    r = r * fma(-w, r, 2.0);
    const double u2 = -n2 * r * r;
    double p = fma(u2, kOddInv[8], kOddInv[7]);
    p = fma(u2, p, kOddInv[6]);
    p = fma(u2, p, kOddInv[5]);
    p = fma(u2, p, kOddInv[4]);
    p = fma(u2, p, kOddInv[3]);
    p = fma(u2, p, kOddInv[2]);
    p = fma(u2, p, kOddInv[1]);
    p = fma(u2, p, 1.0);
    return 2.0 * r * p;
}
// sign-normalised copy (w >= 0) and squared vector norm
__device__ __forceinline__ Quat quat_pos(const Quat& q, double& n2) {
    n2 = fma(q.x, q.x, fma(q.y, q.y, q.z * q.z));
    const int flip = __double2hiint(q.w) & 0x80000000;  // sign bit of w: q and -q are the same rotation
    auto sx = [flip](double v) { return __hiloint2double(__double2hiint(v) ^ flip, __double2loint(v)); };
    return {sx(q.w), sx(q.x), sx(q.y), sx(q.z)};
}
__device__ __forceinline__ void quat_log_any(const Quat& q, double& x, double& y, double& z) {
    double n2;
    const Quat u = quat_pos(q, n2);
    const double k = (n2 <= kSmallLog * u.w * u.w) ? log_scale_small(n2, u.w) : log_scale_general(n2, u.w);
    x = k * u.x; y = k * u.y; z = k * u.z;
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
// SE(3) families.  A lane evaluates TWO particles per iteration (n, n+32): the exponentials of both are
// taken on the polynomial path when every rotation vector of the warp's 64 particles is principal
// (warp-uniform vote), the logarithms on the small-angle path when every residual rotation is small --
// straight-line Float64 code whose two dependency chains interleave; otherwise the general functions run.
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
// measurement offsets L z of one particle (z: its 8 normals, 6 used)
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
// the lane's pair of particles for iteration `it`: slots 2 it and 2 it + 1 (particles lane + 32 slot)
struct Pair {
    int n[2];
    bool live[2];
};
__device__ __forceinline__ Pair pair_of(int it, int lane, int Npad) {
    Pair pr;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int nn = lane + 32 * (2 * it + j);
        pr.live[j] = nn < Npad;
        pr.n[j] = pr.live[j] ? nn : lane;  // dead slots re-read particle `lane`; their outputs are masked
    }
    return pr;
}
template <bool kSample>
__device__ __forceinline__ void meas6(const RowSE3& row, const EvalParams& P, const FactorView& V, int f, int lane,
                                      int slot, int n, float (&m)[6]) {
    if (!kSample) {
        load6(V.meas + 6 * n, m);
    } else {  // two Philox blocks per particle (8 normals, 6 used)
        float z[8];
        normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(2 * slot), z);
        normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(2 * slot + 1), z + 4);
        sample6(row, z, m);
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
        const double dax = ap[0] - aq[0], day = ap[1] - aq[1], daz = ap[2] - aq[2];  // anchor delta (exact Float64)
        float st[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) st[i] = 0.f;
        for (int it = 0; 64 * it < Npad; ++it) {
            const Pair pr = pair_of(it, lane, Npad);
            float p[2][6], q[2][6], m[2][6];
            double X[2][6], wp[2][3], wq[2][3], t2p[2], t2q[2], t2m[2];
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                load6(Pp + 6 * pr.n[j], p[j]);
                load6(Qp + 6 * pr.n[j], q[j]);
                meas6<kSample>(row, P, V, f, lane, 2 * it + j, pr.n[j], m[j]);
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
                if (flags & ROME_B200_PROPOSAL_FWD) {  // q = p o Exp(X): coordinates as offsets from q's anchor
                    double ox, oy, oz;
                    quat_log_any(Rh[j], ox, oy, oz);
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
                    const float o[6] = {(float)(((double)q[j][0] - dax) - bx), (float)(((double)q[j][1] - day) - by),
                                        (float)(((double)q[j][2] - daz) - bz), (float)(ox - ap[3]),
                                        (float)(oy - ap[4]),                   (float)(oz - ap[5])};
                    if (live) store6_global(P.prop_bwd + fo + 6 * n, o);
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
        for (int it = 0; 64 * it < Npad; ++it) {
            const Pair pr = pair_of(it, lane, Npad);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = pr.n[j];
                const bool live = pr.live[j];
                float p[6], m[6];
                load6(Pp + 6 * n, p);
                meas6<kSample>(row, P, V, f, lane, 2 * it + j, n, m);
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
