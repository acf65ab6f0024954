// se3_common.cuh -- Float64 quaternion arithmetic for SO(3) and the row helpers shared by the SE(3) families
// (fam_se3.cu: Pose3Pose3, PriorPose3; fam_se3_partial.cu: Pose3Pose3XYYaw, Pose3Pose3Rotation, Pose3Pose3UnitTrans).
#pragma once
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
static __device__ __noinline__ double2 exp_scale_general(double t2) {
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
static __device__ __noinline__ double log_scale_general(double n2, double w) {
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

// ---- analytic Jacobian blocks (perturbations t <- t + dt, R <- R Exp(delta)) -----------------
__device__ __forceinline__ void quat_to_rot(const Quat& q, double (&R)[9]) {
    const double xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z, xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z,
                 wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    R[0] = 1.0 - 2.0 * (yy + zz); R[1] = 2.0 * (xy - wz);       R[2] = 2.0 * (xz + wy);
    R[3] = 2.0 * (xy + wz);       R[4] = 1.0 - 2.0 * (xx + zz); R[5] = 2.0 * (yz - wx);
    R[6] = 2.0 * (xz - wy);       R[7] = 2.0 * (yz + wx);       R[8] = 1.0 - 2.0 * (xx + yy);
}
// inverse right Jacobian of SO(3) at the rotation vector w, sgn = +1;  sgn = -1 gives the inverse LEFT Jacobian
//   Jr^-1(w) = I + 1/2 [w]x + c [w]x^2,  c = 1/t^2 - (1 + cos t) / (2 t sin t)  (series below t = 0.05),  Jl^-1(w) = Jr^-1(-w)
__device__ __forceinline__ void so3_jinv(double wx, double wy, double wz, double sgn, double (&J)[9]) {
    const double t2 = wx * wx + wy * wy + wz * wz;
    double c;
    if (t2 < 2.5e-3) {
        c = 1.0 / 12.0 + t2 * (1.0 / 720.0 + t2 * (1.0 / 30240.0));
    } else {
        const double t = sqrt(t2);
        double st, ct;
        sincos(t, &st, &ct);
        c = 1.0 / t2 - (1.0 + ct) / (2.0 * t * st);
    }
    const double h = 0.5 * sgn;
    // [w]x^2 = w w' - |w|^2 I
    J[0] = 1.0 + c * (wx * wx - t2); J[1] = -h * wz + c * wx * wy;     J[2] = h * wy + c * wx * wz;
    J[3] = h * wz + c * wx * wy;     J[4] = 1.0 + c * (wy * wy - t2); J[5] = -h * wx + c * wy * wz;
    J[6] = -h * wy + c * wx * wz;    J[7] = h * wx + c * wy * wz;     J[8] = 1.0 + c * (wz * wz - t2);
}
__device__ __forceinline__ void store9_global(float* p, const double (&M)[9], double scale) {
#pragma unroll
    for (int i = 0; i < 9; ++i) __stcs(p + i, (float)(scale * M[i]));
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
        // dead slots re-read a particle that exists (lane, or 0 when Npad < 32); their outputs are masked, but a NaN/Inf
        // picked up beyond the block would survive the zero mask of the statistics
        pr.n[j] = pr.live[j] ? nn : (lane < Npad ? lane : 0);
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

// measurements of the lane's PAIR of particles of iteration `it` (Pose3Pose3, PriorPose3): three Philox blocks
// (3 it, 3 it + 1, 3 it + 2) give the 12 normals of the two particles -- no wasted draws (meas6 spends two blocks on six)
template <bool kSample>
__device__ __forceinline__ void meas6_pair(const RowSE3& row, const EvalParams& P, const FactorView& V, int f, int lane,
                                           int it, const Pair& pr, float (&m)[2][6]) {
    if (!kSample) {
        load6(V.meas + 6 * pr.n[0], m[0]);
        load6(V.meas + 6 * pr.n[1], m[1]);
    } else {
        float z[12];
#pragma unroll
        for (int b = 0; b < 3; ++b)
            normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(3 * it + b), z + 4 * b);
        sample6(row, z, m[0]);
        sample6(row, z + 6, m[1]);
    }
}

}  // namespace rome
