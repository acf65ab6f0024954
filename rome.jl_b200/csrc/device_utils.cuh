// device_utils.cuh -- sm_100a building blocks shared by the factor kernels:
// TMA 1-D bulk copy + mbarrier (per-factor table staging), cache-hinted vector loads/stores,
// Philox4x32-10 sampler, Float64 wrap helpers, and the halving-butterfly warp reduction used for
// the per-factor statistics.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rome {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 6.28318530717958647692;
constexpr double kTwoPiLo = 2.4492935982947064e-16;  // 2*pi - double(2*pi)
constexpr double kInvTwoPi = 0.15915494309189533577;

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk) -- SASS: SYNCS.* / UBLKCP
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(phase)
            : "memory");
    } while (!ok);
}
// global -> shared 1-D bulk copy, completion signalled on `bar` (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared -> global 1-D bulk store (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk stores have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// all but the newest n (0..2) of this thread's bulk-store groups are complete (written, not merely read)
__device__ __forceinline__ void tma_store_wait_pending(int n) {
    if (n <= 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    else if (n == 1) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
}

// back-off inside a spin loop on shared / global memory
__device__ __forceinline__ void spin_pause() { __nanosleep(32); }
// system-scope release store / acquire load (flags in peer-GPU memory over NVLink)
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// global memory access with cache hints
// ---------------------------------------------------------------------------------------------
// streaming read (touched once): bypass L1 allocation
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// particle read (re-used by neighbouring factors): read-only path, keep in L1/L2
__device__ __forceinline__ float4 ld_reuse4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming store
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ float& f4(float4& v, int j) { return reinterpret_cast<float*>(&v)[j]; }
__device__ __forceinline__ float f4c(const float4& v, int j) { return reinterpret_cast<const float*>(&v)[j]; }

// ---------------------------------------------------------------------------------------------
// Float64 angle helpers
// ---------------------------------------------------------------------------------------------
// a - 2*pi*rint(a/2pi) in [-pi, pi]; equals atan(sin a, cos a) of src/factors/Pose2D.jl:64 away from
// the branch cut, and |.| = pi on it.
__device__ __forceinline__ double wrap_pi(double a) {
    const double k = rint(a * kInvTwoPi);
    return fma(-k, kTwoPiLo, fma(-k, kTwoPi, a));
}
// Manifolds.sym_rem (src/factors/BearingRange2D.jl:61): [-pi, pi], x ~ pi -> -pi
__device__ __forceinline__ double sym_rem(double x) {
    const double r = wrap_pi(x);
    return (fabs(x - kPi) <= 1.4901161193847656e-08 * fmax(fabs(x), kPi)) ? -kPi : r;
}

// sin/cos of a SMALL angle |x| <= pi/4 (the float32 heading offset of a particle from its variable's
// anchor heading): fdlibm __kernel_sin/__kernel_cos minimax polynomials, coefficients in constant memory so
// that every DFMA takes its coefficient as a constant-bank operand.  |error| < 1e-16.
__constant__ double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
__constant__ double kCosC[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};
// atan(u)/u = sum (-u^2)^k / (2k+1): reciprocal odd numbers in constant memory (DFMA constant-bank operands)
__constant__ double kOddInv[9] = {1.0,        1.0 / 3.0,  1.0 / 5.0,  1.0 / 7.0, 1.0 / 9.0,
                                  1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0, 1.0 / 17.0};
constexpr double kSmallAngle = 0.78;
__device__ __forceinline__ void sincos_small(double x, double& s, double& c) {
    const double z = x * x;
    double ps = fma(z, kSinC[5], kSinC[4]);
    double pc = fma(z, kCosC[5], kCosC[4]);
    ps = fma(z, ps, kSinC[3]); pc = fma(z, pc, kCosC[3]);
    ps = fma(z, ps, kSinC[2]); pc = fma(z, pc, kCosC[2]);
    ps = fma(z, ps, kSinC[1]); pc = fma(z, pc, kCosC[1]);
    ps = fma(z, ps, kSinC[0]); pc = fma(z, pc, kCosC[0]);
    s = fma(x * z, ps, x);
    c = fma(z * z, pc, fma(z, -0.5, 1.0));
}
// ---------------------------------------------------------------------------------------------
// float32 per-particle helpers (the default arithmetic of the per-particle path: every per-particle
// quantity is a SMALL offset, so float32 keeps ~1e-8 absolute; the large parts are per-factor Float64)
// ---------------------------------------------------------------------------------------------
// sin(x) and cos(x) - 1 for |x| <= 0.78, Taylor to x^9 / x^10: truncation < 2e-9 relative
__device__ __forceinline__ void sincosm1_small_f(float x, float& s, float& cm1) {
    const float z = x * x;
    float ps = fmaf(z, 2.7557319224e-6f, -1.9841269841e-4f);
    float pc = fmaf(z, -2.7557319224e-7f, 2.4801587302e-5f);
    ps = fmaf(z, ps, 8.3333333333e-3f);
    pc = fmaf(z, pc, -1.3888888889e-3f);
    ps = fmaf(z, ps, -1.6666666667e-1f);
    pc = fmaf(z, pc, 4.1666666667e-2f);
    s = fmaf(x * z, ps, x);
    pc = fmaf(z, pc, -0.5f);
    cm1 = z * pc;
}
// a - 2 pi rint(a / 2 pi) in float32; rint by the 1.5 * 2^23 magic constant (two full-rate FADDs instead of FRND),
// 2 pi split into float(2 pi) + remainder so that the reduction of a small multiple is exact to ~1e-8
__device__ __forceinline__ float wrap_pi_f(float a) {
    const float k = __fadd_rn(__fmaf_rn(a, 0.15915494309f, 12582912.f), -12582912.f);
    return fmaf(-k, -1.7484555e-7f, fmaf(-k, 6.2831854820f, a));
}
// split a Float64 into float32 hi + lo (hi = the value with its low 29 mantissa bits cleared, exactly a float32 for
// magnitudes in float32's normal range): one widening conversion less than (float)(x - (double)(float)x)
__device__ __forceinline__ void split_f64(double x, float& hi, float& lo) {
    const double h = __hiloint2double(__double2hiint(x), __double2loint(x) & (int)0xE0000000);
    hi = (float)h;
    lo = (float)(x - h);
}
// sin/cos of any float32 angle with full float32 accuracy: polynomial for small angles, libdevice otherwise
__device__ __forceinline__ void sincos_any_f(float x, float& s, float& c) {
    if (fabsf(x) <= 0.78f) {
        float cm1;
        sincosm1_small_f(x, s, cm1);
        c = 1.f + cm1;
    } else {
        sincosf(x, &s, &c);
    }
}

// sqrt of a positive normal double in float32 range: MUFU.RSQ seed (relative error ~2^-22) followed by two
// coupled Goldschmidt steps (error -> 1.5 e^2 each); result within ~1 ulp.  No branches, no division.
__device__ __forceinline__ double sqrt_seeded(double a) {
    float y0;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)a));
    double g = a * (double)y0, h = 0.5 * (double)y0;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    return fma(g, r, g);
}
// A rotation has the rotation vectors w (1 - 2 pi k / |w|), k integer.  Particle coordinates are stored as offsets from
// the variable's anchor, so the representative CLOSEST to the anchor is the one to keep (near |w| = pi the principal
// vector of a neighbouring rotation sits on the far side of the ball).  Replaces (x, y, z) by the better of k = 0, 1.
__device__ __forceinline__ void closest_rotvec(double& x, double& y, double& z, double ax, double ay, double az) {
    const double t2 = x * x + y * y + z * z;
    if (t2 < 1e-12) return;
    const double f = 1.0 - kTwoPi * rsqrt(t2);
    const double bx = x * f, by = y * f, bz = z * f;
    const double d0 = (x - ax) * (x - ax) + (y - ay) * (y - ay) + (z - az) * (z - az);
    const double d1 = (bx - ax) * (bx - ax) + (by - ay) * (by - ay) + (bz - az) * (bz - az);
    if (d1 < d0) { x = bx; y = by; z = bz; }
}
// sin/cos of (anchor + x) given (ca, sa) = cos/sin(anchor): angle addition for small x, general path otherwise
__device__ __forceinline__ void sincos_anchored(double anchor, double ca, double sa, double x, double& s, double& c) {
    if (fabs(x) <= kSmallAngle) {
        double sx, cx;
        sincos_small(x, sx, cx);
        s = fma(sa, cx, ca * sx);
        c = fma(ca, cx, -sa * sx);
    } else {
        sincos(anchor + x, &s, &c);
    }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG + Box-Muller; host twin: oracle/rome_oracle.c rome_oracle_normal4
// counter = (particle, factor, stream, block), key = seed
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
// Box-Muller on the special-function unit (MUFU.LG2 / RSQ / SIN / COS): the draws differ from the Float64 host
// twin by <= ~1e-4 sigma (tests/test_gpu_parity_raw.py); exact parity of a sweep is taken on the samples the
// kernel writes back (ROME_B200_WRITE_MEAS), never on re-derived ones.
//   u1 = ((a >> 9) + 0.5) / 2^23 in (0,1), built from the bit pattern of a float in [1,2);  radius = sqrt(-2 ln u1)
//   u2 likewise;  angle = 2 pi u2 evaluated as -(cos, sin)(2 pi u2 - pi) to keep the MUFU argument in [-pi, pi]
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
#ifdef ROME_B200_ACCURATE_SAMPLER
    const float u1 = (static_cast<float>(a >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float u2 = (static_cast<float>(b >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
#else
    const float f1 = __uint_as_float(0x3f800000u | (a >> 9));  // 1 + (a>>9)/2^23
    const float f2 = __uint_as_float(0x3f800000u | (b >> 9));
    const float u1 = f1 - (1.0f - 5.9604644775390625e-08f);    // exact: (a>>9)/2^23 + 2^-24
    float lg, rs, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    const float t = fmaxf(lg * -1.3862943611198906f, 1e-12f);  // -2 ln u1 = -2 ln2 * lg2 u1
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(t));
    const float rad = t * rs;
    // 2 pi u2 - pi = 2 pi (f2 - 1 + 2^-24) - pi
    __sincosf(fmaf(f2, 6.28318530717958647692f, -9.42477758f), &s, &c);
    s = -s; c = -c;
#endif
    z0 = rad * c;
    z1 = rad * s;
}
__device__ __forceinline__ void normal4(uint32_t seed_lo, uint32_t seed_hi, uint32_t stream, uint32_t factor,
                                        uint32_t particle, uint32_t block, float z[4]) {
    const uint4 x = philox4x32_10(make_uint4(particle, factor, stream, block), seed_lo, seed_hi);
    box_muller(x.x, x.y, z[0], z[1]);
    box_muller(x.z, x.w, z[2], z[3]);
}

// ---------------------------------------------------------------------------------------------
// warp reductions: halving butterfly ("reduce-scatter" over lanes).  K values per lane in,
// after log2(K) exchange steps every lane owns ONE fully reduced value: K=16 -> value index
// lane>>1 (16 shuffles instead of 80), K=32 -> value index lane (31 shuffles instead of 160).
// ---------------------------------------------------------------------------------------------
template <int K, int BIT>
struct HalvingStep {
    static __device__ __forceinline__ float run(float (&v)[K], int lane) {
        constexpr int H = K / 2;
        float w[H];
        const bool up = (lane & BIT) != 0;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            const float send = up ? v[i] : v[i + H];
            const float keep = up ? v[i + H] : v[i];
            w[i] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
        }
        return HalvingStep<H, BIT / 2>::run(w, lane);
    }
};
template <int BIT>
struct HalvingStep<1, BIT> {
    static __device__ __forceinline__ float run(float (&v)[1], int) {
        float x = v[0];
#pragma unroll
        for (int b = BIT; b >= 1; b >>= 1) x += __shfl_xor_sync(0xffffffffu, x, b);
        return x;
    }
};
// 16 values per lane -> lane holds total of value (lane >> 1)
__device__ __forceinline__ float warp_reduce_scatter16(float (&v)[16], int lane) {
    return HalvingStep<16, 16>::run(v, lane);
}
// 32 values per lane -> lane holds total of value `lane`
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
    return HalvingStep<32, 16>::run(v, lane);
}

}  // namespace rome
