// product_kernels.cu -- the step AFTER the convolution path: the new belief of a variable is the
// product of the proposal densities its factors produced.  Every proposal is a kernel density estimate over its N
// proposal particles with a per-dimension rule-of-thumb bandwidth; N samples of the product of the k KDEs are drawn
// with a Gibbs sampler over the component labels (the restated core of ApproxManifoldProducts.manifoldProduct /
// KernelDensityEstimate's product sampler; callers: IIF propagateBelief, test/testBearingRange2D.jl:275,296,342,
// test/testBasicPose2Conv.jl:25-34) and written straight into the device particle store, so particles never leave
// the GPU between sweeps.
//
//   one CTA per variable; lane = one chain = one output particle; the CTA's warps split the blocks of 32 chains.
//   k = 2 (the usual interior pose of a chain: forward + backward proposal) is sampled EXACTLY: the marginal weights
//       W_a = sum_b Normal(x_0a - x_1b; 0, h_0^2 + h_1^2) of the N^2-component product mixture are computed once per
//       variable (log-sum-exp), a chain draws a from their CDF and then b | a.
//   k > 2: the remaining proposals enter one at a time conditioned on the components already chosen, followed by
//       `iters` Gibbs sweeps; in a sweep, for every proposal j
//       mu_-j, var_-j  <- precision-weighted mean / variance of the other proposals' selected components
//       l_j            <- categorical draw over i = 1..N with log-weight -1/2 sum_dim (x_ji - mu_-j)^2 / (h_j^2 + var_-j),
//                          taken in ONE pass with the Gumbel-max trick (no normalisation, no underflow)
//   then x ~ Normal(mu_all, var_all).  Heading dimensions (Pose2 theta) use wrapped differences.
// Proposal rows are read through the read-only path: all lanes of a warp read the same component in the categorical
// scan (a broadcast), the k rows of a variable (k x Npad x d floats) stay L1-resident.
//
// STAGED path (the usual case: the k rows of the variable fit kProdRowFloats of shared memory and no heading offset is
// beyond 1.5 rad, so no difference needs wrapping): the rows are copied once into shared memory, coordinate-major
// ([source][dim][Nst], padding components parked at 1e18 = weight 0), and every N-component scan reads four components
// per 16-byte broadcast load.  Weights are evaluated in scaled coordinates, w_i = 2^-(sum_c (s_c x_ic - s_c mu_c)^2),
// s_c = sqrt(log2(e) / 2 / (h_c^2 + var_c)): two FFMA per dimension and ONE special-function instruction per
// component.  The categorical draw is a one-pass weighted reservoir (component i replaces the choice with probability
// w_i / (w_1 + .. + w_i): exactly w_i / sum w overall), the pair stage sums the weights directly.  Weights are at most
// 1, so only far-apart densities can defeat this (every weight underflows): a draw whose weight sum is below 1e-30 and a
// pair stage whose total is below 1e-25 are redone by the log-domain code of the general path (Gumbel-max draw,
// log-sum-exp marginals), which is also what variables that do not fit or need wrapped differences use throughout.
#include <cuda_runtime.h>

#include "../../include/rome_b200.h"
#include "device_utils.cuh"
#include "se3_common.cuh"
#include "tables.h"

namespace rome {

constexpr int kProdWarps = 4;

__device__ __forceinline__ uint32_t xorshift32(uint32_t& s) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return s;
}
// per-chain stream for the Gumbel noise: 32-bit LCG (Numerical Recipes constants), one IMAD per draw; only its top 23
// bits are used and every chain is seeded by its own Philox block, which is ample for a categorical draw over N
// uniform in (0,1) with 23 random bits
__device__ __forceinline__ float u01(uint32_t& s) {
    s = s * 1664525u + 1013904223u;
    return __uint_as_float(0x3f800000u | (s >> 9)) - (1.0f - 5.9604644775390625e-08f);
}
__device__ __forceinline__ float lg2f(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2f(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int b = 16; b >= 1; b >>= 1) v += __shfl_xor_sync(0xffffffffu, v, b);
    return v;
}

// D = coordinate dimension, WRAP = index of the heading coordinate or -1
template <int D, int WRAP>
struct ProdOps {
    // precision-weighted fusion of the selected components of sources [0, upto) except `skip`
    static __device__ __forceinline__ void fuse(const ProductParams& P, const float (*bw)[D], const uint16_t (*lab)[32],
                                                int s0, int upto, int skip, int lane /* of the chain */, int row_floats,
                                                float (&mu)[D], float (&vr)[D]) {
        float lam[D], s[D], ref = 0.f;
        bool have_ref = false;
#pragma unroll
        for (int c = 0; c < D; ++c) { lam[c] = 0.f; s[c] = 0.f; }
#pragma unroll 1
        for (int jj = 0; jj < upto; ++jj) {
            if (jj == skip) continue;
            const float* row = P.bufs[P.src_buf[s0 + jj]] + (size_t)P.src_row[s0 + jj] * row_floats;
            const int l = lab[jj][lane];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                float x = __ldg(row + l * D + c);
                const float w = bw[jj][c];
                if (c == WRAP) {  // unwrap every heading towards the first one so the weighted mean is well defined
                    if (!have_ref) { ref = x; have_ref = true; }
                    x = ref + wrap_pi_f(x - ref);
                }
                lam[c] += w;
                s[c] = fmaf(w, x, s[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) { vr[c] = 1.0f / lam[c]; mu[c] = s[c] * vr[c]; }
    }
    // categorical draw over the N components of source j given (mu, vr) of the others:
    //   argmax_i  q_i + Gumbel_i,  q_i = -1/2 sum (x_ji - mu)^2 / (h_j^2 + vr)      (one pass, no normalisation)
    static __device__ __forceinline__ int draw(const ProductParams& P, const float (*bw)[D], int s0, int j, int row_floats,
                                               const float (&mu)[D], const float (&vr)[D], uint32_t& rng) {
        float c2[D];
#pragma unroll
        for (int c = 0; c < D; ++c) c2[c] = -0.5f / (1.0f / bw[j][c] + vr[c]);
        const float* row = P.bufs[P.src_buf[s0 + j]] + (size_t)P.src_row[s0 + j] * row_floats;
        float best = -3.0e38f;
        int arg = 0;
        const float* x = row;
#pragma unroll 1
        for (int i = 0; i < P.N; ++i, x += D) {  // running pointer: the D loads take immediate offsets
            float q = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) {
                float dlt = __ldg(x + c) - mu[c];
                if (c == WRAP) dlt = wrap_pi_f(dlt);
                q = fmaf(c2[c] * dlt, dlt, q);
            }
            const float key = q - 0.6931471805599453f * lg2f(-lg2f(u01(rng)));  // + Gumbel(0,1) up to a constant
            if (key > best) { best = key; arg = i; }
        }
        return arg;
    }
    // log2 of the pair-stage marginals W_a = sum_b w_ab by log-sum-exp (rows in global memory, wrapped heading differences)
    static __device__ __forceinline__ void pair_log(const ProductParams& P, const float (*bw)[D], const float* r0,
                                                    const float* r1, float* cdf) {
        float c2[D];
#pragma unroll
        for (int c = 0; c < D; ++c) c2[c] = -0.5f * 1.4426950408889634f / (1.0f / bw[0][c] + 1.0f / bw[1][c]);
        for (int a = threadIdx.x; a < P.N; a += blockDim.x) {
            float xa[D];
#pragma unroll
            for (int c = 0; c < D; ++c) xa[c] = __ldg(r0 + a * D + c);
            float m = -3.0e38f, sum = 0.f;
            const float* xb = r1;
#pragma unroll 1
            for (int b = 0; b < P.N; ++b, xb += D) {
                float q = 0.f;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    float dlt = __ldg(xb + c) - xa[c];
                    if (c == WRAP) dlt = wrap_pi_f(dlt);
                    q = fmaf(c2[c] * dlt, dlt, q);
                }
                const float m2 = fmaxf(m, q);
                sum = fmaf(sum, exp2f(m - m2), exp2f(q - m2));
                m = m2;
            }
            cdf[a] = m + lg2f(sum);
        }
    }
    // ---- staged path: rows in shared memory, coordinate-major: xs[(j * D + c) * Nst + i] ---------------------------
    static __device__ __forceinline__ void fuse_s(const float* xs, int Nst, const float (*bw)[D],
                                                  const uint16_t (*lab)[32], int upto, int skip, int cl, float (&mu)[D],
                                                  float (&vr)[D]) {
        float lam[D], s[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { lam[c] = 0.f; s[c] = 0.f; }
        for (int jj = 0; jj < upto; ++jj) {
            if (jj == skip) continue;
            const float* x = xs + (size_t)jj * D * Nst + lab[jj][cl];
#pragma unroll
            for (int c = 0; c < D; ++c) {
                const float w = bw[jj][c];
                lam[c] += w;
                s[c] = fmaf(w, x[c * Nst], s[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < D; ++c) { vr[c] = __fdividef(1.0f, lam[c]); mu[c] = s[c] * vr[c]; }
    }
    // sum_i w_i and a component drawn with probability w_i / sum: a one-pass weighted reservoir over BLOCKS of four
    // components (one uniform per block: block B replaces the choice with probability W_B / (W_1 + .. + W_B)), then one
    // component of the chosen block by its four recomputed weights.
    // w_i = 2^-(sum_c (sc_c x_ic + nm_c)^2), nm_c = -sc_c mu_c.  xj: the source's rows, N4: N rounded up to 4
    static __device__ __forceinline__ void weights4(const float* xj, int Nst, int i, const float (&sc)[D],
                                                    const float (&nm)[D], float (&w)[4]) {
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(xj + c * Nst + i);
            const float d0 = fmaf(x.x, sc[c], nm[c]), d1 = fmaf(x.y, sc[c], nm[c]);
            const float d2 = fmaf(x.z, sc[c], nm[c]), d3 = fmaf(x.w, sc[c], nm[c]);
            q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
        }
        w[0] = ex2f(-q0); w[1] = ex2f(-q1); w[2] = ex2f(-q2); w[3] = ex2f(-q3);
    }
    // `r` lanes share a chain (r = 1, or a power of two for the trailing block of chains, warp-uniform): sub-lane `sub`
    // scans the blocks sub, sub + r, ..., the r reservoirs are merged pairwise (the partner's choice is taken with
    // probability S' / (S + S')) and every lane of the group returns the group's component and weight sum.
    static __device__ __forceinline__ int draw_s(const float* xj, int Nst, int N4, const float (&sc)[D],
                                                 const float (&nm)[D], uint32_t& rng, float& Ssum, int r, int sub,
                                                 int lane) {
        float S = 0.f;
        int blk = 4 * sub < N4 ? 4 * sub : 0;
        for (int i = 4 * sub; i < N4; i += 4 * r) {
            float w[4];
            weights4(xj, Nst, i, sc, nm, w);
            const float W = (w[0] + w[1]) + (w[2] + w[3]);
            S += W;
            rng = rng * 1664525u + 1013904223u;
            const float u = __uint_as_float(0x3f800000u | (rng >> 9));  // [1, 2)
            if (fmaf(u, S, -S) < W) blk = i;                            // (u - 1) S < W: probability W / S
        }
        rng = rng * 1664525u + 1013904223u;
        float u = __uint_as_float(0x3f800000u | (rng >> 9));
        if (r > 1) {
            for (int o = 1; o < r; o <<= 1) {
                const float S2 = __shfl_down_sync(0xffffffffu, S, o);
                const int blk2 = __shfl_down_sync(0xffffffffu, blk, o);
                rng = rng * 1664525u + 1013904223u;
                const float um = __uint_as_float(0x3f800000u | (rng >> 9));
                S += S2;                                   // (only the group's first lane ends with the right values)
                if (fmaf(um, S, -S) < S2) blk = blk2;
            }
            const int lead = lane & ~(r - 1);
            S = __shfl_sync(0xffffffffu, S, lead);
            blk = __shfl_sync(0xffffffffu, blk, lead);
            u = __shfl_sync(0xffffffffu, u, lead);
        }
        Ssum = S;
        float w[4];
        weights4(xj, Nst, blk, sc, nm, w);
        const float W = (w[0] + w[1]) + (w[2] + w[3]);
        const float tgt = fmaf(u, W, -W);  // uniform in [0, W)
        int t = 0;
        if (tgt >= w[0]) t = 1;
        if (tgt >= w[0] + w[1]) t = 2;
        if (tgt >= (w[0] + w[1]) + w[2]) t = 3;
        // a padding component (weight 0) can only be reached through rounding of the partial sums: step back to a live one
        if (t == 3 && !(w[3] > 0.f)) t = 2;
        if (t == 2 && !(w[2] > 0.f)) t = 1;
        if (t == 1 && !(w[1] > 0.f)) t = 0;
        return blk + t;
    }
};

constexpr int kProdMaxN = 1024;
constexpr int kProdRowFloats = 5120;  // shared pool per variable: k * D * Nst floats of staged rows + the labels (20 KB)  // components per proposal for which the exact pair stage is used

// bandwidth of dimension c of one source (a warp's task): h = std * bw_scale, circular statistics for the heading.
// Element (component i, dimension c) sits at base[i * si + c * sd]: (D, 1) for a row in global memory, (1, Nst) staged.
template <int WRAP>
__device__ __forceinline__ void bandwidth_task(const float* base, int si, int sd, int N, int c, int lane, float bw_scale,
                                               float& h, float& tmax) {
    const bool circ = c == WRAP;
    const float* x0 = base + c * sd;
    float a = 0.f, b = 0.f, tm = 0.f;
    for (int i = lane; i < N; i += 32) {
        const float x = x0[i * si];
        if (circ) { float sn, cs; __sincosf(x, &sn, &cs); a += cs; b += sn; tm = fmaxf(tm, fabsf(x)); }
        else a += x;
    }
    a = warp_sum(a);
    float mean;
    if (circ) {
        b = warp_sum(b);
        mean = atan2f(b, a);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
    } else {
        mean = a / (float)N;
    }
    float acc = 0.f;
    for (int i = lane; i < N; i += 32) {
        float dlt = x0[i * si] - mean;
        if (circ) dlt = wrap_pi_f(dlt);
        acc = fmaf(dlt, dlt, acc);
    }
    const float var = warp_sum(acc) / (float)max(N - 1, 1);
    h = fmaxf(sqrtf(var) * bw_scale, 1e-6f);
    tmax = tm;
}

template <int D, int WRAP>
__global__ void __launch_bounds__(kProdWarps * 32, 8) product_kernel(const __grid_constant__ ProductParams P) {
    // one CTA per variable (grid-stride); its warps share the bandwidths and the pair-stage CDF and split the chains
    __shared__ float s_bw[ROME_B200_MAX_PRODUCT_SOURCES][D];                   // 1 / h^2 per source and dimension
    __shared__ float s_cdf[kProdMaxN];                                         // pair stage: CDF over source-0 components
    __shared__ float s_tmax[ROME_B200_MAX_PRODUCT_SOURCES];                    // largest |heading offset| per source
    __shared__ const float* s_row[ROME_B200_MAX_PRODUCT_SOURCES];              // the sources' rows in global memory
    // one pool: the staged rows [source][dim][Nst], then the chains' labels [warp][source][32] (uint16); 20 KB keep
    // eight CTAs resident per SM (64 registers per thread)
    __shared__ __align__(16) float s_rows[kProdRowFloats];
    using Ops = ProdOps<D, WRAP>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blocks = (P.Npad + 31) / 32;
    const int row_floats = D * P.Npad;
    const int N4 = (P.N + 3) & ~3, Nst = N4;
    const float kLog2e = 1.4426950408889634f;
    for (int v = blockIdx.x; v < P.nvars; v += gridDim.x) {
        const int s0 = P.var_off[v], k = P.var_off[v + 1] - s0;
        if (k == 0) continue;  // no proposal: the belief is left as it is
        float* dst = reinterpret_cast<float*>(P.store + (size_t)v * var_block_bytes(D, P.Npad) + var_header_bytes(D));
        if (k == 1) {  // product of one density: adopt its particles
            const float* row = P.bufs[P.src_buf[s0]] + (size_t)P.src_row[s0] * row_floats;
            for (int n = threadIdx.x; n < P.Npad; n += blockDim.x) {
#pragma unroll
                for (int c = 0; c < D; ++c) dst[n * D + c] = n < P.N ? __ldg(row + n * D + c) : 0.f;
            }
            continue;
        }
        __syncthreads();  // the previous variable's shared state is no longer read
        if (threadIdx.x < k)
            s_row[threadIdx.x] = P.bufs[P.src_buf[s0 + threadIdx.x]] + (size_t)P.src_row[s0 + threadIdx.x] * row_floats;
        __syncthreads();
        bool vec = true;  // every row 16-byte aligned: staged by 16-byte loads
        for (int j = 0; j < k; ++j) vec = vec && (reinterpret_cast<uintptr_t>(s_row[j]) & 15) == 0;
        // ---- staging: the k rows, coordinate-major; padding components sit at 1e18 (weight 0).  Every thread issues
        //      its (up to four) 16-byte loads before it scatters the first one, so one round trip to L2 covers a batch
        const int lab_floats = kProdWarps * k * 16;  // 32 uint16 per (warp, source)
        const bool fits = P.N <= kProdMaxN && k * D * Nst + lab_floats <= kProdRowFloats;
        if (fits) {
            const int ND = P.N * D;
            if (vec) {
                const int R4 = (ND + 3) >> 2, total = k * R4;
                for (int t0 = 0; t0 < total; t0 += 4 * blockDim.x) {
                    float4 val[4];
                    int jq[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t0 + u * blockDim.x + threadIdx.x;
                        jq[u] = -1;
                        if (t < total) {
                            const int j = t / R4, q = t - j * R4;
                            jq[u] = (j << 16) | q;
                            val[u] = __ldg(reinterpret_cast<const float4*>(s_row[j]) + q);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (jq[u] < 0) continue;
                        const int j = jq[u] >> 16, q = jq[u] & 0xffff;
                        float* xj = s_rows + j * D * Nst;
                        const float e[4] = {val[u].x, val[u].y, val[u].z, val[u].w};
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const int idx = 4 * q + w, i = idx / D, c = idx - i * D;
                            if (idx < ND) xj[c * Nst + i] = e[w];
                        }
                    }
                }
            } else {
                for (int j = 0; j < k; ++j) {
                    const float* row = s_row[j];
                    float* xj = s_rows + j * D * Nst;
                    for (int idx = threadIdx.x; idx < ND; idx += blockDim.x) {
                        const int i = idx / D, c = idx - i * D;
                        xj[c * Nst + i] = __ldg(row + idx);
                    }
                }
            }
            const int npadc = Nst - P.N;
            for (int t = threadIdx.x; t < k * D * npadc; t += blockDim.x) {
                const int jc = t / npadc, i = P.N + (t - jc * npadc);
                s_rows[jc * Nst + i] = 1e18f;
            }
            __syncthreads();
        }
        // ---- Pose3 on the manifold (ROME_B200_PRODUCT_MANIFOLD): the staged rotation-vector offsets become tangent
        //      coordinates at the anchor rotation, xi = Log(R_anchor^-1 Exp(anchor_w + offset)); bandwidths, weights and
        //      fusion below then work in that tangent space and the sample is retracted, R_anchor Exp(xi)
        bool man = false;
        const double* const hdr = reinterpret_cast<const double*>(P.store + (size_t)v * var_block_bytes(D, P.Npad));
        if constexpr (D == 6) {
            man = P.manifold && fits;
            if (man) {
                const double manA[3] = {hdr[3], hdr[4], hdr[5]};
                const Quat qa = qconj(quat_exp<false>(manA[0], manA[1], manA[2]));
                for (int t = threadIdx.x; t < k * P.N; t += blockDim.x) {
                    const int j = t / P.N, i = t - j * P.N;
                    float* x = s_rows + j * D * Nst + i;
                    const Quat q = quat_exp<false>(manA[0] + (double)x[3 * Nst], manA[1] + (double)x[4 * Nst],
                                                   manA[2] + (double)x[5 * Nst]);
                    double ex, ey, ez;
                    quat_log_any(qmul(qa, q), ex, ey, ez);
                    x[3 * Nst] = (float)ex; x[4 * Nst] = (float)ey; x[5 * Nst] = (float)ez;
                }
                __syncthreads();
            }
        }
        // ---- per-source bandwidths; the k * D (source, dimension) tasks are dealt to the warps
        for (int t = warp; t < k * D; t += kProdWarps) {
            const int j = t / D, c = t - j * D;
            float h, tm;
            bandwidth_task<WRAP>(fits ? s_rows + j * D * Nst : s_row[j], fits ? 1 : D, fits ? Nst : 1, P.N, c, lane,
                                 P.bw_scale, h, tm);
            if (lane == 0) {
                s_bw[j][c] = 1.0f / (h * h);
                if (c == WRAP) s_tmax[j] = tm;
                if (P.bw_out) P.bw_out[(size_t)(s0 + j) * D + c] = h;
            }
        }
        __syncthreads();
        // staged path: rows fit and no difference of headings needs wrapping (CTA-uniform)
        bool staged = fits;
        if (WRAP >= 0 && staged) {
            float tm = 0.f;
            for (int j = 0; j < k; ++j) tm = fmaxf(tm, s_tmax[j]);
            staged = tm < 1.5f;
        }
        // ---- exact pair stage for sources 0 and 1: the product of two KDEs is a mixture of N^2 Gaussians with weights
        //      w_ab = Normal(x_0a - x_1b; 0, h_0^2 + h_1^2); marginal W_a = sum_b w_ab -> CDF over a (shared by all
        //      chains of the variable); a chain draws a ~ W, then b | a.
        const bool pair = P.N <= kProdMaxN;
        for (int pass = pair ? 0 : 2; pass < 2; ++pass) {
            const bool logw = !staged || pass == 1;
            if (!logw) {  // weights summed directly in scaled coordinates
                float sc[D];
#pragma unroll
                for (int c = 0; c < D; ++c) sc[c] = sqrtf(0.5f * kLog2e / (1.0f / s_bw[0][c] + 1.0f / s_bw[1][c]));
                const float* x1 = s_rows + D * Nst;
                for (int a = threadIdx.x; a < P.N; a += blockDim.x) {
                    float na[D];
#pragma unroll
                    for (int c = 0; c < D; ++c) na[c] = -sc[c] * s_rows[c * Nst + a];
                    float W = 0.f;
                    for (int b = 0; b < N4; b += 4) {
                        float w[4];
                        Ops::weights4(x1, Nst, b, sc, na, w);
                        W += (w[0] + w[1]) + (w[2] + w[3]);
                    }
                    s_cdf[a] = W;
                }
            } else {  // general path: log-sum-exp keeps far-apart proposals finite; s_cdf[a] = log2 W_a
                Ops::pair_log(P, s_bw, s_row[0], s_row[1], s_cdf);
            }
            __syncthreads();
            if (warp == 0) {  // inclusive prefix sum over a (warp scan per chunk of 32 with a running carry)
                float wmax = 0.f;
                if (logw) {
                    wmax = -3.0e38f;
                    for (int a = lane; a < P.N; a += 32) wmax = fmaxf(wmax, s_cdf[a]);
#pragma unroll
                    for (int b = 16; b >= 1; b >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, b));
                }
                float carry = 0.f;
#pragma unroll 1
                for (int base = 0; base < P.N; base += 32) {
                    const int a = base + lane;
                    float x = a < P.N ? (logw ? exp2f(s_cdf[a] - wmax) : s_cdf[a]) : 0.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float y = __shfl_up_sync(0xffffffffu, x, o);
                        if (lane >= o) x += y;
                    }
                    x += carry;
                    if (a < P.N) s_cdf[a] = x;
                    carry = __shfl_sync(0xffffffffu, x, 31);
                }
            }
            __syncthreads();
            // far-apart proposals: every direct weight underflowed -> the log-domain marginals in a second pass
            if (logw || s_cdf[P.N - 1] >= 1e-25f) break;
            __syncthreads();  // (everyone has read the total before it is rewritten)
        }
        uint16_t (*const lab)[32] =
            reinterpret_cast<uint16_t (*)[32]>(s_rows + (fits ? k * D * Nst : 0)) + warp * k;  // this warp's [source][32]
        // ---- chains: one output particle each; blocks of 32 chains are dealt to the warps round-robin.  In the trailing
        //      block (N = 100: 4 chains) r = 32 / chains lanes share a chain and split its component scans (draw_s)
        for (int n = P.N + threadIdx.x; n < P.Npad; n += blockDim.x) {
#pragma unroll
            for (int c = 0; c < D; ++c) dst[n * D + c] = 0.f;
        }
        for (int blk = warp; blk < blocks; blk += kProdWarps) {
            const int cnt = min(32, P.N - blk * 32);
            if (cnt <= 0) continue;
            int r = 1;
            if (staged) while (2 * r * cnt <= 32) r *= 2;
            const int sub = lane & (r - 1), lead = lane & ~(r - 1), cl = lane / r;
            const int n = blk * 32 + cl;
            uint32_t rng;
            {
                const uint4 x = philox4x32_10(make_uint4((uint32_t)(blk * 32 + lane), (uint32_t)v, P.stream_id, 0x50524f44u), P.seed_lo, P.seed_hi);
                rng = x.x | 1u;
            }
            float mu[D], vr[D];
            if (pair) {
                const float target = u01(rng) * s_cdf[P.N - 1];
                int lo = 0, hi = P.N - 1;  // first a with cdf[a] >= target
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (s_cdf[mid] < target) lo = mid + 1; else hi = mid;
                }
                if (r > 1) lo = __shfl_sync(0xffffffffu, lo, lead);
                lab[0][cl] = (uint16_t)lo;
            } else {
                lab[0][lane] = (uint16_t)(xorshift32(rng) % (uint32_t)P.N);
                lab[1][lane] = (uint16_t)(xorshift32(rng) % (uint32_t)P.N);
            }
            // sources `first` .. k - 1 enter one at a time conditioned on the components already chosen (with the exact
            // pair stage: b | a first), then -- unless the pair stage already sampled the whole product exactly --
            // `iters` Gibbs sweeps over all sources.  ONE fuse / draw site serves every step (instruction-cache footprint)
            const int first = pair ? 1 : 2, nseq = k - first;
            const int nsteps = nseq + ((k > 2 || !pair) ? P.iters * k : 0);
            for (int s = 0, g = 0; s < nsteps; ++s) {
                int j, upto, skip;
                if (s < nseq) { j = first + s; upto = j; skip = -1; }
                else { j = g; upto = k; skip = g; g = g + 1 == k ? 0 : g + 1; }
                int arg = -1;
                if (staged) {
                    Ops::fuse_s(s_rows, Nst, s_bw, lab, upto, skip, cl, mu, vr);
                    float sc[D], nm[D], S;
#pragma unroll
                    for (int c = 0; c < D; ++c) {  // sqrt(log2(e) / 2 / (h_j^2 + var))
                        sc[c] = rsqrt_fast((__fdividef(1.0f, s_bw[j][c]) + vr[c]) * (2.0f / kLog2e));
                        nm[c] = -sc[c] * mu[c];
                    }
                    arg = Ops::draw_s(s_rows + j * D * Nst, Nst, N4, sc, nm, rng, S, r, sub, lane);
                    if (S < 1e-30f) arg = -1;  // every weight underflowed: the log-domain scan decides
                } else {
                    Ops::fuse(P, s_bw, lab, s0, upto, skip, cl, row_floats, mu, vr);
                }
                if (arg < 0) arg = Ops::draw(P, s_bw, s0, j, row_floats, mu, vr, rng);
                if (r > 1) arg = __shfl_sync(0xffffffffu, arg, lead);
                lab[j][cl] = (uint16_t)arg;
            }
            // the sample: Normal(fused mean, fused variance) of the chosen components
            if (staged) Ops::fuse_s(s_rows, Nst, s_bw, lab, k, -1, cl, mu, vr);
            else Ops::fuse(P, s_bw, lab, s0, k, -1, cl, row_floats, mu, vr);
            float z[8];
            const uint4 a = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)v, P.stream_id, 0x50524f45u), P.seed_lo, P.seed_hi);
            box_muller(a.x, a.y, z[0], z[1]); box_muller(a.z, a.w, z[2], z[3]);
            if (D > 4) {
                const uint4 b = philox4x32_10(make_uint4((uint32_t)n, (uint32_t)v, P.stream_id, 0x50524f46u), P.seed_lo, P.seed_hi);
                box_muller(b.x, b.y, z[4], z[5]); box_muller(b.z, b.w, z[6], z[7]);
            }
            if (sub == 0 && n < P.N) {
                float xo[D];
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    xo[c] = fmaf(sqrtf(vr[c]), z[c], mu[c]);
                    if (c == WRAP) xo[c] = wrap_pi_f(xo[c]);
                }
                if constexpr (D == 6) {
                    if (man) {  // retraction: rotation vector of R_anchor Exp(xi), the representative closest to the anchor
                        const double manA[3] = {hdr[3], hdr[4], hdr[5]};
                        const Quat q = qmul(quat_exp<false>(manA[0], manA[1], manA[2]),
                                            quat_exp<false>((double)xo[3], (double)xo[4], (double)xo[5]));
                        double wx, wy, wz;
                        quat_log_any(q, wx, wy, wz);
                        closest_rotvec(wx, wy, wz, manA[0], manA[1], manA[2]);
                        xo[3] = (float)(wx - manA[0]); xo[4] = (float)(wy - manA[1]); xo[5] = (float)(wz - manA[2]);
                    }
                }
#pragma unroll
                for (int c = 0; c < D; ++c) dst[n * D + c] = xo[c];
            }
            __syncwarp();
        }
    }
}

// move every variable's anchor onto its first particle (offsets stay small after the belief has moved)
template <int D, int WRAP>
__global__ void reanchor_kernel(unsigned char* store, int nvars, int N, int Npad) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    unsigned char* blk = store + (size_t)v * var_block_bytes(D, Npad);
    double* hdr = reinterpret_cast<double*>(blk);
    float* off = reinterpret_cast<float*>(blk + var_header_bytes(D));
    float o0[D];
#pragma unroll
    for (int c = 0; c < D; ++c) o0[c] = off[c];
    __syncwarp();
    for (int n = lane; n < N; n += 32) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            float x = off[n * D + c] - o0[c];
            if (c == WRAP) x = wrap_pi_f(x);
            off[n * D + c] = x;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double a = hdr[c] + (double)o0[c];
            if (c == WRAP) a = wrap_pi(a);
            hdr[c] = a;
        }
        if (WRAP >= 0) {
            hdr[D] = cos(hdr[WRAP]); hdr[D + 1] = sin(hdr[WRAP]);
            *reinterpret_cast<float2*>(hdr + D + 2) = make_float2((float)hdr[D], (float)hdr[D + 1]);
        }
    }
}

int launch_product(int d, int wrap_dim, const void* params, int num_sms, void* stream) {
    const ProductParams& p = *static_cast<const ProductParams*>(params);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (p.nvars == 0) return 0;
    // one CTA per variable: the cost of a variable grows with its number of proposals (one N x N scan for two, a dozen
    // for four), so the hardware block scheduler balances the SMs better than a static grid-stride assignment would
    const int grid = p.nvars;
    (void)num_sms;
    if (d == 3 && wrap_dim == 2) product_kernel<3, 2><<<grid, kProdWarps * 32, 0, s>>>(p);
    else if (d == 3) product_kernel<3, -1><<<grid, kProdWarps * 32, 0, s>>>(p);
    else if (d == 2) product_kernel<2, -1><<<grid, kProdWarps * 32, 0, s>>>(p);
    else if (d == 6) product_kernel<6, -1><<<grid, kProdWarps * 32, 0, s>>>(p);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_reanchor(int d, int wrap_dim, unsigned char* store, int nvars, int N, int Npad, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (nvars == 0) return 0;
    const int grid = (nvars + 7) / 8;
    if (d == 3 && wrap_dim == 2) reanchor_kernel<3, 2><<<grid, 256, 0, s>>>(store, nvars, N, Npad);
    else if (d == 3) reanchor_kernel<3, -1><<<grid, 256, 0, s>>>(store, nvars, N, Npad);
    else if (d == 2) reanchor_kernel<2, -1><<<grid, 256, 0, s>>>(store, nvars, N, Npad);
    else if (d == 6) reanchor_kernel<6, -1><<<grid, 256, 0, s>>>(store, nvars, N, Npad);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
size_t product_params_size() { return sizeof(ProductParams); }

}  // namespace rome
