// factor_kernels.cu -- launch planning + family dispatch of the factor-residual kernels, and the layout
// conversion kernels (reference Float64 arrays <-> particle store).  The family kernels themselves live in
// fam_pose2.cu, fam_bearingrange.cu, fam_point2.cu, fam_se3.cu (one translation unit each, compiled in
// parallel) on top of the shared TMA producer/consumer pipeline eval_pipeline.cuh.
#include <cstdlib>

#include "eval_pipeline.cuh"

namespace rome {

int launch_pose2pose2(const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_priorpose2(const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_bearingrange(const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_point2(int family, const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_pose3pose3(const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_priorpose3(const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_point3(int family, const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_pose3_partial(int family, const EvalParams&, const LaunchPlan&, int, cudaStream_t);
int launch_pose3_ternary(int family, const EvalParams&, const LaunchPlan&, int, cudaStream_t);

// =============================================================================================
// host side: launch planning + dispatch
// =============================================================================================
struct FamDims {
    int row_bytes, d0, d1, dm, dr, dfwd, d2;  // d2: dimension of a third variable (0: none)
};
static FamDims fam_dims(int family) {
    switch (family) {
        case ROME_B200_POSE2POSE2: return {(int)sizeof(RowSE2), 3, 3, 3, 3, 3};
        case ROME_B200_PRIORPOSE2: return {(int)sizeof(RowSE2), 3, 0, 3, 3, 3};
        case ROME_B200_BEARINGRANGE: return {(int)sizeof(RowBR), 3, 2, 2, 2, 2};
        case ROME_B200_POSE3POSE3: return {(int)sizeof(RowSE3), 6, 6, 6, 6, 6};
        case ROME_B200_PRIORPOINT2: return {(int)sizeof(RowPT2), 2, 0, 2, 2, 2};
        case ROME_B200_POINT2POINT2: return {(int)sizeof(RowPT2), 2, 2, 2, 2, 2};
        case ROME_B200_POSE2POINT2: return {(int)sizeof(RowPT2), 3, 2, 2, 2, 2};
        case ROME_B200_POSE2POINT2RANGE: return {(int)sizeof(RowS1), 3, 2, 1, 1, 0};
        case ROME_B200_POINT2POINT2RANGE: return {(int)sizeof(RowS1), 2, 2, 1, 1, 0};
        case ROME_B200_POSE2POINT2BEARING: return {(int)sizeof(RowS1), 3, 2, 1, 1, 0};
        case ROME_B200_PRIORPOINT3: return {(int)sizeof(RowSE2), 3, 0, 3, 3, 3};
        case ROME_B200_POINT3POINT3: return {(int)sizeof(RowSE2), 3, 3, 3, 3, 3};
        case ROME_B200_POSE3POSE3XYYAW:
        case ROME_B200_POSE3POSE3ROTATION: return {(int)sizeof(RowSE2), 6, 6, 3, 3, 0};
        case ROME_B200_POSE3POSE3UNITTRANS: return {(int)sizeof(RowSE3), 6, 6, 6, 6, 0};
        case ROME_B200_POSE3POSE3ROTOFFSET: return {(int)sizeof(RowSE3), 6, 6, 6, 6, 6, 3};
        case ROME_B200_POSE3POSE3TRANSFORM: return {(int)sizeof(RowSE3), 6, 6, 6, 6, 6, 6};
        default: return {(int)sizeof(RowSE3), 6, 0, 6, 6, 6};
    }
}

// stages a 2-CTAs/SM configuration must reach before it is preferred over 1 CTA/SM with a deeper ring
// (tuning knob for experiments: ROME_B200_MIN_STAGES_2CTA)
static int min_stages_2cta() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("ROME_B200_MIN_STAGES_2CTA");
        v = e ? atoi(e) : 2;
        if (v < 2) v = 2;
    }
    return v;
}

// tile size of the hot variants: 8 factors x 2 CTAs per SM; ROME_B200_TILE=4 selects 4 factors x 4 CTAs per SM (measured on
// the bench workload: the kernel alone is 1 us faster -- its last, partly filled round costs half as much -- but a step
// with a second, tiny family kernel is 0.8 us slower; profiles/r02_analysis.md)
static int tile_choice() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("ROME_B200_TILE");
        v = (e && atoi(e) == 4) ? 4 : 8;
    }
    return v;
}

// Pipeline depth vs L1.  Shared memory and L1 are one 256 KB array per SM; the carve-out steps are ... 132, 164, 196,
// 228 KB.  The kernels keep a few loop variables in local memory (spilled around the factor body), and with the
// largest carve-out the 28 KB of L1 left no longer hold the resident threads' stack frames (Pose3: 384 threads x 128 B),
// so every spill access goes to L2.  Measured (profiles/r02_analysis.md): Pose3Pose3 50.8 -> 47.1 us and BearingRange
// at N = 200 23.6 -> 22.6 us with two stages instead of three, Pose2Pose2 at N = 100 unchanged between 2, 3 and 4.
// Rule: the deepest ring (>= 2) whose resident CTAs stay within the 164 KB carve-out; ROME_B200_MAX_STAGES overrides.
static int trim_stages(int stages, int ctas, int fixed_bytes, int stage_bytes) {
    static int cap = 0;
    if (!cap) {
        const char* e = getenv("ROME_B200_MAX_STAGES");
        cap = e ? atoi(e) : -1;
    }
    if (cap >= 2) return stages < cap ? stages : cap;
    while (stages > 2 && ctas * (fixed_bytes + stages * stage_bytes + 1024) > 164 * 1024) --stages;
    return stages;
}

// pipeline selection for the Pose3 families: per-warp pipelines unless ROME_B200_PIPELINE=cta
static int pipeline_choice() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("ROME_B200_PIPELINE");
        v = (e && e[0] == 'c') ? 0 : 1;
    }
    return v;
}

int plan_launch(int family, uint32_t flags, int Npad, int smem_per_sm, int smem_per_cta_max, LaunchPlan* plan) {
    const FamDims fd = fam_dims(family);
    const bool sample = (flags & ROME_B200_SAMPLE) != 0;
    const bool se3 = fd.d0 == 6;  // Pose3 families: one CTA per SM (register budget)
    if (fd.dfwd == 0) flags &= ~ROME_B200_PROPOSAL_FWD;
    const uint32_t out_flags = flags & ~kSchedFlags;
    // 3: RESIDUAL|STATS compiled in, forward rows only for the factors that have a destination (ROUTED_ONLY)
    // (the plain RESIDUAL|STATS variant carries no multi-GPU code: with barrier flags it runs as the routed variant too)
    const uint32_t multi = flags & (ROME_B200_ROUTED_ONLY | ROME_B200_BARRIER_WAIT | ROME_B200_BARRIER_SIGNAL);
    const int hot = out_flags == kHot1 ? ((multi && fd.dfwd) ? 3 : 1)
                                       : out_flags == kHot2 ? ((flags & ROME_B200_ROUTED_ONLY) ? 3 : 2) : 0;
    plan->pipeline = 0;
    if (se3 && fd.d2 == 0 && pipeline_choice() == 1) {  // (families with a third variable are built for the CTA pipeline)
        // per-warp pipelines: W warps per CTA, each with its own ring of `stages` slots + output slice
        const int W = 12;
        const int out_warp = (fd.dr + (hot == 1 ? 0 : fd.dfwd)) * Npad * 4;
        const SlotLayout L = slot_layout(fd.row_bytes, fd.d0, fd.d1, fd.dm, sample, Npad);
        const int bar_bytes = (W * kMaxStages * 8 + 127) / 128 * 128;
        for (int ctas = (se3 ? 1 : 2); ctas >= 1; --ctas) {
            const int budget = (smem_per_sm / ctas) - 1024;
            const int cap = budget < smem_per_cta_max ? budget : smem_per_cta_max;
            int stages = (cap - bar_bytes - W * out_warp) / (W * L.bytes);
            if (stages > 4) stages = 4;
            if (stages >= 2) {
                stages = trim_stages(stages, ctas, bar_bytes + W * out_warp, W * L.bytes);
                plan->pipeline = 1; plan->ft = W; plan->variant = hot; plan->stages = stages;
                plan->stage_bytes = L.bytes; plan->out_warp_bytes = out_warp;
                plan->smem_bytes = bar_bytes + W * (stages * L.bytes + out_warp);
                plan->ctas_per_sm = ctas;
                return 0;
            }
        }
    }
    // Optional (ROME_B200_TILE=4) for the hot flag sets of the SE(2)-sized families: 4-factor tiles, FOUR CTAs (4 consumer
    // warps + producer) per SM -- the same 16 consumer warps per SM as two 8-factor CTAs, but a persistent grid's last,
    // partly filled round costs half as much (12 000 factors: 3 000 tiles on 592 CTAs instead of 1 500 on 296).
    if (!se3 && hot != 0 && tile_choice() == 4) {
        const int ft = 4, ctas = 4;
        const int out_warp = (fd.dr + (hot == 1 ? 0 : fd.dfwd)) * Npad * 4;
        const StageLayout L = stage_layout(ft, fd.row_bytes, fd.d0, fd.d1, fd.dm, sample, Npad);
        const int budget = (smem_per_sm / ctas) - 1024;
        const int cap = budget < smem_per_cta_max ? budget : smem_per_cta_max;
        int stages = (cap - kBarrierBytes - ft * out_warp) / L.bytes;
        if (stages > 4) stages = 4;
        if (stages >= 3) {
            plan->ft = ft; plan->variant = hot; plan->stages = stages; plan->stage_bytes = L.bytes;
            plan->out_warp_bytes = out_warp;
            plan->smem_bytes = kBarrierBytes + stages * L.bytes + ft * out_warp;
            plan->ctas_per_sm = ctas;
            return 0;
        }
    }
    static const int fts[3] = {8, 2, 1};
    for (int k = 0; k < 3; ++k) {
        const int ft = fts[k];
        const int variant = ft == 8 ? hot : 0;  // compile-time flag variants exist for the 8- and 4-factor tiles only
        // per-warp output slice: residual rows, then forward-proposal rows (the generic variant reserves both)
        const int out_warp = (fd.dr + (variant == 1 ? 0 : fd.dfwd)) * Npad * 4;
        const StageLayout L = stage_layout(ft, fd.row_bytes, fd.d0, fd.d1, fd.dm, sample, Npad, fd.d2);
        // prefer 2 CTAs/SM (register cap of the SE(2) kernels allows it) with >= 3 stages each, else 1 CTA/SM
        for (int ctas = (se3 ? 1 : 2); ctas >= 1; --ctas) {
            const int budget = (smem_per_sm / ctas) - 1024;  // 1 KB per CTA is reserved by the system
            const int cap = budget < smem_per_cta_max ? budget : smem_per_cta_max;
            int stages = (cap - kBarrierBytes - ft * out_warp) / L.bytes;
            if (stages > kMaxStages) stages = kMaxStages;
            const int need = ctas == 2 ? min_stages_2cta() : 2;
            if (stages >= need) {
                if (ctas == 2 && stages > 4) stages = 4;
                stages = trim_stages(stages, ctas, kBarrierBytes + ft * out_warp, L.bytes);
                plan->ft = ft; plan->variant = variant; plan->stages = stages;
                plan->stage_bytes = L.bytes;
                plan->out_warp_bytes = out_warp;
                plan->smem_bytes = kBarrierBytes + stages * L.bytes + ft * out_warp;
                plan->ctas_per_sm = ctas;
                return 0;
            }
        }
    }
    return (int)cudaErrorInvalidConfiguration;  // N too large for the shared-memory pipeline
}

int launch_eval(int family, const EvalParams& p, const LaunchPlan& plan, int grid, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (family) {
        case ROME_B200_POSE2POSE2: return launch_pose2pose2(p, plan, grid, s);
        case ROME_B200_PRIORPOSE2: return launch_priorpose2(p, plan, grid, s);
        case ROME_B200_BEARINGRANGE: return launch_bearingrange(p, plan, grid, s);
        case ROME_B200_POSE3POSE3: return launch_pose3pose3(p, plan, grid, s);
        case ROME_B200_PRIORPOSE3: return launch_priorpose3(p, plan, grid, s);
        case ROME_B200_PRIORPOINT2:
        case ROME_B200_POINT2POINT2:
        case ROME_B200_POSE2POINT2:
        case ROME_B200_POSE2POINT2RANGE:
        case ROME_B200_POINT2POINT2RANGE:
        case ROME_B200_POSE2POINT2BEARING: return launch_point2(family, p, plan, grid, s);
        case ROME_B200_PRIORPOINT3:
        case ROME_B200_POINT3POINT3: return launch_point3(family, p, plan, grid, s);
        case ROME_B200_POSE3POSE3XYYAW:
        case ROME_B200_POSE3POSE3ROTATION:
        case ROME_B200_POSE3POSE3UNITTRANS: return launch_pose3_partial(family, p, plan, grid, s);
        case ROME_B200_POSE3POSE3ROTOFFSET:
        case ROME_B200_POSE3POSE3TRANSFORM: return launch_pose3_ternary(family, p, plan, grid, s);
    }
    return (int)cudaErrorInvalidValue;
}

// =============================================================================================
// layout conversion: reference layout (Float64 particle-major [var][N][d]) <-> particle store blocks
// one warp per variable; anchor = first particle; heading offsets (wrap_dim) wrapped to [-pi, pi]
// =============================================================================================
template <int D>
__global__ void pack_kernel(int nvars, int N, int Npad, int wrap_dim, const double* __restrict__ coords,
                            unsigned char* __restrict__ store) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const double* src = coords + (size_t)v * N * D;
    unsigned char* blk = store + (size_t)v * var_block_bytes(D, Npad);
    double a[D];
#pragma unroll
    for (int c = 0; c < D; ++c) a[c] = src[c];
    double* hdr = reinterpret_cast<double*>(blk);
    if (lane < var_header_bytes(D) / 8) {
        double h = lane < D ? src[lane] : 0.0;
        if (wrap_dim >= 0 && lane == D) h = cos(src[wrap_dim]);      // Pose2 header: {x, y, theta, cos, sin, 0}
        if (wrap_dim >= 0 && lane == D + 1) h = sin(src[wrap_dim]);
        hdr[lane] = h;
        // last header slot of a Pose2 block: float32 copies of (cos, sin) for the float32 per-particle path
        if (wrap_dim >= 0 && lane == D + 2)
            *reinterpret_cast<float2*>(hdr + lane) = make_float2((float)cos(src[wrap_dim]), (float)sin(src[wrap_dim]));
    }
    float* dst = reinterpret_cast<float*>(blk + var_header_bytes(D));
    for (int n = lane; n < Npad; n += 32) {
        double x[D];
#pragma unroll
        for (int c = 0; c < D; ++c) x[c] = n < N ? src[(size_t)n * D + c] : a[c];
        if (D == 6) closest_rotvec(x[3], x[4], x[5], a[3], a[4], a[5]);  // Pose3: rotation vector nearest the anchor's
        if (D == 3 && wrap_dim == -2) closest_rotvec(x[0], x[1], x[2], a[0], a[1], a[2]);  // Rotation3 likewise
#pragma unroll
        for (int c = 0; c < D; ++c) {
            double o = x[c] - a[c];
            if (c == wrap_dim) o = wrap_pi(o);
            dst[(size_t)n * D + c] = (float)o;
        }
    }
}
// anchored upload: the caller supplies the Float64 anchors [nvars][D] and float32 offsets [nvars][N][D] (half the bytes
// of the Float64 coordinates); this kernel only lays them out as blocks {header, [Npad][D] offsets}
template <int D>
__global__ void pack_anchored_kernel(int nvars, int N, int Npad, int wrap_dim, const double* __restrict__ anchors,
                                     const float* __restrict__ offs, unsigned char* __restrict__ store) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const double* a = anchors + (size_t)v * D;
    unsigned char* blk = store + (size_t)v * var_block_bytes(D, Npad);
    double* hdr = reinterpret_cast<double*>(blk);
    if (lane < var_header_bytes(D) / 8) {
        double h = lane < D ? a[lane] : 0.0;
        if (wrap_dim >= 0 && lane == D) h = cos(a[wrap_dim]);
        if (wrap_dim >= 0 && lane == D + 1) h = sin(a[wrap_dim]);
        hdr[lane] = h;
        if (wrap_dim >= 0 && lane == D + 2)
            *reinterpret_cast<float2*>(hdr + lane) = make_float2((float)cos(a[wrap_dim]), (float)sin(a[wrap_dim]));
    }
    const float* src = offs + (size_t)v * N * D;
    float* dst = reinterpret_cast<float*>(blk + var_header_bytes(D));
    for (int i = lane; i < Npad * D; i += 32) dst[i] = i < N * D ? src[i] : 0.f;
}
template <int D>
__global__ void unpack_kernel(int nvars, int N, int Npad, int wrap_dim, const unsigned char* __restrict__ store,
                              double* __restrict__ coords) {
    const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (v >= nvars) return;
    const unsigned char* blk = store + (size_t)v * var_block_bytes(D, Npad);
    const double* hdr = reinterpret_cast<const double*>(blk);
    const float* src = reinterpret_cast<const float*>(blk + var_header_bytes(D));
    double* dst = coords + (size_t)v * N * D;
    for (int n = lane; n < N; n += 32) {
        double x[D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
            x[c] = hdr[c] + (double)src[(size_t)n * D + c];
            if (c == wrap_dim) x[c] = wrap_pi(x[c]);
        }
        if (D == 6) closest_rotvec(x[3], x[4], x[5], 0.0, 0.0, 0.0);  // Pose3: report the principal rotation vector
        if (D == 3 && wrap_dim == -2) closest_rotvec(x[0], x[1], x[2], 0.0, 0.0, 0.0);  // Rotation3 likewise
#pragma unroll
        for (int c = 0; c < D; ++c) dst[(size_t)n * D + c] = x[c];
    }
}
__global__ void adopt_kernel(int d, int Npad, unsigned char* __restrict__ store, int var,
                             const float* __restrict__ prop, int factor) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float* dst = reinterpret_cast<float*>(store + (size_t)var * var_block_bytes(d, Npad) + var_header_bytes(d));
    if (i < d * Npad) dst[i] = prop[(size_t)factor * d * Npad + i];
}

int launch_pack(int d, int wrap_dim, int nvars, int N, int Npad, const double* coords, unsigned char* store,
                void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) pack_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, store);
    else if (d == 2) pack_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, store);
    else if (d == 6) pack_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, coords, store);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_pack_anchored(int d, int wrap_dim, int nvars, int N, int Npad, const double* anchors, const float* offs,
                         unsigned char* store, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) pack_anchored_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, anchors, offs, store);
    else if (d == 2) pack_anchored_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, anchors, offs, store);
    else if (d == 6) pack_anchored_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, anchors, offs, store);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_unpack(int d, int wrap_dim, int nvars, int N, int Npad, const unsigned char* store, double* coords,
                  void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = (nvars + 7) / 8;
    if (nvars == 0) return 0;
    if (d == 3) unpack_kernel<3><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, store, coords);
    else if (d == 2) unpack_kernel<2><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, store, coords);
    else if (d == 6) unpack_kernel<6><<<grid, 256, 0, s>>>(nvars, N, Npad, wrap_dim, store, coords);
    else return (int)cudaErrorInvalidValue;
    return (int)cudaGetLastError();
}
int launch_adopt(int d, int Npad, unsigned char* store, int var, const float* prop, int factor, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = d * Npad;
    adopt_kernel<<<(n + 255) / 256, 256, 0, s>>>(d, Npad, store, var, prop, factor);
    return (int)cudaGetLastError();
}

}  // namespace rome
