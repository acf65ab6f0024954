// tables.h -- per-factor table rows (AoS, sized/aligned for 1-D TMA bulk copies), the particle-store
// block layout and the kernel parameter block shared by the host API (rome_b200_api.cu) and the
// kernels (factor_kernels.cu).
#pragma once

#include <stddef.h>
#include <stdint.h>

#include "../../include/rome_b200.h"

namespace rome {

// Pose2Pose2 / PriorPose2: MvNormal(mu, Sigma) -> mu (Float64) + lower Cholesky factor (float32).
// src/factors/Pose2D.jl:30-32, src/factors/PriorPose2.jl:13-15
struct alignas(64) RowSE2 {
    int32_t ip, iq;  // variable indices (iq = -1 for a prior)
    double mu[3];
    float L[6];  // L00 L10 L11 L20 L21 L22
    float pad[2];
};
static_assert(sizeof(RowSE2) == 64, "RowSE2 must be 64 B");

// Pose2Point2BearingRange: Normal(mu_b, sig_b), Normal(mu_r, sig_r). src/factors/BearingRange2D.jl:10-13
struct alignas(32) RowBR {
    int32_t ip, iq;  // iq = landmark index
    double mu_b, mu_r;
    float sig_b, sig_r;
};
static_assert(sizeof(RowBR) == 32, "RowBR must be 32 B");

// Pose3Pose3 / PriorPose3. src/factors/Pose3Pose3.jl:9-11, src/factors/Pose3D.jl:9-11
struct alignas(32) RowSE3 {
    int32_t ip, iq;
    double mu[6];
    float L[21];  // row-major lower triangle: L00 L10 L11 L20 L21 L22 ...
    int32_t ir;   // third variable (Pose3Pose3RotOffset / Pose3Pose3Transform), 0 otherwise
    float pad[4];
};
static_assert(sizeof(RowSE3) == 160, "RowSE3 must be 160 B");

// 2-D Gaussian point factors (PriorPoint2, Point2Point2, Pose2Point2): src/factors/Point2D.jl, Pose2Point2.jl
struct alignas(16) RowPT2 {
    int32_t ip, iq;
    double mu[2];
    float L[3];  // L00 L10 L11
    float pad[3];
};
static_assert(sizeof(RowPT2) == 48, "RowPT2 must be 48 B");
// scalar Normal factors (ranges, bearing): src/factors/Range2D.jl, Bearing2D.jl
struct alignas(32) RowS1 {
    int32_t ip, iq;
    double mu;
    float sigma;
    float pad[3];
};
static_assert(sizeof(RowS1) == 32, "RowS1 must be 32 B");

// Particle store: one contiguous BLOCK per variable so that a single 1-D TMA bulk copy brings a whole
// variable (anchor + all particles) into shared memory:
//     [ anchor header ][ Npad x d float32 offsets, particle-major ]
//   Pose2 (d=3): anchor = {x, y, theta, cos(theta), sin(theta), 0}  (48 B; the heading's cos/sin are computed once
//                when the particles are packed so the kernels only evaluate small-angle polynomials)
//   Point2 (d=2): {x, y} (16 B);  Pose3 (d=6): {x, y, z, wx, wy, wz} (48 B)
__host__ __device__ inline int var_header_bytes(int d) { return d == 2 ? 16 : 48; }
__host__ __device__ inline int var_block_bytes(int d, int Npad) { return var_header_bytes(d) + d * Npad * 4; }

struct EvalParams {
    const void* rows;  // factor table (device)
    int first, count;  // factor range [first, first+count)
    int N, Npad;
    const unsigned char* v0;  // particle store of the first variable's type
    const unsigned char* v1;  // second variable's type (may alias v0; unused for priors)
    const unsigned char* v2;  // third variable's type (families with three variables), else null
    const float* meas;
    float* meas_out;
    float* res;
    float* prop_fwd;
    float* prop_bwd;
    float* stats;
    float* jac;
    uint32_t flags;
    uint32_t seed_lo, seed_hi, stream_id;
    int stages;          // pipeline depth
    int stage_bytes;     // shared memory per stage
    int out_warp_bytes;  // shared-memory output staging per consumer warp (residual rows [+ forward proposal rows])
    int n_peers;         // forward proposals are also stored to these peer-GPU buffers (fused all-gather)
    float* peer_fwd[7];
    // owner-sharded exchange: per-factor destination of the forward / backward proposal row (device pointers, possibly
    // into a peer GPU's memory; 0 = the default row of prop_fwd / prop_bwd); arrays indexed by factor id, or null
    const unsigned long long* fwd_dst;
    const unsigned long long* bwd_dst;
    int fwd_dst_lo, fwd_dst_hi;  // every non-null entry of fwd_dst lies in [lo, hi): the array is only read there
    // rank barrier fused into the launch (ROME_B200_BARRIER_WAIT / _SIGNAL): this rank's state words, the slot it owns
    // in every peer's state, the number of peers, and the give-up limit of the wait in clock cycles
    uint32_t* bar_state;
    uint32_t* bar_peer[7];
    int bar_n;
    int bar_lo, bar_hi;  // BARRIER_WAIT is passed right before the first factor of [bar_lo, bar_hi) is fetched
    long long bar_timeout;
};
// flags that change scheduling / routing only, never which outputs a launch computes (ignored when a compile-time
// flag variant is selected)
constexpr uint32_t kSchedFlags = ROME_B200_SAMPLE | ROME_B200_INDEPENDENT | ROME_B200_ROUTED_ONLY |
                                 ROME_B200_BARRIER_WAIT | ROME_B200_BARRIER_SIGNAL;
// row of the backward proposal of factor f (row_floats = dbwd * Npad)
__host__ __device__ inline float* bwd_row(const EvalParams& P, int f, size_t row_floats) {
    if (P.bwd_dst) {
        const unsigned long long d = P.bwd_dst[f];
        if (d) return reinterpret_cast<float*>(d);
    }
    return P.prop_bwd + (size_t)f * row_floats;
}

struct ProductParams {
    unsigned char* store;        // particle store of the variable type (updated in place)
    const int32_t* var_off;      // [nvars + 1] CSR offsets into the source arrays
    const int32_t* src_buf;      // [nsrc] index into bufs[]
    const int32_t* src_row;      // [nsrc] proposal row (factor index) inside that buffer
    const float* bufs[ROME_B200_MAX_PRODUCT_BUFFERS];
    float* bw_out;               // optional [nsrc][D] bandwidths (diagnostics), may be null
    int nvars, N, Npad, iters;
    uint32_t seed_lo, seed_hi, stream_id;
    float bw_scale;              // rule-of-thumb factor (4 / ((d + 2) N))^(1 / (d + 4))
    int manifold;                // Pose3: rotation part multiplied in the tangent space at the anchor rotation
};

// launch geometry chosen on the host for (family, sample, Npad)
struct LaunchPlan {
    int ft;           // factors per tile == consumer warps per CTA (8, 2 or 1)
    int variant;      // 0: runtime flags; 1: RESIDUAL|STATS; 2: RESIDUAL|STATS|PROPOSAL_FWD (compile-time flags);
                      // 3: RESIDUAL|STATS compile-time + forward rows of the routed factors (ROUTED_ONLY launches)
    int stages;
    int stage_bytes;
    int out_warp_bytes;
    int smem_bytes;   // dynamic shared memory per CTA
    int ctas_per_sm;
    int pipeline;     // 0: producer warp + CTA-wide stages; 1: per-warp pipelines (no producer warp)
};

int plan_launch(int family, uint32_t flags, int Npad, int smem_per_sm, int smem_per_cta_max, LaunchPlan* plan);
int launch_eval(int family, const EvalParams& p, const LaunchPlan& plan, int grid, void* stream);
int launch_pack(int d, int wrap_dim, int nvars, int N, int Npad, const double* coords, unsigned char* store,
                void* stream);
int launch_pack_anchored(int d, int wrap_dim, int nvars, int N, int Npad, const double* anchors, const float* offs,
                         unsigned char* store, void* stream);
int launch_unpack(int d, int wrap_dim, int nvars, int N, int Npad, const unsigned char* store, double* coords,
                  void* stream);
int launch_adopt(int d, int Npad, unsigned char* store, int var, const float* prop, int factor, void* stream);
int launch_product(int d, int wrap_dim, const void* params, int num_sms, void* stream);
int launch_reanchor(int d, int wrap_dim, unsigned char* store, int nvars, int N, int Npad, void* stream);
int launch_halo_push(const unsigned char* store, int block_bytes, int n, const int32_t* src_var,
                     const unsigned long long* dst_blocks, void* stream);
int launch_peer_signal(uint32_t* const* slots, int n_peers, uint32_t* epoch, void* stream);
int launch_peer_barrier(uint32_t* const* slots, int n_peers, uint32_t* state, long long max_cycles, void* stream);
int launch_peer_wait(const uint32_t* flags, int n, uint32_t* epoch, uint32_t* status, long long max_cycles, void* stream);

}  // namespace rome
