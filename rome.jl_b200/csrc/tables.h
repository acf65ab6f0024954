// tables.h -- per-factor table rows (AoS, sized/aligned for 1-D TMA bulk copies) and the kernel
// parameter block shared by the host API (rome_b200_api.cu) and the kernels (factor_kernels.cu).
#pragma once

#include <stdint.h>

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
    int32_t ip, il;
    double mu_b, mu_r;
    float sig_b, sig_r;
};
static_assert(sizeof(RowBR) == 32, "RowBR must be 32 B");

// Pose3Pose3 / PriorPose3. src/factors/Pose3Pose3.jl:9-11, src/factors/Pose3D.jl:9-11
struct alignas(32) RowSE3 {
    int32_t ip, iq;
    double mu[6];
    float L[21];  // row-major lower triangle: L00 L10 L11 L20 L21 L22 ...
    float pad[5];
};
static_assert(sizeof(RowSE3) == 160, "RowSE3 must be 160 B");

struct EvalParams {
    const void* rows;  // factor table (device)
    int first, count;  // factor range [first, first+count)
    int N, Npad;
    const float* v0;  // offsets of the first variable's type  [nvars][d][Npad]
    const double* a0; // anchors                                  [nvars][d]
    const float* v1;  // second variable's type (may alias v0)
    const double* a1;
    const float* meas;
    float* meas_out;
    float* res;
    float* prop_fwd;
    float* prop_bwd;
    float* stats;
    float* jac;
    uint32_t flags;
    uint32_t seed_lo, seed_hi, stream_id;
};

constexpr int kWarpsPerCta = 8;        // one factor per warp per tile
constexpr int kThreads = kWarpsPerCta * 32;

// launchers (factor_kernels.cu); return cudaError_t as int
int launch_eval(int family, const EvalParams& p, int grid, void* stream);
int launch_pack(int d, int wrap_dim, int nvars, int N, int Npad, const double* coords, float* offsets,
                double* anchors, void* stream);
int launch_unpack(int d, int wrap_dim, int nvars, int N, int Npad, const float* offsets, const double* anchors,
                  double* coords, void* stream);
int launch_adopt(int d, int Npad, float* offsets, int var, const float* prop, int factor, void* stream);
int max_resident_ctas(int family, bool sample);

}  // namespace rome
