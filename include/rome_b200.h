/*
 * rome_b200.h -- C ABI of librome_b200.so: the B200 (sm_100a) implementation of RoME.jl's
 * factor-residual / convolution hot path (BASELINE.json north_star).
 *
 * What it replaces.  In the reference, IncrementalInference calls the factor functor
 *   (cf::CalcFactor{<:F})(meas, vars...)                 once per particle per optimiser step
 * and `getSample(cf)` once per particle per convolution.  This library
 * evaluates the same arithmetic for ALL factors of one family x ALL particles per call:
 *
 *   family ROME_B200_POSE2POSE2    src/factors/Pose2D.jl:51-67  (+ _compose/_vee PriorPose2.jl:19-25)
 *   family ROME_B200_PRIORPOSE2    src/factors/PriorPose2.jl:37-47
 *   family ROME_B200_BEARINGRANGE  src/factors/BearingRange2D.jl:48-64, getSample :17-27
 *   family ROME_B200_POSE3POSE3    src/factors/Pose3Pose3.jl:17-29
 *   family ROME_B200_PRIORPOSE3    src/factors/Pose3D.jl:15-19
 *   default getSample on the `.Z` field (IIF sampleTangent/samplePoint; Pose2D.jl:31,
 *   PriorPose2.jl:14, Pose3Pose3.jl:10, Pose3D.jl:10)   -> ROME_B200_SAMPLE
 *
 * The reference-side binding (Julia `ccall`) is shown in INTEGRATION.md and
 * rome.jl_b200/julia/RoMEB200.jl; tests drive the identical ABI through ctypes.
 *
 * Conventions
 *   - plain C, no exceptions; every call returns 0 (ROME_B200_OK) or a negative status;
 *     rome_b200_last_error() gives the message.  NaN/Inf inputs propagate (no trapping),
 *     as in the reference.
 *   - one rome_b200_ctx per host thread / GPU; calls on one ctx are not re-entrant; different
 *     contexts are independent (no global mutable state).
 *   - the caller owns every buffer it passes; the ctx owns particle storage, factor tables
 *     and staging scratch only.
 *   - variable coordinates use the reference's own layout: Float64, particle-major
 *     `[nvars][N][d]` (DFG `vecval`), d = 3 (Pose2: x,y,theta), 2 (Point2), 6 (Pose3: x,y,z,
 *     rotation vector), 3 (Point3) -- src/variables/VariableTypes.jl:13,23,35,47.
 *   - device layout ("anchored float32"): value = anchor(Float64, per variable / per factor
 *     mean) + offset(float32).  Every row set is PARTICLE-MAJOR like the reference's arrays:
 *     particles: one block per variable {anchor header, [Npad][d] float32 offsets};
 *     measurements [nF][Npad][dm] offsets from the factor mean; residuals [nF][Npad][dr] float32;
 *     proposals [nF][Npad][dv] offsets from the TARGET variable's anchor; jac [nF][Npad][dj].
 *     Npad = N rounded up to a multiple of 8; padding particles hold 0 on input and are ignored.
 *     meas, res and prop_fwd must be 16-byte aligned (they move through 1-D TMA bulk copies).
 *     Arithmetic: Float64 per factor, float32 per particle on the (small) offsets; ROME_B200_PRECISE selects
 *     Float64 throughout.
 */
#ifndef ROME_B200_H
#define ROME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROME_B200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define ROME_B200_API __attribute__((visibility("default")))
#else
#define ROME_B200_API
#endif

typedef struct rome_b200_ctx rome_b200_ctx;

enum rome_b200_status {
    ROME_B200_OK = 0,
    ROME_B200_BAD_ARG = -1,
    ROME_B200_CUDA_ERROR = -2,
    ROME_B200_SHAPE_MISMATCH = -3,
    ROME_B200_NOT_SET = -4,
    ROME_B200_NO_DEVICE = -5
};

/* variable types: src/variables/VariableTypes.jl:35 (Pose2), :13 (Point2), :47 (Pose3) */
/* Point3: src/variables/VariableTypes.jl:23 */
enum rome_b200_vartype {
    ROME_B200_POSE2 = 0,
    ROME_B200_POINT2 = 1,
    ROME_B200_POSE3 = 2,
    ROME_B200_POINT3 = 3,
    ROME_B200_ROTATION3 = 4, /* src/variables/VariableTypes.jl:50 (Rotation3, SO(3)): rotation-vector coordinates */
    ROME_B200_NVARTYPES = 5
};

enum rome_b200_family {
    ROME_B200_POSE2POSE2 = 0,
    ROME_B200_PRIORPOSE2 = 1,
    ROME_B200_BEARINGRANGE = 2,
    ROME_B200_POSE3POSE3 = 3,
    ROME_B200_PRIORPOSE3 = 4,
    /* next-row families, same kernel template */
    ROME_B200_PRIORPOINT2 = 5,        /* src/factors/Point2D.jl:8-18      r = m - x                       */
    ROME_B200_POINT2POINT2 = 6,       /* src/factors/Point2D.jl:25-35     r = m - (xj - xi)               */
    ROME_B200_POSE2POINT2 = 7,        /* src/factors/Pose2Point2.jl:9-40  r = l - (p.t + R_p m)           */
    ROME_B200_POSE2POINT2RANGE = 8,   /* src/factors/Range2D.jl:43-54     r = rho - |l - p.t|             */
    ROME_B200_POINT2POINT2RANGE = 9,  /* src/factors/Range2D.jl:5-18      r = rho - |xj - xi|             */
    ROME_B200_POSE2POINT2BEARING = 10,/* src/factors/Bearing2D.jl:13-32   r = sym_rem(b - atan(R_p'(l-p.t))) */
    ROME_B200_PRIORPOINT3 = 11,        /* src/factors/Point3D.jl:7-20         r = m - x                         */
    ROME_B200_POINT3POINT3 = 12,       /* src/factors/Point3Point3.jl:4-15    r = m - (xj - xi)                 */
    ROME_B200_POSE3POSE3XYYAW = 13,    /* src/factors/PartialPose3.jl:103-134 SE(2) residual of (x, y, yaw)     */
    ROME_B200_POSE3POSE3ROTATION = 14, /* src/factors/PartialPose3.jl:198-226 r = m - Log(R_p' R_q)             */
    ROME_B200_POSE3POSE3UNITTRANS = 15,/* src/factors/Pose3Pose3.jl:100-116   Pose3Pose3, unit translation part */
    /* families with a THIRD variable (rome_b200_set_factors_ternary) */
    ROME_B200_POSE3POSE3ROTOFFSET = 16,/* src/factors/Pose3Pose3.jl:57-78  qhat = p o (m.t, bRa Exp(m.w)), bRa: Rotation3 */
    ROME_B200_POSE3POSE3TRANSFORM = 17,/* src/factors/Pose3Pose3.jl:80-95  qhat = p o Delta o exp(m), Delta: Pose3        */
    ROME_B200_NFAMILIES = 18
};

/* eval flags */
#define ROME_B200_RESIDUAL 1u     /* write the residual coordinates                                  */
#define ROME_B200_PROPOSAL_FWD 2u /* closed-form root for the LAST variable (q / landmark / prior var) */
#define ROME_B200_PROPOSAL_BWD 4u /* closed-form root for the FIRST variable (Pose2Pose2, Pose3Pose3; BearingRange:
                                     the root that keeps the pose particle's current heading)          */
#define ROME_B200_STATS 8u        /* per-factor statistics (warp-shuffle reductions)                 */
#define ROME_B200_SAMPLE 16u      /* getSample fused in-kernel (Philox4x32-10); `meas` is not read   */
#define ROME_B200_WRITE_MEAS 32u  /* with SAMPLE: also store the drawn measurement offsets            */
#define ROME_B200_JACOBIAN 64u    /* compact analytic Jacobian blocks, dj floats per particle (rome_b200_family_dims):
                                     Pose2Pose2 (4): d r_t/d theta_p = (-ry, rx), (cos, sin) of theta_p [d r/d m = R(theta_p) (+) 1];
                                     BearingRange (4): d r1/d l, d r2/d l;
                                     Pose3Pose3 (36), perturbations t <- t + dt, R <- R Exp(delta), four row-major 3x3 blocks
                                       A = d r_t/d delta_p = -R_p [m_t]x, B = d r_w/d delta_p = Jr^-1(r_w) M',
                                       C = d r_w/d delta_q = -Jl^-1(r_w), R_p = d r_t/d m_t
                                       (d r_t/d t_p = I, d r_t/d t_q = -I, d r_t/d delta_q = 0, d r_w/d m_w = B M);
                                     PriorPose3 (9): d r_w/d delta_p = -Jl^-1(r_w) (d r_t/d t_p = -I)          */
/* Scheduling hint: this launch reads nothing that was written by launches issued on the same ctx stream since the
 * last launch WITHOUT this flag (e.g. the family kernels of one Gibbs sweep: all read the same particle state and
 * write different buffers).  It is then launched with programmatic dependent launch and may start on SMs the
 * preceding kernel has already vacated.  Results are identical; only the overlap changes. */
#define ROME_B200_INDEPENDENT 128u
/* Deconvolution (IIF approxDeconv, test/testBasicPose2Conv.jl:51-56): write to `meas_out` the measurement under which
 * each particle pair (or prior particle) has ZERO residual, as offsets from the factor mean -- Pose2Pose2:
 * (R_p'(t_q - t_p), wrap(th_q - th_p)); BearingRange: (bearing, range) of the landmark seen from the pose;
 * Pose3Pose3: (R_p'(t_q - t_p), Log(R_p' R_q)); priors: the particle itself.  Excludes WRITE_MEAS. */
#define ROME_B200_DECONV 256u
/* Per-particle arithmetic in Float64.  Default (flag absent): everything LARGE -- anchors, factor means, the lever arm
 * R(anchor) mu_t, the anchors' own residual -- is folded per factor in Float64, and the per-particle part, which then
 * only involves small offsets, runs in float32 (|error| ~ 2^-24 x the offsets' magnitude, i.e. ~1e-8 on SLAM-sized
 * spreads; tests/test_gpu_parity_raw.py states the bar).  With this flag the whole chain is Float64 on
 * anchor + offset and the only rounding is that of the float32 output.  Jacobian and deconvolution outputs always take
 * the Float64 chain. */
#define ROME_B200_PRECISE 512u
/* Owner-sharded multi-GPU sweeps (see rome_b200_set_proposal_destinations / rome_b200_set_step_barrier below):
 * ROUTED_ONLY    with PROPOSAL_FWD: only factors that HAVE a destination write their forward row (and accumulate
 *                proposal statistics); the others behave as if PROPOSAL_FWD were absent -- one launch covers a rank's
 *                interior and cut factors.
 * BARRIER_WAIT   this launch is the first of a step: before it fetches particle blocks it waits (on the device) until
 *                every peer has signalled the end of its previous step.
 * BARRIER_SIGNAL this launch is the last of a step: when its grid has finished, the next epoch is published to every
 *                peer. */
#define ROME_B200_ROUTED_ONLY 1024u
#define ROME_B200_BARRIER_WAIT 2048u
#define ROME_B200_BARRIER_SIGNAL 4096u

/* Buffers of one eval call.  Unused members may be NULL.  `_host` entry points take host
 * pointers with the same shapes; plain entry points take device pointers. */
typedef struct rome_b200_buffers {
    const float* meas; /* in : [nF][Npad][dm] offsets from the factor mean (ignored with SAMPLE) */
    float* meas_out;   /* out: same shape, with SAMPLE|WRITE_MEAS                               */
    float* res;        /* out: [nF][Npad][dr]                                                   */
    float* prop_fwd;   /* out: [nF][Npad][dv_last]  offsets from the last variable's anchor      */
    float* prop_bwd;   /* out: [nF][Npad][dv_first] offsets from the first variable's anchor     */
    float* stats;      /* out: [nF][nstats], nstats = 16 or 32 (rome_b200_family_dims)          */
    float* jac;        /* out: [nF][Npad][dj] compact Jacobian entries                           */
} rome_b200_buffers;

/* ---- life cycle -------------------------------------------------------------------------- */
ROME_B200_API int rome_b200_version(void);
/* Fails with ROME_B200_NO_DEVICE when no CUDA device is usable: there is NO CPU fallback. */
ROME_B200_API int rome_b200_create(int device, rome_b200_ctx** out);
ROME_B200_API int rome_b200_destroy(rome_b200_ctx* ctx);
ROME_B200_API const char* rome_b200_last_error(const rome_b200_ctx* ctx); /* ctx may be NULL (creation errors) */
/* Use a caller-owned cudaStream_t for all work of this ctx (NULL -> the ctx's own non-blocking stream; pass
 * cudaStreamLegacy / cudaStreamPerThread explicitly to run on a default stream). */
ROME_B200_API int rome_b200_set_stream(rome_b200_ctx* ctx, void* cuda_stream);
ROME_B200_API int rome_b200_synchronize(rome_b200_ctx* ctx);
/* family/vartype dimensions: dm (measurement), dr (residual), stats width, jacobian rows */
ROME_B200_API int rome_b200_family_dims(int family, int* dm, int* dr, int* nstats, int* dj);
ROME_B200_API int rome_b200_vartype_dim(int vartype);
ROME_B200_API int rome_b200_npad(int N);
/* Launch geometry the library would choose on a B200 for (family, flags, N) -- pure host arithmetic, no device needed:
 * consumer warps per CTA (= factors per tile), pipeline stages, CTAs per SM, dynamic shared memory per CTA, and the
 * pipeline kind (0 producer-warp, 1 per-warp).  ROME_B200_SHAPE_MISMATCH when N is too large for the shared-memory
 * pipeline of that family. */
ROME_B200_API int rome_b200_plan_query(int family, uint32_t flags, int N, int* warps, int* stages, int* ctas_per_sm,
                                       int* smem_bytes, int* pipeline);

/* ---- variables: replaces the per-variable `Vector{ArrayPartition}` particle storage -------- */
/* Upload Float64 coordinates [nvars][N][d] (host); converts on device to anchored float32 SoA.
 * anchor of a variable = its first particle.  The variable types a factor family touches must
 * hold the same N when that family is evaluated (checked by rome_b200_eval). */
ROME_B200_API int rome_b200_set_particles(rome_b200_ctx* ctx, int vartype, int nvars, int N, const double* coords_host);
/* Same in the anchored form the device keeps: Float64 anchors [nvars][d] + float32 offsets [nvars][N][d] (offset =
 * coordinate - anchor, heading offsets wrapped to [-pi, pi], Pose3 rotation vectors taken nearest the anchor's) -- half
 * the bytes over PCIe for a host that already holds its particles relative to a per-variable reference point. */
ROME_B200_API int rome_b200_set_particles_anchored(rome_b200_ctx* ctx, int vartype, int nvars, int N,
                                                   const double* anchors_host, const float* offsets_host);
ROME_B200_API int rome_b200_get_particles(rome_b200_ctx* ctx, int vartype, double* coords_host);
/* Device view of the particle store (zero-copy interop): nvars contiguous blocks of `block_bytes`, each
 * { anchor header of `header_bytes` (Pose2: x,y,theta,cos,sin) }{ [Npad][d] float32 offsets }. */
ROME_B200_API int rome_b200_particles_device(rome_b200_ctx* ctx, int vartype, void** d_store, int* block_bytes,
                                             int* header_bytes, int* nvars, int* N, int* Npad);
/* Replace the particles of variable `var` by proposal row `factor` of a device proposal buffer
 * (offsets from that variable's anchor) -- used to chain convolutions (graph init). */
ROME_B200_API int rome_b200_adopt_proposal(rome_b200_ctx* ctx, int vartype, int var, const float* d_prop, int factor);

/* ---- factors: replaces the factor structs' belief fields ----------------------------------- */
/* Pose2Pose2(MvNormal(mu, cov)): src/factors/Pose2D.jl:30-32.  cov row-major [nF][3][3]. */
ROME_B200_API int rome_b200_set_factors_pose2pose2(rome_b200_ctx* ctx, int nF, const int32_t* ip, const int32_t* iq,
                                     const double* mu, const double* cov);
/* PriorPose2(MvNormal(mu, cov)): src/factors/PriorPose2.jl:13-15 */
ROME_B200_API int rome_b200_set_factors_priorpose2(rome_b200_ctx* ctx, int nF, const int32_t* ip, const double* mu,
                                     const double* cov);
/* Pose2Point2BearingRange(Normal(mu_b, sig_b), Normal(mu_r, sig_r)): BearingRange2D.jl:10-13.
 * bearing/range: [nF][2] = (mean, standard deviation). */
ROME_B200_API int rome_b200_set_factors_bearingrange(rome_b200_ctx* ctx, int nF, const int32_t* ip, const int32_t* il,
                                       const double* bearing, const double* range);
/* Pose3Pose3(MvNormal(mu, cov)): src/factors/Pose3Pose3.jl:9-11.  cov row-major [nF][6][6]. */
ROME_B200_API int rome_b200_set_factors_pose3pose3(rome_b200_ctx* ctx, int nF, const int32_t* ip, const int32_t* iq,
                                     const double* mu, const double* cov);
/* PriorPose3(MvNormal(mu, cov)): src/factors/Pose3D.jl:9-11 */
ROME_B200_API int rome_b200_set_factors_priorpose3(rome_b200_ctx* ctx, int nF, const int32_t* ip, const double* mu,
                                     const double* cov);
/* 2-D Gaussian point factors: family in {PRIORPOINT2 (i1 = NULL), POINT2POINT2, POSE2POINT2};
 * mu [nF][2], cov row-major [nF][2][2] */
ROME_B200_API int rome_b200_set_factors_point2(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0, const int32_t* i1,
                                               const double* mu, const double* cov);
/* scalar factors with a Normal belief: family in {POSE2POINT2RANGE, POINT2POINT2RANGE, POSE2POINT2BEARING};
 * belief [nF][2] = (mean, standard deviation) */
ROME_B200_API int rome_b200_set_factors_scalar(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0, const int32_t* i1,
                                               const double* belief);
/* Any family whose belief is ONE MvNormal(mu[dm], cov[dm][dm]) (every family except BEARINGRANGE and the scalar
 * ones): i1 = NULL for priors.  The typed entry points above are thin wrappers of this one. */
ROME_B200_API int rome_b200_set_factors_gaussian(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0,
                                                 const int32_t* i1, const double* mu, const double* cov);
/* Families with a third variable (Pose3Pose3RotOffset: i2 indexes Rotation3 variables; Pose3Pose3Transform: i2 indexes
 * Pose3 variables): one MvNormal(mu[6], cov[6x6]) per factor like rome_b200_set_factors_pose3pose3.  The residual and
 * the forward proposal (onto the SECOND variable) are closed-form; convolutions onto the first or third variable are
 * left to the caller's numeric solver, as in the reference. */
ROME_B200_API int rome_b200_set_factors_ternary(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0,
                                                const int32_t* i1, const int32_t* i2, const double* mu,
                                                const double* cov);
ROME_B200_API int rome_b200_num_factors(rome_b200_ctx* ctx, int family);

/* ---- the hot path --------------------------------------------------------------------------- */
/* Evaluate factors [first, first+count) of `family` (count < 0: through the last factor) for all
 * N particles.  Buffers are indexed by GLOBAL factor id, so ranks of a multi-GPU job that split the
 * factor list write disjoint slices of identically shaped buffers.  Asynchronous on the ctx stream.
 * (seed, stream_id, factor, particle) keys the in-kernel sampler. */
ROME_B200_API int rome_b200_eval(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed, uint32_t stream_id,
                   int first, int count, const rome_b200_buffers* device_buffers);
/* Same with HOST buffers: copies `meas` in (unless SAMPLE), launches, copies every requested output
 * back, and synchronises.  This is the call a host-resident caller (Julia/IIF) makes. */
ROME_B200_API int rome_b200_eval_host(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed, uint32_t stream_id,
                        int first, int count, const rome_b200_buffers* host_buffers);

/* Same without the final synchronisation: copies and kernels are only enqueued on the ctx stream; host buffers must
 * be pinned (rome_b200_malloc_host) and stay untouched until rome_b200_synchronize(ctx).  Lets a caller overlap the
 * upload of one context with the download of another (two contexts, two streams). */
ROME_B200_API int rome_b200_eval_host_async(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed,
                                            uint32_t stream_id, int first, int count,
                                            const rome_b200_buffers* host_buffers);

/* ---- the step after the path: belief update = product of the proposal densities -------------
 * Replaces the host hop  proposals -> AMP.manifoldProduct / manikde! -> new particles  (IIF propagateBelief; callers
 * test/testBearingRange2D.jl:275,296,342, test/testBasicPose2Conv.jl:25-34): every proposal row is a kernel density
 * estimate (per-dimension rule-of-thumb bandwidth from its own particles, circular statistics for headings); N samples
 * of the product of a variable's k proposal KDEs are drawn by Gibbs sampling over the component labels and written
 * into the device particle store of that variable type, so particles stay on the GPU between sweeps.
 * The plan lists, per variable (CSR `var_offsets[nvars+1]`), its sources: buffer index `src_buf` into the
 * `d_prop_bufs` array given to rome_b200_product, and proposal row `src_row` (the factor index) inside that buffer.
 * Rows must hold offsets from THAT variable's anchor (prop_fwd of factors whose last variable it is, prop_bwd of
 * factors whose first variable it is) and `src_row` must lie inside the buffer passed at that index (not checked: the
 * library does not know the buffers' sizes).  Variables without sources keep their particles; one source is adopted as is. */
#define ROME_B200_MAX_PRODUCT_SOURCES 32 /* proposals per variable */
#define ROME_B200_MAX_PRODUCT_BUFFERS 16 /* distinct proposal buffers per call */
#define ROME_B200_PRODUCT_REANCHOR 1u    /* afterwards move every anchor onto the variable's new first particle.  Proposal rows are
                                          * offsets from the anchor their evaluation saw: rows written BEFORE a re-anchoring must not
                                          * be multiplied again afterwards (evaluate again first) */
#define ROME_B200_PRODUCT_MANIFOLD 2u    /* Pose3: the rotation part of every proposal is taken to the tangent space at the
                                          * variable's anchor rotation, xi = Log(R_anchor^-1 R) (how ApproxManifoldProducts -- absent
                                          * third-party package -- treats group-valued KDE points; the reference's own manifold
                                          * glue for it: src/services/FixmeManifolds.jl:22-38), the product is
                                          * sampled there and retracted, R_anchor Exp(xi), instead of multiplying rotation-vector
                                          * offsets as Euclidean coordinates.  Other variable types ignore the flag; a variable whose proposal rows
                                          * do not fit the kernel's 20 KB shared-memory pool (more than 7 proposals at N = 100) is
                                          * multiplied in chart coordinates as without the flag. */
ROME_B200_API int rome_b200_set_product_plan(rome_b200_ctx* ctx, int vartype, int nvars, const int32_t* var_offsets,
                                             const int32_t* src_buf, const int32_t* src_row);
/* gibbs_iters <= 0 selects the default (2; only variables with more than two proposals iterate).  d_bw_out: optional device [nsrc][d] bandwidths (diagnostics). */
ROME_B200_API int rome_b200_product(rome_b200_ctx* ctx, int vartype, int n_bufs, const float* const* d_prop_bufs,
                                    uint64_t seed, uint32_t stream_id, int gibbs_iters, uint32_t flags, float* d_bw_out);
ROME_B200_API int rome_b200_reanchor(rome_b200_ctx* ctx, int vartype);

/* ---- multi-GPU: proposals written straight into the peers' buffers (fused compute + all-gather) ------------
 * With peers set for a family, every evaluation with ROME_B200_PROPOSAL_FWD stores each factor's forward-proposal
 * rows not only to `prop_fwd` but also, from the same shared-memory slice and by the same warp-local TMA bulk
 * stores, to the identically indexed buffer of every peer (device pointers into the peers' memory, obtained with
 * rome_b200_ipc_import over NVLink).  Ranks that split the factor list therefore end the kernel holding all
 * proposals -- no separate all-gather pass; only a stream-ordered barrier between the ranks is still needed.
 * A sweep LOOP needs TWO barriers per sweep (or receive buffers alternating with the sweep's parity): one after the
 * evaluation (every peer's rows have landed before the belief update reads them) and one after the belief update
 * (no peer's next evaluation may store into a buffer this rank's product is still reading) -- the order
 * OwnerShardedSolver.sweep keeps: eval, barrier A, product + halo push, barrier B.
 * n_peers <= 7; n_peers = 0 clears. */
ROME_B200_API int rome_b200_set_peer_proposals(rome_b200_ctx* ctx, int family, int n_peers, float* const* peer_prop_fwd);
/* ---- multi-GPU, owner-sharded (the exchange of a sweep whose variables AND factors are partitioned) ---------------
 * Every rank owns a contiguous range of variables and evaluates the factors placed on it; only what a cut edge needs
 * crosses NVLink: the proposal row of a factor whose target variable lives on another rank, and the particle blocks of
 * the foreign ("halo") variables its factors read.
 * rome_b200_set_proposal_destinations: rows[f] = device address the forward (direction 0) / backward (direction 1)
 * proposal row of factor f is written to INSTEAD of row f of prop_fwd / prop_bwd, or NULL for that default; typically
 * a row of a receive buffer in the owner's memory (rome_b200_ipc_import).  Forward rows leave through the kernel's own
 * TMA bulk stores, backward rows through its streaming stores -- the transfer rides on the evaluation, tile by tile.
 * nF must equal the number of factors of the family; nF = 0 clears.
 * rome_b200_set_halo_plan / rome_b200_push_halo: copy the particle blocks (anchor header + offsets) of the listed local
 * variables to the given device addresses (slots in the peers' particle stores, rome_b200_particles_device of the peer
 * + slot * block_bytes, imported through CUDA IPC); run it after the belief update of a sweep.
 * Order both against the peers with rome_b200_peer_signal / rome_b200_peer_wait. */
ROME_B200_API int rome_b200_set_proposal_destinations(rome_b200_ctx* ctx, int family, int direction, int nF,
                                                      void* const* rows);
/* Variables [0, n_owned) of the type are this rank's own: rome_b200_product / rome_b200_reanchor update only those (the
 * product plan then covers n_owned variables); the slots behind them are halo copies written by their owners' pushes.
 * n_owned < 0: all variables (default). */
ROME_B200_API int rome_b200_set_owned_variables(rome_b200_ctx* ctx, int vartype, int n_owned);
ROME_B200_API int rome_b200_set_halo_plan(rome_b200_ctx* ctx, int vartype, int n, const int32_t* src_var,
                                          void* const* dst_blocks);
ROME_B200_API int rome_b200_push_halo(rome_b200_ctx* ctx, int vartype);
/* Rank barrier FUSED into the evaluation kernels (flags ROME_B200_BARRIER_WAIT / ROME_B200_BARRIER_SIGNAL): `d_state`
 * and `peer_slots` as for rome_b200_peer_signal.  Protocol per step and rank: the FIRST launch carries BARRIER_WAIT, the
 * LAST launch (one that is not empty) carries BARRIER_SIGNAL; a single launch may carry both.  Every rank must execute
 * the same number of steps.  The wait target is this rank's own signal count, so rome_b200_peer_signal / _wait and the
 * fused flags must not be interleaved on one state buffer except that a completed signal + wait PAIR of the kernels may
 * precede the first fused step (set-up barrier).  n_peers = 0 clears. */
ROME_B200_API int rome_b200_set_step_barrier(rome_b200_ctx* ctx, void* d_state, uint32_t* const* peer_slots, int n_peers);
/* Only factors [first, first + count) of the family depend on the peers (they read halo blocks / write into peer memory:
 * a rank's cut factors, uploaded as one block of the table): a launch flagged BARRIER_WAIT evaluates everything else
 * without waiting and passes the barrier only before it fetches the first factor of that range, so the barrier's latency
 * hides behind the interior work.  count < 0 restores the default: every factor depends on the peers. */
ROME_B200_API int rome_b200_set_barrier_range(rome_b200_ctx* ctx, int family, int first, int count);
/* CUDA IPC plumbing for buffers allocated with rome_b200_malloc_device (64-byte opaque handles).  Only such buffers may
 * be exported: rome_b200_malloc_device hands out whole 2 MiB blocks, so the handle (which names the driver's block) and
 * the buffer coincide; a pointer into a packed small cudaMalloc allocation would be opened at the wrong address. */
ROME_B200_API int rome_b200_ipc_export(rome_b200_ctx* ctx, void* dev_ptr, unsigned char handle[64]);
ROME_B200_API int rome_b200_ipc_import(rome_b200_ctx* ctx, const unsigned char handle[64], void** dev_ptr);
ROME_B200_API int rome_b200_ipc_close(rome_b200_ctx* ctx, void* dev_ptr);

/* Stream-ordered barrier between the ranks, carried by the GPUs over NVLink peer memory (closes the fused exchange: no
 * NCCL kernel, no host round trip).  `d_state` = one device buffer per rank of ROME_B200_PEER_STATE_WORDS uint32 words
 * from rome_b200_malloc_device, ZEROED by the caller (rome_b200_memcpy_h2d): words [0, 8) are the flag slots the peers
 * write, word 8 / 9 the signal / wait epochs, word 10 the give-up status.  `peer_slots[r]` = device pointer to the slot
 * THIS rank owns inside peer r's state buffer (rome_b200_ipc_import of the peer's buffer + 4 * my_slot).
 * rome_b200_peer_signal: after everything enqueued so far on the ctx stream, publish the next epoch to every peer.
 * rome_b200_peer_wait: the stream continues once every one of the `n_slots` listed local slots has reached this rank's
 * next wait epoch; gives up after ~2 s (a peer died) and sets the status word, read by rome_b200_peer_status.
 * rome_b200_peer_barrier: both in ONE launch (signal, then wait on local slots 0..n_peers-1) -- the closing barrier of a
 * step for hosts without NCCL (what `bench.py --barrier flags` and OwnerShardedSolver use; validated at 2, 4 and 8 GPUs).
 * Epochs live on the device, so the calls may be captured in a CUDA graph and replayed. */
#define ROME_B200_PEER_STATE_WORDS 16
ROME_B200_API int rome_b200_peer_signal(rome_b200_ctx* ctx, void* d_state, uint32_t* const* peer_slots, int n_peers);
ROME_B200_API int rome_b200_peer_wait(rome_b200_ctx* ctx, void* d_state, const int32_t* slots, int n_slots);
ROME_B200_API int rome_b200_peer_barrier(rome_b200_ctx* ctx, void* d_state, uint32_t* const* peer_slots, int n_peers);
ROME_B200_API int rome_b200_peer_status(rome_b200_ctx* ctx, void* d_state, int* gave_up);

/* ---- CUDA-graph capture of a sweep (several eval calls replayed with one launch) ------------- */
ROME_B200_API int rome_b200_graph_begin(rome_b200_ctx* ctx);
ROME_B200_API int rome_b200_graph_end(rome_b200_ctx* ctx, int* graph_id);
ROME_B200_API int rome_b200_graph_launch(rome_b200_ctx* ctx, int graph_id);

/* ---- memory helpers for callers without a CUDA runtime binding ------------------------------- */
ROME_B200_API int rome_b200_malloc_device(rome_b200_ctx* ctx, size_t bytes, void** out);
ROME_B200_API int rome_b200_free_device(rome_b200_ctx* ctx, void* p);
ROME_B200_API int rome_b200_malloc_host(rome_b200_ctx* ctx, size_t bytes, void** out); /* pinned */
ROME_B200_API int rome_b200_free_host(rome_b200_ctx* ctx, void* p);
ROME_B200_API int rome_b200_memcpy_h2d(rome_b200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
ROME_B200_API int rome_b200_memcpy_d2h(rome_b200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- instrumentation --------------------------------------------------------------------------- */
/* Number of hot-path kernel launches issued by this ctx since creation (graph replays count the
 * kernels inside the graph). */
ROME_B200_API uint64_t rome_b200_launch_count(const rome_b200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ROME_B200_H */
