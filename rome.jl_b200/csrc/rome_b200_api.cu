// rome_b200_api.cu -- the C ABI (include/rome_b200.h): context, particle/factor stores, eval
// entry points, CUDA-graph capture.  No CPU fallback: every compute entry point needs a device.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rome_b200.h"
#include "tables.h"

using namespace rome;

namespace {

struct VarStore {
    int nvars = 0, N = 0, Npad = 0;
    unsigned char* store = nullptr;  // nvars blocks of var_block_bytes(d, Npad): {anchor f64[d] (padded), d rows x Npad f32}
    size_t cap = 0;                  // bytes
};
struct FactorStore {
    int nF = 0;
    void* rows = nullptr;
    size_t cap = 0;
    int max_i0 = -1, max_i1 = -1, max_i2 = -1;
};
struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};
struct ProductPlan {
    int nvars = 0, nsrc = 0, max_buf = -1;
    Scratch off, buf, row;  // device copies of the CSR arrays
};

const int kVarDim[ROME_B200_NVARTYPES] = {3, 2, 6, 3, 3};
// heading coordinate that wraps (Pose2), -1: none, -2: the three coordinates are a rotation vector (Rotation3; the
// rotation part of Pose3 is recognised by d = 6)
const int kWrapDim[ROME_B200_NVARTYPES] = {2, -1, -1, -1, -2};
// third variable type of a family (-1: none)
inline int fam_vt2(int family) {
    return family == ROME_B200_POSE3POSE3ROTOFFSET ? ROME_B200_ROTATION3
         : family == ROME_B200_POSE3POSE3TRANSFORM ? ROME_B200_POSE3 : -1;
}
// family -> (first variable type, second variable type or -1, dm, dr, nstats, dj, row bytes, prop dims fwd/bwd)
struct FamInfo {
    int vt0, vt1, dm, dr, nstats, dj, row_bytes, dfwd, dbwd;
};
const FamInfo kFam[ROME_B200_NFAMILIES] = {
    {ROME_B200_POSE2, ROME_B200_POSE2, 3, 3, 16, 4, (int)sizeof(RowSE2), 3, 3},
    {ROME_B200_POSE2, -1, 3, 3, 16, 0, (int)sizeof(RowSE2), 3, 0},
    {ROME_B200_POSE2, ROME_B200_POINT2, 2, 2, 16, 4, (int)sizeof(RowBR), 2, 3},
    {ROME_B200_POSE3, ROME_B200_POSE3, 6, 6, 32, 36, (int)sizeof(RowSE3), 6, 6},
    {ROME_B200_POSE3, -1, 6, 6, 32, 9, (int)sizeof(RowSE3), 6, 0},
    {ROME_B200_POINT2, -1, 2, 2, 16, 0, (int)sizeof(RowPT2), 2, 0},                // PriorPoint2
    {ROME_B200_POINT2, ROME_B200_POINT2, 2, 2, 16, 0, (int)sizeof(RowPT2), 2, 0},  // Point2Point2
    {ROME_B200_POSE2, ROME_B200_POINT2, 2, 2, 16, 0, (int)sizeof(RowPT2), 2, 0},   // Pose2Point2
    {ROME_B200_POSE2, ROME_B200_POINT2, 1, 1, 16, 0, (int)sizeof(RowS1), 0, 0},    // Pose2Point2Range
    {ROME_B200_POINT2, ROME_B200_POINT2, 1, 1, 16, 0, (int)sizeof(RowS1), 0, 0},   // Point2Point2Range
    {ROME_B200_POSE2, ROME_B200_POINT2, 1, 1, 16, 0, (int)sizeof(RowS1), 0, 0},    // Pose2Point2Bearing
    {ROME_B200_POINT3, -1, 3, 3, 16, 0, (int)sizeof(RowSE2), 3, 0},                // PriorPoint3
    {ROME_B200_POINT3, ROME_B200_POINT3, 3, 3, 16, 0, (int)sizeof(RowSE2), 3, 0},  // Point3Point3
    {ROME_B200_POSE3, ROME_B200_POSE3, 3, 3, 16, 0, (int)sizeof(RowSE2), 0, 0},    // Pose3Pose3XYYaw
    {ROME_B200_POSE3, ROME_B200_POSE3, 3, 3, 16, 0, (int)sizeof(RowSE2), 0, 0},    // Pose3Pose3Rotation
    {ROME_B200_POSE3, ROME_B200_POSE3, 6, 6, 32, 0, (int)sizeof(RowSE3), 0, 0},    // Pose3Pose3UnitTrans
    {ROME_B200_POSE3, ROME_B200_POSE3, 6, 6, 32, 0, (int)sizeof(RowSE3), 6, 0},    // Pose3Pose3RotOffset (+ Rotation3)
    {ROME_B200_POSE3, ROME_B200_POSE3, 6, 6, 32, 0, (int)sizeof(RowSE3), 6, 0},    // Pose3Pose3Transform (+ Pose3)
};

thread_local std::string g_create_error;

}  // namespace

struct rome_b200_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    VarStore vars[ROME_B200_NVARTYPES];
    FactorStore fac[ROME_B200_NFAMILIES];
    int smem_per_sm = 0, smem_per_cta_max = 0;
    Scratch stage_dev, stage_host;            // particle upload/download staging
    Scratch out_dev[ROME_B200_NFAMILIES][8];  // eval_host device mirrors per family: meas, meas_out, res, fwd, bwd, stats, jac
    ProductPlan plan[ROME_B200_NVARTYPES];
    int n_peers[ROME_B200_NFAMILIES] = {};
    float* peers[ROME_B200_NFAMILIES][7] = {};
    Scratch row_dst[ROME_B200_NFAMILIES][2];     // per-factor proposal row destinations (fwd, bwd), device arrays
    int row_dst_n[ROME_B200_NFAMILIES][2] = {};  // number of factors they cover (0 = not set)
    int row_dst_lo[ROME_B200_NFAMILIES][2] = {}; // factors [lo, hi) hold every non-null destination (a rank's cut block):
    int row_dst_hi[ROME_B200_NFAMILIES][2] = {}; // the kernels read the destination array only inside this range
    Scratch halo_src[ROME_B200_NVARTYPES], halo_dst[ROME_B200_NVARTYPES];
    int halo_n[ROME_B200_NVARTYPES] = {};
    uint32_t* bar_state = nullptr;   // fused step barrier (rome_b200_set_step_barrier)
    uint32_t* bar_peer[7] = {};
    int bar_n = 0;
    int bar_lo[ROME_B200_NFAMILIES] = {};    // factors of a family that depend on the peers (rome_b200_set_barrier_range);
    int bar_hi[ROME_B200_NFAMILIES];         // default [0, INT_MAX): all
    int owned[ROME_B200_NVARTYPES] = {-1, -1, -1, -1};  // variables [0, owned) are updated by product / reanchor (-1: all)
    std::vector<cudaGraphExec_t> graphs;
    std::vector<uint64_t> graph_kernels;
    bool capturing = false;
    uint64_t capture_kernels = 0;
    uint64_t launches = 0;
};

namespace {

int fail(rome_b200_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_error = msg;
    return code;
}
int cuda_fail(rome_b200_ctx* c, cudaError_t e, const char* what) {
    return fail(c, ROME_B200_CUDA_ERROR, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

int bind(rome_b200_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    return 0;
}
int grow_dev(rome_b200_ctx* ctx, Scratch& s, size_t bytes) {
    if (bytes <= s.cap) return 0;
    if (s.p) CK(cudaFree(s.p));
    s.p = nullptr; s.cap = 0;
    CK(cudaMalloc(&s.p, bytes));
    s.cap = bytes;
    return 0;
}
int grow_host(rome_b200_ctx* ctx, Scratch& s, size_t bytes) {
    if (bytes <= s.cap) return 0;
    if (s.p) CK(cudaFreeHost(s.p));
    s.p = nullptr; s.cap = 0;
    CK(cudaMallocHost(&s.p, bytes));
    s.cap = bytes;
    return 0;
}
bool is_pinned_or_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// lower Cholesky factor of a symmetric d x d matrix (row-major), Float64; false if not positive definite
bool cholesky(const double* A, int d, double* L) {
    for (int i = 0; i < d * d; ++i) L[i] = 0.0;
    for (int j = 0; j < d; ++j) {
        double s = A[j * d + j];
        for (int k = 0; k < j; ++k) s -= L[j * d + k] * L[j * d + k];
        if (!(s > 0.0)) return false;
        L[j * d + j] = std::sqrt(s);
        for (int i = j + 1; i < d; ++i) {
            double t = 0.5 * (A[i * d + j] + A[j * d + i]);
            for (int k = 0; k < j; ++k) t -= L[i * d + k] * L[j * d + k];
            L[i * d + j] = t / L[j * d + j];
        }
    }
    return true;
}

int upload_rows(rome_b200_ctx* ctx, int family, const void* host_rows, int nF, int max_i0, int max_i1) {
    if (int e = bind(ctx)) return e;
    FactorStore& fs = ctx->fac[family];
    const size_t bytes = (size_t)nF * kFam[family].row_bytes;
    if (bytes > fs.cap) {
        if (fs.rows) CK(cudaFree(fs.rows));
        fs.rows = nullptr; fs.cap = 0;
        CK(cudaMalloc(&fs.rows, bytes ? bytes : 256));
        fs.cap = bytes ? bytes : 256;
    }
    if (bytes) {
        CK(cudaMemcpyAsync(fs.rows, host_rows, bytes, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));  // host_rows is a temporary
    }
    fs.nF = nF; fs.max_i0 = max_i0; fs.max_i1 = max_i1; fs.max_i2 = -1;
    return 0;
}

template <class Row, int D>
int set_gaussian_factors(rome_b200_ctx* ctx, int family, int nF, const int32_t* ip, const int32_t* iq,
                         const double* mu, const double* cov, const int32_t* ir = nullptr) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (nF < 0 || (nF > 0 && (!ip || !mu || !cov))) return fail(ctx, ROME_B200_BAD_ARG, "null factor arrays");
    std::vector<Row> rows((size_t)nF);
    int m0 = -1, m1 = -1, m2 = -1;
    double L[D * D];
    for (int f = 0; f < nF; ++f) {
        Row& r = rows[f];
        std::memset(&r, 0, sizeof(Row));
        r.ip = ip[f];
        r.iq = iq ? iq[f] : -1;
        if (r.ip < 0 || (iq && r.iq < 0)) return fail(ctx, ROME_B200_BAD_ARG, "negative variable index");
        if (r.ip > m0) m0 = r.ip;
        if (r.iq > m1) m1 = r.iq;
        if constexpr (D == 6) {
            if (ir) {
                if (ir[f] < 0) return fail(ctx, ROME_B200_BAD_ARG, "negative variable index");
                r.ir = ir[f];
                if (r.ir > m2) m2 = r.ir;
            }
        }
        for (int i = 0; i < D; ++i) r.mu[i] = mu[(size_t)f * D + i];
        if (!cholesky(cov + (size_t)f * D * D, D, L)) {
            char b[96];
            std::snprintf(b, sizeof b, "covariance of factor %d is not positive definite", f);
            return fail(ctx, ROME_B200_BAD_ARG, b);
        }
        int k = 0;
        for (int i = 0; i < D; ++i)
            for (int j = 0; j <= i; ++j) r.L[k++] = (float)L[i * D + j];
    }
    if (int e = upload_rows(ctx, family, rows.data(), nF, m0, m1)) return e;
    ctx->fac[family].max_i2 = m2;
    return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

int rome_b200_version(void) { return ROME_B200_VERSION; }

int rome_b200_create(int device, rome_b200_ctx** out) {
    if (!out) return fail(nullptr, ROME_B200_BAD_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, ROME_B200_NO_DEVICE,
                    std::string("no CUDA device (librome_b200 has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, ROME_B200_BAD_ARG, "device index out of range");
    rome_b200_ctx* ctx = new (std::nothrow) rome_b200_ctx();
    if (!ctx) return fail(nullptr, ROME_B200_BAD_ARG, "out of host memory");
    ctx->device = device;
    for (int& h : ctx->bar_hi) h = 0x7fffffff;
    if ((e = cudaSetDevice(device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, ROME_B200_CUDA_ERROR, std::string("context creation: ") + cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    int cc_major = 0;
    cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device);
    cudaDeviceGetAttribute(&ctx->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    cudaDeviceGetAttribute(&ctx->smem_per_cta_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (cc_major != 10) {  // the library carries sm_100a code only (TMA bulk copies, mbarrier pipeline)
        cudaStreamDestroy(ctx->own_stream);
        delete ctx;
        return fail(nullptr, ROME_B200_NO_DEVICE, "device is not sm_100 (B200): librome_b200 is built for sm_100a only");
    }
    *out = ctx;
    return ROME_B200_OK;
}

int rome_b200_destroy(rome_b200_ctx* ctx) {
    if (!ctx) return ROME_B200_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto g : ctx->graphs) cudaGraphExecDestroy(g);
    for (auto& v : ctx->vars) cudaFree(v.store);
    for (auto& f : ctx->fac) cudaFree(f.rows);
    cudaFree(ctx->stage_dev.p);
    cudaFreeHost(ctx->stage_host.p);
    for (auto& f : ctx->out_dev) for (auto& s : f) cudaFree(s.p);
    for (auto& pl : ctx->plan) { cudaFree(pl.off.p); cudaFree(pl.buf.p); cudaFree(pl.row.p); }
    for (auto& f : ctx->row_dst) for (auto& s : f) cudaFree(s.p);
    for (auto& s : ctx->halo_src) cudaFree(s.p);
    for (auto& s : ctx->halo_dst) cudaFree(s.p);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return ROME_B200_OK;
}

const char* rome_b200_last_error(const rome_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int rome_b200_set_stream(rome_b200_ctx* ctx, void* cuda_stream) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "cannot change stream during graph capture");
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return ROME_B200_OK;
}

int rome_b200_synchronize(rome_b200_ctx* ctx) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaStreamSynchronize(ctx->stream));
    return ROME_B200_OK;
}

int rome_b200_family_dims(int family, int* dm, int* dr, int* nstats, int* dj) {
    if (family < 0 || family >= ROME_B200_NFAMILIES) return ROME_B200_BAD_ARG;
    if (dm) *dm = kFam[family].dm;
    if (dr) *dr = kFam[family].dr;
    if (nstats) *nstats = kFam[family].nstats;
    if (dj) *dj = kFam[family].dj;
    return ROME_B200_OK;
}
int rome_b200_vartype_dim(int vartype) {
    return (vartype < 0 || vartype >= ROME_B200_NVARTYPES) ? ROME_B200_BAD_ARG : kVarDim[vartype];
}
int rome_b200_npad(int N) { return N <= 0 ? ROME_B200_BAD_ARG : (N + 7) / 8 * 8; }
int rome_b200_plan_query(int family, uint32_t flags, int N, int* warps, int* stages, int* ctas_per_sm, int* smem_bytes,
                         int* pipeline) {
    if (family < 0 || family >= ROME_B200_NFAMILIES || N <= 0) return ROME_B200_BAD_ARG;
    LaunchPlan plan;
    // B200 (sm_100): 228 KB shared memory per SM, 227 KB opt-in maximum per CTA
    if (plan_launch(family, flags, rome_b200_npad(N), 233472, 232448, &plan)) return ROME_B200_SHAPE_MISMATCH;
    if (warps) *warps = plan.ft;
    if (stages) *stages = plan.stages;
    if (ctas_per_sm) *ctas_per_sm = plan.ctas_per_sm;
    if (smem_bytes) *smem_bytes = plan.smem_bytes;
    if (pipeline) *pipeline = plan.pipeline;
    return ROME_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int rome_b200_set_particles(rome_b200_ctx* ctx, int vartype, int nvars, int N, const double* coords_host) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    if (nvars < 0 || N <= 0 || (nvars > 0 && !coords_host)) return fail(ctx, ROME_B200_BAD_ARG, "bad particle shape");
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "set_particles during graph capture");
    if (int e = bind(ctx)) return e;
    const int d = kVarDim[vartype], Npad = rome_b200_npad(N);
    VarStore& vs = ctx->vars[vartype];
    const size_t store_bytes = (size_t)nvars * var_block_bytes(d, Npad);
    if (store_bytes > vs.cap) {
        if (vs.store) CK(cudaFree(vs.store));
        vs.store = nullptr; vs.cap = 0;
        // whole 2 MiB blocks: the store may be exported through CUDA IPC (halo blocks pushed by peer GPUs), and an IPC
        // handle names the driver's allocation block, see rome_b200_malloc_device
        const size_t kBlock = size_t(2) << 20;
        const size_t rounded = (store_bytes + kBlock - 1) / kBlock * kBlock;
        CK(cudaMalloc(&vs.store, rounded));
        vs.cap = rounded;
    }
    vs.nvars = nvars; vs.N = N; vs.Npad = Npad;
    if (nvars == 0) return ROME_B200_OK;
    const size_t in_bytes = (size_t)nvars * N * d * sizeof(double);
    if (int e = grow_dev(ctx, ctx->stage_dev, in_bytes)) return e;
    const void* src = coords_host;
    if (!is_pinned_or_device(coords_host)) {  // pageable caller memory: stage through pinned memory
        if (int e = grow_host(ctx, ctx->stage_host, in_bytes)) return e;
        CK(cudaStreamSynchronize(ctx->stream));  // staging buffer may still be in flight
        std::memcpy(ctx->stage_host.p, coords_host, in_bytes);
        src = ctx->stage_host.p;
    }
    CK(cudaMemcpyAsync(ctx->stage_dev.p, src, in_bytes, cudaMemcpyDefault, ctx->stream));
    int e = launch_pack(d, kWrapDim[vartype], nvars, N, Npad, static_cast<const double*>(ctx->stage_dev.p), vs.store,
                        ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "pack kernel");
    return ROME_B200_OK;
}

int rome_b200_set_particles_anchored(rome_b200_ctx* ctx, int vartype, int nvars, int N, const double* anchors_host,
                                     const float* offsets_host) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    if (nvars < 0 || N <= 0 || (nvars > 0 && (!anchors_host || !offsets_host))) return fail(ctx, ROME_B200_BAD_ARG, "bad particle shape");
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "set_particles during graph capture");
    if (int e = bind(ctx)) return e;
    const int d = kVarDim[vartype], Npad = rome_b200_npad(N);
    VarStore& vs = ctx->vars[vartype];
    const size_t store_bytes = (size_t)nvars * var_block_bytes(d, Npad);
    if (store_bytes > vs.cap) {
        if (vs.store) CK(cudaFree(vs.store));
        vs.store = nullptr; vs.cap = 0;
        const size_t kBlock = size_t(2) << 20;
        const size_t rounded = (store_bytes + kBlock - 1) / kBlock * kBlock;
        CK(cudaMalloc(&vs.store, rounded));
        vs.cap = rounded;
    }
    vs.nvars = nvars; vs.N = N; vs.Npad = Npad;
    if (nvars == 0) return ROME_B200_OK;
    const size_t a_bytes = ((size_t)nvars * d * sizeof(double) + 255) / 256 * 256;
    const size_t o_bytes = (size_t)nvars * N * d * sizeof(float);
    if (int e = grow_dev(ctx, ctx->stage_dev, a_bytes + o_bytes)) return e;
    unsigned char* sd = static_cast<unsigned char*>(ctx->stage_dev.p);
    const void *sa = anchors_host, *so = offsets_host;
    if (!is_pinned_or_device(anchors_host) || !is_pinned_or_device(offsets_host)) {  // pageable: stage through pinned memory
        if (int e = grow_host(ctx, ctx->stage_host, a_bytes + o_bytes)) return e;
        CK(cudaStreamSynchronize(ctx->stream));
        unsigned char* sh = static_cast<unsigned char*>(ctx->stage_host.p);
        std::memcpy(sh, anchors_host, (size_t)nvars * d * sizeof(double));
        std::memcpy(sh + a_bytes, offsets_host, o_bytes);
        sa = sh; so = sh + a_bytes;
    }
    CK(cudaMemcpyAsync(sd, sa, (size_t)nvars * d * sizeof(double), cudaMemcpyDefault, ctx->stream));
    CK(cudaMemcpyAsync(sd + a_bytes, so, o_bytes, cudaMemcpyDefault, ctx->stream));
    int e = launch_pack_anchored(d, kWrapDim[vartype], nvars, N, Npad, reinterpret_cast<const double*>(sd),
                                 reinterpret_cast<const float*>(sd + a_bytes), vs.store, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "pack kernel");
    return ROME_B200_OK;
}

int rome_b200_get_particles(rome_b200_ctx* ctx, int vartype, double* coords_host) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES || !coords_host) return fail(ctx, ROME_B200_BAD_ARG, "bad argument");
    VarStore& vs = ctx->vars[vartype];
    if (vs.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of this variable type are not set");
    if (int e = bind(ctx)) return e;
    const int d = kVarDim[vartype];
    const size_t bytes = (size_t)vs.nvars * vs.N * d * sizeof(double);
    if (int e = grow_dev(ctx, ctx->stage_dev, bytes)) return e;
    int e = launch_unpack(d, kWrapDim[vartype], vs.nvars, vs.N, vs.Npad, vs.store,
                          static_cast<double*>(ctx->stage_dev.p), ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "unpack kernel");
    CK(cudaMemcpyAsync(coords_host, ctx->stage_dev.p, bytes, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ROME_B200_OK;
}

int rome_b200_particles_device(rome_b200_ctx* ctx, int vartype, void** d_store, int* block_bytes, int* header_bytes,
                               int* nvars, int* N, int* Npad) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    VarStore& vs = ctx->vars[vartype];
    if (vs.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of this variable type are not set");
    if (d_store) *d_store = vs.store;
    if (block_bytes) *block_bytes = var_block_bytes(kVarDim[vartype], vs.Npad);
    if (header_bytes) *header_bytes = var_header_bytes(kVarDim[vartype]);
    if (nvars) *nvars = vs.nvars;
    if (N) *N = vs.N;
    if (Npad) *Npad = vs.Npad;
    return ROME_B200_OK;
}

int rome_b200_adopt_proposal(rome_b200_ctx* ctx, int vartype, int var, const float* d_prop, int factor) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES || !d_prop || factor < 0)
        return fail(ctx, ROME_B200_BAD_ARG, "bad argument");
    VarStore& vs = ctx->vars[vartype];
    if (var < 0 || var >= vs.nvars) return fail(ctx, ROME_B200_BAD_ARG, "variable index out of range");
    if (int e = bind(ctx)) return e;
    int e = launch_adopt(kVarDim[vartype], vs.Npad, vs.store, var, d_prop, factor, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "adopt kernel");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int rome_b200_set_factors_pose2pose2(rome_b200_ctx* ctx, int nF, const int32_t* ip, const int32_t* iq,
                                     const double* mu, const double* cov) {
    if (ctx && nF > 0 && !iq) return fail(ctx, ROME_B200_BAD_ARG, "iq is NULL");
    return set_gaussian_factors<RowSE2, 3>(ctx, ROME_B200_POSE2POSE2, nF, ip, iq, mu, cov);
}
int rome_b200_set_factors_priorpose2(rome_b200_ctx* ctx, int nF, const int32_t* ip, const double* mu,
                                     const double* cov) {
    return set_gaussian_factors<RowSE2, 3>(ctx, ROME_B200_PRIORPOSE2, nF, ip, nullptr, mu, cov);
}
int rome_b200_set_factors_pose3pose3(rome_b200_ctx* ctx, int nF, const int32_t* ip, const int32_t* iq,
                                     const double* mu, const double* cov) {
    if (ctx && nF > 0 && !iq) return fail(ctx, ROME_B200_BAD_ARG, "iq is NULL");
    return set_gaussian_factors<RowSE3, 6>(ctx, ROME_B200_POSE3POSE3, nF, ip, iq, mu, cov);
}
int rome_b200_set_factors_priorpose3(rome_b200_ctx* ctx, int nF, const int32_t* ip, const double* mu,
                                     const double* cov) {
    return set_gaussian_factors<RowSE3, 6>(ctx, ROME_B200_PRIORPOSE3, nF, ip, nullptr, mu, cov);
}
int rome_b200_set_factors_bearingrange(rome_b200_ctx* ctx, int nF, const int32_t* ip, const int32_t* il,
                                       const double* bearing, const double* range) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (nF < 0 || (nF > 0 && (!ip || !il || !bearing || !range))) return fail(ctx, ROME_B200_BAD_ARG, "null factor arrays");
    std::vector<RowBR> rows((size_t)nF);
    int m0 = -1, m1 = -1;
    for (int f = 0; f < nF; ++f) {
        RowBR& r = rows[f];
        r.ip = ip[f]; r.iq = il[f];
        if (r.ip < 0 || r.iq < 0) return fail(ctx, ROME_B200_BAD_ARG, "negative variable index");
        if (!(bearing[2 * f + 1] > 0.0) || !(range[2 * f + 1] > 0.0))
            return fail(ctx, ROME_B200_BAD_ARG, "standard deviation must be positive");
        if (r.ip > m0) m0 = r.ip;
        if (r.iq > m1) m1 = r.iq;
        r.mu_b = bearing[2 * f]; r.sig_b = (float)bearing[2 * f + 1];
        r.mu_r = range[2 * f]; r.sig_r = (float)range[2 * f + 1];
    }
    return upload_rows(ctx, ROME_B200_BEARINGRANGE, rows.data(), nF, m0, m1);
}
int rome_b200_set_factors_point2(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0, const int32_t* i1,
                                 const double* mu, const double* cov) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family != ROME_B200_PRIORPOINT2 && family != ROME_B200_POINT2POINT2 && family != ROME_B200_POSE2POINT2)
        return fail(ctx, ROME_B200_BAD_ARG, "family is not a 2-D Gaussian point factor");
    const bool binary = family != ROME_B200_PRIORPOINT2;
    if (nF < 0 || (nF > 0 && (!i0 || !mu || !cov || (binary && !i1)))) return fail(ctx, ROME_B200_BAD_ARG, "null factor arrays");
    std::vector<RowPT2> rows((size_t)nF);
    int m0 = -1, m1 = -1;
    double L[4];
    for (int f = 0; f < nF; ++f) {
        RowPT2& r = rows[f];
        std::memset(&r, 0, sizeof r);
        r.ip = i0[f]; r.iq = binary ? i1[f] : -1;
        if (r.ip < 0 || (binary && r.iq < 0)) return fail(ctx, ROME_B200_BAD_ARG, "negative variable index");
        if (r.ip > m0) m0 = r.ip;
        if (r.iq > m1) m1 = r.iq;
        r.mu[0] = mu[2 * f]; r.mu[1] = mu[2 * f + 1];
        if (!cholesky(cov + (size_t)f * 4, 2, L)) return fail(ctx, ROME_B200_BAD_ARG, "covariance is not positive definite");
        r.L[0] = (float)L[0]; r.L[1] = (float)L[2]; r.L[2] = (float)L[3];
    }
    return upload_rows(ctx, family, rows.data(), nF, m0, m1);
}
int rome_b200_set_factors_scalar(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0, const int32_t* i1,
                                 const double* belief) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family != ROME_B200_POSE2POINT2RANGE && family != ROME_B200_POINT2POINT2RANGE &&
        family != ROME_B200_POSE2POINT2BEARING)
        return fail(ctx, ROME_B200_BAD_ARG, "family is not a scalar factor");
    if (nF < 0 || (nF > 0 && (!i0 || !i1 || !belief))) return fail(ctx, ROME_B200_BAD_ARG, "null factor arrays");
    std::vector<RowS1> rows((size_t)nF);
    int m0 = -1, m1 = -1;
    for (int f = 0; f < nF; ++f) {
        RowS1& r = rows[f];
        std::memset(&r, 0, sizeof r);
        r.ip = i0[f]; r.iq = i1[f];
        if (r.ip < 0 || r.iq < 0) return fail(ctx, ROME_B200_BAD_ARG, "negative variable index");
        if (!(belief[2 * f + 1] > 0.0)) return fail(ctx, ROME_B200_BAD_ARG, "standard deviation must be positive");
        if (r.ip > m0) m0 = r.ip;
        if (r.iq > m1) m1 = r.iq;
        r.mu = belief[2 * f]; r.sigma = (float)belief[2 * f + 1];
    }
    return upload_rows(ctx, family, rows.data(), nF, m0, m1);
}
int rome_b200_set_factors_gaussian(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0, const int32_t* i1,
                                   const double* mu, const double* cov) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family < 0 || family >= ROME_B200_NFAMILIES) return fail(ctx, ROME_B200_BAD_ARG, "bad family");
    const FamInfo& fi = kFam[family];
    const bool binary = fi.vt1 >= 0;
    if (fam_vt2(family) >= 0) return fail(ctx, ROME_B200_BAD_ARG, "family has a third variable: use rome_b200_set_factors_ternary");
    if (binary && nF > 0 && !i1) return fail(ctx, ROME_B200_BAD_ARG, "second variable index array is NULL");
    if (fi.row_bytes == (int)sizeof(RowSE3))
        return set_gaussian_factors<RowSE3, 6>(ctx, family, nF, i0, binary ? i1 : nullptr, mu, cov);
    if (fi.row_bytes == (int)sizeof(RowSE2) && fi.dm == 3)
        return set_gaussian_factors<RowSE2, 3>(ctx, family, nF, i0, binary ? i1 : nullptr, mu, cov);
    if (fi.row_bytes == (int)sizeof(RowPT2)) return rome_b200_set_factors_point2(ctx, family, nF, i0, i1, mu, cov);
    return fail(ctx, ROME_B200_BAD_ARG, "family does not hold a single MvNormal belief");
}
int rome_b200_set_factors_ternary(rome_b200_ctx* ctx, int family, int nF, const int32_t* i0, const int32_t* i1,
                                  const int32_t* i2, const double* mu, const double* cov) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family < 0 || family >= ROME_B200_NFAMILIES || fam_vt2(family) < 0)
        return fail(ctx, ROME_B200_BAD_ARG, "family has no third variable");
    if (nF > 0 && (!i1 || !i2)) return fail(ctx, ROME_B200_BAD_ARG, "second / third variable index array is NULL");
    return set_gaussian_factors<RowSE3, 6>(ctx, family, nF, i0, i1, mu, cov, i2);
}
int rome_b200_num_factors(rome_b200_ctx* ctx, int family) {
    if (!ctx || family < 0 || family >= ROME_B200_NFAMILIES) return ROME_B200_BAD_ARG;
    return ctx->fac[family].nF;
}

// ---------------------------------------------------------------------------------------------
static int check_eval(rome_b200_ctx* ctx, int family, uint32_t flags, int first, int& count,
                      const rome_b200_buffers* b) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family < 0 || family >= ROME_B200_NFAMILIES) return fail(ctx, ROME_B200_BAD_ARG, "bad family");
    if (!b) return fail(ctx, ROME_B200_BAD_ARG, "buffers is NULL");
    const FamInfo& fi = kFam[family];
    FactorStore& fs = ctx->fac[family];
    if (fs.rows == nullptr) return fail(ctx, ROME_B200_NOT_SET, "factors of this family are not set");
    if (count < 0) count = fs.nF - first;
    if (first < 0 || count < 0 || first + count > fs.nF) return fail(ctx, ROME_B200_BAD_ARG, "factor range out of bounds");
    const VarStore& v0 = ctx->vars[fi.vt0];
    if (v0.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of the first variable type are not set");
    if (fs.max_i0 >= v0.nvars) return fail(ctx, ROME_B200_SHAPE_MISMATCH, "factor refers to a variable index beyond the particle store");
    if (fi.vt1 >= 0) {
        const VarStore& v1 = ctx->vars[fi.vt1];
        if (v1.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of the second variable type are not set");
        if (fs.max_i1 >= v1.nvars) return fail(ctx, ROME_B200_SHAPE_MISMATCH, "factor refers to a variable index beyond the particle store");
        if (v1.N != v0.N) return fail(ctx, ROME_B200_SHAPE_MISMATCH, "particle counts differ between variable types");
    }
    if (const int vt2 = fam_vt2(family); vt2 >= 0) {
        const VarStore& v2 = ctx->vars[vt2];
        if (v2.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of the third variable type are not set");
        if (fs.max_i2 < 0) return fail(ctx, ROME_B200_NOT_SET, "factors of this family must be set with rome_b200_set_factors_ternary");
        if (fs.max_i2 >= v2.nvars) return fail(ctx, ROME_B200_SHAPE_MISMATCH, "factor refers to a variable index beyond the particle store");
        if (v2.N != v0.N) return fail(ctx, ROME_B200_SHAPE_MISMATCH, "particle counts differ between variable types");
    }
    if (!(flags & ROME_B200_SAMPLE) && !b->meas) return fail(ctx, ROME_B200_BAD_ARG, "meas is NULL and SAMPLE is not set");
    if ((flags & ROME_B200_WRITE_MEAS) && (!(flags & ROME_B200_SAMPLE) || !b->meas_out))
        return fail(ctx, ROME_B200_BAD_ARG, "WRITE_MEAS needs SAMPLE and meas_out");
    if (flags & ROME_B200_DECONV) {
        if (family > ROME_B200_PRIORPOSE3) return fail(ctx, ROME_B200_BAD_ARG, "this family has no closed-form deconvolution");
        if ((flags & ROME_B200_WRITE_MEAS) || !b->meas_out)
            return fail(ctx, ROME_B200_BAD_ARG, "DECONV needs meas_out and excludes WRITE_MEAS");
    }
    if ((flags & ROME_B200_RESIDUAL) && !b->res) return fail(ctx, ROME_B200_BAD_ARG, "res is NULL");
    if ((flags & ROME_B200_PROPOSAL_FWD) && fi.dfwd == 0)
        return fail(ctx, ROME_B200_BAD_ARG, "this family has no closed-form forward proposal");
    if ((flags & ROME_B200_PROPOSAL_FWD) && !b->prop_fwd) return fail(ctx, ROME_B200_BAD_ARG, "prop_fwd is NULL");
    // meas / res / prop_fwd rows move through 1-D TMA bulk copies: 16-byte alignment is required
    if (((flags & ROME_B200_RESIDUAL) && ((uintptr_t)b->res & 15)) ||
        ((flags & ROME_B200_PROPOSAL_FWD) && ((uintptr_t)b->prop_fwd & 15)) ||
        (!(flags & ROME_B200_SAMPLE) && ((uintptr_t)b->meas & 15)))
        return fail(ctx, ROME_B200_BAD_ARG, "meas, res and prop_fwd must be 16-byte aligned");
    if (flags & ROME_B200_PROPOSAL_BWD) {
        if (fi.dbwd == 0) return fail(ctx, ROME_B200_BAD_ARG, "this family has no closed-form backward proposal");
        if (!b->prop_bwd) return fail(ctx, ROME_B200_BAD_ARG, "prop_bwd is NULL");
    }
    if ((flags & ROME_B200_STATS) && !b->stats) return fail(ctx, ROME_B200_BAD_ARG, "stats is NULL");
    if (flags & ROME_B200_JACOBIAN) {
        if (fi.dj == 0) return fail(ctx, ROME_B200_BAD_ARG, "this family has no Jacobian output");
        if (!b->jac) return fail(ctx, ROME_B200_BAD_ARG, "jac is NULL");
    }
    return ROME_B200_OK;
}

int rome_b200_eval(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed, uint32_t stream_id, int first,
                   int count, const rome_b200_buffers* b) {
    if (int e = check_eval(ctx, family, flags, first, count, b)) return e;
    if (count == 0) return ROME_B200_OK;
    if (int e = bind(ctx)) return e;
    const FamInfo& fi = kFam[family];
    const VarStore& v0 = ctx->vars[fi.vt0];
    const VarStore& v1 = ctx->vars[fi.vt1 >= 0 ? fi.vt1 : fi.vt0];
    EvalParams p;
    p.rows = ctx->fac[family].rows;
    p.first = first; p.count = count;
    p.N = v0.N; p.Npad = v0.Npad;
    p.v0 = v0.store; p.v1 = v1.store;
    p.v2 = fam_vt2(family) >= 0 ? ctx->vars[fam_vt2(family)].store : nullptr;
    p.meas = b->meas; p.meas_out = b->meas_out; p.res = b->res; p.prop_fwd = b->prop_fwd; p.prop_bwd = b->prop_bwd;
    p.stats = b->stats; p.jac = b->jac;
    p.n_peers = (flags & ROME_B200_PROPOSAL_FWD) ? ctx->n_peers[family] : 0;
    for (int r = 0; r < 7; ++r) p.peer_fwd[r] = ctx->peers[family][r];
    p.fwd_dst = p.bwd_dst = nullptr;
    p.fwd_dst_lo = p.fwd_dst_hi = 0;
    for (int dir = 0; dir < 2; ++dir) {
        const int n = ctx->row_dst_n[family][dir];
        if (!n) continue;
        if (n != ctx->fac[family].nF)
            return fail(ctx, ROME_B200_SHAPE_MISMATCH, "proposal destinations were set for a different number of factors");
        (dir ? p.bwd_dst : p.fwd_dst) = static_cast<const unsigned long long*>(ctx->row_dst[family][dir].p);
        if (dir == 0) { p.fwd_dst_lo = ctx->row_dst_lo[family][0]; p.fwd_dst_hi = ctx->row_dst_hi[family][0]; }
    }
    p.bar_state = ctx->bar_state; p.bar_n = ctx->bar_n; p.bar_timeout = 0; p.bar_lo = ctx->bar_lo[family]; p.bar_hi = ctx->bar_hi[family];
    for (int r = 0; r < 7; ++r) p.bar_peer[r] = ctx->bar_peer[r];
    if (flags & (ROME_B200_BARRIER_WAIT | ROME_B200_BARRIER_SIGNAL)) {
        if (!ctx->bar_state || ctx->bar_n == 0)
            return fail(ctx, ROME_B200_NOT_SET, "BARRIER_WAIT / BARRIER_SIGNAL need rome_b200_set_step_barrier");
        int clock_khz = 0;
        cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, ctx->device);
        p.bar_timeout = 2LL * 1000LL * (clock_khz > 0 ? clock_khz : 1965000);  // ~2 s
    }
    p.flags = flags;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.stream_id = stream_id;
    LaunchPlan plan;
    if (plan_launch(family, flags, v0.Npad, ctx->smem_per_sm, ctx->smem_per_cta_max, &plan))
        return fail(ctx, ROME_B200_SHAPE_MISMATCH, "N is too large for the shared-memory pipeline of this family");
    p.stages = plan.stages; p.stage_bytes = plan.stage_bytes; p.out_warp_bytes = plan.out_warp_bytes;
    const int nTiles = (count + plan.ft - 1) / plan.ft;
    // (leaving one SM free for the small INDEPENDENT kernels launched next to a full persistent grid was measured: the
    // dependent launches that follow get 1.5-2 us slower, profiles/r02_analysis.md)
    const int resident = ctx->num_sms * plan.ctas_per_sm;
    const int grid = nTiles < resident ? nTiles : resident;
    int e = launch_eval(family, p, plan, grid, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "eval kernel launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}

static int eval_host_impl(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed, uint32_t stream_id, int first,
                          int count, const rome_b200_buffers* hb, bool sync) {
    if (int e = check_eval(ctx, family, flags, first, count, hb)) return e;
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "eval_host during graph capture");
    if (count == 0) return ROME_B200_OK;
    if (int e = bind(ctx)) return e;
    const FamInfo& fi = kFam[family];
    const int Npad = ctx->vars[fi.vt0].Npad;
    // per-factor strides (floats) of each buffer; device mirrors cover factors [first, first+count)
    const size_t sm = (size_t)fi.dm * Npad, sr = (size_t)fi.dr * Npad, sf = (size_t)fi.dfwd * Npad,
                 sb = (size_t)fi.dbwd * Npad, ss = (size_t)fi.nstats, sj = (size_t)fi.dj * Npad;
    rome_b200_buffers db;
    std::memset(&db, 0, sizeof db);
    auto mirror = [&](int slot, size_t stride, float** out) -> int {
        if (int e = grow_dev(ctx, ctx->out_dev[family][slot], (size_t)count * stride * sizeof(float))) return e;
        // kernels index buffers by global factor id: bias the mirror base by -first
        *out = static_cast<float*>(ctx->out_dev[family][slot].p) - (size_t)first * stride;
        return 0;
    };
    float* tmp = nullptr;
    if (!(flags & ROME_B200_SAMPLE)) {
        if (int e = mirror(0, sm, &tmp)) return e;
        CK(cudaMemcpyAsync(ctx->out_dev[family][0].p, hb->meas + (size_t)first * sm, (size_t)count * sm * sizeof(float),
                           cudaMemcpyDefault, ctx->stream));
        db.meas = tmp;
    }
    if (flags & (ROME_B200_WRITE_MEAS | ROME_B200_DECONV)) { if (int e = mirror(1, sm, &db.meas_out)) return e; }
    if (flags & ROME_B200_RESIDUAL) { if (int e = mirror(2, sr, &db.res)) return e; }
    if (flags & ROME_B200_PROPOSAL_FWD) { if (int e = mirror(3, sf, &db.prop_fwd)) return e; }
    if (flags & ROME_B200_PROPOSAL_BWD) { if (int e = mirror(4, sb, &db.prop_bwd)) return e; }
    if (flags & ROME_B200_STATS) { if (int e = mirror(5, ss, &db.stats)) return e; }
    if (flags & ROME_B200_JACOBIAN) { if (int e = mirror(6, sj, &db.jac)) return e; }
    if (int e = rome_b200_eval(ctx, family, flags, seed, stream_id, first, count, &db)) return e;
    auto back = [&](int slot, size_t stride, float* host) -> int {
        CK(cudaMemcpyAsync(host + (size_t)first * stride, ctx->out_dev[family][slot].p, (size_t)count * stride * sizeof(float),
                           cudaMemcpyDefault, ctx->stream));
        return 0;
    };
    if (flags & (ROME_B200_WRITE_MEAS | ROME_B200_DECONV)) { if (int e = back(1, sm, hb->meas_out)) return e; }
    if (flags & ROME_B200_RESIDUAL) { if (int e = back(2, sr, hb->res)) return e; }
    if (flags & ROME_B200_PROPOSAL_FWD) { if (int e = back(3, sf, hb->prop_fwd)) return e; }
    if (flags & ROME_B200_PROPOSAL_BWD) { if (int e = back(4, sb, hb->prop_bwd)) return e; }
    if (flags & ROME_B200_STATS) { if (int e = back(5, ss, hb->stats)) return e; }
    if (flags & ROME_B200_JACOBIAN) { if (int e = back(6, sj, hb->jac)) return e; }
    if (sync) CK(cudaStreamSynchronize(ctx->stream));
    return ROME_B200_OK;
}

int rome_b200_eval_host(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed, uint32_t stream_id, int first,
                        int count, const rome_b200_buffers* hb) {
    return eval_host_impl(ctx, family, flags, seed, stream_id, first, count, hb, true);
}
int rome_b200_eval_host_async(rome_b200_ctx* ctx, int family, uint32_t flags, uint64_t seed, uint32_t stream_id,
                              int first, int count, const rome_b200_buffers* hb) {
    return eval_host_impl(ctx, family, flags, seed, stream_id, first, count, hb, false);
}

// ---------------------------------------------------------------------------------------------
int rome_b200_set_product_plan(rome_b200_ctx* ctx, int vartype, int nvars, const int32_t* var_offsets,
                               const int32_t* src_buf, const int32_t* src_row) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    if (nvars < 0 || (nvars > 0 && !var_offsets)) return fail(ctx, ROME_B200_BAD_ARG, "bad plan arrays");
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "set_product_plan during graph capture");
    const int nsrc = nvars ? var_offsets[nvars] : 0;
    if (nvars && (var_offsets[0] != 0 || nsrc < 0 || (nsrc > 0 && (!src_buf || !src_row))))
        return fail(ctx, ROME_B200_BAD_ARG, "bad plan arrays");
    int max_buf = -1;
    for (int v = 0; v < nvars; ++v) {
        const int k = var_offsets[v + 1] - var_offsets[v];
        if (k < 0) return fail(ctx, ROME_B200_BAD_ARG, "var_offsets must be non-decreasing");
        if (k > ROME_B200_MAX_PRODUCT_SOURCES) return fail(ctx, ROME_B200_BAD_ARG, "too many proposals for one variable");
    }
    for (int i = 0; i < nsrc; ++i) {
        if (src_buf[i] < 0 || src_buf[i] >= ROME_B200_MAX_PRODUCT_BUFFERS || src_row[i] < 0)
            return fail(ctx, ROME_B200_BAD_ARG, "source buffer index / row out of range");
        if (src_buf[i] > max_buf) max_buf = src_buf[i];
    }
    if (int e = bind(ctx)) return e;
    ProductPlan& pl = ctx->plan[vartype];
    if (int e = grow_dev(ctx, pl.off, (size_t)(nvars + 1) * 4)) return e;
    if (int e = grow_dev(ctx, pl.buf, (size_t)(nsrc ? nsrc : 1) * 4)) return e;
    if (int e = grow_dev(ctx, pl.row, (size_t)(nsrc ? nsrc : 1) * 4)) return e;
    if (nvars) CK(cudaMemcpyAsync(pl.off.p, var_offsets, (size_t)(nvars + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (nsrc) {
        CK(cudaMemcpyAsync(pl.buf.p, src_buf, (size_t)nsrc * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(pl.row.p, src_row, (size_t)nsrc * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));  // the caller's arrays may be temporaries
    pl.nvars = nvars; pl.nsrc = nsrc; pl.max_buf = max_buf;
    return ROME_B200_OK;
}

int rome_b200_product(rome_b200_ctx* ctx, int vartype, int n_bufs, const float* const* d_prop_bufs, uint64_t seed,
                      uint32_t stream_id, int gibbs_iters, uint32_t flags, float* d_bw_out) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    const ProductPlan& pl = ctx->plan[vartype];
    VarStore& vs = ctx->vars[vartype];
    if (vs.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of this variable type are not set");
    const int nown = (ctx->owned[vartype] >= 0 && ctx->owned[vartype] < vs.nvars) ? ctx->owned[vartype] : vs.nvars;
    if (pl.nvars != nown) return fail(ctx, ROME_B200_NOT_SET, "no product plan for this variable type (or its size differs)");
    if (n_bufs < 0 || n_bufs > ROME_B200_MAX_PRODUCT_BUFFERS || pl.max_buf >= n_bufs || (n_bufs > 0 && !d_prop_bufs))
        return fail(ctx, ROME_B200_BAD_ARG, "the plan refers to more proposal buffers than were passed");
    for (int b = 0; b <= pl.max_buf; ++b)
        if (!d_prop_bufs[b]) return fail(ctx, ROME_B200_BAD_ARG, "proposal buffer is NULL");
    if (int e = bind(ctx)) return e;
    const int d = kVarDim[vartype];
    ProductParams p;
    std::memset(&p, 0, sizeof p);
    p.store = vs.store;
    p.var_off = static_cast<const int32_t*>(pl.off.p);
    p.src_buf = static_cast<const int32_t*>(pl.buf.p);
    p.src_row = static_cast<const int32_t*>(pl.row.p);
    for (int b = 0; b < n_bufs; ++b) p.bufs[b] = d_prop_bufs[b];
    p.bw_out = d_bw_out;
    p.nvars = nown; p.N = vs.N; p.Npad = vs.Npad;
    p.iters = gibbs_iters > 0 ? gibbs_iters : 2;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32); p.stream_id = stream_id;
    p.bw_scale = (float)std::pow(4.0 / ((d + 2.0) * vs.N), 1.0 / (d + 4.0));
    p.manifold = ((flags & ROME_B200_PRODUCT_MANIFOLD) && vartype == ROME_B200_POSE3) ? 1 : 0;
    int e = launch_product(d, kWrapDim[vartype] < 0 ? -1 : kWrapDim[vartype], &p, ctx->num_sms, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "product kernel launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    if (flags & ROME_B200_PRODUCT_REANCHOR) return rome_b200_reanchor(ctx, vartype);
    return ROME_B200_OK;
}

int rome_b200_reanchor(rome_b200_ctx* ctx, int vartype) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    VarStore& vs = ctx->vars[vartype];
    if (vs.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of this variable type are not set");
    if (int e = bind(ctx)) return e;
    const int nown = (ctx->owned[vartype] >= 0 && ctx->owned[vartype] < vs.nvars) ? ctx->owned[vartype] : vs.nvars;
    int e = launch_reanchor(kVarDim[vartype], kWrapDim[vartype] < 0 ? -1 : kWrapDim[vartype], vs.store, nown, vs.N, vs.Npad, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "reanchor kernel launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int rome_b200_set_peer_proposals(rome_b200_ctx* ctx, int family, int n_peers, float* const* peer_prop_fwd) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family < 0 || family >= ROME_B200_NFAMILIES) return fail(ctx, ROME_B200_BAD_ARG, "bad family");
    if (n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_prop_fwd)) return fail(ctx, ROME_B200_BAD_ARG, "bad peer list");
    for (int r = 0; r < n_peers; ++r)
        if (!peer_prop_fwd[r] || ((uintptr_t)peer_prop_fwd[r] & 15))
            return fail(ctx, ROME_B200_BAD_ARG, "peer buffers must be non-NULL and 16-byte aligned");
    ctx->n_peers[family] = n_peers;
    for (int r = 0; r < 7; ++r) ctx->peers[family][r] = r < n_peers ? peer_prop_fwd[r] : nullptr;
    return ROME_B200_OK;
}
int rome_b200_set_proposal_destinations(rome_b200_ctx* ctx, int family, int direction, int nF, void* const* rows) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family < 0 || family >= ROME_B200_NFAMILIES || (direction != 0 && direction != 1))
        return fail(ctx, ROME_B200_BAD_ARG, "bad family / direction");
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "set_proposal_destinations during graph capture");
    if (nF == 0) { ctx->row_dst_n[family][direction] = 0; return ROME_B200_OK; }
    if (nF < 0 || !rows) return fail(ctx, ROME_B200_BAD_ARG, "bad destination array");
    if (nF != ctx->fac[family].nF) return fail(ctx, ROME_B200_SHAPE_MISMATCH, "one destination per factor of the family is required");
    const FamInfo& fi = kFam[family];
    if ((direction == 0 ? fi.dfwd : fi.dbwd) == 0) return fail(ctx, ROME_B200_BAD_ARG, "this family has no proposal in that direction");
    for (int f = 0; f < nF; ++f)
        if ((uintptr_t)rows[f] & 15) return fail(ctx, ROME_B200_BAD_ARG, "destination rows must be 16-byte aligned");
    if (int e = bind(ctx)) return e;
    Scratch& sc = ctx->row_dst[family][direction];
    if (int e = grow_dev(ctx, sc, (size_t)nF * 8)) return e;
    static_assert(sizeof(void*) == 8, "64-bit pointers");
    CK(cudaMemcpyAsync(sc.p, rows, (size_t)nF * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // the caller's array may be a temporary
    ctx->row_dst_n[family][direction] = nF;
    int lo = nF, hi = 0;
    for (int f = 0; f < nF; ++f)
        if (rows[f]) { if (f < lo) lo = f; hi = f + 1; }
    ctx->row_dst_lo[family][direction] = lo < hi ? lo : 0;
    ctx->row_dst_hi[family][direction] = lo < hi ? hi : 0;
    return ROME_B200_OK;
}
int rome_b200_set_step_barrier(rome_b200_ctx* ctx, void* d_state, uint32_t* const* peer_slots, int n_peers) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (n_peers == 0) { ctx->bar_state = nullptr; ctx->bar_n = 0; return ROME_B200_OK; }
    if (!d_state || n_peers < 0 || n_peers > 7 || !peer_slots) return fail(ctx, ROME_B200_BAD_ARG, "bad peer list");
    for (int r = 0; r < n_peers; ++r)
        if (!peer_slots[r]) return fail(ctx, ROME_B200_BAD_ARG, "peer slot is NULL");
    ctx->bar_state = static_cast<uint32_t*>(d_state);
    ctx->bar_n = n_peers;
    for (int r = 0; r < 7; ++r) ctx->bar_peer[r] = r < n_peers ? peer_slots[r] : nullptr;
    return ROME_B200_OK;
}
int rome_b200_set_barrier_range(rome_b200_ctx* ctx, int family, int first, int count) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (family < 0 || family >= ROME_B200_NFAMILIES || first < 0) return fail(ctx, ROME_B200_BAD_ARG, "bad family / range");
    ctx->bar_lo[family] = count < 0 ? 0 : first;
    ctx->bar_hi[family] = count < 0 ? 0x7fffffff : first + count;
    return ROME_B200_OK;
}
int rome_b200_set_owned_variables(rome_b200_ctx* ctx, int vartype, int n_owned) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    ctx->owned[vartype] = n_owned < 0 ? -1 : n_owned;
    return ROME_B200_OK;
}
int rome_b200_set_halo_plan(rome_b200_ctx* ctx, int vartype, int n, const int32_t* src_var, void* const* dst_blocks) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "set_halo_plan during graph capture");
    if (n == 0) { ctx->halo_n[vartype] = 0; return ROME_B200_OK; }
    if (n < 0 || !src_var || !dst_blocks) return fail(ctx, ROME_B200_BAD_ARG, "bad halo arrays");
    const VarStore& vs = ctx->vars[vartype];
    for (int i = 0; i < n; ++i) {
        if (src_var[i] < 0 || src_var[i] >= vs.nvars) return fail(ctx, ROME_B200_BAD_ARG, "halo source variable out of range");
        if (!dst_blocks[i] || ((uintptr_t)dst_blocks[i] & 15)) return fail(ctx, ROME_B200_BAD_ARG, "halo destinations must be non-NULL and 16-byte aligned");
    }
    if (int e = bind(ctx)) return e;
    if (int e = grow_dev(ctx, ctx->halo_src[vartype], (size_t)n * 4)) return e;
    if (int e = grow_dev(ctx, ctx->halo_dst[vartype], (size_t)n * 8)) return e;
    CK(cudaMemcpyAsync(ctx->halo_src[vartype].p, src_var, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->halo_dst[vartype].p, dst_blocks, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->halo_n[vartype] = n;
    return ROME_B200_OK;
}
int rome_b200_push_halo(rome_b200_ctx* ctx, int vartype) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (vartype < 0 || vartype >= ROME_B200_NVARTYPES) return fail(ctx, ROME_B200_BAD_ARG, "bad vartype");
    const int n = ctx->halo_n[vartype];
    if (n == 0) return ROME_B200_OK;
    const VarStore& vs = ctx->vars[vartype];
    if (vs.nvars == 0) return fail(ctx, ROME_B200_NOT_SET, "particles of this variable type are not set");
    if (int e = bind(ctx)) return e;
    int e = launch_halo_push(vs.store, var_block_bytes(kVarDim[vartype], vs.Npad), n,
                             static_cast<const int32_t*>(ctx->halo_src[vartype].p),
                             static_cast<const unsigned long long*>(ctx->halo_dst[vartype].p), ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "halo push launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}
int rome_b200_ipc_export(rome_b200_ctx* ctx, void* dev_ptr, unsigned char handle[64]) {
    if (!ctx || !dev_ptr || !handle) return ROME_B200_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (int e = bind(ctx)) return e;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, dev_ptr));
    std::memcpy(handle, &h, 64);
    return ROME_B200_OK;
}
int rome_b200_ipc_import(rome_b200_ctx* ctx, const unsigned char handle[64], void** dev_ptr) {
    if (!ctx || !dev_ptr || !handle) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ROME_B200_OK;
}
int rome_b200_ipc_close(rome_b200_ctx* ctx, void* dev_ptr) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaIpcCloseMemHandle(dev_ptr));
    return ROME_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int rome_b200_peer_signal(rome_b200_ctx* ctx, void* d_state, uint32_t* const* peer_slots, int n_peers) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (!d_state || n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_slots)) return fail(ctx, ROME_B200_BAD_ARG, "bad peer list");
    for (int r = 0; r < n_peers; ++r)
        if (!peer_slots[r]) return fail(ctx, ROME_B200_BAD_ARG, "peer slot is NULL");
    if (int e = bind(ctx)) return e;
    uint32_t* st = static_cast<uint32_t*>(d_state);
    int e = launch_peer_signal(peer_slots, n_peers, st + 8, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "peer signal launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}
int rome_b200_peer_wait(rome_b200_ctx* ctx, void* d_state, const int32_t* slots, int n_slots) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (!d_state || n_slots < 0 || n_slots > 8 || (n_slots > 0 && !slots)) return fail(ctx, ROME_B200_BAD_ARG, "bad slot list");
    // the listed slots must be a prefix-free set of [0, 8); the kernel polls slots 0..n-1 of a compacted view, so the
    // caller's slot numbering is required to be 0..n_slots-1 (ranks number their peers densely)
    for (int i = 0; i < n_slots; ++i)
        if (slots[i] != i) return fail(ctx, ROME_B200_BAD_ARG, "slots must be numbered 0..n_slots-1");
    if (int e = bind(ctx)) return e;
    uint32_t* st = static_cast<uint32_t*>(d_state);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, ctx->device);
    const long long max_cycles = 2LL * 1000LL * (clock_khz > 0 ? clock_khz : 1965000);  // ~2 s
    int e = launch_peer_wait(st, n_slots, st + 9, st + 10, max_cycles, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "peer wait launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}
int rome_b200_peer_barrier(rome_b200_ctx* ctx, void* d_state, uint32_t* const* peer_slots, int n_peers) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (!d_state || n_peers < 0 || n_peers > 7 || (n_peers > 0 && !peer_slots)) return fail(ctx, ROME_B200_BAD_ARG, "bad peer list");
    for (int r = 0; r < n_peers; ++r)
        if (!peer_slots[r]) return fail(ctx, ROME_B200_BAD_ARG, "peer slot is NULL");
    if (int e = bind(ctx)) return e;
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, ctx->device);
    const long long max_cycles = 2LL * 1000LL * (clock_khz > 0 ? clock_khz : 1965000);  // ~2 s
    int e = launch_peer_barrier(peer_slots, n_peers, static_cast<uint32_t*>(d_state), max_cycles, ctx->stream);
    if (e) return cuda_fail(ctx, (cudaError_t)e, "peer barrier launch");
    if (ctx->capturing) ctx->capture_kernels++; else ctx->launches++;
    return ROME_B200_OK;
}
int rome_b200_peer_status(rome_b200_ctx* ctx, void* d_state, int* gave_up) {
    if (!ctx || !d_state || !gave_up) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    uint32_t v = 0;
    CK(cudaMemcpyAsync(&v, static_cast<uint32_t*>(d_state) + 10, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *gave_up = (int)v;
    return ROME_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int rome_b200_graph_begin(rome_b200_ctx* ctx) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "graph capture already active");
    if (int e = bind(ctx)) return e;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    ctx->capture_kernels = 0;
    return ROME_B200_OK;
}
int rome_b200_graph_end(rome_b200_ctx* ctx, int* graph_id) {
    if (!ctx || !graph_id) return ROME_B200_BAD_ARG;
    if (!ctx->capturing) return fail(ctx, ROME_B200_BAD_ARG, "no graph capture active");
    ctx->capturing = false;
    cudaGraph_t g = nullptr;
    CK(cudaStreamEndCapture(ctx->stream, &g));
    cudaGraphExec_t ge = nullptr;
    cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaGraphInstantiate");
    ctx->graphs.push_back(ge);
    ctx->graph_kernels.push_back(ctx->capture_kernels);
    *graph_id = (int)ctx->graphs.size() - 1;
    return ROME_B200_OK;
}
int rome_b200_graph_launch(rome_b200_ctx* ctx, int graph_id) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (graph_id < 0 || graph_id >= (int)ctx->graphs.size()) return fail(ctx, ROME_B200_BAD_ARG, "bad graph id");
    if (int e = bind(ctx)) return e;
    CK(cudaGraphLaunch(ctx->graphs[graph_id], ctx->stream));
    ctx->launches += ctx->graph_kernels[graph_id];
    return ROME_B200_OK;
}

// ---------------------------------------------------------------------------------------------
int rome_b200_malloc_device(rome_b200_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    // Whole 2 MiB blocks: the driver packs smaller allocations into shared 2 MiB slabs, and a CUDA IPC handle always
    // names the slab -- a peer that opens it gets the slab's base, not this buffer.  Buffers from this call are the ones
    // callers export (rome_b200_ipc_export), so each owns its block(s) and base == pointer.
    const size_t kBlock = size_t(2) << 20;
    const size_t rounded = ((bytes ? bytes : 1) + kBlock - 1) / kBlock * kBlock;
    CK(cudaMalloc(out, rounded));
    return ROME_B200_OK;
}
int rome_b200_free_device(rome_b200_ctx* ctx, void* p) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaFree(p));
    return ROME_B200_OK;
}
int rome_b200_malloc_host(rome_b200_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaMallocHost(out, bytes ? bytes : 1));
    return ROME_B200_OK;
}
int rome_b200_free_host(rome_b200_ctx* ctx, void* p) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaFreeHost(p));
    return ROME_B200_OK;
}
int rome_b200_memcpy_h2d(rome_b200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ROME_B200_OK;
}
int rome_b200_memcpy_d2h(rome_b200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (!ctx) return ROME_B200_BAD_ARG;
    if (int e = bind(ctx)) return e;
    CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ROME_B200_OK;
}

uint64_t rome_b200_launch_count(const rome_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
