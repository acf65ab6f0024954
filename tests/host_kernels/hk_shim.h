// hk_shim.h -- TEST INFRASTRUCTURE (CPU suite only; never part of the product path).
// Lets the SOURCE TEXT of the factor-family kernels (csrc/fam_*.cu, se3_common.cuh, the arithmetic parts of
// device_utils.cuh / eval_pipeline.cuh, the pack/unpack kernels) compile as host C++: CUDA vocabulary as no-ops, the few
// intrinsics the families use, and a 32-lane warp emulator (one fiber per lane, switched at every warp collective) so
// that __any_sync / __all_sync / __shfl_xor_sync behave as on the device.  What is NOT emulated: the TMA/mbarrier
// pipeline (the harness hands every factor its inputs directly) and the MUFU approximations (the accurate sampler
// branch is compiled; sqrt_seeded's seed is the exact reciprocal square root).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static const
#define __restrict__
#define ROME_B200_ACCURATE_SAMPLER 1

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
typedef void* cudaStream_t;

// streaming stores are plain stores here
static inline void __stcs(float* p, float v) { *p = v; }
static inline void __stcs(float2* p, float2 v) { *p = v; }
static inline void __stcs(float4* p, float4 v) { *p = v; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline void sincospif(float x, float* s, float* c) {
    *s = (float)std::sin(3.14159265358979323846 * (double)x);
    *c = (float)std::cos(3.14159265358979323846 * (double)x);
}
static inline int __double2hiint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline int __double2loint(double v) { int64_t b; std::memcpy(&b, &v, 8); return (int)(b & 0xffffffff); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double v; std::memcpy(&v, &b, 8); return v;
}
// the one PTX statement in the compiled regions: sqrt_seeded's `rsqrt.approx.ftz.f32` seed
#define asm(...) (y0 = 1.0f / std::sqrt((float)a))

// kernels launched as grids (pack / unpack): the harness sets these before calling the kernel body per lane
static uint3 threadIdx, blockIdx, blockDim;

// ---- warp emulator: 32 fibers, round-robin, switched at collectives ------------------------------------------------
namespace hk {
constexpr int kLanes = 32;
constexpr size_t kStack = 256 * 1024;
static ucontext_t g_sched, g_lane[kLanes];
static char* g_stack[kLanes];
static bool g_done[kLanes];
static int g_cur = 0;
static uint32_t g_x[2][kLanes];
static unsigned g_phase[kLanes];
static std::function<void(int)> g_body;

static void trampoline() {
    g_body(g_cur);
    g_done[g_cur] = true;
    swapcontext(&g_lane[g_cur], &g_sched);
}
// every lane deposits a word, then reads the words of ALL lanes of the same collective (double-buffered by parity)
static inline const uint32_t* collective(uint32_t v) {
    const int lane = g_cur;
    const unsigned k = g_phase[lane]++;
    g_x[k & 1][lane] = v;
    swapcontext(&g_lane[lane], &g_sched);  // resumed once every lane has deposited
    g_cur = lane;
    return g_x[k & 1];
}
template <class F>
static void run_warp(F&& f) {
    g_body = f;
    for (int l = 0; l < kLanes; ++l) {
        if (!g_stack[l]) g_stack[l] = (char*)std::malloc(kStack);
        g_done[l] = false;
        g_phase[l] = 0;
        getcontext(&g_lane[l]);
        g_lane[l].uc_stack.ss_sp = g_stack[l];
        g_lane[l].uc_stack.ss_size = kStack;
        g_lane[l].uc_link = nullptr;
        makecontext(&g_lane[l], trampoline, 0);
    }
    for (bool any = true; any;) {
        any = false;
        for (int l = 0; l < kLanes; ++l)
            if (!g_done[l]) {
                g_cur = l;
                swapcontext(&g_sched, &g_lane[l]);
                any = any || !g_done[l];
            }
    }
}
}  // namespace hk

static inline int __any_sync(unsigned, int pred) {
    const uint32_t* x = hk::collective(pred ? 1u : 0u);
    uint32_t r = 0;
    for (int l = 0; l < hk::kLanes; ++l) r |= x[l];
    return (int)r;
}
static inline int __all_sync(unsigned, int pred) {
    const uint32_t* x = hk::collective(pred ? 1u : 0u);
    uint32_t r = 1;
    for (int l = 0; l < hk::kLanes; ++l) r &= x[l];
    return (int)r;
}
static inline float __shfl_xor_sync(unsigned, float v, int bit) {
    uint32_t u; std::memcpy(&u, &v, 4);
    const int lane = hk::g_cur;
    const uint32_t* x = hk::collective(u);
    float r; std::memcpy(&r, &x[lane ^ bit], 4);
    return r;
}
using std::fabs; using std::fmax; using std::fma; using std::rint; using std::sqrt; using std::atan2; using std::sin; using std::cos;
