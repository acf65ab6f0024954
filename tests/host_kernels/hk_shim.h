// hk_shim.h -- TEST INFRASTRUCTURE (CPU suite only; never part of the product path).
// Lets the SOURCE TEXT of the factor-family kernels (csrc/fam_*.cu, se3_common.cuh, the arithmetic parts of
// device_utils.cuh / eval_pipeline.cuh, the pack/unpack kernels) compile as host C++: CUDA vocabulary as no-ops, the few
// intrinsics the kernels use, and a thread emulator (one fiber per CUDA thread of a block, switched at every warp or block
// collective) so that __any_sync / __all_sync / __shfl_*_sync / __syncwarp / __syncthreads behave as on the device.  What is NOT emulated: the TMA/mbarrier
// pipeline (the harness hands every factor its inputs directly) and the MUFU approximations (the accurate sampler
// branch is compiled; the `*.approx.ftz.f32` PTX statements are replaced by the exact functions in build.py).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

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
struct uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
typedef void* cudaStream_t;

// streaming stores are plain stores here
static inline void __stcs(float* p, float v) { *p = v; }
static inline void __stcs(float2* p, float2 v) { *p = v; }
static inline void __stcs(float4* p, float4 v) { *p = v; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdividef(float a, float b) { return a / b; }
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

// ---- thread emulator: one fiber per CUDA thread of ONE block, round-robin, switched at every collective --------------
// Warp collectives (__any_sync, __all_sync, __shfl_*_sync, __syncwarp) rendezvous the 32 fibers of a warp, __syncthreads
// all fibers of the block; a fiber that arrives early keeps yielding until its rendezvous is complete, so warps may
// run ahead of each other exactly as far as the barriers of the kernel allow.
struct dim3_ { unsigned x, y, z; };
static dim3_ blockIdx = {0, 1, 1}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};
namespace hk {
constexpr int kMaxThreads = 416;  // eval_kernel: 8 + 1 warps, eval_kernel_w: 12 warps, product_kernel: 4 warps
constexpr size_t kStack = 256 * 1024;
static ucontext_t g_sched, g_fiber[kMaxThreads];
static char* g_stack[kMaxThreads];
static bool g_done[kMaxThreads];
static int g_cur = 0, g_nthreads = 32;
static uint32_t g_x[2][kMaxThreads];
static unsigned g_warp_phase[kMaxThreads], g_warp_deposits[kMaxThreads / 32];
static unsigned g_block_phase[kMaxThreads], g_block_arrivals;
static std::function<void(int)> g_body;
alignas(128) static unsigned char g_smem[240 * 1024];  // the running block's dynamic shared memory
static unsigned long long g_progress = 0;  // bumped by everything that can unblock a waiting fiber
// Asynchronous bulk copies: with g_async_max > 0 a copy is not performed when it is issued but a pseudo-random number of
// scheduler rounds later (1 .. g_async_max), like the copy engine working behind the issuing thread's back; loads
// complete their bytes on the mbarrier only then, stores read shared memory only then -- unless the issuing thread waits
// for its bulk group first.  g_async_max = 0: copies are instantaneous.
struct PendingCopy { void* dst; const void* src; uint32_t bytes; uint64_t* bar; int owner; unsigned long long due; };
static std::vector<PendingCopy> g_pending;
static unsigned long long g_round = 0;
static unsigned g_async_max = 0, g_async_state = 12345u;
static inline unsigned async_delay() {
    g_async_state = g_async_state * 1664525u + 1013904223u;
    return 1u + (g_async_state >> 8) % g_async_max;
}
static void complete_copy(const PendingCopy& c);   // defined with the mbarrier emulation below
static inline void run_due_copies(bool all_of_owner = false, int owner = -1) {
    for (size_t i = 0; i < g_pending.size();) {
        const PendingCopy c = g_pending[i];
        if (all_of_owner ? (c.owner == owner && c.bar == nullptr) : c.due <= g_round) {
            g_pending.erase(g_pending.begin() + i);
            complete_copy(c);
        } else {
            ++i;
        }
    }
}
static bool g_deadlock = false;

static inline dim3_ tid3() { return {(unsigned)g_cur, 0, 0}; }
static void trampoline() {
    g_body(g_cur);
    g_done[g_cur] = true;
    swapcontext(&g_fiber[g_cur], &g_sched);
}
static inline void yield_() {
    const int t = g_cur;
    swapcontext(&g_fiber[t], &g_sched);
    g_cur = t;
}
// every lane of the warp deposits a word, then reads the words of the warp's 32 lanes (double-buffered by parity)
static inline const uint32_t* collective(uint32_t v) {
    const int t = g_cur, w = t >> 5;
    const unsigned k = g_warp_phase[t]++;
    g_x[k & 1][t] = v;
    ++g_warp_deposits[w];
    ++g_progress;
    while (g_warp_deposits[w] < 32u * (k + 1)) yield_();
    return &g_x[k & 1][w << 5];
}
static inline void block_barrier() {
    const int t = g_cur;
    const unsigned k = g_block_phase[t]++;
    ++g_block_arrivals;
    ++g_progress;
    while (g_block_arrivals < (unsigned)g_nthreads * (k + 1)) yield_();
}
template <class F>
static void run_block(int nthreads, F&& f) {
    g_body = f;
    g_nthreads = nthreads;
    g_block_arrivals = 0;
    for (int w = 0; w < kMaxThreads / 32; ++w) g_warp_deposits[w] = 0;
    for (int t = 0; t < nthreads; ++t) {
        if (!g_stack[t]) g_stack[t] = (char*)std::malloc(kStack);
        g_done[t] = false;
        g_warp_phase[t] = g_block_phase[t] = 0;
        getcontext(&g_fiber[t]);
        g_fiber[t].uc_stack.ss_sp = g_stack[t];
        g_fiber[t].uc_stack.ss_size = kStack;
        g_fiber[t].uc_link = nullptr;
        makecontext(&g_fiber[t], trampoline, 0);
    }
    g_deadlock = false;
    g_pending.clear();
    unsigned long long idle_rounds = 0;
    for (bool any = true; any;) {
        any = false;
        const unsigned long long before = g_progress;
        ++g_round;
        run_due_copies();
        if (!g_pending.empty()) ++g_progress;   // a copy still in flight will unblock somebody
        for (int t = 0; t < nthreads; ++t)
            if (!g_done[t]) {
                g_cur = t;
                swapcontext(&g_sched, &g_fiber[t]);
                any = any || !g_done[t];
                if (g_done[t]) ++g_progress;
            }
        // a round in which no fiber finished, no rendezvous completed and no barrier word changed is a deadlock of the
        // emulated kernel (e.g. an mbarrier whose expected byte count never arrives): give up instead of spinning
        idle_rounds = (g_progress == before) ? idle_rounds + 1 : 0;
        if (idle_rounds > 4) { g_deadlock = true; return; }
    }
    if (!g_pending.empty()) g_deadlock = true;   // the block exited with bulk copies in flight (no wait on their group)
}
template <class F>
static void run_warp(F&& f) { run_block(32, f); }
}  // namespace hk
#define threadIdx (hk::tid3())

static inline int __any_sync(unsigned, int pred) {
    const uint32_t* x = hk::collective(pred ? 1u : 0u);
    uint32_t r = 0;
    for (int l = 0; l < 32; ++l) r |= x[l];
    return (int)r;
}
static inline int __all_sync(unsigned, int pred) {
    const uint32_t* x = hk::collective(pred ? 1u : 0u);
    uint32_t r = 1;
    for (int l = 0; l < 32; ++l) r &= x[l];
    return (int)r;
}
static inline float hk_shfl(float v, int src_lane_fn(int lane, int arg), int arg) {
    uint32_t u; std::memcpy(&u, &v, 4);
    const int lane = hk::g_cur & 31;
    const uint32_t* x = hk::collective(u);
    const int src = src_lane_fn(lane, arg);
    float r; std::memcpy(&r, &x[(src < 0 || src > 31) ? lane : src], 4);
    return r;
}
static inline float __shfl_xor_sync(unsigned, float v, int bit) { return hk_shfl(v, [](int l, int a) { return l ^ a; }, bit); }
static inline float __shfl_up_sync(unsigned, float v, int d) { return hk_shfl(v, [](int l, int a) { return l - a; }, d); }
static inline float __shfl_sync(unsigned, float v, int src) { return hk_shfl(v, [](int, int a) { return a & 31; }, src); }
static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return hk::collective(v)[src & 31]; }
static inline int __shfl_sync(unsigned, int v, int src) { return (int)hk::collective((uint32_t)v)[src & 31]; }
static inline float __shfl_down_sync(unsigned, float v, int d) { return hk_shfl(v, [](int l, int a) { return l + a; }, d); }
static inline int __shfl_down_sync(unsigned, int v, int d) {
    const int lane = hk::g_cur & 31, src = lane + d;
    return (int)hk::collective((uint32_t)v)[src > 31 ? lane : src];
}
static inline void __syncwarp(unsigned = 0xffffffffu) { hk::collective(0u); }
static inline void __syncthreads() { hk::block_barrier(); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
#define __align__(n) __attribute__((aligned(n)))

// ---- mbarrier + 1-D bulk copies (device_utils.cuh's PTX wrappers, host semantics) -----------------------------------
// The 64-bit barrier word holds {phase bit, pending arrivals, pending transaction bytes, arrival count of a phase}; a
// bulk load is performed at once (one legal schedule of the asynchronous copy) and completes its bytes on the barrier;
// a phase completes when both pending counts reach zero.  try_wait.parity(p) succeeds once the phase of parity p is over.
static long long hk_clock_ticks = 0;
static inline long long clock64() { return ++hk_clock_ticks; }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
static inline uint32_t atomicExch(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = v; return o; }
namespace rome {
static inline void spin_pause() { hk::yield_(); }
static inline void st_release_sys_u32(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; ++hk::g_progress; }
static inline uint32_t ld_acquire_sys_u32(const uint32_t* p) { hk::yield_(); return *(volatile const uint32_t*)p; }
struct HkBar { uint8_t phase, expected; int16_t pending; int32_t tx; };
static_assert(sizeof(HkBar) <= 8, "an emulated mbarrier must fit the kernel's 8-byte barrier slot");
static inline void hk_bar_check(HkBar* b) {
    if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = (int16_t)b->expected; }
    ++hk::g_progress;
}
static inline void mbar_init(uint64_t* bar, uint32_t count) {
    HkBar* b = reinterpret_cast<HkBar*>(bar);
    b->phase = 0; b->expected = (uint8_t)count; b->pending = (int16_t)count; b->tx = 0;
}
static inline void fence_mbar_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    HkBar* b = reinterpret_cast<HkBar*>(bar);
    b->tx += (int32_t)bytes; b->pending -= 1;
    hk_bar_check(b);
}
static inline void mbar_arrive(uint64_t* bar) {
    HkBar* b = reinterpret_cast<HkBar*>(bar);
    b->pending -= 1;
    hk_bar_check(b);
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    const HkBar* b = reinterpret_cast<const HkBar*>(bar);
    while (b->phase == parity) hk::yield_();
}
static inline void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    if ((bytes & 15u) || (reinterpret_cast<uintptr_t>(smem_dst) & 15u) || (reinterpret_cast<uintptr_t>(gmem_src) & 15u)) std::abort();
    const hk::PendingCopy c = {smem_dst, gmem_src, bytes, bar, hk::g_cur, hk::g_round + (hk::g_async_max ? hk::async_delay() : 0)};
    if (hk::g_async_max) hk::g_pending.push_back(c); else hk::complete_copy(c);
}
static inline void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    if ((bytes & 15u) || (reinterpret_cast<uintptr_t>(gmem_dst) & 15u) || (reinterpret_cast<uintptr_t>(smem_src) & 15u)) std::abort();
    const hk::PendingCopy c = {gmem_dst, smem_src, bytes, nullptr, hk::g_cur, hk::g_round + (hk::g_async_max ? hk::async_delay() : 0)};
    if (hk::g_async_max) hk::g_pending.push_back(c); else hk::complete_copy(c);
}
static inline void tma_store_commit() {}
// waiting for the thread's bulk groups: its stores in flight are carried out now (they have read their source, and --
// for wait_all -- written their destination; the emulation does not distinguish the two)
static inline void tma_store_wait_read() { hk::run_due_copies(true, hk::g_cur); }
static inline void tma_store_wait_all() { hk::run_due_copies(true, hk::g_cur); }
static inline void tma_store_wait_pending(int) { hk::run_due_copies(true, hk::g_cur); }
}  // namespace rome
namespace hk {
static void complete_copy(const PendingCopy& c) {
    std::memcpy(c.dst, c.src, c.bytes);
    if (c.bar) {
        rome::HkBar* b = reinterpret_cast<rome::HkBar*>(c.bar);
        b->tx -= (int32_t)c.bytes;
        rome::hk_bar_check(b);
    }
    ++g_progress;
}
}  // namespace hk
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorInvalidConfiguration = 9 };
#define __sincosf(x, s, c) sincosf((x), (s), (c))
static inline int max(int a, int b) { return a > b ? a : b; }
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__
using std::fabs; using std::fmax; using std::fma; using std::rint; using std::sqrt; using std::atan2; using std::sin; using std::cos;
