// eval_pipeline.cuh -- the persistent TMA producer/consumer pipeline shared by every factor family:
//   persistent CTAs = FT consumer warps (one factor each per tile of FT factors) + 1 producer warp;
//   the producer stages, per tile and S tiles ahead, everything the consumers read into shared memory
//   with 1-D TMA bulk copies (cp.async.bulk, completion on an mbarrier per stage): the tile's factor-table
//   rows {var ids, mu (f64), chol(Sigma) (f32)}, its measurement block, and -- gathered by variable id --
//   one contiguous particle block {anchor (f64), Npad x d float32 offsets} per factor slot;
//   consumers never issue a global load: they wait on the stage's "full" barrier, compute in Float64 on
//   the anchored float32 data (DESIGN.md "precision"), write residual / proposal rows into the warp's
//   shared-memory slice, flush it with a warp-local TMA bulk store and release the stage through its
//   "empty" barrier; statistics are reduced with a halving-butterfly of __shfl_xor_sync (device_utils.cuh).
// A family (fam_*.cu) supplies `struct Fam { Row, D0, D1, DM, DR, DFWD, kMinCtas, factor<kStatic,kSample>() }`
// and instantiates launch_family<Fam>.
#pragma once
#include <cuda_runtime.h>

#include "../../include/rome_b200.h"
#include "device_utils.cuh"
#include "tables.h"

namespace rome {

// =============================================================================================
// SE(2) statistics accumulator: 16 additive values per factor
//   0..2 sum r | 3..8 sum r r' (11 12 13 22 23 33) | 9,10 sum proposal (dx,dy) | 11,12 sum cos/sin of the
//   proposal heading offset | 13..15 sum dx^2, dx dy, dy^2     (offsets from the target anchor)
// =============================================================================================
__device__ __forceinline__ void acc_res3(float (&st)[16], float m, float r1, float r2, float r3) {
    r1 *= m; r2 *= m; r3 *= m;
    st[0] += r1; st[1] += r2; st[2] += r3;
    st[3] = fmaf(r1, r1, st[3]); st[4] = fmaf(r1, r2, st[4]); st[5] = fmaf(r1, r3, st[5]);
    st[6] = fmaf(r2, r2, st[6]); st[7] = fmaf(r2, r3, st[7]); st[8] = fmaf(r3, r3, st[8]);
}
__device__ __forceinline__ void acc_prop2(float (&st)[16], float m, float dx, float dy) {
    dx *= m; dy *= m;
    st[9] += dx; st[10] += dy;
    st[13] = fmaf(dx, dx, st[13]); st[14] = fmaf(dx, dy, st[14]); st[15] = fmaf(dy, dy, st[15]);
}
__device__ __forceinline__ void acc_heading(float (&st)[16], float m, float dth) {
    float s, c;
    sincos_any_f(dth, s, c);
    st[11] = fmaf(m, c, st[11]);
    st[12] = fmaf(m, s, st[12]);
}
__device__ __forceinline__ void write_stats16(float (&st)[16], float* stats, int f, int lane) {
    const float tot = warp_reduce_scatter16(st, lane);
    if ((lane & 1) == 0) stats[(size_t)f * 16 + (lane >> 1)] = tot;
}
// What a consumer warp sees of its factor: inputs already in shared memory, plus its private output slice.
struct FactorView {
    const unsigned char* b0;  // particle block of the first variable  {anchor, rows}
    const unsigned char* b1;  // particle block of the second variable (nullptr for priors)
    const unsigned char* b2;  // particle block of the third variable (families with three variables, else nullptr)
    const float* meas;        // [dm][Npad] measurement offsets (nullptr with SAMPLE)
    float* out_res;           // [dr][Npad] residual rows, flushed by a warp-local TMA bulk store
    float* out_fwd;           // [dfwd][Npad] forward-proposal rows, same
    bool fwd_on;              // false: this factor writes no forward row (ROME_B200_ROUTED_ONLY and no destination)
};
constexpr uint32_t kHot1 = ROME_B200_RESIDUAL | ROME_B200_STATS;
constexpr uint32_t kHot2 = ROME_B200_RESIDUAL | ROME_B200_STATS | ROME_B200_PROPOSAL_FWD;

// =============================================================================================
// SE(2) families.  Rows are particle-major ([Npad][d], the reference's own `vecval` order): lane l owns
// particles l, l+32, l+64, ...; consecutive lanes read consecutive 12-B (8-B) records -> bank-conflict free.
// The slot loop is unrolled by four: one Philox/Box-Muller batch serves four particles of a lane and the
// Float64 chains of the four slots interleave.  kStatic != 0 fixes the output flags at compile time.
// =============================================================================================
// normals for the lane's slots [4g, 4g+4) of factor f: D normals per particle, 4 per Philox call
template <int D>
__device__ __forceinline__ void normals_for_group(const EvalParams& P, int f, int lane, int g, float (&z)[4 * D]) {
#pragma unroll
    for (int b = 0; b < D; ++b)
        normal4(P.seed_lo, P.seed_hi, P.stream_id, (uint32_t)f, (uint32_t)lane, (uint32_t)(g * D + b), &z[4 * b]);
}

// Slot loop.  A group = the lane's 4 slots {n0, n0+32, n0+64, n0+96}.  A group with at least three live slots
// is evaluated as branch-free straight-line code: first the four slot BODIES (loads + arithmetic into
// registers), then the four slot STORES -- no shared-memory store sits between the loads of different slots, so
// the Float64 chains of the four particles may interleave.  Slots 0-1 are live for every lane; a lane whose 3rd or
// 4th particle is beyond Npad reads particle `lane` instead (in range: this path needs Npad > 64) and has its
// stores/statistics masked; in the trailing partial group dead lanes read particle 0 -- with Npad < 32 particle `lane`
// would lie beyond the block, and a NaN/Inf found there survives the multiplication by the zero mask.  FASTCOND
// (warp-uniform) selects the variant whose body may assume kFast (e.g. small heading offsets -> polynomial
// sin/cos without a fallback branch).  The trailing partial group is evaluated with warp-uniform guards.
// The family defines ROME_SLOT_DECL (per-group register arrays) and ROME_SLOT_STORE (uses k, n, live).
#define ROME_SLOT_LOOP(FASTCOND, ...)                                                          \
    for (int g = 0, n0 = lane; n0 < Npad; ++g, n0 += 128) {                                    \
        float z[4 * DZ];                                                                       \
        if (kSample) normals_for_group<DZ>(P, f, lane, g, z);                                  \
        ROME_SLOT_DECL                                                                         \
        if (n0 - lane + 64 < Npad) { /* at least three live slots: straight-line code for all four */ \
            const int n2 = (n0 + 64 < Npad) ? n0 + 64 : lane;                                  \
            const int n3 = (n0 + 96 < Npad) ? n0 + 96 : lane;                                  \
            if (FASTCOND) {                                                                    \
                constexpr bool kFast = true; (void)kFast;                                      \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                \
                    const int nn = n0 + 32 * k;                                                \
                    const bool live = (k < 2) || nn < Npad;                                    \
                    const int n = (k < 2) ? nn : (k == 2 ? n2 : n3);                           \
                    __VA_ARGS__                                                                \
                }                                                                              \
            } else {                                                                           \
                constexpr bool kFast = false; (void)kFast;                                     \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                \
                    const int nn = n0 + 32 * k;                                                \
                    const bool live = (k < 2) || nn < Npad;                                    \
                    const int n = (k < 2) ? nn : (k == 2 ? n2 : n3);                           \
                    __VA_ARGS__                                                                \
                }                                                                              \
            }                                                                                  \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                    \
                const int nn = n0 + 32 * k;                                                    \
                const bool live = (k < 2) || nn < Npad;                                        \
                const int n = (k < 2) ? nn : (k == 2 ? n2 : n3);                               \
                ROME_SLOT_STORE                                                                \
            }                                                                                  \
        } else {                                                                               \
            constexpr bool kFast = false; (void)kFast;                                         \
            _Pragma("unroll") for (int k = 0; k < 2; ++k) {                                    \
                const int nn = n0 + 32 * k;                                                    \
                if (n0 - lane + 32 * k < Npad) {                                               \
                    const bool live = nn < Npad;                                               \
                    const int n = live ? nn : 0;  /* dead lanes re-read particle 0 (always in range) */ \
                    __VA_ARGS__                                                                \
                    ROME_SLOT_STORE                                                            \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    }

// =============================================================================================
// stage layout (shared by host planning and the kernel)
// =============================================================================================
struct StageLayout {
    int rows_off, v0_off, v1_off, v2_off, meas_off, bytes;
    int b0, b1, b2, mb;  // bytes of one slot-0 / slot-1 / slot-2 block, one factor's measurement block
};
__host__ __device__ inline StageLayout stage_layout(int ft, int row_bytes, int d0, int d1, int dm, bool sample,
                                                    int Npad, int d2 = 0) {
    StageLayout L;
    L.b0 = var_block_bytes(d0, Npad);
    L.b1 = d1 ? var_block_bytes(d1, Npad) : 0;
    L.b2 = d2 ? var_block_bytes(d2, Npad) : 0;
    L.mb = sample ? 0 : dm * Npad * 4;
    L.rows_off = 0;
    L.v0_off = (ft * row_bytes + 127) / 128 * 128;
    L.v1_off = L.v0_off + ft * L.b0;
    L.v2_off = L.v1_off + ft * L.b1;
    L.meas_off = L.v2_off + ft * L.b2;
    L.bytes = (L.meas_off + ft * L.mb + 127) / 128 * 128;
    return L;
}
constexpr int kBarrierBytes = 128;
constexpr int kMaxStages = 6;

// =============================================================================================
// Order in which a persistent CTA visits its tiles (tile b + i * grid, i = 0 .. n-1).  A launch that takes part in the
// rank barrier visits the tiles holding peer-dependent factors (a contiguous range of tile indices) FIRST:
//   * they are the only tiles that must wait for the peers (BARRIER_WAIT) -- and what they wait for, the peers' signal of
//     the PREVIOUS step, was published early in that step (next point), i.e. long ago;
//   * once the rows they store into peer memory have landed, this rank's signal can go out (BARRIER_SIGNAL) while the
//     interior tiles, the bulk of the launch, are still being evaluated: the latency of the NVLink stores, of the
//     system-scope fence and of the flag's flight all hide behind interior work instead of following the kernel.
// =============================================================================================
struct TileOrder {
    int n, i0, nc;     // tiles of this CTA; its peer-dependent tiles are i0 .. i0 + nc - 1
    int cut_ctas;      // CTAs of the grid that hold peer-dependent tiles
    __device__ __forceinline__ void init(int nTiles, int grid, int b, bool front, int c0, int c1 /* tile range */) {
        n = nTiles > b ? (nTiles - b + grid - 1) / grid : 0;
        i0 = 0; nc = 0; cut_ctas = 0;
        if (c1 > nTiles - 1) c1 = nTiles - 1;
        if (front && c1 >= c0) {
            cut_ctas = (c1 - c0 + 1) < grid ? (c1 - c0 + 1) : grid;
            if (n > 0) {
                int lo = c0 <= b ? 0 : (c0 - b + grid - 1) / grid;        // first i with b + i grid >= c0
                int hi = c1 < b ? -1 : (c1 - b) / grid;                    // last i with b + i grid <= c1
                if (hi > n - 1) hi = n - 1;
                if (hi >= lo) { i0 = lo; nc = hi - lo + 1; }
            }
        }
    }
    // visiting order: ONE interior tile, then the peer-dependent tiles, then the other interior tiles (the wait for the
    // peers' flags, issued by the fetching warp before the first peer-dependent tile, overlaps that first interior tile)
    __device__ __forceinline__ int lead() const { return (nc > 0 && n > nc) ? 1 : 0; }
    __device__ __forceinline__ int interior(int m) const { return m < i0 ? m : m + nc; }
    __device__ __forceinline__ int at(int j) const {  // position j of the visiting order -> i
        const int ld = lead();
        if (j < ld) return interior(0);
        if (j < ld + nc) return i0 + (j - ld);
        return interior(j - nc);
    }
    // position at which the early signal is prepared: two tiles after the last peer-dependent one (its rows have had
    // two tile times to land), or the CTA's last position
    __device__ __forceinline__ int signal_pos() const {
        const int p = lead() + nc + 1;
        return p < n ? p : n - 1;
    }
};
// owner of the barrier range of a launch, relative to its factor range: tile range [c0, c1] (c1 < c0: none)
__device__ __forceinline__ void barrier_tiles(const EvalParams& P, int ft, int& c0, int& c1) {
    const int lo = P.bar_lo - P.first, hi = P.bar_hi - P.first;
    if (!(P.flags & (ROME_B200_BARRIER_WAIT | ROME_B200_BARRIER_SIGNAL)) || hi <= 0 || lo >= P.count || hi <= lo) {
        c0 = 0; c1 = -1;
        return;
    }
    c0 = (lo > 0 ? lo : 0) / ft;
    c1 = ((hi < P.count ? hi : P.count) - 1) / ft;
}

// =============================================================================================
// rank barrier fused into the evaluation kernels (owner-sharded multi-GPU sweeps; the state buffer is the one of
// rome_b200_peer_signal / _wait: words [0, 8) flag slots written by the peers, word 8 this rank's signal epoch, word 10
// give-up status, word 12 the CTA counter of the signalling launch).
//   ROME_B200_BARRIER_WAIT   (a launch of the next step): before the first factor of [bar_lo, bar_hi) -- the family's cut
//       factors; default: every factor -- is fetched, the fetching warp polls the local flag slots until each peer has
//       signalled as often as this rank has (word 8): the peers' previous step, with its stores into this GPU's memory,
//       is complete.
//   ROME_B200_BARRIER_SIGNAL: as soon as the launch's peer-dependent tiles have landed (they are visited first, see
//       TileOrder) -- or, for a launch without any, when its grid has finished -- the next epoch is published to the slot
//       this rank owns in every peer's state (st.release.sys after a system-scope fence: the rows this grid stored into
//       peer memory are visible before the flag).
// No extra kernel, no host round trip: the barrier costs the NVLink latency of a 4-byte store.
// =============================================================================================
// executed by ONE WARP (the warp that fetches particle blocks), right before the first factor of [P.bar_lo, P.bar_hi) is
// fetched: the other factors -- a rank's interior factors -- neither read halo blocks nor write into peer memory, so
// their evaluation overlaps the barrier's latency
__device__ __forceinline__ void fused_barrier_wait(const EvalParams& P, int lane) {
    const long long t0 = clock64();
    if (lane < P.bar_n) {
        const uint32_t target = *reinterpret_cast<volatile const uint32_t*>(P.bar_state + 8);
        for (;;) {
            const uint32_t v = ld_acquire_sys_u32(P.bar_state + lane);
            if ((int32_t)(v - target) >= 0) break;  // wrap-safe "v >= target"
            if (clock64() - t0 > P.bar_timeout) {    // a peer died: give up instead of hanging the GPU
                atomicExch(P.bar_state + 10, 1u);
                break;
            }
        }
    }
    __syncwarp();
    if (lane == 0) {  // instrumentation (words 13 / 14 of the state): cycles spent at the barrier, number of passes
        atomicAdd(P.bar_state + 13, (uint32_t)(clock64() - t0));
        atomicAdd(P.bar_state + 14, 1u);
    }
}
__device__ __forceinline__ void fused_barrier_signal(const EvalParams& P) {
    __syncthreads();  // every warp of this CTA has waited for its bulk stores
    if (threadIdx.x < 32) {
        uint32_t done = 0;
        if (threadIdx.x == 0) {
            __threadfence();
            done = atomicAdd(P.bar_state + 12, 1u);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done == gridDim.x - 1) {  // the last CTA of the grid: lane r publishes the epoch to peer r, all in parallel
            const long long t0 = clock64();
            const uint32_t e = *reinterpret_cast<volatile uint32_t*>(P.bar_state + 8) + 1;
            __syncwarp();
            if (threadIdx.x == 0) {
                P.bar_state[12] = 0;
                P.bar_state[8] = e;
            }
            if (threadIdx.x < (unsigned)P.bar_n) st_release_sys_u32(P.bar_peer[threadIdx.x], e);  // fence.sys + store each
            __syncwarp();
            if (threadIdx.x == 0) atomicAdd(P.bar_state + 15, (uint32_t)(clock64() - t0));  // instrumentation: publish cycles
        }
    }
}

// Early signal (see TileOrder), in two halves.
//  (1) every consumer warp of a CTA that holds peer-dependent tiles, two tiles after the last of them: wait until the
//      bulk-store groups of those tiles are complete (`pending` = groups committed since: they may still be in flight) and
//      count the warp in the CTA's shared-memory counter.  Nothing else: the arithmetic warps never pay for a fence.
//  (2) the warp that has time -- the producer warp, once it has issued the CTA's last tile; in the per-warp pipeline the
//      warp itself -- makes those rows visible system-wide (fence.sys is cumulative over what it observed through the
//      counter), counts the CTA in the grid's counter, and the last one publishes the epoch to every peer in parallel.
__device__ __forceinline__ void early_signal_arrive(int lane, int pending, int* cta_count) {
    if (lane == 0) {
        tma_store_wait_pending(pending);
        __threadfence_block();
        atomicAdd(cta_count, 1);
    }
}
__device__ __forceinline__ void early_signal_publish(const EvalParams& P, int lane, uint32_t expected) {
    uint32_t pub = 0;
    if (lane == 0) {
        __threadfence_system();
        pub = (atomicAdd(P.bar_state + 12, 1u) == expected - 1) ? 1u : 0u;
    }
    pub = __shfl_sync(0xffffffffu, pub, 0);
    if (pub) {
        const uint32_t e = *reinterpret_cast<volatile uint32_t*>(P.bar_state + 8) + 1;
        __syncwarp();
        if (lane == 0) {
            P.bar_state[12] = 0;
            P.bar_state[8] = e;
        }
        if (lane < P.bar_n) st_release_sys_u32(P.bar_peer[lane], e);
    }
}

// =============================================================================================
// persistent producer/consumer pipeline
//   smem: [full[], empty[] mbarriers, early-signal counter | S input stages | FT per-warp output slices]
// =============================================================================================
// (A warpgroup register re-allocation -- 384 threads, setmaxnreg.inc 104 for the two consumer warpgroups, .dec 24 for the
// producer's -- was measured on B200 and is SLOWER than this 288-thread form at 96 registers: 16.6 vs 15.8 us per
// launch of the bench workload; requesting more registers than the CTA was launched with hangs.  profiles/r02_analysis.md)
template <int FT>
constexpr int eval_threads() { return (FT + 1) * 32; }
// dimension of a family's THIRD variable (0: none) -- families declare `static constexpr int D2` only when they have one
template <class Fam, class = void>
struct FamD2 { static constexpr int value = 0; };
template <class Fam>
struct FamD2<Fam, decltype((void)Fam::D2)> { static constexpr int value = Fam::D2; };
// resident CTAs per SM the kernel is compiled for: the 4-factor tile runs twice as many CTAs as the 8-factor tile
template <class Fam, int FT>
constexpr int eval_min_ctas() { return FT == 4 ? 2 * Fam::kMinCtas : Fam::kMinCtas; }
// kRouted (the compile-time RESIDUAL|STATS variant of a ROUTED_ONLY launch): factors without a destination run the
// compile-time RESIDUAL|STATS body; the few factors WITH one (a rank's cut factors) run the compile-time
// RESIDUAL|STATS|PROPOSAL_FWD body, which also writes their forward row -- the common path keeps the instruction count
// of the plain variant.  (An out-of-line run-time-flag body for the routed factors was measured: 6x slower per factor,
// and the cut factors sit in the last tiles, i.e. on the kernel's critical path.)
template <class Fam, uint32_t kStatic, bool kSample, int FT, bool kRouted = false>
__global__ void __launch_bounds__(eval_threads<FT>(), eval_min_ctas<Fam, FT>()) eval_kernel(const __grid_constant__ EvalParams P) {
    using Row = typename Fam::Row;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kMaxStages;
    unsigned char* stage0 = smem + kBarrierBytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = P.stages;
    const int nTiles = (P.count + FT - 1) / FT;
    constexpr int D2 = FamD2<Fam>::value;
    const StageLayout L = stage_layout(FT, (int)sizeof(Row), Fam::D0, Fam::D1, Fam::DM, kSample, P.Npad, D2);
    const Row* __restrict__ table = reinterpret_cast<const Row*>(P.rows) + P.first;
    const uint32_t flags = kStatic ? kStatic : P.flags;

    // The producer warp holds the variable ids of a CHUNK of 32/FT tiles at once (lane l <-> tile l/FT of the
    // chunk, factor l%FT): one global-load latency per chunk instead of one per tile; the first chunk is
    // requested before the barrier initialisation is published.
    constexpr int TPC = 32 / FT;  // tiles per chunk
    const int jl = lane / FT, fl_in_tile = lane % FT;
    int2 ids_cur = make_int2(0, 0);
    int id2_cur = 0;  // third variable of the factor (families with three variables)
    // Multi-GPU features (per-factor row destinations, peer replication, the fused rank barrier with its tile order) are
    // compiled into the run-time-flag variant, the forward-proposal variant and the routed variant only: the plain
    // RESIDUAL|STATS variant -- the single-GPU hot path -- carries none of their state through its loop (launches that
    // need them are planned onto the routed variant, plan_launch).
    constexpr bool kMulti = kRouted || kStatic == 0u || (kStatic & ROME_B200_PROPOSAL_FWD) != 0u;
    TileOrder ord;  // visiting order of this CTA's tiles (peer-dependent tiles first when the launch takes part in the barrier)
    {
        int c0 = 0, c1 = -1;
        if (kMulti) barrier_tiles(P, FT, c0, c1);
        ord.init(nTiles, (int)gridDim.x, (int)blockIdx.x, kMulti && c1 >= c0, c0, c1);
    }
    // the signal goes out as soon as the peer-dependent tiles have landed (early), or -- a launch without such tiles, or
    // one that replicates its rows to every peer -- when the whole grid has finished
    const bool sig_early = kMulti && (P.flags & ROME_B200_BARRIER_SIGNAL) && ord.cut_ctas > 0 && P.n_peers == 0;
    int* sig_count = reinterpret_cast<int*>(smem + 2 * kMaxStages * 8);
    auto fetch_chunk = [&](int base_j, int& id2) {  // ids of visiting positions base_j .. base_j + TPC - 1
        const int j = base_j + jl;
        int2 ids = make_int2(0, 0);
        id2 = 0;
        if (j < ord.n) {
            const int fl = ((int)blockIdx.x + ord.at(j) * (int)gridDim.x) * FT + fl_in_tile;
            if (fl < P.count) {
                ids = __ldg(reinterpret_cast<const int2*>(table + fl));
                if constexpr (D2 > 0) id2 = __ldg(&table[fl].ir);
            }
        }
        return ids;
    };
    if (warp == FT) ids_cur = fetch_chunk(0, id2_cur);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], FT);
        }
        *sig_count = 0;
        fence_mbar_init();
    }
    __syncthreads();
    // Programmatic dependent launch.  Every launch lets its successor start on SMs this grid has vacated
    // (launch_dependents), and every launch is itself started that way: what it has done so far -- barrier
    // initialisation and the first variable ids, read from the factor table, which no kernel writes -- overlaps the
    // predecessor's tail and the launch latency.  Before the first access to data a predecessor may have written
    // (particle blocks) or may still read (the output buffers) it waits for the predecessor grid to complete; a launch
    // flagged ROME_B200_INDEPENDENT touches no such data and defers that wait to its end.  A dependent launch releases
    // ITS successors only after its own wait: an INDEPENDENT successor starts without waiting, so it must not be able to
    // start before the sweep's first launch has seen the previous sweep complete.
    if (!(P.flags & ROME_B200_INDEPENDENT)) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == FT) {
        // ---------------- producer warp ---------------------------------------------------------------------
        int s = 0;
        uint32_t phase = 1;  // parity of the previous round; the first pass over the ring does not wait
        bool first_round = true;
        bool synced = !kMulti || !(P.flags & ROME_B200_BARRIER_WAIT);  // rank barrier still to be passed?
        for (int base = 0; base < ord.n; base += TPC) {
            int id2_next = 0;
            const int2 ids_next = fetch_chunk(base + TPC, id2_next);  // in flight while this chunk is issued
#pragma unroll 1
            for (int j = 0; j < TPC; ++j) {
                if (base + j >= ord.n) break;
                const int tile = (int)blockIdx.x + ord.at(base + j) * (int)gridDim.x;
                if (!first_round) mbar_wait(&empty[s], phase);
                unsigned char* st = stage0 + (size_t)s * L.bytes;
                const int t0 = tile * FT, nf = min(FT, P.count - tile * FT);
                if (!synced && P.first + t0 + nf > P.bar_lo && P.first + t0 < P.bar_hi) {  // a tile with peer-dependent factors
                    fused_barrier_wait(P, lane);
                    synced = true;
                }
                if (lane == j * FT) {
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(nf * ((int)sizeof(Row) + L.b0 + L.b1 + L.b2 + L.mb)));
                    tma_load_1d(st + L.rows_off, table + t0, (uint32_t)(nf * sizeof(Row)), &full[s]);
                    if (!kSample)
                        tma_load_1d(st + L.meas_off, P.meas + (size_t)(P.first + t0) * Fam::DM * P.Npad,
                                    (uint32_t)(nf * L.mb), &full[s]);
                }
                __syncwarp();
                if (jl == j && fl_in_tile < nf) {
                    tma_load_1d(st + L.v0_off + fl_in_tile * L.b0, P.v0 + (size_t)ids_cur.x * L.b0, (uint32_t)L.b0,
                                &full[s]);
                    if (Fam::D1)
                        tma_load_1d(st + L.v1_off + fl_in_tile * L.b1, P.v1 + (size_t)ids_cur.y * L.b1,
                                    (uint32_t)L.b1, &full[s]);
                    if constexpr (D2 > 0)
                        tma_load_1d(st + L.v2_off + fl_in_tile * L.b2, P.v2 + (size_t)id2_cur * L.b2, (uint32_t)L.b2,
                                    &full[s]);
                }
                if (++s == S) { s = 0; phase ^= 1u; first_round = false; }
            }
            ids_cur = ids_next;
            id2_cur = id2_next;
        }
        if (kMulti && sig_early && ord.nc > 0) {  // early signal, second half: when all consumer warps have reported their rows landed
            if (lane == 0)
                while (*reinterpret_cast<volatile int*>(sig_count) < FT) spin_pause();
            __syncwarp();
            early_signal_publish(P, lane, (uint32_t)ord.cut_ctas);
        }
    } else {
        // ---------------- consumer warps: warp w owns the tile's w-th factor ----------------------------------
        float* out = reinterpret_cast<float*>(stage0 + (size_t)S * L.bytes + (size_t)warp * P.out_warp_bytes);
        const int res_floats = Fam::DR * P.Npad;
        int s = 0;
        uint32_t phase = 0;
        bool wrote_peer = false;  // did this warp store rows into another GPU's memory?
        int after_cut = 0;        // bulk-store groups committed since the peer-dependent tiles
        const int sig_pos = sig_early && ord.nc > 0 ? ord.signal_pos() : -1;
        const int cut_end = ord.lead() + ord.nc;  // first position behind the peer-dependent tiles
        for (int j = 0; j < ord.n; ++j) {
            const int tile = (int)blockIdx.x + ord.at(j) * (int)gridDim.x;
            const int fl = tile * FT + warp;
            const bool mine = fl < P.count;
            const int f = P.first + fl;
            // owner-sharded exchange: this factor's forward row may have its own destination (a peer GPU); requested
            // before the wait for the stage so that the load's latency hides behind it (warp-uniform address)
            unsigned long long fdst = 0;
            // (interior factors -- outside the range that holds every destination -- skip the load: its L2 latency would
            // otherwise be paid by every factor once the ring runs ahead and the wait below returns at once)
            if (kMulti && mine && ((flags & ROME_B200_PROPOSAL_FWD) || kRouted) && P.fwd_dst && f >= P.fwd_dst_lo &&
                f < P.fwd_dst_hi)
                fdst = __ldg(P.fwd_dst + f);
            mbar_wait(&full[s], phase);
            const unsigned char* st = stage0 + (size_t)s * L.bytes;
            if (mine) {
                const Row row = reinterpret_cast<const Row*>(st + L.rows_off)[warp];
                FactorView V;
                V.b0 = st + L.v0_off + warp * L.b0;
                V.b1 = Fam::D1 ? st + L.v1_off + warp * L.b1 : nullptr;
                V.b2 = D2 ? st + L.v2_off + warp * L.b2 : nullptr;
                V.meas = kSample ? nullptr : reinterpret_cast<const float*>(st + L.meas_off + (size_t)warp * L.mb);
                V.out_res = out;
                V.out_fwd = out + res_floats;
                V.fwd_on = !kMulti || !(P.flags & ROME_B200_ROUTED_ONLY) || fdst != 0;
                if (kMulti) wrote_peer = wrote_peer || fdst != 0 || P.n_peers > 0;
                if (flags & (ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD)) {
                    if (lane == 0) tma_store_wait_read();  // the previous tile's rows have left the slice
                    __syncwarp();
                }
                const bool routed = kRouted && fdst != 0;  // warp-uniform
                if (routed) Fam::template factor<kHot2 | (kStatic & ROME_B200_SAMPLE), kSample>(row, P, V, f, lane);
                else Fam::template factor<kStatic, kSample>(row, P, V, f, lane);
                if (flags & (ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD)) {
                    fence_proxy_async();  // generic-proxy writes of the slice -> visible to the bulk-copy engine
                    __syncwarp();
                    if (lane == 0) {
                        if (flags & ROME_B200_RESIDUAL)
                            tma_store_1d(P.res + (size_t)f * res_floats, V.out_res, (uint32_t)(res_floats * 4));
                        if (((flags & ROME_B200_PROPOSAL_FWD) && V.fwd_on) || routed) {
                            const size_t off = (size_t)f * Fam::DFWD * P.Npad;
                            const uint32_t bytes = (uint32_t)(Fam::DFWD * P.Npad * 4);
                            tma_store_1d((kMulti && fdst) ? reinterpret_cast<float*>(fdst) : P.prop_fwd + off, V.out_fwd, bytes);
                            // fused all-gather: the same slice goes to every peer GPU over NVLink
                            if (kMulti)
                                for (int r = 0; r < P.n_peers; ++r) tma_store_1d(P.peer_fwd[r] + off, V.out_fwd, bytes);
                        }
                        tma_store_commit();
                        if (kMulti && j >= cut_end) ++after_cut;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            if (++s == S) { s = 0; phase ^= 1u; }
            if (kMulti && j == sig_pos) early_signal_arrive(lane, after_cut, sig_count);
        }
        if (lane == 0) {
            tma_store_wait_all();
            // end-of-grid signal: rows stored into peer memory are performed system-wide before this CTA reports
            // completion (only the warps that have such rows pay for the system-scope fence)
            if (kMulti && (P.flags & ROME_B200_BARRIER_SIGNAL) && !sig_early && wrote_peer) __threadfence_system();
        }
    }
    // a launch that overlapped its predecessor must not be seen as complete before the predecessor is
    if (P.flags & ROME_B200_INDEPENDENT) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (kMulti && (P.flags & ROME_B200_BARRIER_SIGNAL) && !sig_early) fused_barrier_signal(P);
}
// =============================================================================================
// per-warp pipeline ("warp pipeline"): no producer warp and no CTA-wide barrier.  Warp w of a CTA owns slot w of
// every tile the CTA visits and runs its own S-deep ring: lane s keeps the variable ids of the factor destined for
// stage s (fetched one factor ahead, behind the arithmetic) and issues that factor's four bulk copies (table row,
// the two particle blocks, the measurement block) onto the (warp, stage) mbarrier as soon as the warp has finished
// the factor that occupied the stage.  All registers of the CTA belong to consumer warps, so the SE(3) families run
// 12 warps x 168 registers per SM instead of 8 + producer, and warps never wait for each other.  Built for the Pose3
// families only (Fam::kWarpFT): measured 6 % faster there; for the SE(2) families (96 registers, two CTAs per SM) the
// producer-warp pipeline with its CTA-wide stages is as fast at N = 100 and clearly faster at N = 200.
//   smem: [bar[FT][kMaxStages] | FT x (S slots + output slice)]
// =============================================================================================
struct SlotLayout {
    int row_off, v0_off, v1_off, meas_off, bytes, b0, b1, mb;
};
__host__ __device__ inline SlotLayout slot_layout(int row_bytes, int d0, int d1, int dm, bool sample, int Npad) {
    SlotLayout L;
    L.b0 = var_block_bytes(d0, Npad);
    L.b1 = d1 ? var_block_bytes(d1, Npad) : 0;
    L.mb = sample ? 0 : dm * Npad * 4;
    L.row_off = 0;
    L.v0_off = (row_bytes + 15) / 16 * 16;
    L.v1_off = L.v0_off + L.b0;
    L.meas_off = L.v1_off + L.b1;
    L.bytes = (L.meas_off + L.mb + 127) / 128 * 128;
    return L;
}
template <class Fam, uint32_t kStatic, bool kSample, int FT, bool kRouted = false>
__global__ void __launch_bounds__(FT * 32, Fam::kMinCtas) eval_kernel_w(const __grid_constant__ EvalParams P) {
    using Row = typename Fam::Row;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = P.stages;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem) + warp * kMaxStages;
    const SlotLayout L = slot_layout((int)sizeof(Row), Fam::D0, Fam::D1, Fam::DM, kSample, P.Npad);
    const int warp_bytes = S * L.bytes + P.out_warp_bytes;
    unsigned char* slots = smem + ((FT * kMaxStages * 8 + 127) / 128 * 128) + (size_t)warp * warp_bytes;
    float* out = reinterpret_cast<float*>(slots + (size_t)S * L.bytes);
    const Row* __restrict__ table = reinterpret_cast<const Row*>(P.rows) + P.first;
    const uint32_t flags = kStatic ? kStatic : P.flags;
    const int res_floats = Fam::DR * P.Npad;
    // the i-th factor of this warp: tile blockIdx.x + i * gridDim.x, slot `warp` (-1: none)
    constexpr bool kMulti = kRouted || kStatic == 0u || (kStatic & ROME_B200_PROPOSAL_FWD) != 0u;  // see eval_kernel
    TileOrder ord;
    {
        int c0 = 0, c1 = -1;
        if (kMulti) barrier_tiles(P, FT, c0, c1);
        ord.init((P.count + FT - 1) / FT, (int)gridDim.x, (int)blockIdx.x, kMulti && c1 >= c0, c0, c1);
    }
    const bool sig_early = kMulti && (P.flags & ROME_B200_BARRIER_SIGNAL) && ord.cut_ctas > 0 && P.n_peers == 0;
    auto factor_of = [&](int j) {  // the factor this warp evaluates at visiting position j
        if (j >= ord.n) return -1;
        const int fl = ((int)blockIdx.x + ord.at(j) * (int)gridDim.x) * FT + warp;
        return fl < P.count ? fl : -1;
    };
    auto fetch_ids = [&](int i) {
        const int fl = factor_of(i);
        return fl >= 0 ? __ldg(reinterpret_cast<const int2*>(table + fl)) : make_int2(0, 0);
    };
    auto issue = [&](int i, int s, int2 ids) {  // one lane: bulk copies of the factor of position i into stage s
        if (i >= ord.n) return;
        const int fl = factor_of(i);
        if (fl < 0) {  // this warp has no factor in that tile (the partial last tile): complete the stage's phase as is
            mbar_arrive(&bar[s]);
            return;
        }
        unsigned char* st = slots + (size_t)s * L.bytes;
        fence_proxy_async();
        mbar_arrive_expect_tx(&bar[s], (uint32_t)((int)sizeof(Row) + L.b0 + L.b1 + L.mb));
        tma_load_1d(st + L.row_off, table + fl, (uint32_t)sizeof(Row), &bar[s]);
        tma_load_1d(st + L.v0_off, P.v0 + (size_t)ids.x * L.b0, (uint32_t)L.b0, &bar[s]);
        if (Fam::D1) tma_load_1d(st + L.v1_off, P.v1 + (size_t)ids.y * L.b1, (uint32_t)L.b1, &bar[s]);
        if (!kSample)
            tma_load_1d(st + L.meas_off, P.meas + (size_t)(P.first + fl) * Fam::DM * P.Npad, (uint32_t)L.mb, &bar[s]);
    };
    // prologue: lane s < S fetches the ids of factor s and fills stage s (all stages in flight after one load latency)
    int2 ids = make_int2(0, 0);
    if (lane < S) ids = fetch_ids(lane);
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bar[s], 1);
        fence_mbar_init();
    }
    __syncwarp();
    if (!(P.flags & ROME_B200_INDEPENDENT)) asm volatile("griddepcontrol.wait;" ::: "memory");  // see eval_kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    bool synced = !kMulti || !(P.flags & ROME_B200_BARRIER_WAIT);  // rank barrier still to be passed (every warp fetches for itself)
    auto sync_before = [&](int i) {
        const int fl = factor_of(i);
        if (!synced && fl >= 0 && P.first + fl >= P.bar_lo && P.first + fl < P.bar_hi) {
            fused_barrier_wait(P, lane);
            synced = true;
        }
    };
    if (kMulti)
        for (int i = 0; i < S; ++i) sync_before(i);
    if (lane < S) issue(lane, lane, ids);

    int s = 0;
    uint32_t phase = 0;
    bool wrote_peer = false;
    int after_cut = 0;
    const int sig_pos = sig_early && ord.nc > 0 ? ord.signal_pos() : -1;
    const int cut_end = ord.lead() + ord.nc;
    for (int i = 0; i < ord.n; ++i) {
        const int fl = factor_of(i);
        const int f = P.first + fl;
        unsigned long long fdst = 0;  // owner-sharded exchange: per-factor destination of the forward row
        if (kMulti && fl >= 0 && ((flags & ROME_B200_PROPOSAL_FWD) || kRouted) && P.fwd_dst && f >= P.fwd_dst_lo &&
            f < P.fwd_dst_hi)
            fdst = __ldg(P.fwd_dst + f);
        if (lane == s) ids = fetch_ids(i + S);  // consumed when this factor is done: hidden behind its arithmetic
        mbar_wait(&bar[s], phase);
        if (fl >= 0) {
        const unsigned char* st = slots + (size_t)s * L.bytes;
        const Row row = *reinterpret_cast<const Row*>(st + L.row_off);
        FactorView V;
        V.b0 = st + L.v0_off;
        V.b1 = Fam::D1 ? st + L.v1_off : nullptr;
        V.b2 = nullptr;  // families with a third variable run the CTA pipeline (kWarpFT = 0)
        V.meas = kSample ? nullptr : reinterpret_cast<const float*>(st + L.meas_off);
        V.out_res = out;
        V.out_fwd = out + res_floats;
        V.fwd_on = !kMulti || !(P.flags & ROME_B200_ROUTED_ONLY) || fdst != 0;
        if (kMulti) wrote_peer = wrote_peer || fdst != 0 || P.n_peers > 0;
        if (flags & (ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD)) {
            if (lane == 0) tma_store_wait_read();  // the previous factor's rows have left the slice
            __syncwarp();
        }
        const bool routed = kRouted && fdst != 0;  // warp-uniform
        if (routed) Fam::template factor<kHot2 | (kStatic & ROME_B200_SAMPLE), kSample>(row, P, V, f, lane);
        else Fam::template factor<kStatic, kSample>(row, P, V, f, lane);
        if (flags & (ROME_B200_RESIDUAL | ROME_B200_PROPOSAL_FWD)) {
            fence_proxy_async();  // generic-proxy writes of the slice -> visible to the bulk-copy engine
            __syncwarp();
            if (lane == 0) {
                if (flags & ROME_B200_RESIDUAL)
                    tma_store_1d(P.res + (size_t)f * res_floats, V.out_res, (uint32_t)(res_floats * 4));
                if (((flags & ROME_B200_PROPOSAL_FWD) && V.fwd_on) || routed) {
                    const size_t off = (size_t)f * Fam::DFWD * P.Npad;
                    const uint32_t bytes = (uint32_t)(Fam::DFWD * P.Npad * 4);
                    tma_store_1d((kMulti && fdst) ? reinterpret_cast<float*>(fdst) : P.prop_fwd + off, V.out_fwd, bytes);
                    if (kMulti)
                        for (int r = 0; r < P.n_peers; ++r) tma_store_1d(P.peer_fwd[r] + off, V.out_fwd, bytes);
                }
                tma_store_commit();
                if (kMulti && i >= cut_end) ++after_cut;
            }
        }
        }
        __syncwarp();  // every lane has finished reading stage s
        // early signal: every warp of a CTA with peer-dependent tiles is counted (FT warps per such CTA)
        if (kMulti && i == sig_pos) {
            if (lane == 0) tma_store_wait_pending(after_cut);
            early_signal_publish(P, lane, (uint32_t)(ord.cut_ctas * FT));
        }
        if (kMulti) sync_before(i + S);
        if (lane == s) issue(i + S, s, ids);
        if (++s == S) { s = 0; phase ^= 1u; }
    }
    if (lane == 0) {
        tma_store_wait_all();
        if (kMulti && (P.flags & ROME_B200_BARRIER_SIGNAL) && !sig_early && wrote_peer) __threadfence_system();
    }
    if (P.flags & ROME_B200_INDEPENDENT) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (kMulti && (P.flags & ROME_B200_BARRIER_SIGNAL) && !sig_early) fused_barrier_signal(P);
}

template <class K>
int launch_kernel_cfg(K k, int* configured, int threads, const EvalParams& p, const LaunchPlan& plan, int grid,
                      cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || plan.smem_bytes > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem_bytes);
        if (e != cudaSuccess) return (int)e;
        // shared memory and L1 share one array: ask for no more carve-out than the resident CTAs need, the rest stays L1
        // (it backs the kernels' local memory -- loop state spilled around the factor body)
        const int pct = (int)((100LL * plan.ctas_per_sm * (plan.smem_bytes + 1024) + 233471) / 233472);
        cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
        if (dev >= 0 && dev < 64) configured[dev] = plan.smem_bytes;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, k, p);
}
template <class Fam, uint32_t kStatic, bool kSample, int FT, bool kRouted = false>
int launch_ft(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    static int configured[64] = {0};  // per-instantiation, per-device cache of the opt-in shared memory size
    return launch_kernel_cfg(eval_kernel<Fam, kStatic, kSample, FT, kRouted>, configured, eval_threads<FT>(), p, plan, grid, s);
}
template <class Fam, uint32_t kStatic, bool kSample, int FT, bool kRouted = false>
int launch_ft_w(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    static int configured[64] = {0};
    return launch_kernel_cfg(eval_kernel_w<Fam, kStatic, kSample, FT, kRouted>, configured, FT * 32, p, plan, grid, s);
}
template <class Fam, bool kSample>
int launch_sample(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    constexpr uint32_t smp = kSample ? ROME_B200_SAMPLE : 0u;
    if constexpr (Fam::kWarpFT > 0) {  // families with a per-warp-pipeline build (Fam::kWarpFT warps per CTA)
        if (plan.pipeline == 1) {
            constexpr int W = Fam::kWarpFT;
            if (plan.ft != W) return (int)cudaErrorInvalidValue;
            if (plan.variant == 1) return launch_ft_w<Fam, kHot1 | smp, kSample, W>(p, plan, grid, s);
            if (plan.variant == 2) return launch_ft_w<Fam, kHot2 | smp, kSample, W>(p, plan, grid, s);
            if constexpr (Fam::DFWD > 0) {
                if (plan.variant == 3) return launch_ft_w<Fam, kHot1 | smp, kSample, W, true>(p, plan, grid, s);
            }
            return launch_ft_w<Fam, 0u, kSample, W>(p, plan, grid, s);
        }
    } else if (plan.pipeline == 1) {
        return (int)cudaErrorInvalidValue;
    }
    if (plan.ft == 8) {
        if (plan.variant == 1) return launch_ft<Fam, kHot1 | smp, kSample, 8>(p, plan, grid, s);
        if (plan.variant == 2) return launch_ft<Fam, kHot2 | smp, kSample, 8>(p, plan, grid, s);
        if constexpr (Fam::DFWD > 0) {
            if (plan.variant == 3) return launch_ft<Fam, kHot1 | smp, kSample, 8, true>(p, plan, grid, s);
        }
        return launch_ft<Fam, 0u, kSample, 8>(p, plan, grid, s);
    }
    if constexpr (Fam::kMinCtas == 2) {
        if (plan.ft == 4) {
            if (plan.variant == 1) return launch_ft<Fam, kHot1 | smp, kSample, 4>(p, plan, grid, s);
            if (plan.variant == 2) return launch_ft<Fam, kHot2 | smp, kSample, 4>(p, plan, grid, s);
            if constexpr (Fam::DFWD > 0) {
                if (plan.variant == 3) return launch_ft<Fam, kHot1 | smp, kSample, 4, true>(p, plan, grid, s);
            }
            return (int)cudaErrorInvalidValue;
        }
    }
    if (plan.ft == 2) return launch_ft<Fam, 0u, kSample, 2>(p, plan, grid, s);
    if (plan.ft == 1) return launch_ft<Fam, 0u, kSample, 1>(p, plan, grid, s);
    return (int)cudaErrorInvalidValue;
}
template <class Fam>
int launch_family(const EvalParams& p, const LaunchPlan& plan, int grid, cudaStream_t s) {
    return (p.flags & ROME_B200_SAMPLE) ? launch_sample<Fam, true>(p, plan, grid, s)
                                        : launch_sample<Fam, false>(p, plan, grid, s);
}

}  // namespace rome
