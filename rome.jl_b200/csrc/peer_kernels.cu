// peer_kernels.cu -- stream-ordered barrier between the ranks of a multi-GPU job, carried by the GPUs themselves over
// NVLink peer memory (no NCCL kernel, no host round trip).  It closes the fused exchange of the factor kernels: their
// TMA bulk stores have already put the proposal rows into every peer's buffer; what is left is to tell the peers
// "everything of my step e has landed" and to learn the same from them.
//   signal: e = ++(*epoch); st.release.sys e -> the slot this rank owns in every peer's flag array
//   wait:   e = ++(*epoch); spin (ld.acquire.sys, nanosleep back-off) until every local slot >= e
// Both epochs live in device memory, so a captured CUDA graph can be replayed: each replay continues the sequence.
// One tiny CTA each: it co-resides with the persistent factor kernels instead of queueing behind them.
#include <cuda_runtime.h>
#include <stdint.h>

namespace rome {

struct PeerSlots {
    uint32_t* slot[8];
};

// Both kernels are launched with programmatic dependent launch: they are resident (one warp) while their predecessor
// drains, wait for its COMPLETION (griddepcontrol.wait: all its stores, peer stores included, are performed) and only
// then publish / poll; their own successor may be set up meanwhile (launch_dependents).
__global__ void peer_signal_kernel(PeerSlots peers, int n_peers, uint32_t* epoch) {
    __shared__ uint32_t e;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) e = ++(*epoch);
    __syncthreads();
    if ((int)threadIdx.x < n_peers) {
        __threadfence_system();  // everything this GPU stored before (the proposal rows) is visible before the flag
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.slot[threadIdx.x]), "r"(e) : "memory");
    }
}

// status[0] is set to 1 if the wait gave up (a peer never arrived): the host reads it with rome_b200_peer_status
__global__ void peer_wait_kernel(const uint32_t* flags, int n, uint32_t* epoch, uint32_t* status, long long max_cycles) {
    __shared__ uint32_t e;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (threadIdx.x == 0) e = ++(*epoch);
    __syncthreads();
    if ((int)threadIdx.x < n) {
        const long long t0 = clock64();
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + threadIdx.x) : "memory");
            if ((int32_t)(v - e) >= 0) break;  // wrap-safe "v >= e"
            if (clock64() - t0 > max_cycles) {
                atomicExch(status, 1u);
                break;
            }
            __nanosleep(64);
        }
    }
}

// signal + wait in ONE launch (a step's closing barrier costs one kernel start instead of two): lane r publishes the next
// signal epoch to peer r, then polls local slot r for the next wait epoch.  state = the rank's state words (slots [0, 8),
// word 8 / 9 signal / wait epoch, word 10 give-up status).
__global__ void peer_barrier_kernel(PeerSlots peers, int n_peers, uint32_t* state, long long max_cycles) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t es = 0, ew = 0;
    if (threadIdx.x == 0) {
        es = ++state[8];
        ew = ++state[9];
    }
    es = __shfl_sync(0xffffffffu, es, 0);  // one warp: the epochs travel by shuffle
    ew = __shfl_sync(0xffffffffu, ew, 0);
    if ((int)threadIdx.x < n_peers) {
        __threadfence_system();  // everything this GPU stored before (the proposal rows) is visible before the flag
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.slot[threadIdx.x]), "r"(es) : "memory");
        const long long t0 = clock64();
        uint32_t v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(state + threadIdx.x) : "memory");
            if ((int32_t)(v - ew) >= 0) break;  // wrap-safe "v >= ew"
            if (clock64() - t0 > max_cycles) {
                atomicExch(state + 10, 1u);
                break;
            }
            __nanosleep(64);
        }
    }
}

template <class K, class... A>
static int launch_one_warp_pdl(K k, void* stream, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int)cudaLaunchKernelEx(&cfg, k, args...);
}
int launch_peer_signal(uint32_t* const* slots, int n_peers, uint32_t* epoch, void* stream) {
    PeerSlots p = {};
    for (int i = 0; i < n_peers && i < 8; ++i) p.slot[i] = slots[i];
    return launch_one_warp_pdl(peer_signal_kernel, stream, p, n_peers, epoch);
}
int launch_peer_wait(const uint32_t* flags, int n, uint32_t* epoch, uint32_t* status, long long max_cycles, void* stream) {
    return launch_one_warp_pdl(peer_wait_kernel, stream, flags, n, epoch, status, max_cycles);
}

int launch_peer_barrier(uint32_t* const* slots, int n_peers, uint32_t* state, long long max_cycles, void* stream) {
    PeerSlots p = {};
    for (int i = 0; i < n_peers && i < 8; ++i) p.slot[i] = slots[i];
    return launch_one_warp_pdl(peer_barrier_kernel, stream, p, n_peers, state, max_cycles);
}

// Halo push (owner-sharded sweeps): copy whole particle blocks of variables this rank owns into the particle stores of
// the peers that read them -- one warp per block, 16-byte stores over NVLink peer memory.
__global__ void halo_push_kernel(const unsigned char* __restrict__ store, int block_bytes, int n,
                                 const int32_t* __restrict__ src_var, const unsigned long long* __restrict__ dst_blocks) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint4* src = reinterpret_cast<const uint4*>(store + (size_t)src_var[i] * block_bytes);
    uint4* dst = reinterpret_cast<uint4*>(dst_blocks[i]);
    for (int k = lane; k < block_bytes / 16; k += 32) dst[k] = src[k];
}
int launch_halo_push(const unsigned char* store, int block_bytes, int n, const int32_t* src_var,
                     const unsigned long long* dst_blocks, void* stream) {
    if (n <= 0) return 0;
    halo_push_kernel<<<(n + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(store, block_bytes, n, src_var, dst_blocks);
    return (int)cudaGetLastError();
}

}  // namespace rome
