// sm_peer.cuh -- the collectives of the CG iteration done by the kernels themselves over NVLink peer memory
// (lattice split along x, one process per GPU, every rank's "window" mapped into every other rank with CUDA IPC).
//
// The reference ends every dot product with MPI_Allreduce (include/variables.h:181-192) and starts every stencil with
// blocking halo Send/Recv (src/dirac_operator.cpp:66-88).  Here, inside the CG loop,
//   * the block that finishes a kernel's grid reduction STORES the rank's partial sum into a slot of every rank's
//     window (an all-gather by peer stores) and releases it with an epoch; the kernels that consume the sum read
//     the slots of all ranks in rank order and add them -- every rank forms the same bits, no ncclAllReduce;
//   * the kernel that produces r (k_cg_resid) stores r's two boundary rows on each side straight into the
//     neighbours' ghost rows and the finishing block raises their ghost flags; the boundary bands of the next
//     D D^dagger pass wait on those flags inside the kernel -- no send/recv, no packing, no extra launch.
// Nothing in the loop is a library call, so a batch of iterations of a split lattice is one CUDA graph.
// Slots and ghost rows are double-buffered on the iteration parity: a rank can run at most one iteration ahead of a
// neighbour (it needs that neighbour's sum of the current iteration), so parity suffices.
#pragma once
#include "sm_common.cuh"

namespace sm {

constexpr int kMaxPeers = 8;        // ranks whose sums are gathered through peer memory (one NVSwitch node)

struct SumSlot {                    // 32 bytes: up to 2 doubles and the epoch that releases them
    double v[2];
    unsigned int epoch;
    unsigned int pad[3];
};

// window layout behind the ghost rows and the ghost flags (see sm_ops.cuh): slots[kind 2][parity 2][rank kMaxPeers]
__host__ __device__ inline int sum_slot_index(int kind, int parity, int rank) { return (kind * 2 + parity) * kMaxPeers + rank; }

struct DistLink {
    int on;                          // 0: single tile or NCCL path (sums go to CgState directly)
    int nranks, rank;
    SumSlot* mine;                   // my window's slots (peers store into it)
    SumSlot* peer[kMaxPeers];        // every rank's slots as seen from here (peer[rank] == mine)
    // ghost rows of r: [parity] destination in the -x neighbour's "hi" ghost / the +x neighbour's "lo" ghost,
    // layout [component][2 rows][wt]; the flags that release them there, and my own flags to wait on
    cplx* push_xm_hi[2];
    cplx* push_xp_lo[2];
    unsigned int* flag_xm;           // in the -x neighbour's window: its "hi" flag for r
    unsigned int* flag_xp;           // in the +x neighbour's window: its "lo" flag for r
    const unsigned int* my_flag_lo;  // raised by my -x neighbour
    const unsigned int* my_flag_hi;  // raised by my +x neighbour
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(double* p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// spin until *flag == epoch; a peer that never delivers (a dead rank) traps after ~20 s instead of hanging the GPU
__device__ __forceinline__ void spin_until(const unsigned int* flag, unsigned int epoch) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != epoch) {
        if (clock64() - t0 > 40000000000LL) __trap();
    }
}

__device__ __forceinline__ void st_relaxed_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Warp 0 of the finishing block (all 32 lanes; the sums are valid in lane 0): my partial sums -> slot (kind, parity,
// my rank) of every rank, lane p serving rank p, then ONE system-scope fence for the warp, then the epochs.  (A
// release store per peer would put a system fence in front of every flag: ~8 serial NVLink round trips.)
// `flag_a/flag_b` (may be null): two more flags to raise with the same epoch behind the same fence (ghost rows).
template <int NS>
__device__ __forceinline__ void publish_sums(const DistLink& dl, int kind, int parity, unsigned int epoch, const double (&v)[NS],
                                             unsigned int* flag_a = nullptr, unsigned int* flag_b = nullptr) {
    const int lane = threadIdx.x & 31;
    double w[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < NS && j < 2; j++) w[j] = __shfl_sync(0xffffffffu, v[j], 0);
    SumSlot* s = nullptr;
    if (lane < dl.nranks) {
        s = dl.peer[lane] + sum_slot_index(kind, parity, dl.rank);
#pragma unroll
        for (int j = 0; j < NS && j < 2; j++) st_relaxed_sys(&s->v[j], w[j]);
    }
    __threadfence_system();
    if (s != nullptr) st_relaxed_sys_u32(&s->epoch, epoch);
    if (lane == 30 && flag_a != nullptr) st_relaxed_sys_u32(flag_a, epoch);
    if (lane == 31 && flag_b != nullptr) st_relaxed_sys_u32(flag_b, epoch);
}

// Whole block: the global sums = slots of rank 0, 1, ... added in that order (the same bits on every rank).
// Lanes 0..nranks-1 of warp 0 wait for one rank each; the result reaches every thread through shared memory.
template <int NS>
__device__ __forceinline__ void gather_sums(const DistLink& dl, int kind, int parity, unsigned int epoch, double (&out)[NS]) {
    __shared__ double s_g[2];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid < 32) {
        double v0 = 0.0, v1 = 0.0;
        if (tid < dl.nranks) {
            const SumSlot* s = dl.mine + sum_slot_index(kind, parity, tid);
            spin_until(&s->epoch, epoch);
            v0 = ld_relaxed_sys(&s->v[0]);
            if (NS > 1) v1 = ld_relaxed_sys(&s->v[1]);
        }
        double a0 = 0.0, a1 = 0.0;
        for (int r = 0; r < dl.nranks; r++) {
            a0 += __shfl_sync(0xffffffffu, v0, r);
            if (NS > 1) a1 += __shfl_sync(0xffffffffu, v1, r);
        }
        if (tid == 0) {
            s_g[0] = a0;
            s_g[1] = a1;
        }
    }
    __syncthreads();
    out[0] = s_g[0];
    if (NS > 1) out[1] = s_g[1];
    __syncthreads();      // s_g may be reused by a second gather
}

// single-thread variant (stopping-rule kernel)
__device__ __forceinline__ double gather_sum1_thread(const DistLink& dl, int kind, int parity, unsigned int epoch) {
    double a = 0.0;
    for (int r = 0; r < dl.nranks; r++) {
        const SumSlot* s = dl.mine + sum_slot_index(kind, parity, r);
        spin_until(&s->epoch, epoch);
        a += ld_relaxed_sys(&s->v[0]);
    }
    return a;
}

}  // namespace sm
