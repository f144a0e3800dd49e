// Micro-benchmark of grid-wide barriers for the resident CG kernels (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/barrier_bench tools/barrier_bench.cu && tools/barrier_bench
// One 512-thread CTA per SM (cooperative launch), N barriers back to back, with and without the
// half-spinor traffic of a stencil phase (4 stores before, 4 neighbour loads after).  Prints ns per barrier.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

struct Args {
    unsigned int* bar;      // counter
    unsigned int* flags;    // one per CTA, 32 uints apart
    double2* hop;           // [2][4][V]
    double2* sink;
    int iters, traffic, V;
};

// V0: what k_cg_coop uses today
__device__ __forceinline__ void bar_counter_fenced(unsigned int* bar, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned int seen;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
}
// V1: release / acquire only
__device__ __forceinline__ void bar_counter(unsigned int* bar, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned int seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}
// V2: relaxed polling, one acquire fence at the end
__device__ __forceinline__ void bar_counter_relaxed(unsigned int* bar, unsigned int& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        unsigned int seen;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
// V3: one flag per CTA (no atomics): CTA b stores the epoch, warp 0 of every CTA polls all flags
__device__ __forceinline__ void bar_flags(unsigned int* flags, unsigned int& epoch) {
    __syncthreads();
    if (threadIdx.x < 32) {
        epoch++;
        if (threadIdx.x == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
        const int nb = gridDim.x;
        bool ok;
        do {
            ok = true;
            for (int b = threadIdx.x; b < nb; b += 32) {
                unsigned int seen;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + b) : "memory");
                ok = ok && (seen >= epoch);
            }
        } while (!__all_sync(0xffffffffu, ok));
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

template <int VAR>
__global__ void __launch_bounds__(512, 1) k_bench(Args a) {
    unsigned int target = 0, epoch = 0;
    const int n = blockIdx.x * 512 + threadIdx.x;
    const int V = a.V;
    int m[4] = {(n + 1) % V, (n + V - 1) % V, (n + 256) % V, (n + V - 256) % V};
    double2 acc = make_double2(0.0, 0.0);
    cg::grid_group grid = cg::this_grid();
    for (int it = 0; it < a.iters; it++) {
        const int buf = it & 1;
        if (a.traffic) {
            double2 v = make_double2(acc.x + it, acc.y);
            for (int k = 0; k < 4; k++) __stcg(a.hop + (size_t)(buf * 4 + k) * V + n, v);
        }
        if (VAR == 0) bar_counter_fenced(a.bar, target);
        if (VAR == 1) bar_counter(a.bar, target);
        if (VAR == 2) bar_counter_relaxed(a.bar, target);
        if (VAR == 3) bar_flags(a.flags, epoch);
        if (VAR == 4) grid.sync();
        if (a.traffic) {
            for (int k = 0; k < 4; k++) {
                const double2 h = __ldcg(a.hop + (size_t)(buf * 4 + k) * V + m[k]);
                acc.x += h.x;
                acc.y += h.y;
            }
        }
    }
    if (acc.x == 12345.678) a.sink[n] = acc;
}

template <int VAR>
static void run(const char* name, Args a, int blocks) {
    for (int traffic = 0; traffic < 2; traffic++) {
        a.traffic = traffic;
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaMemset(a.bar, 0, 128));
            CK(cudaMemset(a.flags, 0, sizeof(unsigned int) * 1024));
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0));
            CK(cudaEventCreate(&e1));
            void* params[] = {&a};
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((const void*)k_bench<VAR>, dim3(blocks), dim3(512), params, 0, 0));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        printf("%-28s blocks %3d traffic %d : %7.1f ns per barrier\n", name, blocks, traffic, best * 1e6 / a.iters);
    }
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    Args a{};
    a.iters = 2000;
    CK(cudaMalloc(&a.bar, 128));
    CK(cudaMalloc(&a.flags, sizeof(unsigned int) * 1024));
    for (int blocks : {sms, 128, 32}) {
        a.V = blocks * 512;
        CK(cudaMalloc(&a.hop, sizeof(double2) * 8 * a.V));
        CK(cudaMalloc(&a.sink, sizeof(double2) * a.V));
        CK(cudaMemset(a.hop, 0, sizeof(double2) * 8 * a.V));
        run<0>("counter, fences (current)", a, blocks);
        run<1>("counter, release/acquire", a, blocks);
        run<2>("counter, relaxed poll", a, blocks);
        run<3>("flag per CTA", a, blocks);
        run<4>("cooperative_groups grid.sync", a, blocks);
        CK(cudaFree(a.hop));
        CK(cudaFree(a.sink));
    }
    return 0;
}
