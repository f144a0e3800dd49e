// sm_common.cuh -- complex helpers, block reductions and the "last block finishes the sum"
// protocol shared by every kernel of libschwinger_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sm {

typedef double2 cplx;   // (re, im): same bytes as the reference's std::complex<double>

__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// conj(a) * b
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
// a * conj(b)
__device__ __forceinline__ cplx cmul_conj(cplx a, cplx b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cscale(double s, cplx a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
// a / b (plain formula; the reference goes through libgcc's __divdc3, equal to rounding)
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
    const double inv = 1.0 / (b.x * b.x + b.y * b.y);
    return make_double2((a.x * b.x + a.y * b.y) * inv, (a.y * b.x - a.x * b.y) * inv);
}

// single-precision twins (the opt-in mixed-precision solver stores its inner vectors as float2)
typedef float2 cplxf;
__device__ __forceinline__ cplxf cmul(cplxf a, cplxf b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplxf cmulc(cplxf a, cplxf b) { return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ cplxf cmul_conj(cplxf a, cplxf b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
__device__ __forceinline__ cplxf cadd(cplxf a, cplxf b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplxf csub(cplxf a, cplxf b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplxf cscale(float s, cplxf a) { return make_float2(s * a.x, s * a.y); }

// scalar type of a complex type, and a constructor that works for both
template <typename C> struct RealOf;
template <> struct RealOf<cplx> { typedef double type; };
template <> struct RealOf<cplxf> { typedef float type; };
template <typename C>
__device__ __forceinline__ C mkc(typename RealOf<C>::type x, typename RealOf<C>::type y) {
    C r;
    r.x = x;
    r.y = y;
    return r;
}

// 16-byte global accesses.  Streaming data that no other thread of the SM re-reads goes past L1.
__device__ __forceinline__ cplx ldg(const cplx* p) { return __ldg(p); }
__device__ __forceinline__ cplx ld_stream(const cplx* p) {
    cplx r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(cplx* p, cplx v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

__device__ __forceinline__ cplxf ld_stream(const cplxf* p) {
    cplxf r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(cplxf* p, cplxf v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// Programmatic dependent launch (opt-in, SM_PDL=1; sm_cg.cuh): the two kernels of a one-pass CG iteration are launched
// with cudaLaunchAttributeProgrammaticStreamSerialization, so the blocks of the next kernel may become resident -- and
// run their prologue -- while the previous kernel drains.  `pdl_wait` returns once every grid this one depends on has
// completed and its memory is visible: it stands before the first read of anything a predecessor wrote and before the
// first global write.  Both are no-ops in a kernel that was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

constexpr int kBlock = 256;          // threads per block for every kernel
constexpr int kWarps = kBlock / 32;
constexpr int kMaxSums = 4;          // doubles reduced per kernel (<= 2 complex numbers)
constexpr int kWilsonBoundaryBlocks = 64;   // most blocks k_wilson_boundary is launched with

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic grid reduction of NS doubles per thread.
//   1. warp shuffle -> shared -> warp 0: one partial per block, stored at partials[NS*block + j]
//   2. the block that takes the last ticket re-reads every partial in a fixed order and writes
//      result[j].  Blocks/partials are bounded by the persistent grid (a few thousand at most).
// Returns true in ALL threads of the finishing block (result[] is then valid in thread 0 only
// through the returned values in `v`).
// A reduction may span several launches (interior + boundary blocks of one pass): they share the
// ticket, and pass the total block count and their global block index explicitly.
// `sys_scope`: the block's earlier stores include stores into a peer GPU's memory that the finishing block is about
// to release with a flag there, so the fence before the ticket must be system-wide.
template <int NS>
__device__ __forceinline__ bool grid_reduce(double (&v)[NS], double* __restrict__ partials, unsigned int* ticket,
                                            int nblocks = -1, int bid = -1, bool sys_scope = false) {
    __shared__ double s_part[kWarps][NS];
    __shared__ bool s_last;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int nthreads = blockDim.x * blockDim.y;      // a multiple of 32, at most kBlock
    const int nwarps = nthreads >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    if (nblocks < 0) {
        nblocks = gridDim.x * gridDim.y;
        bid = blockIdx.y * gridDim.x + blockIdx.x;
    }
#pragma unroll
    for (int j = 0; j < NS; j++) v[j] = warp_sum(v[j]);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NS; j++) s_part[warp][j] = v[j];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < NS; j++) {
            double x = (lane < nwarps) ? s_part[lane][j] : 0.0;
            x = warp_sum(x);
            if (lane == 0) partials[NS * bid + j] = x;
        }
        if (lane == 0) {
            if (sys_scope) __threadfence_system();
            else __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            s_last = (t == (unsigned int)(nblocks - 1));
        }
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
#pragma unroll
    for (int j = 0; j < NS; j++) {
        double x = 0.0;
        for (int b = tid; b < nblocks; b += nthreads) x += __ldcg(&partials[NS * b + j]);
        v[j] = warp_sum(x);
    }
    __syncthreads();   // s_part reuse
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NS; j++) s_part[warp][j] = v[j];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int j = 0; j < NS; j++) {
            double x = (lane < nwarps) ? s_part[lane][j] : 0.0;
            v[j] = warp_sum(x);
        }
        if (lane == 0) *ticket = 0u;   // ready for the next launch
    }
    return true;
}

}  // namespace sm
