// sm_cluster_cg.cuh -- the whole conjugate gradient of a SMALL lattice in one kernel launch.
//
// Small lattices are latency-bound, not bandwidth-bound: a CG iteration on 64 x 64 touches 128 KiB
// per field, so what limits a kernel-per-pass design is the launch and dependency latency of
// ~4000 tiny kernels per HMC trajectory.  Here every lattice site is owned by one thread for the
// entire solve: r, d and the links the site needs stay in REGISTERS, neighbours exchange
// pre-projected half-spinors (each site publishes 4 complex numbers per stencil, the same rank-1
// trick as the halo exchange), and the two global sums of an iteration are reduced in-kernel.
// One iteration = 2 exchanges + 2 sums in 3 barriers (the exchange of r shares its barrier with the
// |r|^2 sum); nothing is re-read from HBM.
//
// Two transports for the exchange/barrier, same body:
//   ClusterComm : one thread-block cluster (<= 16 CTAs x 256 threads = 4096 sites, e.g. 64 x 64,
//                 BASELINE configs[0]); half-spinors and partial sums travel through DISTRIBUTED
//                 SHARED MEMORY (cluster.map_shared_rank), barrier = cluster.sync (~0.3 us).
//   GridComm    : a cooperative grid (one 512-thread CTA per SM, up to ~75 k sites, e.g. 256 x 256,
//                 BASELINE configs[1]); half-spinors travel through L2-resident global buffers,
//                 barrier = one release/acquire counter in L2 (cooperative launch for co-residency).
// The algorithm is exactly src/conjugate_gradient.cpp:4-67 (x0 = phi, complex alpha, recursive
// residual, ||r|| < tol ||phi||).
#pragma once
#include <cooperative_groups.h>

#include "sm_fused.cuh"

namespace sm {

namespace cgx = cooperative_groups;

constexpr int kClusterMaxCtas = 16;
constexpr int kClusterThreads = 256;
constexpr int kCoopThreads = 512;
constexpr int kGridSyncMaxCtas = 160;   // CTAs whose partial sums GridSync::sum_end adds (5 per lane)

struct ResidentCgArgs {
    const cplx* U;
    const cplx* phi;
    cplx* x;
    const cplx* x0;     // start vector (opt-in chronological guess); null: x_0 = phi as the reference (conjugate_gradient.cpp:16)
    int wx, wt, V;
    double mass;
    double sR_edge, sL_edge;
    double tol;
    int max_iter;
    CgState* st;
    // GridComm only
    cplx* hop;          // [buffer][4 kinds][V]: 2 buffers (k_cg_coop), 4 (k_cg_cols)
    double* wsum;       // [2 slots][2 values][blocks * warps]
    unsigned int* bar;  // monotonic arrival counter of the grid barrier (zeroed before the launch)
};

// ---- transport 1: thread-block cluster, distributed shared memory -------------------------------
struct ClusterShared {
    double2 hop[2][4][kClusterThreads];        // [buffer][kind][site slot]
    double wpart[2][kClusterThreads / 32];     // [value][warp]   scratch of the CTA-level step
    double wsum[2][2];                         // [slot][value]   this CTA's partial, read by the whole cluster
};

struct ClusterComm {
    static constexpr int kThreads = kClusterThreads;
    static constexpr bool kXInRegisters = true;      // 1 CTA per SM, registers to spare
    cgx::cluster_group cluster;
    ClusterShared* sh;
    const double2 *q_tp, *q_tm, *q_xp, *q_xm;   // remote views of hop[0][kind][neighbour slot]

    __device__ ClusterComm(ClusterShared* s) : cluster(cgx::this_cluster()), sh(s) {}
    __device__ int site() const { return (int)cluster.block_rank() * kThreads + (int)threadIdx.x; }
    __device__ void bind(int m_tp, int m_tm, int m_xp, int m_xm) {
        q_tp = cluster.map_shared_rank(&sh->hop[0][0][0], m_tp / kThreads) + (m_tp % kThreads);
        q_tm = cluster.map_shared_rank(&sh->hop[0][1][0], m_tm / kThreads) + (m_tm % kThreads);
        q_xp = cluster.map_shared_rank(&sh->hop[0][2][0], m_xp / kThreads) + (m_xp % kThreads);
        q_xm = cluster.map_shared_rank(&sh->hop[0][3][0], m_xm / kThreads) + (m_xm % kThreads);
    }
    __device__ void put(int buf, int kind, cplx v) { sh->hop[buf][kind][threadIdx.x] = v; }
    __device__ cplx get_tp(int buf) const { return q_tp[buf * 4 * kThreads]; }
    __device__ cplx get_tm(int buf) const { return q_tm[buf * 4 * kThreads]; }
    __device__ cplx get_xp(int buf) const { return q_xp[buf * 4 * kThreads]; }
    __device__ cplx get_xm(int buf) const { return q_xm[buf * 4 * kThreads]; }
    __device__ void barrier() { cluster.sync(); }

    // sum over all threads of the cluster, identical result everywhere: warp shuffle -> CTA partial in
    // shared memory -> [barrier] -> every warp adds the <= 16 CTA partials (one DSMEM load per lane).
    // Split in two so a caller can share the barrier with a half-spinor exchange.
    template <int NV>
    __device__ void sum_begin(int slot, double (&v)[NV]) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int W = kThreads / 32;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            v[j] = warp_sum(v[j]);
            if (lane == 0) sh->wpart[j][warp] = v[j];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < NV; j++) {
                double x = (lane < W) ? sh->wpart[j][lane] : 0.0;
                x = warp_sum(x);
                if (lane == 0) sh->wsum[slot][j] = x;
            }
        }
    }
    template <int NV>
    __device__ void sum_end(int slot, double (&v)[NV]) {
        const int lane = threadIdx.x & 31;
        const int nb = (int)cluster.num_blocks();
#pragma unroll
        for (int j = 0; j < NV; j++) {
            double acc = 0.0;
            if (lane < nb) acc = *cluster.map_shared_rank(&sh->wsum[slot][j], lane);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            v[j] = acc;
        }
    }
    template <int NV>
    __device__ void sum(int slot, double (&v)[NV]) {
        sum_begin<NV>(slot, v);
        cluster.sync();
        sum_end<NV>(slot, v);
    }
};

// ---- transport 2: cooperative grid, L2-resident global buffers ----------------------------------
// barrier and sums of a cooperative grid of THREADS-thread CTAs (one per SM)
template <int THREADS>
struct GridSync {
    double* wsum;
    unsigned int* bar;
    unsigned int target;                             // arrivals expected at the next barrier

    __device__ GridSync(double* wsum_, unsigned int* bar_) : wsum(wsum_), bar(bar_), target(0) {}
    // Grid barrier on one monotonic counter: thread 0 of every CTA arrives with a release and spins with acquire
    // loads until all CTAs of this round have arrived (the cooperative launch guarantees they are co-resident).
    // The release / acquire pair orders the whole CTA's traffic: the block barriers put every thread's stores
    // before thread 0's release and every thread's later loads after its acquire (causality order is transitive
    // across bar.sync), so no separate __threadfence is needed -- tools/barrier_bench.cu: 1.27 us against 1.55 us
    // with explicit fences and 1.24 us for cooperative_groups' grid.sync (148 CTAs), 2.3 / 2.7 / 2.6 us with the
    // half-spinor stores before and loads after.
    __device__ void barrier() {
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

    // warp shuffle -> CTA partial (shared) -> L2 -> [grid barrier] -> every warp adds the <= 160 CTA
    // partials with independent loads, in a fixed order
    template <int NV>
    __device__ void sum_begin(int slot, double (&v)[NV]) {
        __shared__ double wpart[2][THREADS / 32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int W = THREADS / 32;
        const int nb = (int)gridDim.x;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            v[j] = warp_sum(v[j]);
            if (lane == 0) wpart[j][warp] = v[j];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int j = 0; j < NV; j++) {
                double x = (lane < W) ? wpart[j][lane] : 0.0;
                x = warp_sum(x);
                if (lane == 0) __stcg(&wsum[(size_t)(slot * 2 + j) * nb + blockIdx.x], x);
            }
        }
    }
    template <int NV>
    __device__ void sum_end(int slot, double (&v)[NV]) {
        const int lane = threadIdx.x & 31;
        const int nb = (int)gridDim.x;
#pragma unroll
        for (int j = 0; j < NV; j++) {
            const double* part = wsum + (size_t)(slot * 2 + j) * nb;
            double p[5];
            static_assert(kGridSyncMaxCtas == 5 * 32, "five partials per lane");
#pragma unroll
            for (int i = 0; i < 5; i++) p[i] = (lane + 32 * i < nb) ? __ldcg(part + lane + 32 * i) : 0.0;
            double acc = ((p[0] + p[1]) + (p[2] + p[3])) + p[4];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            v[j] = acc;
        }
    }
    template <int NV>
    __device__ void sum(int slot, double (&v)[NV]) {
        sum_begin<NV>(slot, v);
        barrier();
        sum_end<NV>(slot, v);
    }
};

struct GridComm : GridSync<kCoopThreads> {
    static constexpr int kThreads = kCoopThreads;
    static constexpr bool kXInRegisters = false;     // 128-register budget: x is updated in place in L2
    cplx* hop;
    int V, n;
    int m_tp, m_tm, m_xp, m_xm;

    __device__ GridComm(const ResidentCgArgs& a)
        : GridSync<kCoopThreads>(a.wsum, a.bar), hop(a.hop), V(a.V), n((int)blockIdx.x * kThreads + (int)threadIdx.x) {}
    __device__ int site() const { return n; }
    __device__ void bind(int tp, int tm, int xp, int xm) { m_tp = tp; m_tm = tm; m_xp = xp; m_xm = xm; }
    __device__ void put(int buf, int kind, cplx v) {
        if (n < V) __stcg(&hop[(size_t)(buf * 4 + kind) * V + n], v);
    }
    __device__ cplx get_tp(int buf) const { return __ldcg(&hop[(size_t)(buf * 4 + 0) * V + m_tp]); }
    __device__ cplx get_tm(int buf) const { return __ldcg(&hop[(size_t)(buf * 4 + 1) * V + m_tm]); }
    __device__ cplx get_xp(int buf) const { return __ldcg(&hop[(size_t)(buf * 4 + 2) * V + m_xp]); }
    __device__ cplx get_xm(int buf) const { return __ldcg(&hop[(size_t)(buf * 4 + 3) * V + m_xm]); }
};

// ---- the solve ------------------------------------------------------------------------------------
template <class Comm>
__device__ __forceinline__ void resident_cg(const ResidentCgArgs& a, Comm& comm) {
    const int n = comm.site();
    const bool active = n < a.V;
    const int wt = a.wt, wx = a.wx, V = a.V;

    int m_tp = 0, m_tm = 0, m_xp = 0, m_xm = 0;
    double sR = 1.0, sL = 1.0;
    if (active) {
        const int x = n / wt, t = n - x * wt;
        m_tp = nb_tp(n, t, wt);
        m_tm = nb_tm(n, t, wt);
        m_xp = nb_xp(n, x, wx, wt);
        m_xm = nb_xm(n, x, wx, wt);
        sR = (t == wt - 1) ? a.sR_edge : 1.0;
        sL = (t == 0) ? a.sL_edge : 1.0;
    }
    comm.bind(m_tp, m_tm, m_xp, m_xm);

    const cplx zero = make_double2(0.0, 0.0);
    cplx u0 = zero, u1 = zero, f0 = zero, f1 = zero;
    if (active) {
        u0 = a.U[n];
        u1 = a.U[V + n];
        f0 = a.phi[n];
        f1 = a.phi[V + n];
    }

    // A stencil application = every site publishes four half-spinors of its spinor, a barrier, every site
    // fetches the four that point at it and combines them with its own value.
    struct Halves {
        cplx tp, xp, tm, xm;
    };
    auto publish = [&](auto hop_tag, int buf, cplx p0, cplx p1) {
        using H = decltype(hop_tag);
        comm.put(buf, 0, H::from_tp(p0, p1));
        comm.put(buf, 1, cmulc(u0, H::from_tm(p0, p1)));
        comm.put(buf, 2, H::from_xp(p0, p1));
        comm.put(buf, 3, cmulc(u1, H::from_xm(p0, p1)));
    };
    auto fetch = [&](int buf) {
        Halves h;
        h.tp = comm.get_tp(buf);
        h.xp = comm.get_xp(buf);
        h.tm = comm.get_tm(buf);
        h.xm = comm.get_xm(buf);
        return h;
    };
    auto combine = [&](auto hop_tag, const Halves& h, cplx p0, cplx p1, cplx& o0, cplx& o1) {
        using H = decltype(hop_tag);
        cplx a0, a1;
        H::add_tp(cscale(sR, cmul(u0, h.tp)), a0, a1);
        H::add_xp(cmul(u1, h.xp), a0, a1);
        H::add_tm(cscale(sL, h.tm), a0, a1);
        H::add_xm(h.xm, a0, a1);
        o0 = make_double2(a.mass * p0.x - 0.5 * a0.x, a.mass * p0.y - 0.5 * a0.y);
        o1 = make_double2(a.mass * p1.x - 0.5 * a1.x, a.mass * p1.y - 0.5 * a1.y);
    };
    // Buffer 0 carries the halves of psi / r, buffer 1 those of t; a buffer (and a sum slot) is rewritten
    // only after a later barrier than the one its readers waited on.

    // x = phi (or the caller's start vector) ; r = phi - D D^dagger x ; d = r   (conjugate_gradient.cpp:16-24).
    cplx g0 = f0, g1 = f1;
    if (a.x0 != nullptr && active) {
        g0 = a.x0[n];
        g1 = a.x0[V + n];
    }
    cplx r0, r1, d0, d1, A0, A1, t0, t1;
    publish(Hop<true>{}, 0, g0, g1);
    comm.barrier();
    combine(Hop<true>{}, fetch(0), g0, g1, t0, t1);
    publish(Hop<false>{}, 1, t0, t1);
    comm.barrier();
    combine(Hop<false>{}, fetch(1), t0, t1, A0, A1);
    r0 = csub(f0, A0);
    r1 = csub(f1, A1);
    if (!active) r0 = r1 = zero;
    d0 = r0;
    d1 = r1;
    cplx xr0 = g0, xr1 = g1;                            // x when it is kept in registers
    if (!Comm::kXInRegisters && active) {               // else x lives in L2: only ever updated in place
        a.x[n] = g0;
        a.x[V + n] = g1;
    }
    double s2[2] = {f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y,
                    r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y};
    // the same barrier serves the two sums and the exchange of the halves of d_0 = r_0
    comm.template sum_begin<2>(1, s2);
    publish(Hop<true>{}, 0, r0, r1);
    comm.barrier();
    comm.template sum_end<2>(1, s2);
    Halves Hd = fetch(0);                               // the neighbours' halves of d_k, kept across iterations
    const double phi_norm = sqrt(s2[0]);
    double rr = s2[1];

    // Three barriers per iteration: the halves of d_{k+1} = r_{k+1} + beta d_k are formed locally from the
    // exchanged halves of r_{k+1} (the projections are linear) and the kept halves of d_k, so the exchange
    // of r shares its barrier with the |r|^2 sum.
    int k = 0, converged = 0;
    while (k < a.max_iter) {
        combine(Hop<true>{}, Hd, d0, d1, t0, t1);       // t = D^dagger d
        publish(Hop<false>{}, 1, t0, t1);
        comm.barrier();
        combine(Hop<false>{}, fetch(1), t0, t1, A0, A1);   // Ad = D t
        if (!active) A0 = A1 = zero;
        // alpha = r_norm2 / dot(d, Ad)   (:33)
        const cplx q0 = cmul_conj(d0, A0), q1 = cmul_conj(d1, A1);
        double dAd[2] = {q0.x + q1.x, q0.y + q1.y};
        cplx x0 = zero, x1 = zero;                      // fetched here so the latency hides behind the sum's barrier
        if (!Comm::kXInRegisters && active) {
            x0 = a.x[n];
            x1 = a.x[V + n];
        }
        comm.template sum<2>(0, dAd);
        const cplx alpha = cdiv(make_double2(rr, 0.0), make_double2(dAd[0], dAd[1]));
        if (Comm::kXInRegisters) {                      // x += alpha d   (:34-36)
            xr0 = cadd(xr0, cmul(alpha, d0));
            xr1 = cadd(xr1, cmul(alpha, d1));
        } else if (active) {
            a.x[n] = cadd(x0, cmul(alpha, d0));
            a.x[V + n] = cadd(x1, cmul(alpha, d1));
        }
        r0 = csub(r0, cmul(alpha, A0));                 // r -= alpha Ad  (:37-41)
        r1 = csub(r1, cmul(alpha, A1));
        double e2[1] = {r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y};
        comm.template sum_begin<1>(1, e2);
        publish(Hop<true>{}, 0, r0, r1);
        comm.barrier();
        comm.template sum_end<1>(1, e2);
        if (sqrt(e2[0]) < a.tol * phi_norm) {           // :45
            converged = 1;
            break;
        }
        const double beta = e2[0] / rr;                 // :51-59
        const Halves Hr = fetch(0);
        Hd.tp = make_double2(Hd.tp.x * beta + Hr.tp.x, Hd.tp.y * beta + Hr.tp.y);
        Hd.xp = make_double2(Hd.xp.x * beta + Hr.xp.x, Hd.xp.y * beta + Hr.xp.y);
        Hd.tm = make_double2(Hd.tm.x * beta + Hr.tm.x, Hd.tm.y * beta + Hr.tm.y);
        Hd.xm = make_double2(Hd.xm.x * beta + Hr.xm.x, Hd.xm.y * beta + Hr.xm.y);
        d0 = make_double2(d0.x * beta + r0.x, d0.y * beta + r0.y);
        d1 = make_double2(d1.x * beta + r1.x, d1.y * beta + r1.y);
        rr = e2[0];
        k++;
    }

    if (Comm::kXInRegisters && active) {
        a.x[n] = xr0;
        a.x[V + n] = xr1;
    }
    if (n == 0) {
        a.st->phi_norm2 = s2[0];
        a.st->rr[0] = rr;
        a.st->iters = k;
        a.st->converged = converged;
        a.st->done = 1;
    }
    comm.barrier();   // nobody leaves while a neighbour may still read what it published
}

__global__ void __launch_bounds__(kClusterThreads, 1) k_cg_cluster(const ResidentCgArgs a) {
    __shared__ ClusterShared sh;
    ClusterComm comm(&sh);
    resident_cg(a, comm);
}

__global__ void __launch_bounds__(kCoopThreads, 1) k_cg_coop(const ResidentCgArgs a) {
    GridComm comm(a);
    resident_cg(a, comm);
}

}  // namespace sm
