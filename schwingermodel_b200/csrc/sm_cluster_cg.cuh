// sm_cluster_cg.cuh -- the whole conjugate gradient of a SMALL lattice in one kernel launch.
//
// Lattices up to 16 x 256 = 4096 sites (64 x 64, BASELINE configs[0]) are latency-bound, not
// bandwidth-bound: a CG iteration touches 128 KiB per field, so what limits a kernel-per-pass
// design is the launch and dependency latency of ~4000 tiny kernels per HMC trajectory.  Here ONE
// thread-block cluster (<= 16 CTAs, one lattice site per thread) keeps x, r, d and the four links
// a site needs in REGISTERS for the entire solve; neighbours exchange pre-projected half-spinors
// through distributed shared memory (each site publishes 4 complex numbers, readers fetch them
// from the owning CTA with cluster.map_shared_rank), and the two global sums of an iteration are
// reduced through DSMEM as well.  An iteration costs 4 cluster barriers and no HBM traffic at all.
// The algorithm is exactly src/conjugate_gradient.cpp:4-67 (x0 = phi, complex alpha, recursive
// residual, ||r|| < tol ||phi||).
#pragma once
#include <cooperative_groups.h>

#include "sm_fused.cuh"

namespace sm {

namespace cgx = cooperative_groups;

constexpr int kClusterMaxCtas = 16;
constexpr int kClusterThreads = 256;

struct ClusterCgArgs {
    const cplx* U;
    const cplx* phi;
    cplx* x;
    int wx, wt, V;
    double mass;
    double sR_edge, sL_edge;
    double tol;
    int max_iter;
    CgState* st;
};

struct ClusterShared {
    double2 hop[2][4][kClusterThreads];                    // [buffer][kind][site slot]
    double wsum[2][2][kClusterThreads / 32];               // [slot][value][warp]
};

// sum over the whole cluster of up to two values per thread; one cluster barrier
template <int NV>
__device__ __forceinline__ void cluster_sum(cgx::cluster_group& cluster, ClusterShared* sh, int slot, double (&v)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = kClusterThreads / 32;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        v[j] = warp_sum(v[j]);
        if (lane == 0) sh->wsum[slot][j][warp] = v[j];
    }
    cluster.sync();
    const int nparts = (int)cluster.num_blocks() * W;     // <= 128 warp partials
#pragma unroll
    for (int j = 0; j < NV; j++) {
        double acc = 0.0;
        for (int p = lane; p < nparts; p += 32) {
            const double* remote = cluster.map_shared_rank(&sh->wsum[slot][j][0], p / W);
            acc += remote[p % W];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        v[j] = acc;                                       // identical in every thread of the cluster
    }
}

__global__ void __launch_bounds__(kClusterThreads, 1) k_cg_cluster(const ClusterCgArgs a) {
    cgx::cluster_group cluster = cgx::this_cluster();
    __shared__ ClusterShared sh;
    const int tid = threadIdx.x;
    const int n = (int)cluster.block_rank() * kClusterThreads + tid;   // this thread's site
    const bool active = n < a.V;
    const int wt = a.wt, wx = a.wx, V = a.V;

    // neighbours: owning CTA and slot there
    int m_tp = n, m_tm = n, m_xp = n, m_xm = n;
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
    // remote views of hop[0][kind][slot]; buffer 1 sits 4*kClusterThreads entries further
    const double2* q_tp = cluster.map_shared_rank(&sh.hop[0][0][0], m_tp / kClusterThreads) + (m_tp % kClusterThreads);
    const double2* q_tm = cluster.map_shared_rank(&sh.hop[0][1][0], m_tm / kClusterThreads) + (m_tm % kClusterThreads);
    const double2* q_xp = cluster.map_shared_rank(&sh.hop[0][2][0], m_xp / kClusterThreads) + (m_xp % kClusterThreads);
    const double2* q_xm = cluster.map_shared_rank(&sh.hop[0][3][0], m_xm / kClusterThreads) + (m_xm % kClusterThreads);
    constexpr int kBuf = 4 * kClusterThreads;

    const cplx zero = make_double2(0.0, 0.0);
    cplx u0 = zero, u1 = zero, f0 = zero, f1 = zero;
    if (active) {
        u0 = a.U[n];
        u1 = a.U[V + n];
        f0 = a.phi[n];
        f1 = a.phi[V + n];
    }

    // one stencil application: publish the four half-spinors of (p0,p1), barrier, gather
    auto publish = [&](auto hop_tag, int buf, cplx p0, cplx p1) {
        using H = decltype(hop_tag);
        sh.hop[buf][0][tid] = H::from_tp(p0, p1);
        sh.hop[buf][1][tid] = cmulc(u0, H::from_tm(p0, p1));
        sh.hop[buf][2][tid] = H::from_xp(p0, p1);
        sh.hop[buf][3][tid] = cmulc(u1, H::from_xm(p0, p1));
    };
    auto gather = [&](auto hop_tag, int buf, cplx p0, cplx p1, cplx& o0, cplx& o1) {
        using H = decltype(hop_tag);
        cplx a0, a1;
        H::add_tp(cscale(sR, cmul(u0, q_tp[buf * kBuf])), a0, a1);
        H::add_xp(cmul(u1, q_xp[buf * kBuf]), a0, a1);
        H::add_tm(cscale(sL, q_tm[buf * kBuf]), a0, a1);
        H::add_xm(q_xm[buf * kBuf], a0, a1);
        o0 = make_double2(a.mass * p0.x - 0.5 * a0.x, a.mass * p0.y - 0.5 * a0.y);
        o1 = make_double2(a.mass * p1.x - 0.5 * a1.x, a.mass * p1.y - 0.5 * a1.y);
    };
    // out = D D^dagger p   (two exchanges, two cluster barriers)
    auto dd = [&](cplx p0, cplx p1, cplx& o0, cplx& o1) {
        cplx t0, t1;
        publish(Hop<true>{}, 0, p0, p1);
        cluster.sync();
        gather(Hop<true>{}, 0, p0, p1, t0, t1);
        publish(Hop<false>{}, 1, t0, t1);
        cluster.sync();
        gather(Hop<false>{}, 1, t0, t1, o0, o1);
    };

    // x = phi ; r = phi - D D^dagger phi ; d = r   (conjugate_gradient.cpp:16-24)
    cplx x0 = f0, x1 = f1, r0, r1, d0, d1, A0, A1;
    dd(x0, x1, A0, A1);
    r0 = csub(f0, A0);
    r1 = csub(f1, A1);
    if (!active) r0 = r1 = zero;
    d0 = r0;
    d1 = r1;
    double s2[2] = {f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y,
                    r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y};
    cluster_sum<2>(cluster, &sh, 1, s2);
    const double phi_norm = sqrt(s2[0]);
    double rr = s2[1];

    int k = 0, converged = 0;
    while (k < a.max_iter) {
        dd(d0, d1, A0, A1);
        if (!active) A0 = A1 = zero;
        // alpha = r_norm2 / dot(d, Ad)
        const cplx q0 = cmul_conj(d0, A0), q1 = cmul_conj(d1, A1);
        double dAd[2] = {q0.x + q1.x, q0.y + q1.y};
        cluster_sum<2>(cluster, &sh, 0, dAd);
        const cplx alpha = cdiv(make_double2(rr, 0.0), make_double2(dAd[0], dAd[1]));
        x0 = cadd(x0, cmul(alpha, d0));
        x1 = cadd(x1, cmul(alpha, d1));
        r0 = csub(r0, cmul(alpha, A0));
        r1 = csub(r1, cmul(alpha, A1));
        double e2[1] = {r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y};
        cluster_sum<1>(cluster, &sh, 1, e2);
        if (sqrt(e2[0]) < a.tol * phi_norm) {
            converged = 1;
            break;
        }
        const double beta = e2[0] / rr;
        d0 = make_double2(d0.x * beta + r0.x, d0.y * beta + r0.y);
        d1 = make_double2(d1.x * beta + r1.x, d1.y * beta + r1.y);
        rr = e2[0];
        k++;
    }

    if (active) {
        a.x[n] = x0;
        a.x[V + n] = x1;
    }
    if (n == 0) {
        a.st->phi_norm2 = s2[0];
        a.st->rr[0] = rr;
        a.st->iters = k;
        a.st->converged = converged;
        a.st->done = 1;
    }
    cluster.sync();   // nobody leaves while a neighbour may still read its shared memory
}

}  // namespace sm
