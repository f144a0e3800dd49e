// sm_column_cg.cuh -- the whole conjugate gradient of a MID-SIZE lattice in one cooperative launch, several sites
// per thread (lattices beyond one site per thread of a full grid: up to S x T x SMs sites, e.g. 512 x 512 =
// BASELINE configs[4] on 148 SMs with S T = 2048).  Same algorithm as sm_cluster_cg.cuh / src/conjugate_gradient.cpp:4-67.
//
// A thread owns S sites that are CONSECUTIVE IN x at one t (a column segment): thread c = g wt + t holds the rows
// g S ... g S + S - 1.  Then
//   * the x neighbours of all but the segment's two end rows are the thread's own slots (registers);
//   * the t neighbours are the adjacent threads' same slots: a warp shuffle, except for lanes 0 / 31 and the row
//     wrap.  Those few halves still travel through L2, fetched by ONE load per lane (lane q S + j loads slot j of
//     "port" q: the warp's <= 2 readers of +t halves and <= 2 readers of -t halves) and handed over through a
//     32-entry shared-memory line per warp, so the edge lanes do not serialise S round trips;
//   * L2 carries 2 row-end halves per thread and stencil instead of 8 S halves (512 x 512: ~50 instead of 256 B
//     per site and iteration).
// Where a site's state lives: d and the vector that crosses a barrier (t = D^dagger d, then A d, then the new r) in
// REGISTERS; r, x and the site's two links in SHARED MEMORY (each thread touches only its own slots, 16-byte accesses
// at consecutive addresses across a warp: a conflict-free register extension of 96 B per site).  The links are kept
// pre-multiplied by -1/2 and by the antiperiodic sign, so a stencil is  (m0 + 2) p + sum of hops.
// A stencil application runs in two passes so that no L2 latency is exposed: everything that comes from registers
// and shuffles first, then the (linear) contributions of the halves that came through L2.
// Three barriers per iteration as in the one-site kernel: a site publishes the halves of r_{k+1} before the |r|^2
// barrier and those of d_k after the previous one, and a reader forms beta halves(d_k) + halves(r_{k+1}) itself.
// Buffers of a.hop: 0 halves of t, 1 halves of r, 2 + (k & 1) halves of d_k.
#pragma once
#include "sm_cluster_cg.cuh"

namespace sm {

constexpr int kColsBufT = 0, kColsBufR = 1, kColsBufD = 2, kColsBuffers = 4;
// r, links, x: 96 B per site; one staging line of 32 halves per warp
constexpr size_t cols_smem_bytes(int slots, int threads) { return sizeof(cplx) * ((size_t)6 * slots + 1) * threads; }

__device__ __forceinline__ cplx shfl_c(cplx v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

template <int S, int T>
__global__ void __launch_bounds__(T, 1) k_cg_cols(const ResidentCgArgs a) {
    static_assert(4 * S <= 32, "one loader lane per (port, slot)");
    extern __shared__ double2 cols_smem[];
    cplx* const sr = cols_smem;          // r       [component][slot][thread]
    cplx* const su = sr + 2 * S * T;     // -1/2 (sign) links  [mu][slot][thread]
    cplx* const sx = su + 2 * S * T;     // x       [component][slot][thread]
    cplx* const stage = sx + 2 * S * T;  // [warp][32] halves fetched from L2 for the warp's edge lanes
    GridSync<T> comm(a.wsum, a.bar);
    const int tid = threadIdx.x, lane = tid & 31;
    const int V = a.V, wt = a.wt, wx = a.wx;
    const size_t Vs = (size_t)V;
    cplx* const hop = a.hop;
    const cplx zero = make_double2(0.0, 0.0);

    const int groups = (wx + S - 1) / S;
    const int c = (int)blockIdx.x * T + tid;
    const bool live = c < groups * wt;
    const int g = live ? c / wt : 0;
    const int t = live ? c - g * wt : 0;
    const int nslots = live ? min(S, wx - g * S) : 0;      // rows g S + j, j < nslots
    const int nb = g * S * wt + t;                         // site of slot j: nb + j wt
    // which of the thread's t neighbours are reached through L2 (the same for all its slots)
    const bool tp_l2 = live && (lane == 31 || t == wt - 1);
    const bool tm_l2 = live && (lane == 0 || t == 0);
    int m_xm = nb - wt, m_xp = nb + nslots * wt;           // beyond the segment's two ends
    if (m_xm < 0) m_xm += V;
    if (m_xp >= V) m_xp -= V;

    // loader role of this lane: port q = lane / S (0, 1: the warp's readers of +t halves; 2, 3: of -t halves)
    int ld_site = -1, ld_kind = 0;
    const cplx *my_tp, *my_tm;           // where this (edge) lane finds the halves fetched for it
    {
        const int m_tp0 = (t == wt - 1) ? nb - (wt - 1) : nb + 1;
        const int m_tm0 = (t == 0) ? nb + (wt - 1) : nb - 1;
        const unsigned mask_tp = __ballot_sync(0xffffffffu, tp_l2), mask_tm = __ballot_sync(0xffffffffu, tm_l2);
        const int q = lane / S, j = lane - q * S;
        unsigned m = (q < 2) ? mask_tp : mask_tm;
        if (q & 1) m &= m - 1;
        const int src = m ? __ffs(m) - 1 : 0;
        const int b_tp = __shfl_sync(0xffffffffu, m_tp0, src), b_tm = __shfl_sync(0xffffffffu, m_tm0, src);
        const int ns = __shfl_sync(0xffffffffu, nslots, src);
        if (q < 4 && m != 0 && j < ns) {
            ld_site = ((q < 2) ? b_tp : b_tm) + j * wt;
            ld_kind = (q < 2) ? 0 : 1;
        }
        const unsigned below = (1u << lane) - 1u;
        my_tp = stage + (tid - lane) + S * __popc(mask_tp & below);
        my_tm = stage + (tid - lane) + S * (2 + __popc(mask_tm & below));
    }

    auto link = [&](int mu, int j) { return su[(mu * S + j) * T + tid]; };

    // the halves of p that are read through L2: both row ends, and the t halves next to a lane-0/31 or wrapped reader
    auto publish = [&](auto hop_tag, int buf, const cplx (&p0)[S], const cplx (&p1)[S]) {
        using H = decltype(hop_tag);
        cplx* const h = hop + (size_t)(buf * 4) * Vs + nb;
        if (nslots == S) {
            __stcg(h + 2 * Vs, H::from_xp(p0[0], p1[0]));
            __stcg(h + 3 * Vs + (S - 1) * wt, cmulc(link(1, S - 1), H::from_xm(p0[S - 1], p1[S - 1])));
        } else if (nslots > 0) {         // a shorter last segment (select chain: no dynamic register index)
            cplx e_last = zero;
#pragma unroll
            for (int j = 0; j < S; j++) {
                const cplx e = cmulc(link(1, j), H::from_xm(p0[j], p1[j]));
                if (j < nslots) e_last = e;
            }
            __stcg(h + 2 * Vs, H::from_xp(p0[0], p1[0]));
            __stcg(h + 3 * Vs + (nslots - 1) * wt, e_last);
        }
        if (tm_l2) {
#pragma unroll
            for (int j = 0; j < S; j++)
                if (j < nslots) __stcg(h + j * wt, H::from_tp(p0[j], p1[j]));
        }
        if (tp_l2) {
#pragma unroll
            for (int j = 0; j < S; j++)
                if (j < nslots) __stcg(h + Vs + j * wt, cmulc(link(0, j), H::from_tm(p0[j], p1[j])));
        }
    };

    // p <- (m0 + 2) p + (hop terms), in place; routed(kind, site) is what was published for `site`
    auto apply = [&](auto hop_tag, auto&& routed, cplx (&p0)[S], cplx (&p1)[S]) {
        using H = decltype(hop_tag);
        constexpr double s = H::si;
        cplx v = zero, in_xm = zero, in_xp = zero;
        if (ld_site >= 0) v = routed(ld_kind, ld_site);
        if (nslots > 0) {
            in_xm = routed(3, m_xm);
            in_xp = routed(2, m_xp);
        }
        // pass 1: own value, t neighbours by shuffle, x neighbours from the thread's own slots (an empty slot is zero)
        cplx e_prev = zero;              // conj(link) x (-x half) of the row below, formed before that row was overwritten
#pragma unroll
        for (int j = 0; j < S; j++) {
            const cplx q0 = p0[j], q1 = p1[j];
            const cplx u0 = link(0, j), u1 = link(1, j);
            cplx h_tp = shfl_c(H::from_tp(q0, q1), lane + 1);
            cplx h_tm = shfl_c(cmulc(u0, H::from_tm(q0, q1)), lane - 1);
            if (tp_l2) h_tp = zero;
            if (tm_l2) h_tm = zero;
            cplx a0 = make_double2(a.mass * q0.x, a.mass * q0.y), a1 = make_double2(a.mass * q1.x, a.mass * q1.y);
            const cplx f = cmul(u0, h_tp);
            a0 = cadd(a0, f);
            a1 = make_double2(a1.x + s * f.x, a1.y + s * f.y);
            if (j + 1 < S) H::add_xp(cmul(u1, H::from_xp(p0[j + 1 < S ? j + 1 : j], p1[j + 1 < S ? j + 1 : j])), a0, a1);
            H::add_tm(h_tm, a0, a1);
            H::add_xm(e_prev, a0, a1);
            e_prev = cmulc(u1, H::from_xm(q0, q1));
            const bool on = j < nslots;
            p0[j] = on ? a0 : zero;
            p1[j] = on ? a1 : zero;
        }
        // pass 2: what came through L2
        if (ld_site >= 0) stage[tid] = v;
        __syncwarp();
        H::add_xm(in_xm, p0[0], p1[0]);                    // zero when the thread holds no site
        if (nslots == S) {
            H::add_xp(cmul(link(1, S - 1), in_xp), p0[S - 1], p1[S - 1]);
        } else {
#pragma unroll
            for (int j = 0; j < S; j++) {
                const cplx f = cmul(link(1, j), in_xp);
                const cplx fz = (j == nslots - 1) ? f : zero;
                H::add_xp(fz, p0[j], p1[j]);
            }
        }
        if (tp_l2) {
#pragma unroll
            for (int j = 0; j < S; j++) {
                if (j < nslots) {
                    const cplx f = cmul(link(0, j), my_tp[j]);
                    p0[j] = cadd(p0[j], f);
                    p1[j] = make_double2(p1[j].x + s * f.x, p1[j].y + s * f.y);
                }
            }
        }
        if (tm_l2) {
#pragma unroll
            for (int j = 0; j < S; j++)
                if (j < nslots) H::add_tm(my_tm[j], p0[j], p1[j]);
        }
    };
    auto from_buf = [&](int buf) {
        const cplx* const h = hop + (size_t)(buf * 4) * Vs;
        return [h, Vs](int kind, int site) { return __ldcg(h + (size_t)kind * Vs + site); };
    };

    cplx d0[S], d1[S], w0[S], w1[S];     // direction; the vector that crosses the next barrier (t, A d, r)

    // x = phi ; r = phi - D D^dagger phi ; d = r   (conjugate_gradient.cpp:16-24)
#pragma unroll
    for (int j = 0; j < S; j++) {
        const int n = nb + j * wt;
        cplx f0 = zero, f1 = zero, u0 = zero, u1 = zero;
        if (j < nslots) {
            const double sg = (t == wt - 1) ? -0.5 * a.sR_edge : -0.5;   // the host checks sR_edge == sL_edge (one tile)
            u0 = cscale(sg, a.U[n]);
            u1 = cscale(-0.5, a.U[Vs + n]);
            f0 = a.x0 ? a.x0[n] : a.phi[n];      // start vector: phi as the reference, or the caller's guess
            f1 = a.x0 ? a.x0[Vs + n] : a.phi[Vs + n];
        }
        su[(0 * S + j) * T + tid] = u0;
        su[(1 * S + j) * T + tid] = u1;
        sx[(0 * S + j) * T + tid] = f0;
        sx[(1 * S + j) * T + tid] = f1;
        w0[j] = f0;
        w1[j] = f1;
    }
    publish(Hop<true>{}, kColsBufD + 1, w0, w1);
    comm.barrier();
    apply(Hop<true>{}, from_buf(kColsBufD + 1), w0, w1);           // t = D^dagger phi
    publish(Hop<false>{}, kColsBufT, w0, w1);
    comm.barrier();
    apply(Hop<false>{}, from_buf(kColsBufT), w0, w1);              // D t
    double s2[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < S; j++) {
        cplx f0 = sx[(0 * S + j) * T + tid], f1 = sx[(1 * S + j) * T + tid];
        if (a.x0 != nullptr) {                                     // x_0 != phi: fetch phi again for r_0 and |phi|
            f0 = f1 = zero;
            if (j < nslots) {
                f0 = a.phi[nb + j * wt];
                f1 = a.phi[Vs + nb + j * wt];
            }
        }
        const cplx r0 = csub(f0, w0[j]), r1 = csub(f1, w1[j]);    // zero in an empty slot
        sr[(0 * S + j) * T + tid] = r0;
        sr[(1 * S + j) * T + tid] = r1;
        d0[j] = r0;
        d1[j] = r1;
        s2[0] += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y;
        s2[1] += r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y;
    }
    publish(Hop<true>{}, kColsBufD, d0, d1);
    comm.template sum<2>(1, s2);          // its barrier also completes the exchange of the halves of d_0
    const double phi_norm = sqrt(s2[0]);
    double rr = s2[1], beta = 0.0;

    int k = 0, converged = 0;
    while (k < a.max_iter) {
        if (k > 0) {                      // d_k = r_k + beta d_{k-1}   (:51-59); its halves are for iteration k + 1
#pragma unroll
            for (int j = 0; j < S; j++) {
                d0[j] = make_double2(d0[j].x * beta + w0[j].x, d0[j].y * beta + w0[j].y);
                d1[j] = make_double2(d1[j].x * beta + w1[j].x, d1[j].y * beta + w1[j].y);
            }
            publish(Hop<true>{}, kColsBufD + (k & 1), d0, d1);
        }
#pragma unroll
        for (int j = 0; j < S; j++) {
            w0[j] = d0[j];
            w1[j] = d1[j];
        }
        if (k == 0) {
            apply(Hop<true>{}, from_buf(kColsBufD), w0, w1);       // t = D^dagger d
        } else {
            const cplx* const hd = hop + (size_t)((kColsBufD + ((k - 1) & 1)) * 4) * Vs;
            const cplx* const hr = hop + (size_t)(kColsBufR * 4) * Vs;
            apply(Hop<true>{}, [hd, hr, Vs, beta](int kind, int site) {
                const cplx dk = __ldcg(hd + (size_t)kind * Vs + site), rk = __ldcg(hr + (size_t)kind * Vs + site);
                return make_double2(dk.x * beta + rk.x, dk.y * beta + rk.y);
            }, w0, w1);
        }
        publish(Hop<false>{}, kColsBufT, w0, w1);
        comm.barrier();
        apply(Hop<false>{}, from_buf(kColsBufT), w0, w1);          // Ad = D t
        double dAd[2] = {0.0, 0.0};       // alpha = r_norm2 / dot(d, Ad)   (:32-33)
#pragma unroll
        for (int j = 0; j < S; j++) {
            const cplx q0 = cmul_conj(d0[j], w0[j]), q1 = cmul_conj(d1[j], w1[j]);
            dAd[0] += q0.x + q1.x;
            dAd[1] += q0.y + q1.y;
        }
        comm.template sum<2>(0, dAd);
        const cplx alpha = cdiv(make_double2(rr, 0.0), make_double2(dAd[0], dAd[1]));
        double e2[1] = {0.0};
#pragma unroll
        for (int j = 0; j < S; j++) {     // x += alpha d ; r -= alpha Ad   (:34-41)
            const int i0 = (0 * S + j) * T + tid, i1 = (1 * S + j) * T + tid;
            sx[i0] = cadd(sx[i0], cmul(alpha, d0[j]));
            sx[i1] = cadd(sx[i1], cmul(alpha, d1[j]));
            const cplx r0 = csub(sr[i0], cmul(alpha, w0[j])), r1 = csub(sr[i1], cmul(alpha, w1[j]));
            sr[i0] = r0;
            sr[i1] = r1;
            w0[j] = r0;
            w1[j] = r1;
            e2[0] += r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y;
        }
        publish(Hop<true>{}, kColsBufR, w0, w1);
        comm.template sum<1>(1, e2);
        if (sqrt(e2[0]) < a.tol * phi_norm) {   // :45
            converged = 1;
            break;
        }
        beta = e2[0] / rr;
        rr = e2[0];
        k++;
    }

#pragma unroll
    for (int j = 0; j < S; j++) {
        if (j < nslots) {
            const int n = nb + j * wt;
            a.x[n] = sx[(0 * S + j) * T + tid];
            a.x[Vs + n] = sx[(1 * S + j) * T + tid];
        }
    }
    if (c == 0) {
        a.st->phi_norm2 = s2[0];
        a.st->rr[0] = rr;
        a.st->iters = k;
        a.st->converged = converged;
        a.st->done = 1;
    }
}

}  // namespace sm
