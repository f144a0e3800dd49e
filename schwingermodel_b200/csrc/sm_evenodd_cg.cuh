// sm_evenodd_cg.cuh -- the even-odd (Schur complement) CG as ONE cooperative launch for lattices whose CG working set
// lives in L2 (opt-in solver SM_SOLVER_EVENODD).
//
// Solves  Dhat Dhat^dagger x = phi  on the even sites,  Dhat = m - (1/4m) H_eo H_oe  (m = m0 + 2, D = m - H/2), with the
// reference CG's start vector, recurrences and stopping rule (src/conjugate_gradient.cpp:4-67).  The graph-replayed form
// (dev_cg_eo: six launches per iteration on parity-masked full-lattice arrays) spends 43 us per iteration at 512^2 on launch
// latencies; here one 512-thread CTA per SM stays resident for the whole solve, every phase is a grid-stride loop over the
// COMPACT index of one parity (no idle lanes), phases are separated by the release/acquire grid barrier of GridSync, and the
// two sums of an iteration are reduced in a fixed order (warp -> CTA -> L2 -> every warp).  Fields live in global memory
// (L2-resident at these sizes); what another CTA wrote before a barrier is read with ld.cg.
// Phases of iteration k (6 grid barriers):
//     W_o  = (D^dagger d)_o                        | barrier
//     t_e  = m d_e - (1/m) (D^dagger W)_e          | barrier
//     W_o  = (D t)_o                               | barrier
//     Ad_e = m t_e - (1/m) (D W)_e ; dot(d, Ad)    | sum (barrier)
//     x += alpha d ; r -= alpha Ad ; |r|^2         | sum (barrier)      -> stopping rule
//     d = r + beta d                               | barrier
#pragma once
#include "sm_cluster_cg.cuh"

namespace sm {

constexpr int kEoThreads = 512;

struct EoCgArgs {
    const cplx* U;
    const cplx* phi;     // even-supported
    cplx* x;             // even-supported on return (the odd sites are zeroed)
    cplx *r, *d, *t, *W, *Ad;   // work fields (2V complex each)
    int wx, wt, V;
    double mass;
    double sR_edge, sL_edge;
    double tol;
    int max_iter;
    CgState* st;
    double* wsum;        // [2 slots][2 values][CTAs]
    unsigned int* bar;   // zeroed before the launch
};

__global__ void __launch_bounds__(kEoThreads, 1) k_cg_eo_coop(const EoCgArgs a) {
    GridSync<kEoThreads> gs(a.wsum, a.bar);
    const int wt = a.wt, V = a.V, ht = wt >> 1, H = V >> 1;
    const int gtid = (int)blockIdx.x * kEoThreads + (int)threadIdx.x, stride = (int)gridDim.x * kEoThreads;
    const double m = a.mass, im = 1.0 / a.mass;
    WilsonArgs w{};
    w.U = a.U;
    w.wx = a.wx;
    w.wt = wt;
    w.V = V;
    w.mass = m;
    w.sR_edge = a.sR_edge;
    w.sL_edge = a.sL_edge;

    // compact index q of parity p  ->  (x, t)
    auto coords = [&](int q, int p, int& x, int& t) {
        x = q / ht;
        t = 2 * (q - x * ht) + ((x + p) & 1);
    };
    // out_o = (D in)_o  or  (D^dagger in)_o  for an even-supported `in` (no self term on the odd sites)
    // (measured: 1024-thread CTAs, or two sites per trip for more loads in flight, are both slower than this plain form)
    auto hop_to_odd = [&](auto dag, const cplx* in, cplx* out) {
        constexpr bool DAG = decltype(dag)::value;
        w.in = in;
        for (int q = gtid; q < H; q += stride) {
            int x, t;
            coords(q, 1, x, t);
            cplx o0, o1;
            wilson_site<DAG, true, false>(w, x, t, o0, o1);
            const int n = x * wt + t;
            __stcg(out + n, o0);
            __stcg(out + V + n, o1);
        }
    };
    // out_e = m v_e - (1/m) (D W)_e  (W odd-supported); optionally the partial of dot(dotv, out)
    auto schur_finish = [&](auto dag, const cplx* v, const cplx* Wf, cplx* out, const cplx* dotv, double (&acc)[2]) {
        constexpr bool DAG = decltype(dag)::value;
        w.in = Wf;
        for (int q = gtid; q < H; q += stride) {
            int x, t;
            coords(q, 0, x, t);
            cplx o0, o1;
            wilson_site<DAG, true, false>(w, x, t, o0, o1);
            const int n = x * wt + t;
            const cplx v0 = __ldcg(v + n), v1 = __ldcg(v + V + n);
            const cplx r0 = make_double2(fma(m, v0.x, -im * o0.x), fma(m, v0.y, -im * o0.y));
            const cplx r1 = make_double2(fma(m, v1.x, -im * o1.x), fma(m, v1.y, -im * o1.y));
            __stcg(out + n, r0);
            __stcg(out + V + n, r1);
            if (dotv != nullptr) {
                const cplx d0 = __ldcg(dotv + n), d1 = __ldcg(dotv + V + n);
                const cplx p0 = cmul_conj(d0, r0), p1 = cmul_conj(d1, r1);
                acc[0] += p0.x + p1.x;
                acc[1] += p0.y + p1.y;
            }
        }
    };
    // Ad = Dhat Dhat^dagger v, dot(v, Ad) summed over the grid (4 barriers)
    auto apply = [&](const cplx* v, double (&dAd)[2]) {
        double none[2] = {0.0, 0.0};
        hop_to_odd(std::true_type{}, v, a.W);
        gs.barrier();
        schur_finish(std::true_type{}, v, a.W, a.t, nullptr, none);
        gs.barrier();
        hop_to_odd(std::false_type{}, a.t, a.W);
        gs.barrier();
        dAd[0] = dAd[1] = 0.0;
        schur_finish(std::false_type{}, a.t, a.W, a.Ad, v, dAd);
        gs.template sum<2>(0, dAd);
    };

    // x = phi ; d = phi   (conjugate_gradient.cpp:16)
    double s2[2] = {0.0, 0.0};
    for (int q = gtid; q < H; q += stride) {
        int x, t;
        coords(q, 0, x, t);
        const int n = x * wt + t;
        const cplx f0 = a.phi[n], f1 = a.phi[V + n];
        a.x[n] = f0;
        a.x[V + n] = f1;
        __stcg(a.d + n, f0);
        __stcg(a.d + V + n, f1);
        s2[0] += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y;
    }
    for (int q = gtid; q < H; q += stride) {     // x may come back holding a completed (even + odd) field of an earlier call
        int x, t;
        coords(q, 1, x, t);
        const int n = x * wt + t;
        a.x[n] = make_double2(0.0, 0.0);
        a.x[V + n] = make_double2(0.0, 0.0);
    }
    gs.barrier();
    double dAd[2];
    apply(a.d, dAd);
    // r = phi - A phi ; d = r   (:17-24)
    for (int q = gtid; q < H; q += stride) {
        int x, t;
        coords(q, 0, x, t);
        const int n = x * wt + t;
        const cplx f0 = a.phi[n], f1 = a.phi[V + n];
        const cplx A0 = __ldcg(a.Ad + n), A1 = __ldcg(a.Ad + V + n);
        const cplx r0 = csub(f0, A0), r1 = csub(f1, A1);
        a.r[n] = r0;
        a.r[V + n] = r1;
        __stcg(a.d + n, r0);
        __stcg(a.d + V + n, r1);
        s2[1] += r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y;
    }
    gs.template sum<2>(1, s2);             // its barrier also publishes d_0
    const double phi_norm = sqrt(s2[0]);
    double rr = s2[1];

    int k = 0, converged = 0;
    while (k < a.max_iter) {
        apply(a.d, dAd);
        const cplx alpha = cdiv(make_double2(rr, 0.0), make_double2(dAd[0], dAd[1]));      // :33
        double e2[1] = {0.0};
        for (int q = gtid; q < H; q += stride) {                                             // :34-41
            int x, t;
            coords(q, 0, x, t);
            const int n = x * wt + t;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int i = c * V + n;
                const cplx dv = __ldcg(a.d + i), Av = __ldcg(a.Ad + i);
                a.x[i] = cadd(a.x[i], cmul(alpha, dv));
                const cplx rv = csub(a.r[i], cmul(alpha, Av));
                a.r[i] = rv;
                e2[0] += rv.x * rv.x + rv.y * rv.y;
            }
        }
        gs.template sum<1>(1, e2);
        if (sqrt(e2[0]) < a.tol * phi_norm) {                                                 // :45
            converged = 1;
            break;
        }
        const double beta = e2[0] / rr;                                                       // :51-59
        for (int q = gtid; q < H; q += stride) {
            int x, t;
            coords(q, 0, x, t);
            const int n = x * wt + t;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int i = c * V + n;
                const cplx rv = a.r[i], dv = __ldcg(a.d + i);
                __stcg(a.d + i, make_double2(dv.x * beta + rv.x, dv.y * beta + rv.y));
            }
        }
        rr = e2[0];
        k++;
        gs.barrier();
    }
    if (gtid == 0) {
        a.st->phi_norm2 = s2[0];
        a.st->rr[0] = rr;
        a.st->iters = k;
        a.st->converged = converged;
        a.st->done = 1;
    }
}

}  // namespace sm
