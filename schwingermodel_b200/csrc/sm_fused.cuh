// sm_fused.cuh -- D D^dagger in ONE pass over HBM (temporal blocking of the two Wilson stencils).
//
// The two-pass form (k_wilson twice) moves 192 B per site-update because the intermediate
// t = D^dagger psi goes out to HBM and comes back.  Here a block owns a strip of columns
// (t direction) and marches down the x rows keeping, per thread (= per column),
//     psi rows j-2, j-1, j      t rows j-3, j-2, j-1      and the links of those rows
// in REGISTERS; the t-direction neighbours travel through a small double-buffered shared-memory
// line of pre-projected half-spinors (the same rank-1 trick the halo exchange uses), one
// __syncthreads per row.  Each step loads row j, forms t(j-1) = D^dagger psi and
// out(j-2) = D t, so psi and U are read once and out is written once: ~96 B per site-update
// plus the halo overhead (2 columns each side of a (BT-4)-wide strip, 4 rows per chunk).
//
// FUSED_CG additionally folds the CG vector updates that touch the same data into the pass
// (src/conjugate_gradient.cpp:32-59):  d_k = r_k + beta d_{k-1} is formed on load (also at the
// halo sites, from r and d_{k-1}), x += alpha_{k-1} d_{k-1} is applied where d_{k-1} is read, and
// dot(d_k, A d_k) is reduced in the epilogue.  With k_cg_resid (r -= alpha A d, |r|^2) one CG
// iteration moves 224 + 96 = 320 B per site instead of 512.
#pragma once
#include "sm_kernels.cuh"

namespace sm {

enum { FUSED_PLAIN = 0, FUSED_DOT = 1, FUSED_CG = 2 };

struct FusedArgs {
    const cplx* U;
    const cplx* in;      // psi (PLAIN/DOT) or d_{k-1} (CG)
    cplx* out;           // D D^dagger psi  (A d_k in CG)
    int wx, wt, V;
    int rows_per_block;
    int cols_per_strip;  // output columns per block (<= blockDim.x - 4)
    double mass;
    double sR_edge, sL_edge;
    double* partials;
    unsigned int* ticket;
    double* sums_out;
    // CG mode
    CgState* st;
    const cplx* r;       // r_k
    cplx* x;             // x_{k-1} -> x_k
    cplx* d_new;         // d_k
    int k;
    double tol;
};

// hop algebra shared by both operators: s = +1 for D^dagger, -1 for D (see k_wilson)
template <bool DAG>
struct Hop {
    static constexpr double s = DAG ? 1.0 : -1.0;
    // half-spinors as seen from the receiving site
    static __device__ __forceinline__ cplx from_tp(cplx p0, cplx p1) { return make_double2(p0.x + s * p1.x, p0.y + s * p1.y); }
    static __device__ __forceinline__ cplx from_xp(cplx p0, cplx p1) { return make_double2(p0.x + s * p1.y, p0.y - s * p1.x); }
    static __device__ __forceinline__ cplx from_tm(cplx p0, cplx p1) { return make_double2(p0.x - s * p1.x, p0.y - s * p1.y); }
    static __device__ __forceinline__ cplx from_xm(cplx p0, cplx p1) { return make_double2(p0.x - s * p1.y, p0.y + s * p1.x); }
    // accumulate the four hop terms v (already multiplied by the link and the sign)
    static __device__ __forceinline__ void add_tp(cplx v, cplx& a0, cplx& a1) { a0 = v; a1 = cscale(s, v); }
    static __device__ __forceinline__ void add_xp(cplx v, cplx& a0, cplx& a1) { a0 = cadd(a0, v); a1.x -= s * v.y; a1.y += s * v.x; }
    static __device__ __forceinline__ void add_tm(cplx v, cplx& a0, cplx& a1) { a0 = cadd(a0, v); a1.x -= s * v.x; a1.y -= s * v.y; }
    static __device__ __forceinline__ void add_xm(cplx v, cplx& a0, cplx& a1) { a0 = cadd(a0, v); a1.x += s * v.y; a1.y -= s * v.x; }
};

__device__ __forceinline__ int wrap_idx(int a, int n) {
    a %= n;
    return a < 0 ? a + n : a;
}

// shared line: 4 complex per column and parity (fwd/bwd half-spinors of psi row and of t row)
template <int MODE>
__global__ void __launch_bounds__(kBlock, 2) k_dd_fused(const FusedArgs a) {
    extern __shared__ double2 s_line[];   // [2 parities][4][BT]
    const int BT = blockDim.x;
    const int tid = threadIdx.x;
    const int wt = a.wt, wx = a.wx, V = a.V;

    double beta = 0.0;
    cplx alpha = make_double2(0.0, 0.0);
    bool first = true;
    if (MODE == FUSED_CG) {
        if (a.st->done) return;
        const int cur = a.k & 1;
        first = (a.k == 0);
        if (!first) {
            if (cg_converged(a.st, cur, a.tol)) {
                if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
                    a.st->iters = a.k - 1;
                    a.st->converged = 1;
                    a.st->done = 1;
                }
                return;
            }
            beta = a.st->rr[cur] / a.st->rr[cur ^ 1];
            alpha = make_double2(a.st->alpha[0], a.st->alpha[1]);
        }
    }

    const int tc = blockIdx.x * a.cols_per_strip - 2 + tid;    // unwrapped column of this thread
    const int t = wrap_idx(tc, wt);
    const bool col_active = (tid < a.cols_per_strip + 4) && (tc <= wt + 1);   // strip + 2 halo columns each side
    const bool col_owner = (tid >= 2) && (tid < a.cols_per_strip + 2) && (tc < wt);
    const double sR = (t == wt - 1) ? a.sR_edge : 1.0;
    const double sL = (t == 0) ? a.sL_edge : 1.0;
    const int xa = blockIdx.y * a.rows_per_block;
    const int xb = min(wx, xa + a.rows_per_block);
    const int tl = (tid == 0) ? 0 : tid - 1, tr = (tid == BT - 1) ? tid : tid + 1;

    const cplx* __restrict__ in0 = a.in;
    const cplx* __restrict__ in1 = a.in + V;
    const cplx* __restrict__ U0 = a.U;
    const cplx* __restrict__ U1 = a.U + V;

    const cplx zero = make_double2(0.0, 0.0);
    // psi rows (m2 = j-2, m1 = j-1), t rows (t3 = j-3, t2 = j-2), links
    cplx pm2_0 = zero, pm2_1 = zero, pm1_0 = zero, pm1_1 = zero;
    cplx t3_0 = zero, t3_1 = zero, t2_0 = zero, t2_1 = zero;
    cplx u0m2 = zero, u0m1 = zero, u1m3 = zero, u1m2 = zero, u1m1 = zero;
    double acc[2] = {0.0, 0.0};

    // row loader: psi (or d_k formed from r and d_{k-1}) and links of row j
    auto load_row = [&](int j, cplx& p0, cplx& p1, cplx& v0, cplx& v1) {
        if (!col_active) {
            p0 = p1 = v0 = v1 = zero;
            return;
        }
        const int x = wrap_idx(j, wx);
        const int n = x * wt + t;
        v0 = ldg(U0 + n);
        v1 = ldg(U1 + n);
        if (MODE != FUSED_CG) {
            p0 = ldg(in0 + n);
            p1 = ldg(in1 + n);
        } else {
            const cplx r0 = ldg(a.r + n), r1 = ldg(a.r + V + n);
            if (first) {
                p0 = r0;
                p1 = r1;
            } else {
                const cplx d0 = ldg(in0 + n), d1 = ldg(in1 + n);
                p0 = make_double2(d0.x * beta + r0.x, d0.y * beta + r0.y);
                p1 = make_double2(d1.x * beta + r1.x, d1.y * beta + r1.y);
                if (col_owner && j >= xa && j < xb) {       // x += alpha_{k-1} d_{k-1}
                    cplx x0 = a.x[n], x1 = a.x[V + n];
                    x0 = cadd(x0, cmul(alpha, d0));
                    x1 = cadd(x1, cmul(alpha, d1));
                    a.x[n] = x0;
                    a.x[V + n] = x1;
                }
            }
            if (col_owner && j >= xa && j < xb) {
                a.d_new[n] = p0;
                a.d_new[V + n] = p1;
            }
        }
    };

    cplx p0, p1, v0, v1;           // row j
    cplx np0, np1, nv0, nv1;       // row j+1 (prefetch)
    load_row(xa - 2, np0, np1, nv0, nv1);

    for (int j = xa - 2; j <= xb + 1; j++) {
        p0 = np0; p1 = np1; v0 = nv0; v1 = nv1;
        if (j < xb + 1) load_row(j + 1, np0, np1, nv0, nv1);

        // publish the t-direction half-spinors of psi row j-1 (for D^dagger) and t row j-2 (for D)
        double2* line = s_line + (j & 1) * 4 * BT;
        line[0 * BT + tid] = Hop<true>::from_tp(pm1_0, pm1_1);                    // read by column t-1
        line[1 * BT + tid] = cmulc(u0m1, Hop<true>::from_tm(pm1_0, pm1_1));      // read by column t+1
        line[2 * BT + tid] = Hop<false>::from_tp(t2_0, t2_1);
        line[3 * BT + tid] = cmulc(u0m2, Hop<false>::from_tm(t2_0, t2_1));
        __syncthreads();

        // t(j-1) = D^dagger psi at row j-1
        cplx tn0, tn1;
        {
            cplx a0, a1;
            Hop<true>::add_tp(cscale(sR, cmul(u0m1, line[0 * BT + tr])), a0, a1);
            Hop<true>::add_xp(cmul(u1m1, Hop<true>::from_xp(p0, p1)), a0, a1);
            Hop<true>::add_tm(cscale(sL, line[1 * BT + tl]), a0, a1);
            Hop<true>::add_xm(cmulc(u1m2, Hop<true>::from_xm(pm2_0, pm2_1)), a0, a1);
            tn0 = make_double2(a.mass * pm1_0.x - 0.5 * a0.x, a.mass * pm1_0.y - 0.5 * a0.y);
            tn1 = make_double2(a.mass * pm1_1.x - 0.5 * a1.x, a.mass * pm1_1.y - 0.5 * a1.y);
        }
        // out(j-2) = D t at row j-2
        if (j >= xa + 2 && col_owner) {
            cplx a0, a1;
            Hop<false>::add_tp(cscale(sR, cmul(u0m2, line[2 * BT + tr])), a0, a1);
            Hop<false>::add_xp(cmul(u1m2, Hop<false>::from_xp(tn0, tn1)), a0, a1);
            Hop<false>::add_tm(cscale(sL, line[3 * BT + tl]), a0, a1);
            Hop<false>::add_xm(cmulc(u1m3, Hop<false>::from_xm(t3_0, t3_1)), a0, a1);
            const cplx o0 = make_double2(a.mass * t2_0.x - 0.5 * a0.x, a.mass * t2_0.y - 0.5 * a0.y);
            const cplx o1 = make_double2(a.mass * t2_1.x - 0.5 * a1.x, a.mass * t2_1.y - 0.5 * a1.y);
            const int n = (j - 2) * wt + t;      // xa <= j-2 < xb: no wrap
            a.out[n] = o0;
            a.out[V + n] = o1;
            if (MODE != FUSED_PLAIN) {           // dot(psi, out) = sum psi conj(out)
                const cplx q0 = cmul_conj(pm2_0, o0), q1 = cmul_conj(pm2_1, o1);
                acc[0] += q0.x + q1.x;
                acc[1] += q0.y + q1.y;
            }
        }
        // rotate the windows
        t3_0 = t2_0; t3_1 = t2_1; t2_0 = tn0; t2_1 = tn1;
        pm2_0 = pm1_0; pm2_1 = pm1_1; pm1_0 = p0; pm1_1 = p1;
        u1m3 = u1m2; u1m2 = u1m1; u1m1 = v1;
        u0m2 = u0m1; u0m1 = v0;
    }

    if (MODE != FUSED_PLAIN) {
        if (grid_reduce<2>(acc, a.partials, a.ticket)) {
            if (tid == 0) {
                a.sums_out[0] = acc[0];
                a.sums_out[1] = acc[1];
            }
        }
    }
}

// r -= alpha A d ; |r|^2 ; alpha kept for the x update that the next fused pass applies
__global__ void __launch_bounds__(kBlock) k_cg_resid(CgState* st, int cur, cplx* __restrict__ r,
                                                     const cplx* __restrict__ Ad, int n_elems, double* partials,
                                                     unsigned int* ticket, double* sums_out) {
    if (st->done) return;
    const cplx alpha = cdiv(make_double2(st->rr[cur], 0.0), make_double2(st->dAd[0], st->dAd[1]));
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx av = ld_stream(Ad + i);
        cplx rv = r[i];
        rv = csub(rv, cmul(alpha, av));
        r[i] = rv;
        acc[0] += rv.x * rv.x + rv.y * rv.y;
    }
    if (grid_reduce<1>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            st->alpha[0] = alpha.x;
            st->alpha[1] = alpha.y;
            st->pending = 1;        // x still lacks alpha_k d_k
            st->pending_buf = cur;  // d_k lives in d buffer (k & 1)
        }
    }
}

// the x update the loop still owes when it stops: x += alpha_K d_K
__global__ void __launch_bounds__(kBlock) k_cg_flush_x(CgState* st, cplx* __restrict__ x, const cplx* __restrict__ d_buf0,
                                                       const cplx* __restrict__ d_buf1, int n_elems) {
    if (!st->pending) return;
    const cplx alpha = make_double2(st->alpha[0], st->alpha[1]);
    const cplx* __restrict__ d = st->pending_buf ? d_buf1 : d_buf0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        x[i] = cadd(x[i], cmul(alpha, ld_stream(d + i)));
    }
}

__global__ void k_cg_clear_pending(CgState* st) { st->pending = 0; }

}  // namespace sm
