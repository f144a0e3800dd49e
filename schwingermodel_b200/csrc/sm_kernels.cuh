// sm_kernels.cuh -- hand-written sm_100a kernels of the Schwinger-model HMC hot path.
//
// Memory-bound FP64 work on a 2-D lattice (96 B per stencil site against ~58 flop): no tensor
// cores.  What matters: 16-byte coalesced accesses along t (the fastest index), every operand
// read from HBM once per pass (neighbour re-reads are served by L1/L2), persistent grids sized
// from the SM count so block partial sums stay few and the last block can finish every
// reduction inside the producing kernel.
//
// Field layout in HBM: one allocation per field, component mu0 at
// elements [0,V) and mu1 at [V,2V), element = complex double (16 B) or double (8 B);
// site n = x*wt + t.  This is the reference's spinor{mu0,mu1} / re_field (include/variables.h:54-141).
#pragma once
#include "sm_common.cuh"
#include "sm_peer.cuh"

namespace sm {

// ----------------------------------------------------------------------------------------------
// Geometry: the arithmetic replacement of periodic_boundary() (include/dirac_operator.h:35-62).
// All kernels take neighbours from these four functions; k_tables emits them as tables so the
// test can compare them bit for bit with the reference's RightPB/LeftPB/x_1_t1/x1_t_1.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int nb_tp(int n, int t, int wt) { return (t == wt - 1) ? n - (wt - 1) : n + 1; }
__device__ __forceinline__ int nb_tm(int n, int t, int wt) { return (t == 0) ? n + (wt - 1) : n - 1; }
__device__ __forceinline__ int nb_xp(int n, int x, int wx, int wt) { return (x == wx - 1) ? n - (wx - 1) * wt : n + wt; }
__device__ __forceinline__ int nb_xm(int n, int x, int wx, int wt) { return (x == 0) ? n + (wx - 1) * wt : n - wt; }

__global__ void k_tables(int wx, int wt, double sR_edge, double sL_edge, int* RightPB, int* LeftPB, double* SignR,
                         double* SignL, int* x_1_t1, int* x1_t_1) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= wx * wt) return;
    const int x = n / wt, t = n - x * wt;
    const int tp = nb_tp(n, t, wt), tm = nb_tm(n, t, wt);
    RightPB[2 * n] = tp;
    RightPB[2 * n + 1] = nb_xp(n, x, wx, wt);
    LeftPB[2 * n] = tm;
    LeftPB[2 * n + 1] = nb_xm(n, x, wx, wt);
    x_1_t1[n] = nb_tp(nb_xm(n, x, wx, wt), t, wt);
    x1_t_1[n] = nb_tm(nb_xp(n, x, wx, wt), t, wt);
    SignR[2 * (2 * n)] = (t == wt - 1) ? sR_edge : 1.0;
    SignR[2 * (2 * n) + 1] = 0.0;
    SignR[2 * (2 * n + 1)] = 1.0;
    SignR[2 * (2 * n + 1) + 1] = 0.0;
    SignL[2 * (2 * n)] = (t == 0) ? sL_edge : 1.0;
    SignL[2 * (2 * n) + 1] = 0.0;
    SignL[2 * (2 * n + 1)] = 1.0;
    SignL[2 * (2 * n + 1) + 1] = 0.0;
}

// ----------------------------------------------------------------------------------------------
// CG scalars, device resident (src/conjugate_gradient.cpp:26-62 keeps them on the host).
// ----------------------------------------------------------------------------------------------
struct CgState {
    double phi_norm2;   // Re dot(phi,phi)
    double rr[2];       // r_norm2, ping-pong on the iteration parity
    double dAd[2];      // dot(d, Ad), complex
    double alpha[2];    // alpha of the last residual update (fused path: x still lacks alpha d)
    int done;           // sticky: set once the stopping rule fired
    int iters;          // the reference's k when it returned
    int converged;      // the reference's return value
    int pending;        // fused path: an x update is owed
    int pending_buf;    // ... with d in ping-pong buffer 0/1
    int k;              // iteration counter kept on the device (one-pass path: kernels are replayed from a CUDA graph)
    int max_iter;       // one-pass path: iteration limit of the running solve (kept here so graphs do not bake it)
    double tol;         // ... and its relative tolerance
    unsigned int epoch_base;   // split lattice with peer-memory sums: epochs of this solve are epoch_base + k (+1)
};

// ----------------------------------------------------------------------------------------------
// Wilson stencil  out = D in  /  out = D^dagger in   (src/dirac_operator.cpp:29-44, :253-268;
// SURVEY.md appendix B).  With s = -1 for D and +1 for D^dagger the four hops are
//   +t : h = psi0 + s psi1          a = sR U0(n) h               out0 += a   out1 += s a
//   +x : h = psi0 - s i psi1        b = U1(n) h                  out0 += b   out1 += s i b
//   -t : h = psi0 - s psi1          c = sL conj(U0(n-t)) h       out0 += c   out1 -= s c
//   -x : h = psi0 + s i psi1        d = conj(U1(n-x)) h          out0 += d   out1 -= s i d
//   out_a = (m0+2) psi_a - 1/2 out_a
// Tiles: blockDim = (TT, TX); blockIdx.x picks a strip of TT sites in t, blockIdx.y a run of
// `rows_per_block` consecutive x rows that the block walks TX rows at a time, so the x-neighbour
// rows it needs next are the ones it just brought into L1/L2.
//
// When the lattice is split over GPUs the off-tile neighbour of a boundary site is not a field
// element but one pre-projected complex number per site received into a ghost line (g_*), the
// same quantities the reference ships in TopRow/BottomRow/LeftCol/RightCol
// (src/dirac_operator.cpp:49-64); a null ghost pointer means "wrap locally".
// ----------------------------------------------------------------------------------------------
// WILSON_EO / WILSON_EO_DOT: building blocks of the even-odd (Schur complement) operator of the opt-in even-odd HMC,
// on full-lattice arrays that are zero on one parity:  out(n) = [parity(n) == eo_keep] (eo_self aux(n) + eo_hop (D in)(n)),
// zero elsewhere (sites of the other parity are not computed); _DOT adds the partial of dot(aux2, out).
enum { WILSON_PLAIN = 0, WILSON_DOT = 1, WILSON_CGINIT = 2, WILSON_EO = 3, WILSON_EO_DOT = 4 };

struct WilsonArgs {
    const cplx* U;
    const cplx* in;
    cplx* out;
    const cplx* aux;   // DOT: d ; CGINIT: phi
    const cplx* aux2;  // CGINIT: the start vector x_0 when it is not phi (opt-in chronological guess), else null
    cplx* r;           // CGINIT outputs
    cplx* d;
    cplx* x;
    int wx, wt, V;
    int rows_per_block;
    double mass;       // m0 + 2
    double sR_edge;    // sign of the +t hop at t = wt-1 (-1 when this tile holds the antiperiodic seam)
    double sL_edge;    // sign of the -t hop at t = 0
    const cplx* g_tp;  // [wx]  projected psi of the +t neighbour tile (column t=0 there)
    const cplx* g_tm;  // [wx]  conj(U0) * projected psi of the -t neighbour tile (column wt-1 there)
    const cplx* g_xp;  // [wt]  projected psi of the +x neighbour tile (row 0 there)
    const cplx* g_xm;  // [wt]  conj(U1) * projected psi of the -x neighbour tile (row wx-1 there)
    double* partials;
    unsigned int* ticket;
    double* sums_out;
    const int* done;   // CG early-out flag (null outside CG)
    int eo_keep;           // EO modes: parity (x + t) & 1 of the sites that are computed
    double eo_self, eo_hop;
    int interior_only;     // split lattice with overlap: skip the sites that read ghost lines ...
    int boundary_blocks;   // ... k_wilson_boundary (this many blocks) computes them and joins the reduction
    int interior_blocks;   // (boundary launch) blocks of the interior launch
};

// COH: the input field is being written by other blocks of the same (cooperative) kernel between grid barriers, so it is
// read with ld.cg (L2) instead of the non-coherent read-only path.  SELF = false: the site's own value is known to be zero
// (even-odd fields, output parity != input parity) and is not read: out = -1/2 sum(hops).
template <bool COH>
__device__ __forceinline__ cplx ld_in(const cplx* p) {
    return COH ? __ldcg(p) : __ldg(p);
}

template <bool DAG, bool COH = false, bool SELF = true>
__device__ __forceinline__ void wilson_site(const WilsonArgs& a, int x, int t, cplx& o0, cplx& o1) {
    constexpr double s = DAG ? 1.0 : -1.0;
    const int wt = a.wt, wx = a.wx, V = a.V;
    const int n = x * wt + t;
    const cplx* __restrict__ in0 = a.in;
    const cplx* __restrict__ in1 = a.in + V;
    const cplx* __restrict__ U0 = a.U;
    const cplx* __restrict__ U1 = a.U + V;

    cplx c0 = make_double2(0.0, 0.0), c1 = c0;
    if (SELF) {
        c0 = ld_in<COH>(in0 + n);
        c1 = ld_in<COH>(in1 + n);
    }
    const cplx u0 = ldg(U0 + n), u1 = ldg(U1 + n);
    cplx acc0, acc1;

    // +t
    {
        cplx h;
        if (t == wt - 1 && a.g_tp) {
            h = ldg(a.g_tp + x);
        } else {
            const int m = nb_tp(n, t, wt);
            const cplx p0 = ld_in<COH>(in0 + m), p1 = ld_in<COH>(in1 + m);
            h = make_double2(p0.x + s * p1.x, p0.y + s * p1.y);
        }
        cplx v = cmul(u0, h);
        if (t == wt - 1) v = cscale(a.sR_edge, v);
        acc0 = v;
        acc1 = cscale(s, v);
    }
    // +x
    {
        cplx h;
        if (x == wx - 1 && a.g_xp) {
            h = ldg(a.g_xp + t);
        } else {
            const int m = nb_xp(n, x, wx, wt);
            const cplx p0 = ld_in<COH>(in0 + m), p1 = ld_in<COH>(in1 + m);
            h = make_double2(p0.x + s * p1.y, p0.y - s * p1.x);
        }
        const cplx v = cmul(u1, h);
        acc0 = cadd(acc0, v);
        acc1.x -= s * v.y;
        acc1.y += s * v.x;
    }
    // -t
    {
        cplx v;
        if (t == 0 && a.g_tm) {
            v = ldg(a.g_tm + x);
        } else {
            const int m = nb_tm(n, t, wt);
            const cplx p0 = ld_in<COH>(in0 + m), p1 = ld_in<COH>(in1 + m);
            const cplx h = make_double2(p0.x - s * p1.x, p0.y - s * p1.y);
            v = cmulc(ldg(U0 + m), h);
        }
        if (t == 0) v = cscale(a.sL_edge, v);
        acc0 = cadd(acc0, v);
        acc1.x -= s * v.x;
        acc1.y -= s * v.y;
    }
    // -x
    {
        cplx v;
        if (x == 0 && a.g_xm) {
            v = ldg(a.g_xm + t);
        } else {
            const int m = nb_xm(n, x, wx, wt);
            const cplx p0 = ld_in<COH>(in0 + m), p1 = ld_in<COH>(in1 + m);
            const cplx h = make_double2(p0.x - s * p1.y, p0.y + s * p1.x);
            v = cmulc(ldg(U1 + m), h);
        }
        acc0 = cadd(acc0, v);
        acc1.x += s * v.y;
        acc1.y -= s * v.x;
    }
    o0 = make_double2(a.mass * c0.x - 0.5 * acc0.x, a.mass * c0.y - 0.5 * acc0.y);
    o1 = make_double2(a.mass * c1.x - 0.5 * acc1.x, a.mass * c1.y - 0.5 * acc1.y);
}

// what a stencil pass does with the site's result besides (or instead of) storing it
template <int MODE, int NS>
__device__ __forceinline__ void wilson_epilogue(const WilsonArgs& a, int n, cplx o0, cplx o1, double (&acc)[NS]) {
    if (MODE == WILSON_PLAIN) {
        a.out[n] = o0;
        a.out[a.V + n] = o1;
    } else if (MODE == WILSON_EO || MODE == WILSON_EO_DOT) {
        cplx v0 = cscale(a.eo_hop, o0), v1 = cscale(a.eo_hop, o1);
        if (a.aux != nullptr) {
            const cplx s0 = ld_stream(a.aux + n), s1 = ld_stream(a.aux + a.V + n);
            v0 = make_double2(fma(a.eo_self, s0.x, v0.x), fma(a.eo_self, s0.y, v0.y));
            v1 = make_double2(fma(a.eo_self, s1.x, v1.x), fma(a.eo_self, s1.y, v1.y));
        }
        a.out[n] = v0;
        a.out[a.V + n] = v1;
        if (MODE == WILSON_EO_DOT) {
            const cplx d0 = ld_stream(a.aux2 + n), d1 = ld_stream(a.aux2 + a.V + n);
            const cplx p0 = cmul_conj(d0, v0), p1 = cmul_conj(d1, v1);
            acc[0] += p0.x + p1.x;
            acc[NS - 1] += p0.y + p1.y;
        }
    } else if (MODE == WILSON_DOT) {
        // Ad = D t, partial of dot(d, Ad) = sum d conj(Ad)   (conjugate_gradient.cpp:32-33)
        a.out[n] = o0;
        a.out[a.V + n] = o1;
        const cplx d0 = ld_stream(a.aux + n), d1 = ld_stream(a.aux + a.V + n);
        const cplx p0 = cmul_conj(d0, o0), p1 = cmul_conj(d1, o1);
        acc[0] += p0.x + p1.x;
        acc[NS - 1] += p0.y + p1.y;
    } else {
        // x = phi ; r = phi - DD^dagger phi ; d = r   (conjugate_gradient.cpp:16-24)
        const cplx f0 = ld_stream(a.aux + n), f1 = ld_stream(a.aux + a.V + n);
        const cplx r0 = csub(f0, o0), r1 = csub(f1, o1);
        if (a.aux2 != nullptr) {
            a.x[n] = a.aux2[n];
            a.x[a.V + n] = a.aux2[a.V + n];
        } else {
            a.x[n] = f0;
            a.x[a.V + n] = f1;
        }
        a.r[n] = r0;
        a.r[a.V + n] = r1;
        a.d[n] = r0;
        a.d[a.V + n] = r1;
        acc[0] += f0.x * f0.x + f0.y * f0.y + f1.x * f1.x + f1.y * f1.y;        // |phi|^2
        acc[NS - 1] += r0.x * r0.x + r0.y * r0.y + r1.x * r1.x + r1.y * r1.y;   // |r|^2
    }
}

// a site whose stencil reads a ghost line (only on a split lattice)
__device__ __forceinline__ bool wilson_on_boundary(const WilsonArgs& a, int x, int t) {
    return (a.g_tp != nullptr && (t == 0 || t == a.wt - 1)) || (a.g_xp != nullptr && (x == 0 || x == a.wx - 1));
}

template <bool DAG, int MODE>
__global__ void __launch_bounds__(kBlock) k_wilson(const WilsonArgs a) {
    if (a.done != nullptr && *a.done) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool t_ok = t < a.wt;
    const int x_begin = blockIdx.y * a.rows_per_block;
    const int x_end = min(a.wx, x_begin + a.rows_per_block);
    constexpr int NS = (MODE == WILSON_PLAIN || MODE == WILSON_EO) ? 1 : 2;
    double acc[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) acc[j] = 0.0;

    if (t_ok) {
        for (int x = x_begin + threadIdx.y; x < x_end; x += blockDim.y) {
            if (a.interior_only && wilson_on_boundary(a, x, t)) continue;   // k_wilson_boundary takes those
            if ((MODE == WILSON_EO || MODE == WILSON_EO_DOT) && ((x + t) & 1) != a.eo_keep) {   // the other parity: zero
                const int n = x * a.wt + t;
                a.out[n] = make_double2(0.0, 0.0);
                a.out[a.V + n] = make_double2(0.0, 0.0);
                continue;
            }
            cplx o0, o1;
            wilson_site<DAG>(a, x, t, o0, o1);
            wilson_epilogue<MODE, NS>(a, x * a.wt + t, o0, o1, acc);
        }
    }
    if (MODE != WILSON_PLAIN && MODE != WILSON_EO) {
        const int nb = gridDim.x * gridDim.y;
        if (grid_reduce<NS>(acc, a.partials, a.ticket, nb + a.boundary_blocks, blockIdx.y * gridDim.x + blockIdx.x)) {
            if (threadIdx.x == 0 && threadIdx.y == 0) {
                a.sums_out[0] = acc[0];
                a.sums_out[1] = acc[1];
            }
        }
    }
}

// The sites of a split lattice that read ghost lines: columns t = 0 and wt-1 (split along t) and rows
// x = 0 and wx-1 (split along x; their corner sites belong to the columns when both are split).  Runs on
// the comm stream after the halo exchange while k_wilson(interior_only) computes everything else; the two
// launches share one reduction (ticket and block numbering).
template <bool DAG, int MODE>
__global__ void __launch_bounds__(kBlock) k_wilson_boundary(const WilsonArgs a) {
    if (a.done != nullptr && *a.done) return;
    constexpr int NS = (MODE == WILSON_PLAIN) ? 1 : 2;
    double acc[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) acc[j] = 0.0;
    const bool split_t = a.g_tp != nullptr, split_x = a.g_xp != nullptr;
    const int ncol = split_t ? 2 * a.wx : 0;
    const int wrow = split_x ? (split_t ? a.wt - 2 : a.wt) : 0;
    const int total = ncol + 2 * wrow;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int x, t;
        if (i < ncol) {
            x = i >> 1;
            t = (i & 1) ? a.wt - 1 : 0;
        } else {
            const int j = i - ncol;
            x = (j < wrow) ? 0 : a.wx - 1;
            t = (j < wrow ? j : j - wrow) + (split_t ? 1 : 0);
        }
        cplx o0, o1;
        wilson_site<DAG>(a, x, t, o0, o1);
        wilson_epilogue<MODE, NS>(a, x * a.wt + t, o0, o1, acc);
    }
    if (MODE != WILSON_PLAIN) {
        if (grid_reduce<NS>(acc, a.partials, a.ticket, a.interior_blocks + (int)gridDim.x, a.interior_blocks + (int)blockIdx.x)) {
            if (threadIdx.x == 0) {
                a.sums_out[0] = acc[0];
                a.sums_out[1] = acc[1];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Halo packing for a split lattice: the four lines a tile sends before a stencil application
// (src/dirac_operator.cpp:49-64 for D, :273-288 for D^dagger), one complex per boundary site.
//   to -t neighbour : psi0 + s psi1              of my column t = 0        (their g_tp)
//   to +t neighbour : conj(U0) (psi0 - s psi1)   of my column t = wt-1     (their g_tm)
//   to -x neighbour : psi0 - s i psi1            of my row x = 0           (their g_xp)
//   to +x neighbour : conj(U1) (psi0 + s i psi1) of my row x = wx-1        (their g_xm)
// ----------------------------------------------------------------------------------------------
struct PackArgs {
    const cplx* U;
    const cplx* in;
    int wx, wt, V;
    cplx* to_tm;   // [wx] or null
    cplx* to_tp;   // [wx]
    cplx* to_xm;   // [wt] or null
    cplx* to_xp;   // [wt]
    const int* done;
};

template <bool DAG>
__global__ void __launch_bounds__(kBlock) k_pack_halo(const PackArgs a) {
    if (a.done != nullptr && *a.done) return;
    constexpr double s = DAG ? 1.0 : -1.0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int wt = a.wt, wx = a.wx, V = a.V;
    if (a.to_tm != nullptr && i < wx) {
        {
            const int n = i * wt;
            const cplx p0 = a.in[n], p1 = a.in[V + n];
            a.to_tm[i] = make_double2(p0.x + s * p1.x, p0.y + s * p1.y);
        }
        {
            const int n = i * wt + wt - 1;
            const cplx p0 = a.in[n], p1 = a.in[V + n];
            a.to_tp[i] = cmulc(a.U[n], make_double2(p0.x - s * p1.x, p0.y - s * p1.y));
        }
    }
    if (a.to_xm != nullptr && i < wt) {
        {
            const int n = i;
            const cplx p0 = a.in[n], p1 = a.in[V + n];
            a.to_xm[i] = make_double2(p0.x + s * p1.y, p0.y - s * p1.x);
        }
        {
            const int n = (wx - 1) * wt + i;
            const cplx p0 = a.in[n], p1 = a.in[V + n];
            a.to_xp[i] = cmulc(a.U[V + n], make_double2(p0.x - s * p1.y, p0.y + s * p1.x));
        }
    }
}

// ----------------------------------------------------------------------------------------------
// BLAS-1 passes of CG fused with their reductions (src/conjugate_gradient.cpp:33-59).
// Flat over the 2V complex elements of a field; persistent grid-stride loops.
// ----------------------------------------------------------------------------------------------
// alpha = r_norm2 / dot(d,Ad) ; x += alpha d ; r -= alpha Ad ; sums_out[0] = sum |r|^2
__global__ void __launch_bounds__(kBlock) k_cg_update(CgState* st, int cur, cplx* __restrict__ x,
                                                      const cplx* __restrict__ d, cplx* __restrict__ r,
                                                      const cplx* __restrict__ Ad, int n_elems, double* partials,
                                                      unsigned int* ticket, double* sums_out) {
    if (st->done) return;
    const cplx alpha = cdiv(make_double2(st->rr[cur], 0.0), make_double2(st->dAd[0], st->dAd[1]));
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx dv = ld_stream(d + i), av = ld_stream(Ad + i);
        cplx xv = x[i], rv = r[i];
        xv = cadd(xv, cmul(alpha, dv));
        rv = csub(rv, cmul(alpha, av));
        x[i] = xv;
        r[i] = rv;
        acc[0] += rv.x * rv.x + rv.y * rv.y;
    }
    if (grid_reduce<1>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            st->k = st->k + 1;      // iterations done (read by the graph-replayed even-odd loop; unused by the two-pass path)
        }
    }
}

// stopping rule of iteration k-1, then beta = err_sqr / r_norm2 ; d = r + beta d   (k >= 1)
__device__ __forceinline__ bool cg_converged(const CgState* st, int cur, double tol) {
    return sqrt(st->rr[cur]) < tol * sqrt(st->phi_norm2);
}

__global__ void __launch_bounds__(kBlock) k_cg_dir(CgState* st, int k, double tol, const cplx* __restrict__ r,
                                                   cplx* __restrict__ d, int n_elems) {
    if (st->done) return;
    const int cur = k & 1;
    if (cg_converged(st, cur, tol)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->iters = k - 1;
            st->converged = 1;
            st->done = 1;
        }
        return;
    }
    const double beta = st->rr[cur] / st->rr[cur ^ 1];
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx rv = ld_stream(r + i);
        cplx dv = d[i];
        dv.x = dv.x * beta + rv.x;
        dv.y = dv.y * beta + rv.y;
        d[i] = dv;
    }
}

// the same with the iteration index read from the device state (kernel arguments independent of the iteration: CUDA graphs)
__global__ void __launch_bounds__(kBlock) k_cg_dir_dev(CgState* st, int cur, const cplx* __restrict__ r,
                                                       cplx* __restrict__ d, int n_elems) {
    if (st->done) return;
    if (cg_converged(st, cur, st->tol)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            st->iters = st->k - 1;
            st->converged = 1;
            st->done = 1;
        }
        return;
    }
    const double beta = st->rr[cur] / st->rr[cur ^ 1];
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx rv = ld_stream(r + i);
        cplx dv = d[i];
        dv.x = dv.x * beta + rv.x;
        dv.y = dv.y * beta + rv.y;
        d[i] = dv;
    }
}

// one thread: stopping rule after iteration k-1 (used at batch ends and after max_iter)
__global__ void k_cg_check(CgState* st, int k, double tol, int max_iter) {
    if (st->done) return;
    if (cg_converged(st, k & 1, tol)) {
        st->iters = k - 1;
        st->converged = 1;
        st->done = 1;
    } else if (k >= max_iter) {
        st->iters = max_iter;
        st->converged = 0;
        st->done = 1;
    }
}

// the same with the device-side iteration counter (one-pass path).  With peer-memory sums the |r|^2 of the last
// iteration is still in the ranks' slots: gather it (kind 1, parity of iteration k-1, epoch base + k).
__global__ void k_cg_check_dev(CgState* st, const DistLink dl) {
    if (st->done) return;
    const int k = st->k, max_iter = st->max_iter;
    if (dl.on && k > 0) st->rr[k & 1] = gather_sum1_thread(dl, 1, (k - 1) & 1, st->epoch_base + (unsigned int)k);
    if (cg_converged(st, k & 1, st->tol)) {
        st->iters = k - 1;
        st->converged = 1;
        st->done = 1;
    } else if (k >= max_iter) {
        st->iters = max_iter;
        st->converged = 0;
        st->done = 1;
    }
}

__global__ void k_cg_reset(CgState* st, double tol, int max_iter, unsigned int epoch_base = 0) {
    st->epoch_base = epoch_base;
    st->tol = tol;
    st->max_iter = max_iter;
    st->done = 0;
    st->iters = 0;
    st->converged = 0;
    st->pending = 0;
    st->pending_buf = 0;
    st->k = 0;
}

// CG start from a finished A phi (even-odd solver): x = phi ; r = phi - A phi ; d = r ; sums |phi|^2, |r|^2
__global__ void __launch_bounds__(kBlock) k_cg_start(const cplx* __restrict__ phi, const cplx* __restrict__ Aphi,
                                                     cplx* __restrict__ x, cplx* __restrict__ r, cplx* __restrict__ d,
                                                     int n_elems, double* partials, unsigned int* ticket, double* sums_out) {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx f = ld_stream(phi + i), av = ld_stream(Aphi + i);
        const cplx rv = csub(f, av);
        x[i] = f;
        r[i] = rv;
        d[i] = rv;
        acc[0] += f.x * f.x + f.y * f.y;
        acc[1] += rv.x * rv.x + rv.y * rv.y;
    }
    if (grid_reduce<2>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            sums_out[1] = acc[1];
        }
    }
}

// keep one parity of a field: f(n) = 0 where (x + t) & 1 != keep
__global__ void __launch_bounds__(kBlock) k_mask_parity(cplx* __restrict__ f, int wt, int V, int keep) {
    const int stride = gridDim.x * blockDim.x;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < V; n += stride) {
        const int x = n / wt, t = n - x * wt;
        if (((x + t) & 1) != keep) {
            f[n] = make_double2(0.0, 0.0);
            f[V + n] = make_double2(0.0, 0.0);
        }
    }
}

// y += x
__global__ void __launch_bounds__(kBlock) k_add_into(cplx* __restrict__ y, const cplx* __restrict__ x, int n_elems) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) y[i] = cadd(y[i], ld_stream(x + i));
}

// chronological start vector (opt-in): g = 2 a - b, the linear extrapolation of the last two solutions along the
// molecular-dynamics trajectory
__global__ void __launch_bounds__(kBlock) k_extrapolate(const cplx* __restrict__ a, const cplx* __restrict__ b,
                                                        cplx* __restrict__ g, int n_elems) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx u = ld_stream(a + i), v = ld_stream(b + i);
        g[i] = make_double2(2.0 * u.x - v.x, 2.0 * u.y - v.y);
    }
}

// dot(x,y) = sum x conj(y) over both components (include/variables.h:181-192)
__global__ void __launch_bounds__(kBlock) k_dot(const cplx* __restrict__ x, const cplx* __restrict__ y, int n_elems,
                                                double* partials, unsigned int* ticket, double* sums_out) {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx p = cmul_conj(ld_stream(x + i), ld_stream(y + i));
        acc[0] += p.x;
        acc[1] += p.y;
    }
    if (grid_reduce<2>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            sums_out[1] = acc[1];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Gauge-field access with optional ghost ring (split lattice): x in [-1,wx], t in [-1,wt].
// Null ghosts = periodic wrap inside the tile (links carry no antiperiodic sign).
//   gx_m / gx_p : rows x=-1 / x=wx, (wt+2) entries per mu indexed [t+1] (corners included)
//   gt_m / gt_p : columns t=-1 / t=wt, wx entries per mu indexed [x]
// ----------------------------------------------------------------------------------------------
struct GaugeView {
    const cplx* U;
    int wx, wt, V;
    const cplx* gx_m;
    const cplx* gx_p;
    const cplx* gt_m;
    const cplx* gt_p;

    __device__ __forceinline__ cplx at(int mu, int x, int t) const {
        if (gx_m == nullptr) {
            if (x < 0) x += wx;
            if (x >= wx) x -= wx;
        } else {
            if (x < 0) return ldg(gx_m + mu * (wt + 2) + t + 1);
            if (x >= wx) return ldg(gx_p + mu * (wt + 2) + t + 1);
        }
        if (gt_m == nullptr) {
            if (t < 0) t += wt;
            if (t >= wt) t -= wt;
        } else {
            if (t < 0) return ldg(gt_m + mu * wx + x);
            if (t >= wt) return ldg(gt_p + mu * wx + x);
        }
        return ldg(U + mu * V + x * wt + t);
    }
};

// Plaquette U01(n) = U0(n) U1(n+t) conj(U0(n+x)) conj(U1(n))  (src/gauge_conf.cpp:41-48) with
// MeasureSp_HMC (:427-437) and Compute_gaugeAction (:441-449) fused as the two sums.
__global__ void __launch_bounds__(kBlock) k_plaquette(const GaugeView g, double beta, cplx* __restrict__ P,
                                                      double* partials, unsigned int* ticket, double* sums_out) {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < g.V; n += stride) {
        const int x = n / g.wt, t = n - x * g.wt;
        const cplx p = cmul_conj(cmul_conj(cmul(g.at(0, x, t), g.at(1, x, t + 1)), g.at(0, x + 1, t)), g.at(1, x, t));
        if (P != nullptr) P[n] = p;
        acc[0] += p.x;
        acc[1] += beta * (1.0 - p.x);
    }
    if (grid_reduce<2>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            sums_out[1] = acc[1];
        }
    }
}

// Staples (src/gauge_conf.cpp:96-127):
//   K0(n) = U1(n) U0(n+x) conj(U1(n+t)) + conj(U1(n-x)) U0(n-x) U1(n-x+t)
//   K1(n) = U0(n) U1(n+t) conj(U0(n+x)) + conj(U0(n-t)) U1(n-t) U0(n+x-t)
__device__ __forceinline__ void staple_site(const GaugeView& g, int x, int t, cplx& K0, cplx& K1) {
    const cplx u0 = g.at(0, x, t), u1 = g.at(1, x, t);
    const cplx u0xp = g.at(0, x + 1, t), u1tp = g.at(1, x, t + 1);
    K0 = cadd(cmul_conj(cmul(u1, u0xp), u1tp), cmul(cmulc(g.at(1, x - 1, t), g.at(0, x - 1, t)), g.at(1, x - 1, t + 1)));
    K1 = cadd(cmul_conj(cmul(u0, u1tp), u0xp), cmul(cmulc(g.at(0, x, t - 1), g.at(1, x, t - 1)), g.at(0, x + 1, t - 1)));
}

__global__ void __launch_bounds__(kBlock) k_staple(const GaugeView g, cplx* __restrict__ K) {
    const int stride = gridDim.x * blockDim.x;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < g.V; n += stride) {
        const int x = n / g.wt, t = n - x * g.wt;
        cplx K0, K1;
        staple_site(g, x, t, K0, K1);
        K[n] = K0;
        K[g.V + n] = K1;
    }
}

// ----------------------------------------------------------------------------------------------
// Forces.  phi_dag_partialD_phi (src/dirac_operator.cpp:493-507; psi = left, chi = right):
//   F0(n) = Im[ U0(n) sR conj(psi0-psi1)(n) (chi0-chi1)(n+t) - conj(U0(n)) sR conj(psi0+psi1)(n+t) (chi0+chi1)(n) ]
//   F1(n) = Im[ U1(n) (conj psi0 - i conj psi1)(n) (chi0 + i chi1)(n+x)
//             + conj(U1(n)) (conj psi0 + i conj psi1)(n+x) (-chi0 + i chi1)(n) ]
// and HMC::Force_G (src/hmc.cpp:32-40):  F_mu(n) += -beta Im( U_mu(n) conj(K_mu(n)) ), fused so
// U, psi, chi are read once and the staples never reach HBM.
// Split lattice: the forward neighbours of the last row/column come as two projected complex
// numbers per site (ghost lines fg_*), see k_pack_force.
// ----------------------------------------------------------------------------------------------
struct ForceArgs {
    GaugeView g;
    const cplx* psi;
    const cplx* chi;
    double* F;
    double beta;
    double sR_edge;
    int fermion, gauge;      // which parts to include
    const cplx* fg_t;        // [2*wx]: (chi0-chi1)(x, t=0 of +t tile), (psi0+psi1)(same)
    const cplx* fg_x;        // [2*wt]: (chi0+i chi1)(x=0 of +x tile, t), (psi0 - i psi1)(same)
};

__global__ void __launch_bounds__(kBlock) k_force(const ForceArgs a) {
    const GaugeView& g = a.g;
    const int wt = g.wt, wx = g.wx, V = g.V;
    const int stride = gridDim.x * blockDim.x;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < V; n += stride) {
        const int x = n / wt, t = n - x * wt;
        const cplx u0 = g.at(0, x, t), u1 = g.at(1, x, t);
        double f0 = 0.0, f1 = 0.0;
        if (a.fermion) {
            const cplx l0 = ldg(a.psi + n), l1 = ldg(a.psi + V + n);
            const cplx q0 = ldg(a.chi + n), q1 = ldg(a.chi + V + n);
            // mu = 0
            cplx hq, hl;   // (chi0-chi1)(n+t), (psi0+psi1)(n+t)
            if (t == wt - 1 && a.fg_t != nullptr) {
                hq = ldg(a.fg_t + x);
                hl = ldg(a.fg_t + wx + x);
            } else {
                const int m = nb_tp(n, t, wt);
                hq = csub(ldg(a.chi + m), ldg(a.chi + V + m));
                hl = cadd(ldg(a.psi + m), ldg(a.psi + V + m));
            }
            const double sR = (t == wt - 1) ? a.sR_edge : 1.0;
            {
                const cplx A = cmul(cmul(u0, cconj(csub(l0, l1))), hq);
                const cplx B = cmul(cmul(cconj(u0), cconj(hl)), cadd(q0, q1));
                f0 = sR * (A.y - B.y);
            }
            // mu = 1
            cplx kq, kl;   // (chi0 + i chi1)(n+x), (psi0 - i psi1)(n+x)
            if (x == wx - 1 && a.fg_x != nullptr) {
                kq = ldg(a.fg_x + t);
                kl = ldg(a.fg_x + wt + t);
            } else {
                const int m = nb_xp(n, x, wx, wt);
                const cplx c0 = ldg(a.chi + m), c1 = ldg(a.chi + V + m);
                const cplx p0 = ldg(a.psi + m), p1 = ldg(a.psi + V + m);
                kq = make_double2(c0.x - c1.y, c0.y + c1.x);
                kl = make_double2(p0.x + p1.y, p0.y - p1.x);
            }
            {
                // conj(psi0) - i conj(psi1) = conj(psi0 + i psi1)
                const cplx w = cconj(make_double2(l0.x - l1.y, l0.y + l1.x));
                const cplx A = cmul(cmul(u1, w), kq);
                // -chi0 + i chi1
                const cplx z = make_double2(-q0.x - q1.y, -q0.y + q1.x);
                const cplx B = cmul(cmul(cconj(u1), cconj(kl)), z);
                f1 = A.y + B.y;
            }
        }
        if (a.gauge) {
            cplx K0, K1;
            staple_site(g, x, t, K0, K1);
            f0 += -a.beta * cmul_conj(u0, K0).y;
            f1 += -a.beta * cmul_conj(u1, K1).y;
        }
        a.F[n] = f0;
        a.F[V + n] = f1;
    }
}

// Leapfrog updates (src/hmc.cpp:69-72, 79-87, 93-101), flat over the 2V links:
//   pi += eps_pi F (skipped when F is null) ; U <- U exp(i eps_u pi)
__global__ void __launch_bounds__(kBlock) k_leap_update(cplx* __restrict__ U, double* __restrict__ pi,
                                                        const double* __restrict__ F, double eps_pi, double eps_u,
                                                        int n_links) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_links; i += stride) {
        double p = pi[i];
        if (F != nullptr) {
            p += eps_pi * F[i];
            pi[i] = p;
        }
        double sn, cs;
        sincos(eps_u * p, &sn, &cs);
        U[i] = cmul(U[i], make_double2(cs, sn));
    }
}

// sum 1/2 pi^2 (src/hmc.cpp:138-144)
__global__ void __launch_bounds__(kBlock) k_kinetic(const double* __restrict__ pi, int n_links, double* partials,
                                                    unsigned int* ticket, double* sums_out) {
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_links; i += stride) {
        const double p = pi[i];
        acc[0] += 0.5 * p * p;
    }
    if (grid_reduce<1>(acc, partials, ticket)) {
        if (threadIdx.x == 0) sums_out[0] = acc[0];
    }
}

// ----------------------------------------------------------------------------------------------
// Gaussian refresh (HMC::RandomPI / RandomCHI, src/hmc.cpp:5-28): pi ~ N(0,1) per link,
// chi re,im ~ N(0, 1/sqrt 2) per spin component.  Counter-based Philox-4x32-10 keyed on
// (seed, trajectory, field) and counted by the GLOBAL element index, so the fields do not depend
// on the GPU decomposition; Box-Muller in FP64 from 2x 53-bit uniforms.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; round++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// two independent N(0,1) from one Philox block
__device__ __forceinline__ void gauss_pair(uint64_t seed, uint64_t stream, uint64_t index, double& g0, double& g1) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[0] << 32) | c[1], b = ((uint64_t)c[2] << 32) | c[3];
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);   // (0,1)
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    g0 = rad * cs;
    g1 = rad * sn;
}

// global element index of local site n (tile origin ox, ot in a lattice with Nt columns)
struct TileMap {
    int wx, wt, ox, ot, Nt;
    long long Vglobal;
    __device__ __forceinline__ long long global_site(int n) const {
        const int x = n / wt, t = n - x * wt;
        return (long long)(ox + x) * Nt + (ot + t);
    }
};

__global__ void __launch_bounds__(kBlock) k_refresh(double* __restrict__ pi, cplx* __restrict__ chi, TileMap m,
                                                    uint64_t seed, uint64_t traj) {
    const int V = m.wx * m.wt;
    const int stride = gridDim.x * blockDim.x;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < V; n += stride) {
        const long long gs = m.global_site(n);
        double a, b;
        gauss_pair(seed, 3 * traj + 0, (uint64_t)gs, a, b);   // pi0(n), pi1(n)
        pi[n] = a;
        pi[V + n] = b;
        const double sd = 0.70710678118654752440;             // 1/sqrt(2) (hmc.cpp:22)
        gauss_pair(seed, 3 * traj + 1, (uint64_t)gs, a, b);
        chi[n] = make_double2(sd * a, sd * b);
        gauss_pair(seed, 3 * traj + 2, (uint64_t)gs, a, b);
        chi[V + n] = make_double2(sd * a, sd * b);
    }
}

}  // namespace sm
