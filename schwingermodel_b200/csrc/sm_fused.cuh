// sm_fused.cuh -- D D^dagger in ONE pass over HBM (temporal blocking of the two Wilson stencils).
//
// The two-pass form (k_wilson twice) moves 192 B per site-update because the intermediate
// t = D^dagger psi goes out to HBM and comes back.  Here a block owns a strip of columns
// (t direction) and marches down the x rows keeping, per thread (= per column),
//     psi rows j, j-1, j-2      t rows j-1, j-2      the finished -x hop term of t row j-3
//     and the links of rows j, j-1, j-2
// in REGISTERS, as rings whose slots are compile-time constants after a 6-fold unroll of the row
// loop (rotating the window costs no instruction).  The t-direction neighbours travel through a
// small double-buffered shared-memory line of pre-projected half-spinors (the same rank-1 trick
// the halo exchange uses), one __syncthreads per row.  Rows are staged from HBM into a
// shared-memory ring with cp.async (16 B per thread and array, L1 bypassed, STAGES-1 rows in
// flight per block) so the loads never occupy registers or stall the step.  Each step takes row
// j, forms t(j-1) = D^dagger psi and out(j-2) = D t, so psi and U are read once and out is
// written once: ~96 B per site-update plus the halo overhead (2 columns each side of a strip,
// 4 rows per chunk).
//
// FUSED_CG additionally folds the CG vector updates that touch the same data into the pass
// (src/conjugate_gradient.cpp:32-59):  d_k = r_k + beta d_{k-1} is formed on load (also at the
// halo sites, from r and d_{k-1}), x += alpha_{k-1} d_{k-1} is applied where d_{k-1} is read, and
// dot(d_k, A d_k) is reduced in the epilogue.  With k_cg_resid (r -= alpha A d, |r|^2) one CG
// iteration moves 224 + 96 = 320 B per site instead of 512.
//
// Everything is templated on the complex type: cplx (double2) is the reference-exact path, cplxf
// (float2) the inner solve of the opt-in mixed-precision CG.
#pragma once
#include "sm_kernels.cuh"

namespace sm {

enum { FUSED_PLAIN = 0, FUSED_DOT = 1, FUSED_CG = 2 };

template <typename C>
struct FusedArgsT {
    const C* U;
    const C* in;      // psi (PLAIN/DOT) or d_{k-1} (CG)
    C* out;           // D D^dagger psi  (A d_k in CG)
    int wx, wt, V;
    int rows_per_block;
    int cols_per_strip;  // output columns per block (<= blockDim.x - 4)
    int nchunks;         // row chunks of the whole pass (for the reduction that spans its launches)
    int chunk_mode;      // 0: all rows, uniform chunks; 1: interior rows [rb, wx-rb); 2: the two boundary bands
    int rb;              // rows per boundary band (split lattice; the only rows that read ghost rows)
    double mass;
    double sR_edge, sL_edge;
    double* partials;
    unsigned int* ticket;
    double* sums_out;
    // CG mode
    CgState* st;
    const C* r;       // r_k
    C* x;             // x_{k-1} -> x_k
    C* d_new;         // d_k
    int first;           // iteration 0: d_0 = r_0, no x update owed
    int cur;             // parity of the iteration (selects rr[], the d ping-pong is in the pointers)
    // lattice split along x (ranks_t == 1): rows -2,-1 ("lo") and wx, wx+1 ("hi") live in ghost
    // arrays laid out [component][2 rows][wt]; null = wrap inside the tile
    const C* gU_lo;
    const C* gU_hi;
    const C* gin_lo;   // ghost rows of `in`
    const C* gin_hi;
    const C* gr_lo;    // CG: ghost rows of r
    const C* gr_hi;
    C* gd_lo;          // CG: ghost rows of d_k, written here for the next iteration
    C* gd_hi;
    DistLink dl;       // CG on a split lattice with peer-memory sums and halos (sm_peer.cuh); dl.on == 0 otherwise
    // lattice split along t (k_dd_tma only): 2-deep ghost COLUMNS, layout [component][wx rows][2]: columns -2,-1 ("lo", from
    // the -t neighbour) and wt, wt+1 ("hi").  With them the ghost ROW arrays are row_w = wt + 4 wide and start at column
    // -2 (they carry the corner entries of a 2-D split); without, row_w = wt.  tg_on == 0: wrap in t inside the tile.
    int tg_on;
    int row_w;
    // which strips a launch covers (k_dd_tma; 0 everywhere else): 0 all (strip = blockIdx.x), 1 the interior strips
    // 1 .. nstrips-2, 2 the two edge strips -- the only ones that read ghost columns; they follow the halo exchange on the
    // comm stream while the interior strips compute.  A pass made of several launches numbers its partial sums explicitly:
    // part_total > 0: this launch's blocks are part_base + (blockIdx.y * gridDim.x + blockIdx.x) of part_total.
    int strip_mode, nstrips;
    int part_base, part_total;
    const C* tgU_lo;
    const C* tgU_hi;
    const C* tgin_lo;  // psi (PLAIN/DOT) or d_{k-1} (CG)
    const C* tgin_hi;
    const C* tgr_lo;   // CG: r
    const C* tgr_hi;
    C* tgd_lo;         // CG: ghost columns of d_k, written by the edge strips for the next iteration
    C* tgd_hi;
};

typedef FusedArgsT<cplx> FusedArgs;

// hop algebra shared by both operators: s = +1 for D^dagger, -1 for D (see k_wilson)
template <bool DAG>
struct Hop {
    static constexpr int si = DAG ? 1 : -1;
    // half-spinors as seen from the receiving site
    template <typename C> static __device__ __forceinline__ C from_tp(C p0, C p1) { const typename RealOf<C>::type s = si; return mkc<C>(p0.x + s * p1.x, p0.y + s * p1.y); }
    template <typename C> static __device__ __forceinline__ C from_xp(C p0, C p1) { const typename RealOf<C>::type s = si; return mkc<C>(p0.x + s * p1.y, p0.y - s * p1.x); }
    template <typename C> static __device__ __forceinline__ C from_tm(C p0, C p1) { const typename RealOf<C>::type s = si; return mkc<C>(p0.x - s * p1.x, p0.y - s * p1.y); }
    template <typename C> static __device__ __forceinline__ C from_xm(C p0, C p1) { const typename RealOf<C>::type s = si; return mkc<C>(p0.x - s * p1.y, p0.y + s * p1.x); }
    // accumulate the four hop terms v (already multiplied by the link and the sign)
    template <typename C> static __device__ __forceinline__ void add_tp(C v, C& a0, C& a1) { const typename RealOf<C>::type s = si; a0 = v; a1 = cscale(s, v); }
    template <typename C> static __device__ __forceinline__ void add_xp(C v, C& a0, C& a1) { const typename RealOf<C>::type s = si; a0 = cadd(a0, v); a1.x -= s * v.y; a1.y += s * v.x; }
    template <typename C> static __device__ __forceinline__ void add_tm(C v, C& a0, C& a1) { const typename RealOf<C>::type s = si; a0 = cadd(a0, v); a1.x -= s * v.x; a1.y -= s * v.y; }
    template <typename C> static __device__ __forceinline__ void add_xm(C v, C& a0, C& a1) { const typename RealOf<C>::type s = si; a0 = cadd(a0, v); a1.x += s * v.y; a1.y -= s * v.x; }
};

// Start of a FUSED_CG pass (iteration k): the stopping rule of iteration k-1 and the scalars beta_k, alpha_{k-1}
// (src/conjugate_gradient.cpp:43-53).  On a split lattice with peer-memory sums |r_k|^2 is still spread over the
// ranks' slots (k_cg_resid of iteration k-1 published it): every block gathers it, the lead block files it in
// CgState.  Returns false when the whole block has to leave (solve finished).  Called by all threads.
template <typename C>
__device__ __forceinline__ bool fused_cg_begin(const FusedArgsT<C>& a, bool lead_thread, typename RealOf<C>::type& beta,
                                               C& alpha, bool& first) {
    typedef typename RealOf<C>::type R;
    if (a.st->done) return false;
    const int cur = a.cur;
    first = (a.first != 0);
    if (first) return true;
    double rr_cur;
    if (a.dl.on) {
        double g[1];
        gather_sums<1>(a.dl, 1, cur ^ 1, a.st->epoch_base + (unsigned int)a.st->k, g);
        rr_cur = g[0];
        if (lead_thread) a.st->rr[cur] = rr_cur;
    } else {
        rr_cur = a.st->rr[cur];
    }
    if (sqrt(rr_cur) < a.st->tol * sqrt(a.st->phi_norm2)) {
        if (lead_thread) {
            a.st->iters = a.st->k - 1;
            a.st->converged = 1;
            a.st->done = 1;
        }
        return false;
    }
    beta = (R)(rr_cur / a.st->rr[cur ^ 1]);
    alpha = mkc<C>((R)a.st->alpha[0], (R)a.st->alpha[1]);
    return true;
}

// blocks that read ghost rows of r wait until both neighbours have delivered r_k (epoch base + k)
template <typename C>
__device__ __forceinline__ void fused_cg_wait_ghosts(const FusedArgsT<C>& a) {
    if (a.dl.on && a.gU_lo != nullptr && a.chunk_mode != 1) {
        if (threadIdx.x == 0) {
            const unsigned int e = a.st->epoch_base + (unsigned int)a.st->k;
            spin_until(a.dl.my_flag_lo, e);
            spin_until(a.dl.my_flag_hi, e);
        }
        __syncthreads();
    }
}

// the finishing block of a FUSED_DOT / FUSED_CG pass files dot(d, A d): CgState (or the caller's buffer), or the ranks'
// slots.  Called by warp 0 of that block (the sums are valid in lane 0).
template <typename C>
__device__ __forceinline__ void fused_sums_out(const FusedArgsT<C>& a, const double (&acc)[2]) {
    if (a.dl.on) {
        publish_sums<2>(a.dl, 0, a.cur, a.st->epoch_base + (unsigned int)a.st->k + 1u, acc);
    } else if ((threadIdx.x & 31) == 0) {
        a.sums_out[0] = acc[0];
        a.sums_out[1] = acc[1];
    }
}

__device__ __forceinline__ int wrap_idx(int a, int n) {
    a %= n;
    return a < 0 ? a + n : a;
}

// cp.async (LDGSTS): 16 bytes global -> shared without passing through registers, L1 bypassed
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// one field element: 16 bytes (double complex, L1 bypassed) or 8 bytes (single complex)
__device__ __forceinline__ void cp_async_elem(cplx* smem_dst, const cplx* gmem_src) { cp_async16(smem_dst, gmem_src); }
__device__ __forceinline__ void cp_async_elem(cplxf* smem_dst, const cplxf* gmem_src) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__host__ __device__ constexpr int fused_arrays(int mode) { return mode == FUSED_CG ? 8 : 4; }
constexpr size_t fused_smem_bytes(int mode, int stages, int BT, size_t elem_bytes = sizeof(cplx)) {
    return elem_bytes * (size_t)BT * (2 * 4 + (size_t)stages * fused_arrays(mode));
}

// Shared memory:  line  [2 parities][4][BT]   t-direction half-spinors of the psi row and the t row
//                 stage [STAGES][NARR][BT]    rows in flight (each thread stages its own column:
//                                             cp.async, no block barrier needed for it)
template <typename C, int MODE, int STAGES>
__global__ void __launch_bounds__(kBlock, 2) k_dd_fused(const FusedArgsT<C> a) {
    typedef typename RealOf<C>::type R;
    static_assert(6 % STAGES == 0, "the stage ring is indexed with the unrolled step: STAGES must divide 6");
    extern __shared__ __align__(16) unsigned char s_raw[];
    C* const s_mem = reinterpret_cast<C*>(s_raw);
    constexpr int NARR = fused_arrays(MODE);
    const int BT = blockDim.x;
    const int tid = threadIdx.x;
    const int wt = a.wt, wx = a.wx, V = a.V;
    C* const s_line = s_mem;
    C* const s_stage = s_mem + 2 * 4 * BT;

    R beta = 0;
    C alpha = mkc<C>(0, 0);
    bool first = true;
    if (MODE == FUSED_CG) {
        if (!fused_cg_begin(a, blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && a.chunk_mode != 1, beta, alpha, first)) return;
        fused_cg_wait_ghosts(a);
    }

    const int tc = blockIdx.x * a.cols_per_strip - 2 + tid;    // unwrapped column of this thread
    const int t = wrap_idx(tc, wt);
    const bool col_active = (tid < a.cols_per_strip + 4) && (tc <= wt + 1);   // strip + 2 halo columns each side
    const bool col_owner = (tid >= 2) && (tid < a.cols_per_strip + 2) && (tc < wt);
    const R sR = (R)((t == wt - 1) ? a.sR_edge : 1.0);
    int chunk, xa, xb;
    if (a.chunk_mode == 0) {
        chunk = blockIdx.y;
        xa = chunk * a.rows_per_block;
        xb = min(wx, xa + a.rows_per_block);
    } else if (a.chunk_mode == 1) {
        chunk = blockIdx.y + 1;
        xa = a.rb + (int)blockIdx.y * a.rows_per_block;
        xb = min(wx - a.rb, xa + a.rows_per_block);
    } else {
        chunk = (blockIdx.y == 0) ? 0 : a.nchunks - 1;
        xa = (blockIdx.y == 0) ? 0 : wx - a.rb;
        xb = xa + a.rb;
    }
    const int tl = (tid == 0) ? 0 : tid - 1, tr = (tid == BT - 1) ? tid : tid + 1;
    const int j_first = xa - 2, j_last = xb + 1;

    const C zero = mkc<C>(0, 0);
    if (!col_active) {
        for (int q = 0; q < STAGES * NARR; q++) s_stage[q * BT + tid] = zero;   // read back as zeros
    }

    // stage row j (ring slot `slot`): psi (or r, d_{k-1}, x) and the links of this thread's column
    const bool split = (a.gU_lo != nullptr);
    auto issue_row = [&](int j, int slot) {
        if (col_active && j <= j_last) {
            C* st = s_stage + slot * NARR * BT + tid;
            const C *pU, *pin, *pr = nullptr;
            int cs;                                   // component stride of the source arrays
            if (split && j < 0) {
                const int o = (j + 2) * wt + t;
                cs = 2 * wt;
                pU = a.gU_lo + o;
                pin = a.gin_lo + o;
                if (MODE == FUSED_CG) pr = a.gr_lo + o;
            } else if (split && j >= wx) {
                const int o = (j - wx) * wt + t;
                cs = 2 * wt;
                pU = a.gU_hi + o;
                pin = a.gin_hi + o;
                if (MODE == FUSED_CG) pr = a.gr_hi + o;
            } else {
                int x = j;                            // -2 <= j <= wx+1: one correction wraps it
                if (x < 0) x += wx;
                else if (x >= wx) x -= wx;
                const int n = x * wt + t;
                cs = V;
                pU = a.U + n;
                pin = a.in + n;
                if (MODE == FUSED_CG) pr = a.r + n;
            }
            cp_async_elem(st + 0 * BT, pU);
            cp_async_elem(st + 1 * BT, pU + cs);
            if (MODE != FUSED_CG) {
                cp_async_elem(st + 2 * BT, pin);
                cp_async_elem(st + 3 * BT, pin + cs);
            } else {
                cp_async_elem(st + 2 * BT, pr);
                cp_async_elem(st + 3 * BT, pr + cs);
                if (!first) {
                    cp_async_elem(st + 4 * BT, pin);
                    cp_async_elem(st + 5 * BT, pin + cs);
                    if (col_owner && j >= xa && j < xb) {
                        const int n = j * wt + t;
                        cp_async_elem(st + 6 * BT, a.x + n);
                        cp_async_elem(st + 7 * BT, a.x + V + n);
                    }
                }
            }
        }
        cp_async_commit();
    };

    // Register rings, indexed with compile-time slots after the 6-fold unroll below (so a "rotation"
    // costs no instruction):  P[.] psi rows j, j-1, j-2 ; Tr[.] t rows j-1, j-2 ; HX[.] the finished
    // -x hop term conj(U1) proj(t) of the row below ; U0w/U1w links of rows j, j-1, j-2.
    // The antiperiodic sign is folded into U0w when the row is taken (rt == 1: the +t hop out of column
    // wt-1 and the -t hop into column 0 are the same link, so one factor serves both directions).
    C P[3][2], Tr[2][2], HX[2], U0w[3], U1w[3];
#pragma unroll
    for (int q = 0; q < 3; q++) P[q][0] = P[q][1] = U0w[q] = U1w[q] = zero;
#pragma unroll
    for (int q = 0; q < 2; q++) Tr[q][0] = Tr[q][1] = HX[q] = zero;
    double acc[2] = {0.0, 0.0};
    const R mass = (R)a.mass, half = (R)0.5;

#pragma unroll
    for (int q = 0; q < STAGES - 1; q++) issue_row(j_first + q, q);

    for (int jb = j_first; jb <= j_last; jb += 6) {
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const int j = jb + s;
            if (j > j_last) break;                         // uniform over the block
            const int c = s % 3, m1 = (s + 2) % 3, m2 = (s + 1) % 3;   // ring slots of rows j, j-1, j-2
            const int tn_i = s % 2, t2_i = (s + 1) % 2;

            issue_row(j + STAGES - 1, (s + STAGES - 1) % STAGES);   // into the stage consumed at step j-1
            cp_async_wait<STAGES - 1>();                            // row j has landed
            const C* st = s_stage + (s % STAGES) * NARR * BT + tid;
            U0w[c] = cscale(sR, st[0 * BT]);
            U1w[c] = st[1 * BT];
            C p0 = st[2 * BT], p1 = st[3 * BT];
            if (MODE == FUSED_CG && !first) {
                // d_k = r_k + beta d_{k-1}  (conjugate_gradient.cpp:54-59), also at the halo sites
                const C d0 = st[4 * BT], d1 = st[5 * BT];
                p0 = mkc<C>(d0.x * beta + p0.x, d0.y * beta + p0.y);
                p1 = mkc<C>(d1.x * beta + p1.x, d1.y * beta + p1.y);
                if (col_owner && j >= xa && j < xb) {            // x += alpha_{k-1} d_{k-1}  (:34-36)
                    const int n = j * wt + t;
                    a.x[n] = cadd(st[6 * BT], cmul(alpha, d0));
                    a.x[V + n] = cadd(st[7 * BT], cmul(alpha, d1));
                }
            }
            if (MODE == FUSED_CG && col_owner) {
                if (j >= xa && j < xb) {
                    const int n = j * wt + t;
                    a.d_new[n] = p0;
                    a.d_new[V + n] = p1;
                } else if (split && j < 0) {             // keep d_k's ghost rows for the next iteration
                    a.gd_lo[(j + 2) * wt + t] = p0;
                    a.gd_lo[2 * wt + (j + 2) * wt + t] = p1;
                } else if (split && j >= wx) {
                    a.gd_hi[(j - wx) * wt + t] = p0;
                    a.gd_hi[2 * wt + (j - wx) * wt + t] = p1;
                }
            }
            P[c][0] = p0;
            P[c][1] = p1;

            // publish the t-direction half-spinors of psi row j-1 (for D^dagger) and t row j-2 (for D)
            C* line = s_line + (s & 1) * 4 * BT;
            line[0 * BT + tid] = Hop<true>::from_tp(P[m1][0], P[m1][1]);                       // read by column t-1
            line[1 * BT + tid] = cmulc(U0w[m1], Hop<true>::from_tm(P[m1][0], P[m1][1]));       // read by column t+1
            line[2 * BT + tid] = Hop<false>::from_tp(Tr[t2_i][0], Tr[t2_i][1]);
            line[3 * BT + tid] = cmulc(U0w[m2], Hop<false>::from_tm(Tr[t2_i][0], Tr[t2_i][1]));
            __syncthreads();

            // t(j-1) = D^dagger psi at row j-1
            C tn0, tn1;
            {
                C a0, a1;
                Hop<true>::add_tp(cmul(U0w[m1], line[0 * BT + tr]), a0, a1);
                Hop<true>::add_xp(cmul(U1w[m1], Hop<true>::from_xp(p0, p1)), a0, a1);
                Hop<true>::add_tm(line[1 * BT + tl], a0, a1);
                Hop<true>::add_xm(cmulc(U1w[m2], Hop<true>::from_xm(P[m2][0], P[m2][1])), a0, a1);
                tn0 = mkc<C>(mass * P[m1][0].x - half * a0.x, mass * P[m1][0].y - half * a0.y);
                tn1 = mkc<C>(mass * P[m1][1].x - half * a1.x, mass * P[m1][1].y - half * a1.y);
            }
            // out(j-2) = D t at row j-2
            if (j >= xa + 2 && col_owner) {
                C a0, a1;
                Hop<false>::add_tp(cmul(U0w[m2], line[2 * BT + tr]), a0, a1);
                Hop<false>::add_xp(cmul(U1w[m2], Hop<false>::from_xp(tn0, tn1)), a0, a1);
                Hop<false>::add_tm(line[3 * BT + tl], a0, a1);
                Hop<false>::add_xm(HX[tn_i], a0, a1);             // conj(U1) proj(t) of row j-3, finished at step j-2
                const C o0 = mkc<C>(mass * Tr[t2_i][0].x - half * a0.x, mass * Tr[t2_i][0].y - half * a0.y);
                const C o1 = mkc<C>(mass * Tr[t2_i][1].x - half * a1.x, mass * Tr[t2_i][1].y - half * a1.y);
                const int n = (j - 2) * wt + t;      // xa <= j-2 < xb: no wrap
                st_stream(a.out + n, o0);
                st_stream(a.out + V + n, o1);
                if (MODE != FUSED_PLAIN) {           // dot(psi, out) = sum psi conj(out)
                    const C q0 = cmul_conj(P[m2][0], o0), q1 = cmul_conj(P[m2][1], o1);
                    acc[0] += (double)q0.x + (double)q1.x;
                    acc[1] += (double)q0.y + (double)q1.y;
                }
            }
            HX[tn_i] = cmulc(U1w[m1], Hop<false>::from_xm(tn0, tn1));   // -x hop term that out(row j) takes at step j+2
            Tr[tn_i][0] = tn0;
            Tr[tn_i][1] = tn1;
        }
    }

    if (MODE != FUSED_PLAIN) {
        if (grid_reduce<2>(acc, a.partials, a.ticket, (int)gridDim.x * a.nchunks, chunk * (int)gridDim.x + (int)blockIdx.x)) {
            if (tid < 32) fused_sums_out(a, acc);
        }
    }
}

// r -= alpha A d ; |r|^2 ; alpha kept for the x update that the next fused pass applies
template <typename C>
__global__ void __launch_bounds__(kBlock) k_cg_resid(CgState* st, int cur, C* __restrict__ r, const C* __restrict__ Ad,
                                                     int n_elems, double* partials, unsigned int* ticket,
                                                     double* sums_out) {
    typedef typename RealOf<C>::type R;
    pdl_wait();                 // (opt-in programmatic dependent launch: pass A's A d and dot(d, A d) are complete)
    pdl_launch_dependents();
    if (st->done) return;
    const cplx alpha = cdiv(make_double2(st->rr[cur], 0.0), make_double2(st->dAd[0], st->dAd[1]));
    const C al = mkc<C>((R)alpha.x, (R)alpha.y);
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const C av = ld_stream(Ad + i);
        C rv = r[i];
        rv = csub(rv, cmul(al, av));
        r[i] = rv;
        acc[0] += (double)rv.x * rv.x + (double)rv.y * rv.y;
    }
    if (grid_reduce<1>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            st->alpha[0] = alpha.x;
            st->alpha[1] = alpha.y;
            st->pending = 1;        // x still lacks alpha_k d_k
            st->pending_buf = cur;  // d_k lives in d buffer (k & 1)
            st->k = st->k + 1;
        }
    }
}

// The same on a lattice split along x with peer-memory sums and halos (sm_peer.cuh): dot(d, A d) is gathered from the
// ranks' slots; the two boundary rows of r on each side are updated FIRST and stored into the neighbours' ghost rows as
// they are formed ([component][2 rows][wt], parity of iteration k+1), then the interior; the finishing block
// publishes |r|^2 and raises the neighbours' ghost flags (epoch base + k + 1).
__global__ void __launch_bounds__(kBlock) k_cg_resid_dist(CgState* st, int cur, cplx* __restrict__ r, const cplx* __restrict__ Ad,
                                                          int wx, int wt, int V, double* partials, unsigned int* ticket,
                                                          const DistLink dl) {
    if (st->done) return;
    double g[2];
    gather_sums<2>(dl, 0, cur, st->epoch_base + (unsigned int)st->k + 1u, g);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->dAd[0] = g[0];
        st->dAd[1] = g[1];
    }
    const cplx alpha = cdiv(make_double2(st->rr[cur], 0.0), make_double2(g[0], g[1]));
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
    cplx* __restrict__ to_xm = dl.push_xm_hi[cur ^ 1];
    cplx* __restrict__ to_xp = dl.push_xp_lo[cur ^ 1];
    const int per_side = 4 * wt;
    for (int i = gtid; i < 2 * per_side; i += stride) {
        const int side = i / per_side, e = i - side * per_side;
        const int comp = e / (2 * wt), off = e - comp * 2 * wt;
        const size_t idx = (size_t)comp * V + (side == 0 ? 0 : (size_t)(wx - 2) * wt) + off;
        const cplx av = ld_stream(Ad + idx);
        cplx rv = r[idx];
        rv = csub(rv, cmul(alpha, av));
        r[idx] = rv;
        (side == 0 ? to_xm : to_xp)[e] = rv;
        acc[0] += rv.x * rv.x + rv.y * rv.y;
    }
    const int vint = (wx - 4) * wt;     // rows 2 .. wx-3 of one component
    for (int i = gtid; i < 2 * vint; i += stride) {
        const int comp = (i >= vint) ? 1 : 0;
        const size_t idx = (size_t)comp * V + 2 * (size_t)wt + (size_t)(i - comp * vint);
        const cplx av = ld_stream(Ad + idx);
        cplx rv = r[idx];
        rv = csub(rv, cmul(alpha, av));
        r[idx] = rv;
        acc[0] += rv.x * rv.x + rv.y * rv.y;
    }
    if (grid_reduce<1>(acc, partials, ticket, -1, -1, true)) {
        if (threadIdx.x < 32) {
            const unsigned int e = st->epoch_base + (unsigned int)st->k + 1u;   // every lane reads k before lane 0 bumps it
            __syncwarp();
            if (threadIdx.x == 0) {
                st->alpha[0] = alpha.x;
                st->alpha[1] = alpha.y;
                st->pending = 1;
                st->pending_buf = cur;
                st->k = st->k + 1;
            }
            // |r|^2 to every rank's slot and the neighbours' ghost flags, behind one system-scope fence (every block
            // fenced its own peer stores before it took its ticket)
            publish_sums<1>(dl, 1, cur, e, acc, dl.flag_xm, dl.flag_xp);
        }
    }
}

// the x update the loop still owes when it stops: x += alpha_K d_K
template <typename C>
__global__ void __launch_bounds__(kBlock) k_cg_flush_x(CgState* st, C* __restrict__ x, const C* __restrict__ d_buf0,
                                                       const C* __restrict__ d_buf1, int n_elems) {
    typedef typename RealOf<C>::type R;
    if (!st->pending) return;
    const C alpha = mkc<C>((R)st->alpha[0], (R)st->alpha[1]);
    const C* __restrict__ d = st->pending_buf ? d_buf1 : d_buf0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        x[i] = cadd(x[i], cmul(alpha, ld_stream(d + i)));
    }
}

// ---- opt-in mixed-precision solver: single-precision inner CG inside a double-precision defect correction ----
__global__ void __launch_bounds__(kBlock) k_to_single(const cplx* __restrict__ in, cplxf* __restrict__ out, int n_elems) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx v = ld_stream(in + i);
        out[i] = make_float2((float)v.x, (float)v.y);
    }
}

// r = phi - A x (true residual, double) ; r32 = r ; e32 = 0 ; sums: |phi|^2, |r|^2
__global__ void __launch_bounds__(kBlock) k_mixed_residual(const cplx* __restrict__ phi, const cplx* __restrict__ Ax,
                                                           cplxf* __restrict__ r32, cplxf* __restrict__ e32, int n_elems,
                                                           double* partials, unsigned int* ticket, double* sums_out) {
    double acc[2] = {0.0, 0.0};
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplx f = ld_stream(phi + i), a = ld_stream(Ax + i);
        const cplx r = csub(f, a);
        r32[i] = make_float2((float)r.x, (float)r.y);
        e32[i] = make_float2(0.f, 0.f);
        acc[0] += f.x * f.x + f.y * f.y;
        acc[1] += r.x * r.x + r.y * r.y;
    }
    if (grid_reduce<2>(acc, partials, ticket)) {
        if (threadIdx.x == 0) {
            sums_out[0] = acc[0];
            sums_out[1] = acc[1];
        }
    }
}

// x += e (double += single)
__global__ void __launch_bounds__(kBlock) k_mixed_correct(cplx* __restrict__ x, const cplxf* __restrict__ e32, int n_elems) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
        const cplxf e = ld_stream(e32 + i);
        cplx v = x[i];
        v.x += (double)e.x;
        v.y += (double)e.y;
        x[i] = v;
    }
}

// inner solve starts from e = 0: r_in = r_outer, |r_in|^2 known; stop at |r_in| < delta |r_outer|
__global__ void k_mixed_begin(CgState* st, const double* sums /* |phi|^2, |r|^2 */, double delta, int max_iter) {
    st->max_iter = max_iter;
    st->done = 0;
    st->iters = 0;
    st->converged = 0;
    st->pending = 0;
    st->pending_buf = 0;
    st->k = 0;
    st->phi_norm2 = sums[1];
    st->rr[0] = sums[1];
    st->tol = delta;
}

// ----------------------------------------------------------------------------------------------
// Halo push over NVLink peer memory (lattice split along x).  Instead of a send/recv pair, the
// rank that owns the boundary rows stores them straight into the neighbour's ghost arrays
// (peer pointers from cudaIpcOpenMemHandle) and then raises an epoch flag in the neighbour's
// memory with a system-scope release; the neighbour's stream waits on its own flag
// (cuStreamWaitValue32) before the boundary bands of the pass start.  Ghosts are double-buffered
// on the epoch parity, so a neighbour that runs one pass ahead never overwrites rows still in use.
//   rows 0,1        -> -x neighbour's "hi" ghost        rows wx-2,wx-1 -> +x neighbour's "lo" ghost
//   ghost layout [component][2 rows][wt], the same as the NCCL path
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_push_rows(const cplx* __restrict__ field, int wx, int wt, int V,
                                                      cplx* __restrict__ peer_xm_hi, cplx* __restrict__ peer_xp_lo,
                                                      unsigned int* flag_xm, unsigned int* flag_xp,
                                                      unsigned int epoch, unsigned int* ticket) {
    const int per_side = 4 * wt;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_side; i += stride) {
        const int side = i / per_side, e = i - side * per_side;
        const int comp = e / (2 * wt), off = e - comp * 2 * wt;
        const cplx v = ld_stream(field + (size_t)comp * V + (side == 0 ? 0 : (size_t)(wx - 2) * wt) + off);
        (side == 0 ? peer_xm_hi : peer_xp_lo)[e] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {           // every block's rows are out: publish the epoch
            *ticket = 0u;
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_xm), "r"(epoch) : "memory");
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_xp), "r"(epoch) : "memory");
        }
    }
}

// The same for a lattice split along t: columns 0,1 -> the -t neighbour's "hi" ghost columns, columns wt-2, wt-1 -> the +t
// neighbour's "lo"; ghost layout [component][wx rows][2] (what k_pack_cols2 + send/recv deliver on the NCCL path).
__global__ void __launch_bounds__(kBlock) k_push_cols(const cplx* __restrict__ field, int wx, int wt, int V,
                                                      cplx* __restrict__ peer_tm_hi, cplx* __restrict__ peer_tp_lo,
                                                      unsigned int* flag_tm, unsigned int* flag_tp,
                                                      unsigned int epoch, unsigned int* ticket) {
    const int per_side = 4 * wx;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per_side; i += stride) {
        const int side = i / per_side, e = i - side * per_side;          // e = (comp, row, c)
        const int comp = e / (2 * wx), q = e - comp * 2 * wx, row = q >> 1, cc = q & 1;
        const cplx v = ld_stream(field + (size_t)comp * V + (size_t)row * wt + (side == 0 ? cc : wt - 2 + cc));
        (side == 0 ? peer_tm_hi : peer_tp_lo)[e] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {           // every block's columns are out: publish the epoch
            *ticket = 0u;
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_tm), "r"(epoch) : "memory");
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_tp), "r"(epoch) : "memory");
        }
    }
}

}  // namespace sm
