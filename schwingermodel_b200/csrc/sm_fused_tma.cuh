// sm_fused_tma.cuh -- the one-pass D D^dagger of sm_fused.cuh with rows staged by TMA bulk copies.
//
// Same blocking as k_dd_fused (a block owns a strip of columns and marches down the x rows; psi rows, t rows and
// links of the three rows in flight live in register rings; the t-direction neighbours travel through a double-
// buffered shared-memory line, one __syncthreads per row).  What changes is everything around the arithmetic:
//   * a row of an array (BT contiguous 16-byte elements = one 4 KB line) reaches shared memory as ONE
//     cp.async.bulk.shared.global (SASS UBLKCP) issued by lane 0 of a warp -- warp w issues array w -- and completes
//     on the stage's mbarrier (complete_tx); the 256 threads no longer compute four global addresses and issue four
//     LDGSTS each per row.  Strips that touch the lattice edge add two 32-byte copies for the wrapped halo columns.
//   * the factor -1/2 of the hop sum and the antiperiodic sign are folded into the links once per row when the row
//     is taken from the stage (exact: a power of two), so  D psi = mass psi + sum(hops)  starts its accumulators with
//     an FMA and the closing  mass psi - 1/2 acc  disappears: 84 FP64 instructions per site-update instead of 90.
//   * the stopping test, split-lattice row selection and wrap logic live in the issuing lane only.
// Step j (after the barrier of step j-1 every thread has finished reading the slot of row j-1):
//     publish the t-direction half-spinors of psi(j-1), t(j-2)    [registers of earlier rows only]
//     wait for row j on its mbarrier                               [overlaps the barrier]
//     __syncthreads
//     issue row j+STAGES-1 into the slot of row j-1
//     take row j, form t(j-1) = D^dagger psi and out(j-2) = D t
// Double precision only (bulk copies need 16-byte granules); the single-precision inner solve of the opt-in mixed
// solver keeps k_dd_fused.
#pragma once
#include "sm_fused.cuh"

namespace sm {

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16; both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned int bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_u32(unsigned int smem_dst, const void* gmem_src, unsigned int bytes, unsigned int bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(bar)
                 : "memory");
}
// wait for the phase with this parity; a copy that never completes (a bug, or a fault on the source) traps after a
// few seconds instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_u32(unsigned int bar, unsigned int parity) {
    unsigned int ok;
    int spins = 0;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1 << 22)) __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

constexpr size_t fused_tma_smem_bytes(int mode, int stages, int BT) {
    return sizeof(cplx) * (size_t)BT * (2 * 4 + (size_t)stages * fused_arrays(mode));
}

template <int MODE, int STAGES>
__global__ void __launch_bounds__(kBlock, 2) k_dd_tma(const FusedArgsT<cplx> a) {
    typedef cplx C;
    typedef double R;
    extern __shared__ __align__(128) unsigned char s_raw_tma[];
    __shared__ __align__(8) uint64_t s_bar[STAGES];
    C* const s_mem = reinterpret_cast<C*>(s_raw_tma);
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
        // (opt-in programmatic dependent launch, sm_common.cuh: nothing above this line touches global memory)
        pdl_wait();
        pdl_launch_dependents();
        // one lead thread per pass: in the launch that holds chunk 0 of strip 0
        const bool lead = blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && a.chunk_mode != 1 && a.strip_mode != 1;
        if (!fused_cg_begin(a, lead, beta, alpha, first)) return;
        fused_cg_wait_ghosts(a);      // peer-memory halos: r_k's ghost rows have arrived (before any copy reads them)
    }

    // columns of this block: the strip + 2 halo columns each side, c_lo <= tc < c_lo + ncols (unwrapped)
    const int strip = (a.strip_mode == 1) ? (int)blockIdx.x + 1
                                          : (a.strip_mode == 2 ? (blockIdx.x == 0 ? 0 : a.nstrips - 1) : (int)blockIdx.x);
    const int c_lo = strip * a.cols_per_strip - 2;
    const int ncols = min(a.cols_per_strip + 4, wt + 2 - c_lo);
    const int tc = c_lo + tid;
    const int t = wrap_idx(tc, wt);
    const bool col_active = tid < ncols;
    const bool col_owner = (tid >= 2) && (tid < a.cols_per_strip + 2) && (tc < wt);
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

    // ---- staging: one lane per array issues that array's row (lane 0 of warp w: array w; with fewer warps than
    //      arrays lane 16 of warp w takes array w + #warps).  Where row j of an array lives -- the tile, wrapped in x
    //      on a single tile, or a 2-row ghost array on a lattice split along x -- is tabulated once in shared memory
    //      as three row pointers per array: row -2 ("lo"), row 0 (tile), row wx ("hi").
    //      With ghost columns (lattice split along t) the two halo columns each side of the tile come from packed
    //      [row][2] arrays (or, in ghost rows of a 2-D split, from the corner entries of the widened ghost-row arrays)
    //      instead of from the wrapped columns of the same row: three pointers per array and region -- main piece,
    //      left piece (column -2), right piece (column wt) -- and their row strides.
    __shared__ const C* s_src[8][3][3];     // [array][region: rows < 0, tile, rows >= wx][piece: main, left, right]
    __shared__ int s_stride[3][2];          // [region][main, side] row stride in elements
    const int nwarps = BT >> 5;
    const int lane = tid & 31, warp = tid >> 5;
    const int my_arr = (NARR > nwarps) ? warp + nwarps * (lane >> 4) : warp;
    const bool issuer = ((NARR > nwarps) ? (lane & 15) == 0 : lane == 0) && my_arr < NARR;
    const bool split = (a.gU_lo != nullptr);
    const bool tg = (a.tg_on != 0);
    const int row_w = tg ? a.row_w : wt;            // width of the ghost-row arrays
    const int row_off = (row_w - wt) >> 1;          // their column of tile column 0 (2 with corners)
    if (tid < NARR) {
        const int kind = tid >> 1, comp = tid & 1;   // kind 0: U ; 1: in (PLAIN/DOT) or r (CG) ; 2: d_{k-1} ; 3: x
        const C *tile, *lo, *hi, *cl, *ch;
        if (kind == 0) { tile = a.U; lo = a.gU_lo; hi = a.gU_hi; cl = a.tgU_lo; ch = a.tgU_hi; }
        else if (kind == 1 && MODE == FUSED_CG) { tile = a.r; lo = a.gr_lo; hi = a.gr_hi; cl = a.tgr_lo; ch = a.tgr_hi; }
        else if (kind == 3) { tile = a.x; lo = hi = cl = ch = nullptr; }   // own rows, owner columns only
        else { tile = a.in; lo = a.gin_lo; hi = a.gin_hi; cl = a.tgin_lo; ch = a.tgin_hi; }
        tile += (size_t)comp * V;
        const bool colg = tg && cl != nullptr;
        if (colg) {
            cl += (size_t)comp * 2 * wx;
            ch += (size_t)comp * 2 * wx;
        }
        // tile rows
        s_src[tid][1][0] = tile;
        s_src[tid][1][1] = colg ? cl : tile + (wt - 2);
        s_src[tid][1][2] = colg ? ch : tile;
        // rows -2, -1 and wx, wx+1: ghost rows of a split along x, else the tile's own rows wx-2.. / 0..
        if (split && lo) {
            const C* g = lo + (size_t)comp * 2 * row_w;
            s_src[tid][0][0] = g + row_off;
            s_src[tid][0][1] = (row_off > 0) ? g : g + (wt - 2);
            s_src[tid][0][2] = (row_off > 0) ? g + row_off + wt : g;
            g = hi + (size_t)comp * 2 * row_w;
            s_src[tid][2][0] = g + row_off;
            s_src[tid][2][1] = (row_off > 0) ? g : g + (wt - 2);
            s_src[tid][2][2] = (row_off > 0) ? g + row_off + wt : g;
        } else {
            s_src[tid][0][0] = tile + (size_t)(wx - 2) * wt;
            s_src[tid][0][1] = colg ? cl + (size_t)(wx - 2) * 2 : tile + (size_t)(wx - 2) * wt + (wt - 2);
            s_src[tid][0][2] = colg ? ch + (size_t)(wx - 2) * 2 : tile + (size_t)(wx - 2) * wt;
            s_src[tid][2][0] = tile;
            s_src[tid][2][1] = colg ? cl : tile + (wt - 2);
            s_src[tid][2][2] = colg ? ch : tile;
        }
    }
    if (tid == 0) {
        const int side_tile = tg ? 2 : wt;
        s_stride[1][0] = wt;
        s_stride[1][1] = side_tile;
        for (int reg = 0; reg < 3; reg += 2) {
            s_stride[reg][0] = split ? row_w : wt;
            s_stride[reg][1] = split ? row_w : side_tile;
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < STAGES; q++) mbar_init(&s_bar[q], (unsigned int)NARR);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const C zero = mkc<C>(0, 0);
    if (!col_active) {
        for (int q = 0; q < STAGES * NARR; q++) s_stage[q * BT + tid] = zero;   // never written by a copy: read back as zeros
    }
    __syncthreads();
    if (split && issuer) asm volatile("fence.proxy.async;" ::: "memory");   // ghost rows written by a peer's stores, read by bulk copies

    // pieces of a row: [main] columns max(c_lo,0) .. min(c_lo+ncols, wt), [left] the wrapped columns c_lo..-1,
    // [right] the wrapped columns wt .. c_lo+ncols-1
    const int m_lo = max(c_lo, 0), m_hi = min(c_lo + ncols, wt);
    const int n_left = (c_lo < 0) ? -c_lo : 0;
    const int n_right = max(c_lo + ncols - wt, 0);
    const unsigned int stage_bytes = (unsigned int)(NARR * BT * (int)sizeof(C));
    const unsigned int my_dst = smem_u32(s_stage) + (unsigned int)(my_arr * BT * (int)sizeof(C));   // slot 0
    const unsigned int bar0 = smem_u32(&s_bar[0]);
    const int my_kind = my_arr >> 1;
    auto issue_row = [&](int j, int slot) {
        if (!issuer || j > j_last) return;
        // kinds 0, 1 every row ; d_{k-1} not in the first iteration ; x on the block's own rows only
        const bool wanted = my_kind < 2 || (!first && (my_kind == 2 || (j >= xa && j < xb)));
        const int reg = (j < 0) ? 0 : (j >= wx ? 2 : 1);
        const int rel = (j < 0) ? j + 2 : (j >= wx ? j - wx : j);
        const C* const* tab = s_src[my_arr][reg];
        const int st_main = s_stride[reg][0];
        const int st_side = (my_kind == 3) ? st_main : s_stride[reg][1];     // x has no ghost columns: its halo is never used
        const C* src = tab[0] + (size_t)rel * st_main;
        const unsigned int bar = bar0 + 8u * slot, dst = my_dst + stage_bytes * slot;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(wanted ? (unsigned int)(ncols * (int)sizeof(C)) : 0u)
                     : "memory");
        if (wanted) {
            bulk_g2s_u32(dst + (unsigned int)((m_lo - c_lo) * (int)sizeof(C)), src + m_lo,
                         (unsigned int)((m_hi - m_lo) * (int)sizeof(C)), bar);
            if (n_left) bulk_g2s_u32(dst, tab[1] + (size_t)rel * st_side, (unsigned int)(n_left * (int)sizeof(C)), bar);
            if (n_right)
                bulk_g2s_u32(dst + (unsigned int)((wt - c_lo) * (int)sizeof(C)), tab[2] + (size_t)rel * st_side,
                             (unsigned int)(n_right * (int)sizeof(C)), bar);
        }
    };

    // Register rings (compile-time slots after the 6-fold unroll): P psi rows j, j-1, j-2 ; Tr t rows j-1, j-2 ;
    // HX the finished -x hop term of the row below ; U0w/U1w links of rows j, j-1, j-2 times -1/2 (U0w also times
    // the antiperiodic sign: the +t hop out of column wt-1 and the -t hop into column 0 are the same link).
    C P[3][2], Tr[2][2], HX[2], U0w[3], U1w[3];
#pragma unroll
    for (int q = 0; q < 3; q++) P[q][0] = P[q][1] = U0w[q] = U1w[q] = zero;
#pragma unroll
    for (int q = 0; q < 2; q++) Tr[q][0] = Tr[q][1] = HX[q] = zero;
    double acc[2] = {0.0, 0.0};
    const R mass = (R)a.mass;
    // the time link of column -1 (a ghost column, or the wrapped column wt-1) carries the sign of the -t hop into column 0,
    // that of column wt-1 the sign of the +t hop out of it; on a tile that is whole in t both are the same link
    const R h0 = (R)(-0.5 * ((tc == -1) ? a.sL_edge : ((t == wt - 1) ? a.sR_edge : 1.0))), h1 = (R)(-0.5);

#pragma unroll
    for (int q = 0; q < STAGES - 1; q++) issue_row(j_first + q, q);

    int slot = 0;                 // stage slot of row j
    unsigned int parity = 0;      // its mbarrier phase
    C* const out_col = a.out + t;     // this thread's column of the output
    int n_out = (xa - 4) * wt;        // row j-2 (stored from j = xa+2 on, i.e. n_out >= xa*wt)
    for (int jb = j_first; jb <= j_last; jb += 6) {
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const int j = jb + s;
            if (j > j_last) break;                         // uniform over the block
            const int c = s % 3, m1 = (s + 2) % 3, m2 = (s + 1) % 3;   // ring slots of rows j, j-1, j-2
            const int tn_i = s % 2, t2_i = (s + 1) % 2;

            // publish the t-direction half-spinors of psi row j-1 (for D^dagger) and t row j-2 (for D)
            C* line = s_line + (s & 1) * 4 * BT;
            line[0 * BT + tid] = Hop<true>::from_tp(P[m1][0], P[m1][1]);                       // read by column t-1
            line[1 * BT + tid] = cmulc(U0w[m1], Hop<true>::from_tm(P[m1][0], P[m1][1]));       // read by column t+1
            line[2 * BT + tid] = Hop<false>::from_tp(Tr[t2_i][0], Tr[t2_i][1]);
            line[3 * BT + tid] = cmulc(U0w[m2], Hop<false>::from_tm(Tr[t2_i][0], Tr[t2_i][1]));
            mbar_wait_u32(bar0 + 8u * slot, parity);                // row j has landed
            __syncthreads();
            {   // every thread is past step j-1: its slot is free for row j + STAGES - 1
                const int pslot = (slot == 0) ? STAGES - 1 : slot - 1;
                issue_row(j + STAGES - 1, pslot);
            }
            const C* st = s_stage + slot * NARR * BT + tid;
            U0w[c] = cscale(h0, st[0 * BT]);
            U1w[c] = cscale(h1, st[1 * BT]);
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
                    a.gd_lo[(j + 2) * row_w + row_off + t] = p0;
                    a.gd_lo[2 * row_w + (j + 2) * row_w + row_off + t] = p1;
                } else if (split && j >= wx) {
                    a.gd_hi[(j - wx) * row_w + row_off + t] = p0;
                    a.gd_hi[2 * row_w + (j - wx) * row_w + row_off + t] = p1;
                }
            } else if (MODE == FUSED_CG && tg && col_active && (tc < 0 || tc >= wt)) {
                // a halo column of the TILE (edge strips of a lattice split along t): d_k's ghost columns, and in the
                // ghost rows of a 2-D split the corner entries
                const int ci = (tc < 0) ? tc + 2 : tc - wt;
                if (j >= xa && j < xb) {
                    C* g = (tc < 0) ? a.tgd_lo : a.tgd_hi;
                    g[j * 2 + ci] = p0;
                    g[2 * wx + j * 2 + ci] = p1;
                } else if (split && row_off > 0 && (j < 0 || j >= wx)) {
                    C* g = (j < 0) ? a.gd_lo : a.gd_hi;
                    const int rr_ = (j < 0) ? j + 2 : j - wx;
                    const int cc = (tc < 0) ? tc + 2 : row_off + tc;
                    g[rr_ * row_w + cc] = p0;
                    g[2 * row_w + rr_ * row_w + cc] = p1;
                }
            }
            P[c][0] = p0;
            P[c][1] = p1;
            slot = (slot == STAGES - 1) ? 0 : slot + 1;
            parity ^= (slot == 0) ? 1u : 0u;

            // t(j-1) = D^dagger psi at row j-1 = mass psi + sum of the (pre-scaled) hop terms
            C tn0, tn1;
            {
                const C v = cmul(U0w[m1], line[0 * BT + tr]);     // +t
                C a0 = mkc<C>(fma(mass, P[m1][0].x, v.x), fma(mass, P[m1][0].y, v.y));
                C a1 = mkc<C>(fma(mass, P[m1][1].x, v.x), fma(mass, P[m1][1].y, v.y));     // D^dagger: +v
                Hop<true>::add_xp(cmul(U1w[m1], Hop<true>::from_xp(p0, p1)), a0, a1);
                Hop<true>::add_tm(line[1 * BT + tl], a0, a1);
                Hop<true>::add_xm(cmulc(U1w[m2], Hop<true>::from_xm(P[m2][0], P[m2][1])), a0, a1);
                tn0 = a0;
                tn1 = a1;
            }
            // out(j-2) = D t at row j-2
            if (j >= xa + 2 && col_owner) {
                const C v = cmul(U0w[m2], line[2 * BT + tr]);     // +t
                C a0 = mkc<C>(fma(mass, Tr[t2_i][0].x, v.x), fma(mass, Tr[t2_i][0].y, v.y));
                C a1 = mkc<C>(fma(mass, Tr[t2_i][1].x, -v.x), fma(mass, Tr[t2_i][1].y, -v.y));   // D: -v
                Hop<false>::add_xp(cmul(U1w[m2], Hop<false>::from_xp(tn0, tn1)), a0, a1);
                Hop<false>::add_tm(line[3 * BT + tl], a0, a1);
                Hop<false>::add_xm(HX[tn_i], a0, a1);             // conj(U1) proj(t) of row j-3, finished at step j-2
                st_stream(out_col + n_out, a0);
                st_stream(out_col + V + n_out, a1);
                if (MODE != FUSED_PLAIN) {           // dot(psi, out) = sum psi conj(out)
                    const C q0 = cmul_conj(P[m2][0], a0), q1 = cmul_conj(P[m2][1], a1);
                    acc[0] += (double)q0.x + (double)q1.x;
                    acc[1] += (double)q0.y + (double)q1.y;
                }
            }
            n_out += wt;
            HX[tn_i] = cmulc(U1w[m1], Hop<false>::from_xm(tn0, tn1));   // -x hop term that out(row j) takes at step j+2
            Tr[tn_i][0] = tn0;
            Tr[tn_i][1] = tn1;
        }
    }

    if (MODE != FUSED_PLAIN) {
        const int nparts = a.part_total > 0 ? a.part_total : (int)gridDim.x * a.nchunks;
        const int part = a.part_total > 0 ? a.part_base + (int)(blockIdx.y * gridDim.x + blockIdx.x) : chunk * (int)gridDim.x + (int)blockIdx.x;
        if (grid_reduce<2>(acc, a.partials, a.ticket, nparts, part)) {
            if (tid < 32) fused_sums_out(a, acc);
        }
    }
}

}  // namespace sm
