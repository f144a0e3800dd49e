// sm_ops.cuh -- operator launches on device fields: Wilson stencil, one-pass D D^dagger (+ 2-row ghosts, peer-memory halos), dot.
// Part of the single translation unit sm_abi.cu (static functions, included in dependency order).
#pragma once
#include "sm_dist.cuh"

// ------------------------------------------------------------------------------------------------
// operator launches on device fields
// ------------------------------------------------------------------------------------------------
template <bool DAG, int MODE>
static int launch_wilson(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, const cplx* aux = nullptr,
                         cplx* r = nullptr, cplx* d = nullptr, cplx* x = nullptr, double* sums_out = nullptr,
                         const int* done = nullptr) {
    // split lattice: pack the halo lines, exchange them on the comm stream while the interior sites compute on
    // the compute stream, then the boundary sites on the comm stream (SM_OVERLAP=0: exchange first, one launch)
    const bool overlap = c->dist() && c->overlap && c->wx >= 4 && c->wt >= 4;
    if (c->dist()) {
        if (overlap) {
            TRY((pack_spinor_halo<DAG>(c, U, in, done)));
            CU(cudaEventRecord(c->ev_ready, c->stream));
            CU(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
            TRY(exchange_spinor_lines(c, c->comm_stream));
        } else {
            TRY((pack_spinor_halo<DAG>(c, U, in, done)));
            TRY(exchange_spinor_lines(c, c->stream));
        }
    }
    WilsonArgs a{};
    a.U = U;
    a.in = in;
    a.out = out;
    a.aux = aux;
    a.aux2 = (MODE == WILSON_CGINIT) ? c->cg_x0 : nullptr;
    a.r = r;
    a.d = d;
    a.x = x;
    a.wx = c->wx;
    a.wt = c->wt;
    a.V = c->V;
    a.rows_per_block = (MODE == WILSON_PLAIN) ? c->rows_per_block_plain : c->rows_per_block;
    a.mass = m0 + 2;
    a.sR_edge = c->sR_edge();
    a.sL_edge = c->sL_edge();
    a.g_tp = c->rt > 1 ? c->g_tp : nullptr;
    a.g_tm = c->rt > 1 ? c->g_tm : nullptr;
    a.g_xp = c->rx > 1 ? c->g_xp : nullptr;
    a.g_xm = c->rx > 1 ? c->g_xm : nullptr;
    a.partials = c->partials;
    a.ticket = c->tickets + TK_WILSON;
    a.sums_out = sums_out;
    a.done = done;
    const dim3 grid = (MODE == WILSON_PLAIN) ? c->wil_grid_plain : c->wil_grid;
    if (overlap) {
        const int nsites = (c->rt > 1 ? 2 * c->wx : 0) + (c->rx > 1 ? 2 * (c->rt > 1 ? c->wt - 2 : c->wt) : 0);
        const int bblocks = std::max(1, std::min(kWilsonBoundaryBlocks, (nsites + kBlock - 1) / kBlock));
        a.interior_blocks = (int)(grid.x * grid.y);
        a.boundary_blocks = bblocks;
        k_wilson_boundary<DAG, MODE><<<bblocks, kBlock, 0, c->comm_stream>>>(a);
        KCHECK();
        CU(cudaEventRecord(c->ev_ghost, c->comm_stream));
        a.interior_only = 1;
        k_wilson<DAG, MODE><<<grid, c->wil_block, 0, c->stream>>>(a);
        KCHECK();
        CU(cudaStreamWaitEvent(c->stream, c->ev_ghost, 0));
        c->launches += 2;
    } else {
        k_wilson<DAG, MODE><<<grid, c->wil_block, 0, c->stream>>>(a);
        KCHECK();
        c->launches++;
    }
    return SM_OK;
}

// ---- even-odd building block (single tile): out = [parity == keep] (self aux + hop D in), see WILSON_EO ----------------
template <bool DAG, int MODE>
static int launch_wilson_eo(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, int keep, const cplx* aux,
                            double self, double hop, const cplx* dot_with = nullptr, double* sums_out = nullptr,
                            const int* done = nullptr) {
    static_assert(MODE == WILSON_EO || MODE == WILSON_EO_DOT, "even-odd modes only");
    if (c->dist()) return fail(SM_ERR_STATE, "the even-odd solver runs on a single tile");
    WilsonArgs a{};
    a.U = U;
    a.in = in;
    a.out = out;
    a.aux = aux;
    a.aux2 = dot_with;
    a.wx = c->wx;
    a.wt = c->wt;
    a.V = c->V;
    a.rows_per_block = c->rows_per_block;
    a.mass = m0 + 2;
    a.sR_edge = c->sR_edge();
    a.sL_edge = c->sL_edge();
    a.partials = c->partials;
    a.ticket = c->tickets + TK_WILSON;
    a.sums_out = sums_out;
    a.done = done;
    a.eo_keep = keep;
    a.eo_self = self;
    a.eo_hop = hop;
    k_wilson<DAG, MODE><<<c->wil_grid, c->wil_block, 0, c->stream>>>(a);
    KCHECK();
    c->launches++;
    return SM_OK;
}

// Schur complement of D on the even sites,  Dhat = m - (1/4m) H_eo H_oe  (m = m0 + 2, D = m - H/2), or its adjoint, applied
// to a field that is zero on the odd sites:  W = [odd] D v ;  out = [even] (m v - (1/m) D W).  `dot_with`: also dot(dot_with, out).
template <bool DAG>
static int dev_Dhat(sm_ctx* c, const cplx* U, const cplx* v, cplx* W, cplx* out, double m0, const cplx* dot_with = nullptr,
                    double* sums_out = nullptr, const int* done = nullptr) {
    const double m = m0 + 2;
    TRY((launch_wilson_eo<DAG, WILSON_EO>(c, U, v, W, m0, 1, nullptr, 0.0, 1.0, nullptr, nullptr, done)));
    if (dot_with != nullptr)
        return launch_wilson_eo<DAG, WILSON_EO_DOT>(c, U, W, out, m0, 0, v, m, -1.0 / m, dot_with, sums_out, done);
    return launch_wilson_eo<DAG, WILSON_EO>(c, U, W, out, m0, 0, v, m, -1.0 / m, nullptr, nullptr, done);
}

static int dev_D(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, bool dagger) {
    if (in == out) return fail(SM_ERR_ARG, "D: in and out must not alias");
    if (dagger) return launch_wilson<true, WILSON_PLAIN>(c, U, in, out, m0);
    return launch_wilson<false, WILSON_PLAIN>(c, U, in, out, m0);
}

// ---- peer-memory window ------------------------------------------------------------------------
// (a window serves a lattice split along ONE axis: ghost rows [comp][2][wt] of a split along x, ghost columns [comp][wx][2] of a
// split along t; "lower" / "upper" neighbour = -x / +x or -t / +t)
static bool win_along_t(const sm_ctx* c) { return c->rt > 1; }
static size_t win_ghost_elems(const sm_ctx* c) { return 4 * (size_t)(win_along_t(c) ? c->wx : c->wt); }
static cplx* win_ghost(const sm_ctx* c, void* base, int kind, int parity, int side) {
    return (cplx*)base + (size_t)((kind * 2 + parity) * 2 + side) * win_ghost_elems(c);
}
static unsigned int* win_flag(const sm_ctx* c, void* base, int kind, int side) {
    return (unsigned int*)((char*)base + sizeof(cplx) * 8 * win_ghost_elems(c)) + kind * 2 + side;
}
static SumSlot* win_slots(const sm_ctx* c, void* base) {
    return (SumSlot*)((char*)base + sizeof(cplx) * 8 * win_ghost_elems(c) + 256);
}
static size_t win_total_bytes(const sm_ctx* c) {
    return sizeof(cplx) * 8 * win_ghost_elems(c) + 256 + sizeof(SumSlot) * 4 * kMaxPeers;
}

// what the CG kernels need to do their own collectives (sm_peer.cuh)
static DistLink dist_link(const sm_ctx* c) {
    DistLink dl{};
    if (!c->peer_sums) return dl;
    dl.on = 1;
    dl.nranks = c->nranks;
    dl.rank = c->rank;
    dl.mine = win_slots(c, c->win);
    for (int r = 0; r < c->nranks; r++) dl.peer[r] = win_slots(c, c->peer_all[r]);
    for (int parity = 0; parity < 2; parity++) {
        dl.push_xm_hi[parity] = win_ghost(c, c->peer_win[0], 1, parity, 1);
        dl.push_xp_lo[parity] = win_ghost(c, c->peer_win[1], 1, parity, 0);
    }
    dl.flag_xm = win_flag(c, c->peer_win[0], 1, 1);
    dl.flag_xp = win_flag(c, c->peer_win[1], 1, 0);
    dl.my_flag_lo = win_flag(c, c->win, 1, 0);
    dl.my_flag_hi = win_flag(c, c->win, 1, 1);
    return dl;
}

// push my boundary rows of `field` into both neighbours' ghosts (epoch parity) and raise their flags
static int p2p_push(sm_ctx* c, const cplx* field, int kind, cudaStream_t st, long long fixed_epoch = -1, int fixed_parity = 0) {
    const unsigned int epoch = fixed_epoch >= 0 ? (unsigned int)fixed_epoch : ++c->p2p_epoch[kind];
    const int parity = fixed_epoch >= 0 ? fixed_parity : (int)(epoch & 1);
    const int n = 8 * c->wt;
    const int blocks = std::max(1, std::min(64, (n + kBlock - 1) / kBlock));
    k_push_rows<<<blocks, kBlock, 0, st>>>(field, c->wx, c->wt, c->V, win_ghost(c, c->peer_win[0], kind, parity, 1),
                                           win_ghost(c, c->peer_win[1], kind, parity, 0),
                                           win_flag(c, c->peer_win[0], kind, 1), win_flag(c, c->peer_win[1], kind, 0),
                                           epoch, c->push_ticket);
    KCHECK();
    c->launches++;
    return SM_OK;
}

// the same for a lattice split along t: my first two columns into the -t neighbour's "hi" ghost columns, my last two into the
// +t neighbour's "lo" (k_push_cols gathers the strided columns and stores them straight into the peers' windows)
static int p2p_push_cols(sm_ctx* c, const cplx* field, int kind, cudaStream_t st) {
    const unsigned int epoch = ++c->p2p_epoch[kind];
    const int parity = (int)(epoch & 1);
    const int n = 8 * c->wx;
    const int blocks = std::max(1, std::min(64, (n + kBlock - 1) / kBlock));
    k_push_cols<<<blocks, kBlock, 0, st>>>(field, c->wx, c->wt, c->V, win_ghost(c, c->peer_win[0], kind, parity, 1),
                                           win_ghost(c, c->peer_win[1], kind, parity, 0),
                                           win_flag(c, c->peer_win[0], kind, 1), win_flag(c, c->peer_win[1], kind, 0),
                                           epoch, c->push_ticket);
    KCHECK();
    c->launches++;
    return SM_OK;
}

// make `st` wait until both neighbours have delivered the current epoch of `kind`
static int p2p_wait(sm_ctx* c, int kind, cudaStream_t st) {
    const unsigned int epoch = c->p2p_epoch[kind];
    for (int side = 0; side < 2; side++) {
        CUresult r = c->wait_value32((CUstream)st, (CUdeviceptr)win_flag(c, c->win, kind, side), epoch,
                                     CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) return fail(SM_ERR_CUDA, "cuStreamWaitValue32 failed (" + std::to_string((int)r) + ")");
    }
    return SM_OK;
}

// two boundary rows of a field (rows 0,1 to the -x neighbour, rows wx-2,wx-1 to the +x neighbour) into
// the [comp][2][wt] ghost arrays; rows are contiguous in HBM, so nothing is packed
static int exchange_rows2(sm_ctx* c, const cplx* field, cplx* lo_dst, cplx* hi_dst, cudaStream_t st = nullptr) {
    if (st == nullptr) st = c->stream;
    const size_t n = 2 * (size_t)c->wt;   // complex per component
    NC(g_nccl.GroupStart());
    for (int comp = 0; comp < 2; comp++) {
        const cplx* f = field + (size_t)comp * c->V;
        NC(g_nccl.Send(f, 2 * n, ncclDouble, c->nb_xm, c->comm, st));
        NC(g_nccl.Send(f + (size_t)(c->wx - 2) * c->wt, 2 * n, ncclDouble, c->nb_xp, c->comm, st));
        NC(g_nccl.Recv(hi_dst + comp * n, 2 * n, ncclDouble, c->nb_xp, c->comm, st));
        NC(g_nccl.Recv(lo_dst + comp * n, 2 * n, ncclDouble, c->nb_xm, c->comm, st));
    }
    NC(g_nccl.GroupEnd());
    return SM_OK;
}

// One-pass D D^dagger on a lattice split along t (and possibly x): 2-deep ghost columns from packed strips, ghost rows
// widened by the corner entries, everything exchanged by NCCL on the compute stream before ONE launch of k_dd_tma (no
// interior/boundary overlap yet).  CG: only r travels; the ghost columns / rows / corners of d_k are written by the pass.
template <int MODE>
static int fused_tsplit_args(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, double* sums_out, const cplx* r,
                             cplx* x, cplx* d_new, int k, FusedArgsT<cplx>& a) {
    a = FusedArgsT<cplx>{};
    a.U = U;
    a.in = in;
    a.out = out;
    a.wx = c->wx;
    a.wt = c->wt;
    a.V = c->V;
    a.rows_per_block = c->fus_rows;
    a.cols_per_strip = c->fus_cols;
    a.mass = m0 + 2;
    a.sR_edge = c->sR_edge();
    a.sL_edge = c->sL_edge();
    a.partials = c->partials;
    a.ticket = c->tickets + TK_WILSON;
    a.sums_out = sums_out;
    a.st = c->cg;
    a.r = r;
    a.x = x;
    a.d_new = d_new;
    a.first = (k == 0);
    a.cur = k & 1;
    a.nchunks = c->fus_grid.y;
    a.chunk_mode = 0;
    a.tg_on = 1;
    a.row_w = c->wt + 4;
    const bool rows = tg_rows(c);
    a.tgU_lo = c->tg_col[0][0];
    a.tgU_hi = c->tg_col[0][1];
    if (rows) {
        a.gU_lo = c->tg_row[0][0];
        a.gU_hi = c->tg_row[0][1];
    }
    if (MODE == FUSED_CG) {
        const int cur = k & 1, dk = 3 + cur, dp = 3 + (cur ^ 1);
        a.tgr_lo = c->tg_col[2][0];
        a.tgr_hi = c->tg_col[2][1];
        a.tgin_lo = c->tg_col[dp][0];
        a.tgin_hi = c->tg_col[dp][1];
        a.tgd_lo = c->tg_col[dk][0];
        a.tgd_hi = c->tg_col[dk][1];
        if (rows) {
            a.gr_lo = c->tg_row[2][0];
            a.gr_hi = c->tg_row[2][1];
            a.gin_lo = c->tg_row[dp][0];
            a.gin_hi = c->tg_row[dp][1];
            a.gd_lo = c->tg_row[dk][0];
            a.gd_hi = c->tg_row[dk][1];
        }
    } else {
        a.tgin_lo = c->tg_col[1][0];
        a.tgin_hi = c->tg_col[1][1];
        if (rows) {
            a.gin_lo = c->tg_row[1][0];
            a.gin_hi = c->tg_row[1][1];
        }
    }
    constexpr int STAGES = (MODE == FUSED_CG) ? 2 : 4;
    const unsigned int attr_bit = 1u << (24 + MODE);
    if (!(c->attr_done & attr_bit)) {
        CU(cudaFuncSetAttribute(k_dd_tma<MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        c->attr_done |= attr_bit;
    }
    a.nstrips = (int)c->fus_grid.x;
    return SM_OK;
}

template <int MODE>
static int launch_fused_tsplit(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, double* sums_out, const cplx* r,
                               cplx* x, cplx* d_new, int k) {
    TRY(tg_alloc(c));
    constexpr int STAGES = (MODE == FUSED_CG) ? 2 : 4;
    const size_t smem = fused_tma_smem_bytes(MODE, STAGES, c->fus_block.x);
    const int S = (int)c->fus_grid.x;
    const bool rows = tg_rows(c);
    const cplx* moving = (MODE == FUSED_CG) ? r : in;
    const int mv = (MODE == FUSED_CG) ? 2 : 1;
    if (c->tg_U_valid_for != U) {
        TRY(tg_exchange(c, U, 0, c->stream));
        c->tg_U_valid_for = U;
    }
    FusedArgsT<cplx> a{};
    // peer-memory windows (t-only splits): the columns are stored straight into the neighbours' HBM by a small kernel of
    // the comm stream and the edge strips wait on flags -- no send/recv kernel that a GPU full of interior blocks would
    // keep from starting (measured on 1 x 2 with NCCL: 0.66-0.69 ms per overlapped pass against 0.57 without overlap)
    const bool push = c->p2p && win_along_t(c) && !rows;
    auto exchange_moving = [&](cudaStream_t st, cudaEvent_t packed) -> int {
        if (!push) return tg_exchange(c, moving, mv, st, packed);
        const int kind = (MODE == FUSED_CG) ? 1 : 0;
        TRY(p2p_push_cols(c, moving, kind, st));
        if (packed != nullptr) CU(cudaEventRecord(packed, st));
        return p2p_wait(c, kind, st);
    };
    auto use_window = [&](FusedArgsT<cplx>& args) {
        if (!push) return;
        const int kind = (MODE == FUSED_CG) ? 1 : 0;
        const int parity = c->p2p_epoch[kind] & 1;
        cplx* lo = win_ghost(c, c->win, kind, parity, 0);
        cplx* hi = win_ghost(c, c->win, kind, parity, 1);
        if (MODE == FUSED_CG) {
            args.tgr_lo = lo;
            args.tgr_hi = hi;
        } else {
            args.tgin_lo = lo;
            args.tgin_hi = hi;
        }
    };
    // Overlap: only the two edge strips read ghost columns and only the boundary bands read ghost rows.  They follow the
    // exchange on the comm stream; the interior strips (interior rows) run meanwhile on the compute stream.
    const bool overlap = c->overlap && S >= 3 && (!rows || c->fus_split_chunks >= 1);
    if (!overlap) {
        TRY(exchange_moving(c->stream, nullptr));
        TRY((fused_tsplit_args<MODE>(c, U, in, out, m0, sums_out, r, x, d_new, k, a)));
        use_window(a);
        k_dd_tma<MODE, STAGES><<<c->fus_grid, c->fus_block, smem, c->stream>>>(a);
        KCHECK();
        c->launches++;
        return SM_OK;
    }
    CU(cudaEventRecord(c->ev_ready, c->stream));
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
    // Peer-memory push: the interior launch is held until the push kernel is through, so that its blocks -- which take every
    // register of every SM -- cannot keep the push from starting (1 x 8 on 8192^2: 0.159 ms per pass).  With the two NCCL
    // phases of a 2-D split the same gate was slower (2 x 4: 0.224 against 0.195 ms), so it is not applied there.
    TRY(exchange_moving(c->comm_stream, push ? c->ev_packed : nullptr));
    if (push) CU(cudaStreamWaitEvent(c->stream, c->ev_packed, 0));
    TRY((fused_tsplit_args<MODE>(c, U, in, out, m0, sums_out, r, x, d_new, k, a)));
    use_window(a);
    const int edge_chunks = (int)c->fus_grid.y;                                   // edge strips: uniform chunks over all rows
    const int int_chunks = rows ? c->fus_split_chunks : (int)c->fus_grid.y;       // interior strips
    const int n_edge = 2 * edge_chunks, n_int = (S - 2) * int_chunks, n_band = rows ? (S - 2) * 2 : 0;
    a.part_total = n_edge + n_int + n_band;
    // edge strips (all rows), then the two boundary bands of the interior strips: comm stream, behind the exchange
    a.strip_mode = 2;
    a.chunk_mode = 0;
    a.part_base = 0;
    k_dd_tma<MODE, STAGES><<<dim3(2, edge_chunks, 1), c->fus_block, smem, c->comm_stream>>>(a);
    KCHECK();
    c->launches++;
    if (rows) {
        a.strip_mode = 1;
        a.chunk_mode = 2;
        a.rb = c->fus_rb;
        a.part_base = n_edge + n_int;
        k_dd_tma<MODE, STAGES><<<dim3(S - 2, 2, 1), c->fus_block, smem, c->comm_stream>>>(a);
        KCHECK();
        c->launches++;
    }
    CU(cudaEventRecord(c->ev_ghost, c->comm_stream));
    // interior strips, interior rows: compute stream, no ghosts
    a.strip_mode = 1;
    a.part_base = n_edge;
    if (rows) {
        a.chunk_mode = 1;
        a.rb = c->fus_rb;
        a.rows_per_block = c->fus_split_rows;
    } else {
        a.chunk_mode = 0;
        a.rows_per_block = c->fus_rows;
    }
    k_dd_tma<MODE, STAGES><<<dim3(S - 2, int_chunks, 1), c->fus_block, smem, c->stream>>>(a);
    KCHECK();
    c->launches++;
    CU(cudaStreamWaitEvent(c->stream, c->ev_ghost, 0));
    return SM_OK;
}

// opt-in (SM_PDL=1): the two kernels of a one-pass CG iteration on a single tile start while their predecessor drains
static bool pdl_ok(const sm_ctx* c) { return c->pdl && !c->dist() && c->fused_tma && !c->self_t && !c->self_x; }

// one-pass D D^dagger (sm_fused.cuh, sm_fused_tma.cuh): a single tile, tiles split along x only (ranks_t == 1, 2-row
// ghosts, peer-memory halos), or -- k_dd_tma only -- tiles split along t as well (launch_fused_tsplit).  C = cplx (double) everywhere except in the
// inner solve of the opt-in mixed-precision CG (C = cplxf, single tile only).
template <typename C, int MODE>
static int launch_fused(sm_ctx* c, const C* U, const C* in, C* out, double m0, double* sums_out = nullptr,
                        const C* r = nullptr, C* x = nullptr, C* d_new = nullptr, int k = 0) {
    constexpr bool kDouble = std::is_same<C, cplx>::value;
    if (!kDouble && c->dist()) return fail(SM_ERR_STATE, "single-precision passes run on a single tile only");
    if constexpr (kDouble) {
        if (tg_cols(c)) return launch_fused_tsplit<MODE>(c, U, in, out, m0, sums_out, r, x, d_new, k);
    }
    FusedArgsT<C> a{};
    a.U = U;
    a.in = in;
    a.out = out;
    a.wx = c->wx;
    a.wt = c->wt;
    a.V = c->V;
    a.rows_per_block = c->fus_rows;
    a.cols_per_strip = c->fus_cols;
    a.mass = m0 + 2;
    a.sR_edge = c->sR_edge();
    a.sL_edge = c->sL_edge();
    a.partials = c->partials;
    a.ticket = c->tickets + TK_WILSON;
    a.sums_out = sums_out;
    a.st = c->cg;
    a.r = r;
    a.x = x;
    a.d_new = d_new;
    a.first = (k == 0);
    a.cur = k & 1;
    a.nchunks = c->fus_grid.y;
    a.chunk_mode = 0;
    // rows in flight per block: as many as 2 blocks per SM leave shared memory for (single precision moves
    // half the bytes per row, so it keeps more rows in flight)
    constexpr int STAGES = kDouble ? ((MODE == FUSED_CG) ? 2 : 3) : 3;
    size_t smem = fused_smem_bytes(MODE, STAGES, c->fus_block.x, sizeof(C));
    // double precision: rows staged by TMA bulk copies (sm_fused_tma.cuh) unless SM_FUSED_TMA=0
    void (*kern)(const FusedArgsT<C>) = k_dd_fused<C, MODE, STAGES>;
    unsigned int attr_bit = 1u << (MODE + (kDouble ? 0 : 4));
    if constexpr (kDouble) {
        if (c->fused_tma) {
            // (a third stage for the CG pass -- 128-thread blocks, 3 per SM -- was measured and dropped: 65-74 us per
            // iteration at 1024^2 against 63.7; profiles/r02_cg_1024_l2_persistence.txt)
            if (MODE != FUSED_CG && c->fused_stages == 4) {
                kern = k_dd_tma<MODE, (MODE == FUSED_CG) ? 2 : 4>;
                smem = fused_tma_smem_bytes(MODE, (MODE == FUSED_CG) ? 2 : 4, c->fus_block.x);
                attr_bit = 1u << (16 + MODE);
            } else {
                kern = k_dd_tma<MODE, STAGES>;
                smem = fused_tma_smem_bytes(MODE, STAGES, c->fus_block.x);
                attr_bit = 1u << (12 + MODE);
            }
        }
    }
    if (!(c->attr_done & attr_bit)) {   // function attributes are per device: once per context and instantiation
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        c->attr_done |= attr_bit;
    }
    bool split_launch = false;
    if constexpr (kDouble) {
        if (c->dist()) {
            if (c->f2_U_valid_for != U) {
                TRY(exchange_rows2(c, U, c->f2_U[0], c->f2_U[1]));
                c->f2_U_valid_for = U;
            }
            a.gU_lo = c->f2_U[0];
            a.gU_hi = c->f2_U[1];
            // the ghost rows this pass needs: psi (PLAIN) or r (CG; d_{k-1} ghosts were written by the previous
            // pass).  Only the two boundary bands read them, so the exchange runs on the comm stream while the
            // interior chunks compute.
            const cplx* moving = (MODE == FUSED_CG) ? r : in;
            const int kind = (MODE == FUSED_CG) ? 1 : 0;
            cplx* dst[2] = {(MODE == FUSED_CG) ? c->f2_r[0] : c->f2_in[0], (MODE == FUSED_CG) ? c->f2_r[1] : c->f2_in[1]};
            split_launch = c->overlap && c->fus_split_chunks >= 1;
            cudaStream_t xs = split_launch ? c->comm_stream : c->stream;
            // CG with peer-memory sums: r's ghost rows were stored into my window by the neighbours' k_cg_resid (or by
            // the push that starts the solve) and the kernel itself waits for them -- nothing to exchange here
            const bool in_kernel = (MODE == FUSED_CG) && c->peer_sums;
            if (split_launch) {
                CU(cudaEventRecord(c->ev_ready, c->stream));
                CU(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
            }
            if (in_kernel) {
                const int cur = k & 1;
                dst[0] = win_ghost(c, c->win, 1, cur, 0);
                dst[1] = win_ghost(c, c->win, 1, cur, 1);
                a.dl = dist_link(c);
            } else if (c->p2p) {
                TRY(p2p_push(c, moving, kind, xs));                   // stores into the neighbours' windows
                TRY(p2p_wait(c, kind, xs));                           // ... and waits for theirs in mine
                const int parity = c->p2p_epoch[kind] & 1;
                dst[0] = win_ghost(c, c->win, kind, parity, 0);
                dst[1] = win_ghost(c, c->win, kind, parity, 1);
            } else {
                TRY(exchange_rows2(c, moving, dst[0], dst[1], xs));
            }
            if (MODE == FUSED_CG) {
                const int cur = k & 1;
                a.gin_lo = c->f2_d[cur ^ 1][0];
                a.gin_hi = c->f2_d[cur ^ 1][1];
                a.gd_lo = c->f2_d[cur][0];
                a.gd_hi = c->f2_d[cur][1];
                a.gr_lo = dst[0];
                a.gr_hi = dst[1];
            } else {
                a.gin_lo = dst[0];
                a.gin_hi = dst[1];
            }
        }
    }
    if (split_launch) {
        // boundary bands follow the exchange on the comm stream; the interior runs meanwhile
        a.rb = c->fus_rb;
        a.rows_per_block = c->fus_split_rows;
        a.nchunks = c->fus_split_chunks + 2;
        a.chunk_mode = 2;
        kern<<<dim3(c->fus_grid.x, 2, 1), c->fus_block, smem, c->comm_stream>>>(a);
        KCHECK();
        CU(cudaEventRecord(c->ev_ghost, c->comm_stream));
        a.chunk_mode = 1;
        kern<<<dim3(c->fus_grid.x, c->fus_split_chunks, 1), c->fus_block, smem, c->stream>>>(a);
        KCHECK();
        CU(cudaStreamWaitEvent(c->stream, c->ev_ghost, 0));
        c->launches += 2;
    } else if (kDouble && MODE == FUSED_CG && pdl_ok(c)) {
        void* params[] = {(void*)&a};
        TRY(launch_pdl(c, (const void*)kern, c->fus_grid, c->fus_block, smem, params));
        c->launches++;
    } else {
        kern<<<c->fus_grid, c->fus_block, smem, c->stream>>>(a);
        KCHECK();
        c->launches++;
    }
    return SM_OK;
}

static bool fused_ok(const sm_ctx* c) {
    if (!c->use_fused) return false;
    if (!c->dist()) return !(c->self_t || c->self_x) || (c->fused_tma && c->wx >= 4 && c->wt >= 4);
    if (c->rt == 1) return c->wx >= 4;
    return c->tsplit_onepass && c->fused_tma && c->wx >= 4 && c->wt >= 4;     // ghost columns: k_dd_tma only
}

// D D^dagger: one pass over HBM on a single tile, else D^dagger then D through the context's
// scratch field (the reference's global DTEMP, dirac_operator.cpp:477-480)
static int dev_DDdag(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0) {
    if (in == out) return fail(SM_ERR_ARG, "D D^dagger: in and out must not alias");
    if (fused_ok(c)) return launch_fused<cplx, FUSED_PLAIN>(c, U, in, out, m0);
    TRY(ensure_complex(c, &c->tmp));
    TRY(dev_D(c, U, in, c->tmp, m0, true));
    return dev_D(c, U, c->tmp, out, m0, false);
}

static int dev_dot_async(sm_ctx* c, const cplx* x, const cplx* y, double* d_out2) {
    k_dot<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(x, y, 2 * c->V, c->partials, c->tickets + TK_DOT,
                                                      sum_target(c, d_out2));
    KCHECK();
    c->launches++;
    return sum_finish(c, d_out2, 2);
}
