// sm_dist.cuh -- split lattice: all-reduced sums and the exchange of projected half-spinor halo lines.
// Part of the single translation unit sm_abi.cu (static functions, included in dependency order).
#pragma once
#include "sm_context.cuh"

// ------------------------------------------------------------------------------------------------
// split lattice: halo exchange of projected half-spinors, all-reduce of sums
// ------------------------------------------------------------------------------------------------
// ghost copies of a gauge field go stale whenever the field is written
static void invalidate_gauge_ghosts(sm_ctx* c, const cplx* U) {
    if (c->ghost_valid_for == U) c->ghost_valid_for = nullptr;
    if (c->f2_U_valid_for == U) c->f2_U_valid_for = nullptr;
    if (c->tg_U_valid_for == U) c->tg_U_valid_for = nullptr;
}

static int allreduce_sums(sm_ctx* c, const double* loc, double* glob, int n) {
    NC(g_nccl.AllReduce(loc, glob, (size_t)n, ncclDouble, ncclSum, c->comm, c->stream));
    return SM_OK;
}

// where a reducing kernel should write, and the follow-up that makes it global
static double* sum_target(sm_ctx* c, double* glob) { return c->dist() ? c->sums_loc : glob; }
static int sum_finish(sm_ctx* c, double* glob, int n) {
    if (!c->dist()) return SM_OK;
    return allreduce_sums(c, c->sums_loc, glob, n);
}

// the four projected halo lines of `in` into the send buffers (k_pack_halo)
template <bool DAG>
static int pack_spinor_halo(sm_ctx* c, const cplx* U, const cplx* in, const int* done) {
    PackArgs p{};
    p.U = U;
    p.in = in;
    p.wx = c->wx;
    p.wt = c->wt;
    p.V = c->V;
    p.to_tm = c->rt > 1 ? c->send_tm : nullptr;
    p.to_tp = c->send_tp;
    p.to_xm = c->rx > 1 ? c->send_xm : nullptr;
    p.to_xp = c->send_xp;
    p.done = done;
    const int n = std::max(c->wx, c->wt);
    k_pack_halo<DAG><<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(p);
    KCHECK();
    c->launches++;
    return SM_OK;
}

// send buffers -> the neighbours' ghost lines (one grouped NCCL send/recv) on stream `st`
static int exchange_spinor_lines(sm_ctx* c, cudaStream_t st) {
    NC(g_nccl.GroupStart());
    if (c->rt > 1) {
        NC(g_nccl.Send(c->send_tm, 2 * (size_t)c->wx, ncclDouble, c->nb_tm, c->comm, st));
        NC(g_nccl.Send(c->send_tp, 2 * (size_t)c->wx, ncclDouble, c->nb_tp, c->comm, st));
        NC(g_nccl.Recv(c->g_tp, 2 * (size_t)c->wx, ncclDouble, c->nb_tp, c->comm, st));
        NC(g_nccl.Recv(c->g_tm, 2 * (size_t)c->wx, ncclDouble, c->nb_tm, c->comm, st));
    }
    if (c->rx > 1) {
        NC(g_nccl.Send(c->send_xm, 2 * (size_t)c->wt, ncclDouble, c->nb_xm, c->comm, st));
        NC(g_nccl.Send(c->send_xp, 2 * (size_t)c->wt, ncclDouble, c->nb_xp, c->comm, st));
        NC(g_nccl.Recv(c->g_xp, 2 * (size_t)c->wt, ncclDouble, c->nb_xp, c->comm, st));
        NC(g_nccl.Recv(c->g_xm, 2 * (size_t)c->wt, ncclDouble, c->nb_xm, c->comm, st));
    }
    NC(g_nccl.GroupEnd());
    return SM_OK;
}

// ------------------------------------------------------------------------------------------------
// 2-deep halos of a field for the one-pass D D^dagger on a lattice split along t (and possibly x)
// ------------------------------------------------------------------------------------------------
// my columns 0,1 -> lo[comp][row][2] ; my columns wt-2, wt-1 -> hi[comp][row][2]
__global__ void __launch_bounds__(kBlock) k_pack_cols2(const cplx* __restrict__ f, int wx, int wt, int V,
                                                       cplx* __restrict__ lo, cplx* __restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (comp, row, c)
    if (i >= 4 * wx) return;
    const int comp = i / (2 * wx), e = i - comp * 2 * wx, row = e >> 1, cc = e & 1;
    const cplx* src = f + (size_t)comp * V + (size_t)row * wt;
    lo[i] = src[cc];
    hi[i] = src[wt - 2 + cc];
}

// rows 0,1 and wx-2, wx-1 widened by their ghost columns -> lo / hi [comp][2][W], W = wt + 4
__global__ void __launch_bounds__(kBlock) k_pack_rows2w(const cplx* __restrict__ f, int wx, int wt, int V,
                                                        const cplx* __restrict__ gc_lo, const cplx* __restrict__ gc_hi,
                                                        cplx* __restrict__ lo, cplx* __restrict__ hi) {
    const int W = wt + 4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (comp, r, col)
    if (i >= 4 * W) return;
    const int comp = i / (2 * W), e = i - comp * 2 * W, r = e / W, col = e - r * W;   // col 0..W-1  <->  t = col - 2
    for (int which = 0; which < 2; which++) {
        const int row = which ? wx - 2 + r : r;
        cplx v;
        if (col < 2) v = gc_lo[(size_t)comp * 2 * wx + row * 2 + col];
        else if (col >= wt + 2) v = gc_hi[(size_t)comp * 2 * wx + row * 2 + (col - wt - 2)];
        else v = f[(size_t)comp * V + (size_t)row * wt + (col - 2)];
        (which ? hi : lo)[i] = v;
    }
}

static size_t tg_col_elems(const sm_ctx* c) { return 4 * (size_t)c->wx; }
static size_t tg_row_elems(const sm_ctx* c) { return 4 * ((size_t)c->wt + 4); }
static bool tg_cols(const sm_ctx* c) { return c->rt > 1 || c->self_t; }
static bool tg_rows(const sm_ctx* c) { return c->rx > 1 || c->self_x; }

static int tg_alloc(sm_ctx* c) {
    if (c->tg_sendc) return SM_OK;
    for (int kind = 0; kind < 5; kind++)
        for (int side = 0; side < 2; side++) {
            TRY(dev_alloc(&c->tg_col[kind][side], tg_col_elems(c)));
            TRY(dev_alloc(&c->tg_row[kind][side], tg_row_elems(c)));
            CU(cudaMemsetAsync(c->tg_col[kind][side], 0, sizeof(cplx) * tg_col_elems(c), c->stream));
            CU(cudaMemsetAsync(c->tg_row[kind][side], 0, sizeof(cplx) * tg_row_elems(c), c->stream));
        }
    TRY(dev_alloc(&c->tg_sendc, 2 * tg_col_elems(c)));
    TRY(dev_alloc(&c->tg_sendr, 2 * tg_row_elems(c)));
    return SM_OK;
}

// field -> ghost columns and (widened) ghost rows of `kind`: t direction first, then x including the fresh ghost columns,
// so the corner entries arrive without diagonal messages (the reference sends them separately, gauge_conf.cpp:225-227)
// `packed` (may be null): recorded on `st` right after the column pack, i.e. immediately before the first NCCL kernel.
static int tg_exchange(sm_ctx* c, const cplx* f, int kind, cudaStream_t st, cudaEvent_t packed = nullptr) {
    const size_t nc = tg_col_elems(c), nr = tg_row_elems(c);
    cplx *slo = c->tg_sendc, *shi = c->tg_sendc + nc;
    k_pack_cols2<<<(int)((nc + kBlock - 1) / kBlock), kBlock, 0, st>>>(f, c->wx, c->wt, c->V, slo, shi);
    KCHECK();
    c->launches++;
    if (packed != nullptr) CU(cudaEventRecord(packed, st));
    if (c->rt > 1) {
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(slo, 2 * nc, ncclDouble, c->nb_tm, c->comm, st));      // my first columns  -> their "hi"
        NC(g_nccl.Send(shi, 2 * nc, ncclDouble, c->nb_tp, c->comm, st));      // my last columns   -> their "lo"
        NC(g_nccl.Recv(c->tg_col[kind][1], 2 * nc, ncclDouble, c->nb_tp, c->comm, st));
        NC(g_nccl.Recv(c->tg_col[kind][0], 2 * nc, ncclDouble, c->nb_tm, c->comm, st));
        NC(g_nccl.GroupEnd());
    } else {                                                                   // one tile: its own opposite edges
        CU(cudaMemcpyAsync(c->tg_col[kind][1], slo, sizeof(cplx) * nc, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(c->tg_col[kind][0], shi, sizeof(cplx) * nc, cudaMemcpyDeviceToDevice, st));
    }
    if (!tg_rows(c)) return SM_OK;
    cplx *rlo = c->tg_sendr, *rhi = c->tg_sendr + nr;
    k_pack_rows2w<<<(int)((nr + kBlock - 1) / kBlock), kBlock, 0, st>>>(f, c->wx, c->wt, c->V, c->tg_col[kind][0], c->tg_col[kind][1],
                                                                        rlo, rhi);
    KCHECK();
    c->launches++;
    if (c->rx > 1) {
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(rlo, 2 * nr, ncclDouble, c->nb_xm, c->comm, st));
        NC(g_nccl.Send(rhi, 2 * nr, ncclDouble, c->nb_xp, c->comm, st));
        NC(g_nccl.Recv(c->tg_row[kind][1], 2 * nr, ncclDouble, c->nb_xp, c->comm, st));
        NC(g_nccl.Recv(c->tg_row[kind][0], 2 * nr, ncclDouble, c->nb_xm, c->comm, st));
        NC(g_nccl.GroupEnd());
    } else {
        CU(cudaMemcpyAsync(c->tg_row[kind][1], rlo, sizeof(cplx) * nr, cudaMemcpyDeviceToDevice, st));
        CU(cudaMemcpyAsync(c->tg_row[kind][0], rhi, sizeof(cplx) * nr, cudaMemcpyDeviceToDevice, st));
    }
    return SM_OK;
}
