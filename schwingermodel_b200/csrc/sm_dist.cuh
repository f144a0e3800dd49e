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
