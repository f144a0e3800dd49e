// sm_cg.cuh -- conjugate-gradient drivers: two-pass, one-pass (+ CUDA graphs), mixed precision, cluster / grid resident.
// Part of the single translation unit sm_abi.cu (static functions, included in dependency order).
#pragma once
#include "sm_ops.cuh"
#include "sm_column_cg.cuh"

// conjugate_gradient (src/conjugate_gradient.cpp:4-67) entirely on the device.  The host only
// enqueues batches of iterations and polls a pinned copy of the CG scalars one batch behind, so
// the GPU never waits for it; once the stopping rule has fired every later kernel of the queue
// returns at its first instruction.
static int dev_cg_twopass(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    TRY(ensure_complex(c, &c->tmp));
    TRY(ensure_complex(c, &c->cg_r));
    TRY(ensure_complex(c, &c->cg_d));
    TRY(ensure_complex(c, &c->cg_Ad));
    const int n_elems = 2 * c->V;
    const double tol = c->tol;
    const int max_iter = c->max_iter;
    CgState* st = c->cg;
    const int* done = &st->done;

    k_cg_reset<<<1, 1, 0, c->stream>>>(st, tol, max_iter);
    c->launches++;
    // Ad = DD^dagger x_0 ; r = phi - Ad ; d = r ; x = x_0 (= phi unless a start vector was given) ; |phi|^2, |r|^2
    TRY((launch_wilson<true, WILSON_PLAIN>(c, U, c->cg_x0 ? c->cg_x0 : phi, c->tmp, m0)));
    TRY((launch_wilson<false, WILSON_CGINIT>(c, U, c->tmp, nullptr, m0, phi, c->cg_r, c->cg_d, x,
                                             sum_target(c, &st->phi_norm2))));
    TRY(sum_finish(c, &st->phi_norm2, 2));

    const int batch = 8;
    int k = 0, slot = 0, prev = -1;
    for (;;) {
        const int k_end = std::min(max_iter, k + batch);
        for (; k < k_end; k++) {
            const int cur = k & 1;
            if (k > 0) {
                k_cg_dir<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, k, tol, c->cg_r, c->cg_d, n_elems);
                KCHECK();
                c->launches++;
            }
            TRY((launch_wilson<true, WILSON_PLAIN>(c, U, c->cg_d, c->tmp, m0, nullptr, nullptr, nullptr, nullptr,
                                                   nullptr, done)));
            TRY((launch_wilson<false, WILSON_DOT>(c, U, c->tmp, c->cg_Ad, m0, c->cg_d, nullptr, nullptr, nullptr,
                                                  sum_target(c, st->dAd), done)));
            TRY(sum_finish(c, st->dAd, 2));
            k_cg_update<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, cur, x, c->cg_d, c->cg_r, c->cg_Ad, n_elems,
                                                                    c->partials, c->tickets + TK_UPDATE,
                                                                    sum_target(c, &st->rr[cur ^ 1]));
            KCHECK();
            c->launches++;
            TRY(sum_finish(c, &st->rr[cur ^ 1], 1));
        }
        // stopping rule of the batch's last iteration; at k == max_iter this always sets `done`
        k_cg_check<<<1, 1, 0, c->stream>>>(st, k, tol, max_iter);
        KCHECK();
        c->launches++;
        CU(cudaMemcpyAsync(&c->h->cg[slot], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->ev_poll[slot], c->stream));
        // look one batch behind so the queue never drains while the host waits
        if (prev >= 0) {
            CU(cudaEventSynchronize(c->ev_poll[prev]));
            if (c->h->cg[prev].done) break;
        }
        if (k >= max_iter) break;
        prev = slot;
        slot ^= 1;
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyAsync(&c->h->cg[0], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (converged) *converged = c->h->cg[0].converged;
    if (iterations) *iterations = c->h->cg[0].iters;
    return SM_OK;
}

// The same algorithm on the one-pass D D^dagger: per iteration k
//   A(k): stopping rule of k-1 ; d_k = r_k + beta d_{k-1} ; x += alpha_{k-1} d_{k-1} ; Ad = D D^dagger d_k ; dot(d_k, Ad)
//   B(k): alpha_k = r_norm2 / dot ; r -= alpha_k Ad ; |r|^2
// and one k_cg_flush_x at the end for the x update the loop still owes.  320 B per site and iteration
// (160 B in single precision).  On entry CgState holds |r|^2 in rr[0], the reference norm in phi_norm2, the
// tolerance, k = 0; r holds the residual of the start vector x.
template <typename C>
static int cg_fused_loop(sm_ctx* c, const C* U, C* r, C* x, C* dbuf0, C* dbuf1, C* Ad, double m0, int max_iter) {
    const int n_elems = 2 * c->V;
    CgState* st = c->cg;
    C* dbuf[2] = {dbuf0, dbuf1};

    // one iteration: A(k) then B(k).  On a split lattice either the sums are all-reduced by NCCL after each kernel, or
    // (peer-memory windows connected, sm_peer.cuh) the kernels gather them themselves and B(k) stores r's boundary rows
    // into the neighbours' ghosts: then the loop holds no library call and is captured in a graph like on one tile.
    const bool peer = c->peer_sums && std::is_same<C, cplx>::value;
    const DistLink dl = peer ? dist_link(c) : DistLink{};
    auto iteration = [&](int k) -> int {
        const int cur = k & 1;
        TRY((launch_fused<C, FUSED_CG>(c, U, dbuf[cur ^ 1], Ad, m0, peer ? nullptr : sum_target(c, st->dAd), r, x, dbuf[cur], k)));
        if (!peer) TRY(sum_finish(c, st->dAd, 2));
        if constexpr (std::is_same<C, cplx>::value) {
            if (peer) {
                k_cg_resid_dist<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, cur, r, Ad, c->wx, c->wt, c->V, c->partials,
                                                                            c->tickets + TK_UPDATE, dl);
                KCHECK();
                c->launches++;
                return SM_OK;
            }
        }
        if (std::is_same<C, cplx>::value && pdl_ok(c)) {
            CgState* st_ = st;
            int cur_ = cur, n_ = n_elems;
            C* r_ = r;
            const C* Ad_ = Ad;
            double* partials_ = c->partials;
            unsigned int* ticket_ = c->tickets + TK_UPDATE;
            double* sums_ = sum_target(c, &st->rr[cur ^ 1]);
            void* params[] = {&st_, &cur_, &r_, &Ad_, &n_, &partials_, &ticket_, &sums_};
            TRY(launch_pdl(c, (const void*)k_cg_resid<C>, dim3(c->flat_blocks_c, 1, 1), dim3(kBlock, 1, 1), 0, params));
            c->launches++;
            return sum_finish(c, &st->rr[cur ^ 1], 1);
        }
        k_cg_resid<C><<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, cur, r, Ad, n_elems, c->partials,
                                                                  c->tickets + TK_UPDATE, sum_target(c, &st->rr[cur ^ 1]));
        KCHECK();
        c->launches++;
        return sum_finish(c, &st->rr[cur ^ 1], 1);
    };
    const int batch = 8;   // even: a replayed batch always starts on the same parity

    // a batch of iterations k = 1 + 8 m ... as one CUDA graph (single tile; the iteration index and the
    // tolerance live in CgState, so the nodes are iteration- and tolerance-independent)
    cudaGraphExec_t exec = nullptr;
    int graph_kernels = 0;
    const bool graphs = c->use_graphs && (!c->dist() || peer) && max_iter > batch;
    if (graphs) {
        for (auto& g : c->cg_graphs)
            if (g.U == (const void*)U && g.x == (const void*)x && g.m0 == m0) {
                exec = g.exec;
                graph_kernels = g.kernels;
            }
    }

    int k = 0, slot = 0, prev = -1;
    TRY(iteration(k++));   // k = 0 is special (d_0 = r_0) and also sets the kernel attributes before any capture
    if (graphs && exec == nullptr) {
        const long long l0 = c->launches;
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = SM_OK;
        for (int i = 0; i < batch && rc == SM_OK; i++) rc = iteration(1 + i);
        if (rc == SM_OK) {
            k_cg_check_dev<<<1, 1, 0, c->stream>>>(st, dl);
            c->launches++;
        }
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc != SM_OK) return rc;
        if (e != cudaSuccess) return fail(SM_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        graph_kernels = (int)(c->launches - l0);
        c->launches = l0;
        CU(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        if (c->cg_graphs.size() >= 16) {
            cudaGraphExecDestroy(c->cg_graphs.front().exec);
            c->cg_graphs.erase(c->cg_graphs.begin());
        }
        c->cg_graphs.push_back({(const void*)U, (const void*)x, m0, exec, graph_kernels});
    }
    for (;;) {
        if (exec != nullptr && k + batch <= max_iter) {
            CU(cudaGraphLaunch(exec, c->stream));
            c->launches += graph_kernels;
            k += batch;
        } else {
            const int k_end = std::min(max_iter, k + batch);
            for (; k < k_end; k++) TRY(iteration(k));
            k_cg_check_dev<<<1, 1, 0, c->stream>>>(st, dl);
            KCHECK();
            c->launches++;
        }
        CU(cudaMemcpyAsync(&c->h->cg[slot], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->ev_poll[slot], c->stream));
        if (prev >= 0) {
            CU(cudaEventSynchronize(c->ev_poll[prev]));
            if (c->h->cg[prev].done) break;
        }
        if (k >= max_iter) break;
        prev = slot;
        slot ^= 1;
    }
    k_cg_flush_x<C><<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, x, dbuf[0], dbuf[1], n_elems);
    KCHECK();
    c->launches++;
    CU(cudaMemcpyAsync(&c->h->cg[0], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return SM_OK;
}

// the reference's algorithm, double precision throughout
static int dev_cg_fused(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    TRY(ensure_complex(c, &c->tmp));
    TRY(ensure_complex(c, &c->cg_r));
    TRY(ensure_complex(c, &c->cg_d));
    TRY(ensure_complex(c, &c->cg_d2));
    TRY(ensure_complex(c, &c->cg_Ad));
    CgState* st = c->cg;
    const unsigned int epoch_base = (++c->solve_seq) << 16;     // max_iter <= 10000 per solve (sm_set_cg caps it below 2^16)
    k_cg_reset<<<1, 1, 0, c->stream>>>(st, c->tol, c->max_iter, epoch_base);
    c->launches++;
    // x = x_0 (phi unless a start vector was given) ; r = phi - D D^dagger x_0 ; |phi|^2, |r|^2   (d_0 = r_0 is formed by the first pass)
    TRY((launch_wilson<true, WILSON_PLAIN>(c, U, c->cg_x0 ? c->cg_x0 : phi, c->tmp, m0)));
    TRY((launch_wilson<false, WILSON_CGINIT>(c, U, c->tmp, nullptr, m0, phi, c->cg_r, c->cg_d2, x,
                                             sum_target(c, &st->phi_norm2))));
    TRY(sum_finish(c, &st->phi_norm2, 2));
    // peer-memory halos: r_0's boundary rows into the neighbours' ghosts (parity 0, epoch base + 0); later r_k are
    // stored by k_cg_resid_dist itself
    if (c->peer_sums) TRY(p2p_push(c, c->cg_r, 1, c->stream, (long long)epoch_base, 0));
    // (Measured and dropped: keeping A d -- or A d and a copy of U -- in the persisting part of L2 with an access-policy
    // window made the 1024^2 solve slower, 63.7 -> 92 us per iteration; profiles/r02_cg_1024_l2_persistence.txt.)
    TRY(cg_fused_loop<cplx>(c, U, c->cg_r, x, c->cg_d, c->cg_d2, c->cg_Ad, m0, c->max_iter));
    if (converged) *converged = c->h->cg[0].converged;
    if (iterations) *iterations = c->h->cg[0].iters;
    return SM_OK;
}

// Opt-in (sm_set_solver(SM_SOLVER_MIXED)): defect correction in double precision around an inner CG in
// single precision.  x = phi; repeat { r = phi - A x (double, true residual); stop if |r| < tol |phi|;
// solve A e = r in single precision to a relative delta; x += e }.  Half the bytes per inner iteration.
// The result meets the same residual criterion (checked on the TRUE residual) but is a different iterate
// than the reference's, so dH parity at 1e-8 does not hold: SURVEY 8(f).4 "solver upgrades".
static int dev_cg_mixed(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    TRY(ensure_complex(c, &c->cg_Ad));
    const size_t n2 = 2 * (size_t)c->V;
    if (!c->mx_U) {
        TRY(dev_alloc(&c->mx_U, n2));
        TRY(dev_alloc(&c->mx_r, n2));
        TRY(dev_alloc(&c->mx_e, n2));
        TRY(dev_alloc(&c->mx_d0, n2));
        TRY(dev_alloc(&c->mx_d1, n2));
        TRY(dev_alloc(&c->mx_Ad, n2));
        CU(cudaMemsetAsync(c->mx_d0, 0, sizeof(cplxf) * n2, c->stream));
        CU(cudaMemsetAsync(c->mx_d1, 0, sizeof(cplxf) * n2, c->stream));
    }
    const int n_elems = (int)n2;
    k_to_single<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(U, c->mx_U, n_elems);
    KCHECK();
    c->launches++;
    CU(cudaMemcpyAsync(x, phi, sizeof(cplx) * n2, cudaMemcpyDeviceToDevice, c->stream));   // x0 = phi as the reference
    int total = 0, ok = 0;
    const int max_cycles = 12;
    for (int cycle = 0; cycle < max_cycles; cycle++) {
        TRY((launch_fused<cplx, FUSED_PLAIN>(c, U, x, c->cg_Ad, m0)));
        k_mixed_residual<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(phi, c->cg_Ad, c->mx_r, c->mx_e, n_elems, c->partials,
                                                                     c->tickets + TK_DOT, c->sums + 12);
        KCHECK();
        c->launches++;
        CU(cudaMemcpyAsync(c->h->sums + 12, c->sums + 12, sizeof(double) * 2, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        const double pp = c->h->sums[12], rr = c->h->sums[13];
        if (std::sqrt(rr) < c->tol * std::sqrt(pp)) {
            ok = 1;
            break;
        }
        if (total >= c->max_iter || cycle == max_cycles - 1) break;
        // do not over-solve the last cycle; single precision stalls near 1e-6
        const double need = 0.5 * c->tol * std::sqrt(pp) / std::sqrt(rr);
        const double delta = std::min(0.1, std::max(1e-5, need));
        const int inner_max = std::max(1, c->max_iter - total);
        k_mixed_begin<<<1, 1, 0, c->stream>>>(c->cg, c->sums + 12, delta, inner_max);
        KCHECK();
        c->launches++;
        TRY(cg_fused_loop<cplxf>(c, c->mx_U, c->mx_r, c->mx_e, c->mx_d0, c->mx_d1, c->mx_Ad, m0, inner_max));
        total += c->h->cg[0].converged ? c->h->cg[0].iters + 1 : c->h->cg[0].iters;
        k_mixed_correct<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(x, c->mx_e, n_elems);
        KCHECK();
        c->launches++;
    }
    if (converged) *converged = ok;
    if (iterations) *iterations = total;
    return SM_OK;
}

// small lattices: the whole solve in one launch, every site resident in one thread (sm_cluster_cg.cuh):
// one thread-block cluster up to 4096 sites, a cooperative grid up to one 512-thread CTA per SM
static int resident_args(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, ResidentCgArgs* a) {
    *a = ResidentCgArgs{};
    a->U = U;
    a->phi = phi;
    a->x = x;
    a->x0 = c->cg_x0;
    a->wx = c->wx;
    a->wt = c->wt;
    a->V = c->V;
    a->mass = m0 + 2;
    a->sR_edge = c->sR_edge();
    a->sL_edge = c->sL_edge();
    a->tol = c->tol;
    a->max_iter = c->max_iter;
    a->st = c->cg;
    return SM_OK;
}

static int resident_finish(sm_ctx* c, int* converged, int* iterations) {
    c->launches++;
    CU(cudaMemcpyAsync(&c->h->cg[0], c->cg, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (converged) *converged = c->h->cg[0].converged;
    if (iterations) *iterations = c->h->cg[0].iters;
    return SM_OK;
}

static void cluster_launch_config(sm_ctx* c, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr) {
    int ctas = 1;
    while (ctas * kClusterThreads < c->V) ctas *= 2;
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3(ctas, 1, 1);
    cfg->blockDim = dim3(kClusterThreads, 1, 1);
    cfg->dynamicSmemBytes = 0;
    cfg->stream = c->stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
}

// can this device co-schedule the (non-portable, up to 16 CTAs) cluster the lattice needs?  Asked once per context;
// a MIG slice or a part with fewer SMs per GPC answers no, and the solve takes the cooperative-grid or one-pass path.
static bool cluster_ok(sm_ctx* c) {
    if (c->cluster_ok < 0) {
        c->cluster_ok = 0;
        if (cudaFuncSetAttribute(k_cg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg;
            cudaLaunchAttribute attr[1];
            cluster_launch_config(c, &cfg, attr);
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, k_cg_cluster, &cfg) == cudaSuccess && n >= 1) c->cluster_ok = 1;
        }
        cudaGetLastError();
    }
    return c->cluster_ok == 1;
}

static int dev_cg_cluster(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    ResidentCgArgs a;
    TRY(resident_args(c, U, phi, x, m0, &a));
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cluster_launch_config(c, &cfg, attr);
    CU(cudaLaunchKernelEx(&cfg, k_cg_cluster, a));
    return resident_finish(c, converged, iterations);
}

// how many sites the cooperative-grid solve can hold on this device (0: not available)
static int coop_capacity(sm_ctx* c) {
    if (c->coop_sites < 0) {
        int per_sm = 0, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_coop, kCoopThreads, 0) != cudaSuccess)
            per_sm = 0;
        c->coop_sites = std::min(per_sm * c->sm_count, kGridSyncMaxCtas) * kCoopThreads;
    }
    return c->coop_sites;
}

static int dev_cg_coop(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    ResidentCgArgs a;
    TRY(resident_args(c, U, phi, x, m0, &a));
    const int blocks = (c->V + kCoopThreads - 1) / kCoopThreads;
    if (!c->coop_hop) {
        TRY(dev_alloc(&c->coop_hop, (size_t)8 * c->V));
        TRY(dev_alloc(&c->coop_wsum, (size_t)4 * blocks * (kCoopThreads / 32)));
    }
    if (!c->coop_bar) TRY(dev_alloc(&c->coop_bar, (size_t)32));
    CU(cudaMemsetAsync(c->coop_bar, 0, sizeof(unsigned int) * 32, c->stream));
    a.hop = c->coop_hop;
    a.wsum = c->coop_wsum;
    a.bar = c->coop_bar;
    void* params[] = {&a};
    CU(cudaLaunchCooperativeKernel((const void*)k_cg_coop, dim3(blocks, 1, 1), dim3(kCoopThreads, 1, 1), params, 0,
                                   c->stream));
    return resident_finish(c, converged, iterations);
}

// Lattices beyond one site per thread of a full cooperative grid: k_cg_cols, S rows per thread in CTAs of T threads
// (sm_column_cg.cuh).  The variants compiled: T = 512 with S = 1..4, T = 256 with S = 2..8.
struct ColsVariant {
    int S, T;
    const void* kernel;
};
static const ColsVariant kColsVariants[] = {
    {1, 512, (const void*)k_cg_cols<1, 512>}, {2, 512, (const void*)k_cg_cols<2, 512>},
    {3, 512, (const void*)k_cg_cols<3, 512>}, {4, 512, (const void*)k_cg_cols<4, 512>},
    {2, 256, (const void*)k_cg_cols<2, 256>}, {3, 256, (const void*)k_cg_cols<3, 256>},
    {4, 256, (const void*)k_cg_cols<4, 256>}, {5, 256, (const void*)k_cg_cols<5, 256>},
    {6, 256, (const void*)k_cg_cols<6, 256>}, {7, 256, (const void*)k_cg_cols<7, 256>},
    {8, 256, (const void*)k_cg_cols<8, 256>},
};

static long long cols_blocks(const sm_ctx* c, int S, int T) {
    const long long threads = (long long)((c->wx + S - 1) / S) * c->wt;
    return (threads + T - 1) / T;
}

// does the variant hold this lattice on one co-resident grid?
static bool cols_fits(sm_ctx* c, const ColsVariant& v) {
    const size_t smem = cols_smem_bytes(v.S, v.T);
    int per_sm = 0;
    if (cudaFuncSetAttribute(v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.kernel, v.T, smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    // GridSync::sum_end adds at most kGridSyncMaxCtas CTA partials (5 per lane)
    return per_sm > 0 && cols_blocks(c, v.S, v.T) <= std::min<long long>((long long)per_sm * c->sm_count, kGridSyncMaxCtas);
}

// picks c->cols (variant index, -1: none) once per context.  SM_COLS=S,T forces a variant wherever it fits (also on
// lattices the one-site kernels would take); SM_COLS=0 disables the kernel.
static int cols_plan(sm_ctx* c) {
    if (c->cols_planned) return c->cols;
    c->cols_planned = true;
    c->cols = -1;
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
    // one tile only (the antiperiodic sign is folded into the links once), warps of >= 32 columns with <= 1 row wrap
    if (!coop || c->dist() || c->wt < 32 || c->sR_edge() != c->sL_edge() || !c->cols_enabled) return c->cols;
    const int nvar = (int)(sizeof(kColsVariants) / sizeof(kColsVariants[0]));
    if (c->cols_force_S > 0) {
        for (int i = 0; i < nvar; i++)
            if (kColsVariants[i].S == c->cols_force_S && kColsVariants[i].T == c->cols_force_T && cols_fits(c, kColsVariants[i]))
                c->cols = i;
        return c->cols;
    }
    if (c->V <= kClusterMaxCtas * kClusterThreads) return c->cols;     // one cluster holds it: k_cg_cluster
    // fewest rows per thread that fit, 256-thread CTAs first (255 registers per thread: no spills).  Measured
    // (profiles/r01_sweep_column_cg.txt): 256^2 9.2 us per iteration with 2 x 256 against 10.3 us with one site per
    // thread (k_cg_coop), 128^2 equal; 384^2 11.2 us (4 x 256) and 512^2 14.6 us (8 x 256) against 25.9 / 29.6 us with
    // CUDA graphs of the one-pass kernels.
    for (int T : {256, 512})
        for (int i = 0; i < nvar && c->cols < 0; i++)
            if (kColsVariants[i].T == T && cols_fits(c, kColsVariants[i])) c->cols = i;
    return c->cols;
}

static int dev_cg_cols(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    ResidentCgArgs a;
    TRY(resident_args(c, U, phi, x, m0, &a));
    const ColsVariant& v = kColsVariants[c->cols];
    const int blocks = (int)cols_blocks(c, v.S, v.T);
    if (!c->cols_hop) {
        TRY(dev_alloc(&c->cols_hop, (size_t)kColsBuffers * 4 * c->V));
        TRY(dev_alloc(&c->cols_wsum, (size_t)4 * (c->sm_count + 32)));
    }
    if (!c->coop_bar) TRY(dev_alloc(&c->coop_bar, (size_t)32));
    CU(cudaMemsetAsync(c->coop_bar, 0, sizeof(unsigned int) * 32, c->stream));
    a.hop = c->cols_hop;
    a.wsum = c->cols_wsum;
    a.bar = c->coop_bar;
    void* params[] = {&a};
    CU(cudaLaunchCooperativeKernel(v.kernel, dim3(blocks, 1, 1), dim3(v.T, 1, 1), params, cols_smem_bytes(v.S, v.T), c->stream));
    return resident_finish(c, converged, iterations);
}

// Opt-in even-odd solver: CG for  Dhat Dhat^dagger x = phi  on the even sites, Dhat the Schur complement of
// D (sm_ops.cuh: dev_Dhat).  Same recurrences, start vector (x_0 = phi) and stopping rule as the reference's CG
// (src/conjugate_gradient.cpp:4-67), another operator: its condition number is ~4x smaller near the critical mass
// (lambda_hat = lambda_+ lambda_- / m), so the solve takes a fraction of the iterations.  phi, x: full-lattice arrays,
// zero on the odd sites.  Iterations are launched in batches; the host polls the device scalars one batch behind.
// the whole even-odd solve in one cooperative launch (sm_evenodd_cg.cuh) while the working set of 7 fields lives in L2
static bool eo_coop_ok(sm_ctx* c) {
    if (c->eo_coop < 0) {
        c->eo_coop = 0;
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        const char* e = getenv("SM_EO_COOP");
        if (coop && !(e && atoi(e) == 0) && c->use_cluster && (long long)c->V <= (1LL << 21) &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_eo_coop, kEoThreads, 0) == cudaSuccess && per_sm >= 1)
            c->eo_coop = 1;
        cudaGetLastError();
    }
    return c->eo_coop == 1;
}

static int dev_cg_eo_coop(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    TRY(ensure_complex(c, &c->tmp));
    TRY(ensure_complex(c, &c->eo_t));
    TRY(ensure_complex(c, &c->cg_r));
    TRY(ensure_complex(c, &c->cg_d));
    TRY(ensure_complex(c, &c->cg_Ad));
    if (!c->eo_wsum) TRY(dev_alloc(&c->eo_wsum, (size_t)4 * kGridSyncMaxCtas));
    if (!c->coop_bar) TRY(dev_alloc(&c->coop_bar, (size_t)32));
    CU(cudaMemsetAsync(c->coop_bar, 0, sizeof(unsigned int) * 32, c->stream));
    EoCgArgs a{};
    a.U = U;
    a.phi = phi;
    a.x = x;
    a.r = c->cg_r;
    a.d = c->cg_d;
    a.t = c->eo_t;
    a.W = c->tmp;
    a.Ad = c->cg_Ad;
    a.wx = c->wx;
    a.wt = c->wt;
    a.V = c->V;
    a.mass = m0 + 2;
    a.sR_edge = c->sR_edge();
    a.sL_edge = c->sL_edge();
    a.tol = c->tol;
    a.max_iter = c->max_iter;
    a.st = c->cg;
    a.wsum = c->eo_wsum;
    a.bar = c->coop_bar;
    const int blocks = std::min(c->sm_count, kGridSyncMaxCtas);
    void* params[] = {&a};
    CU(cudaLaunchCooperativeKernel((const void*)k_cg_eo_coop, dim3(blocks, 1, 1), dim3(kEoThreads, 1, 1), params, 0, c->stream));
    return resident_finish(c, converged, iterations);
}

static int dev_cg_eo(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    NvtxRange nvtx("sm:conjugate_gradient(even-odd)");
    if (eo_coop_ok(c)) return dev_cg_eo_coop(c, U, phi, x, m0, converged, iterations);
    TRY(ensure_complex(c, &c->tmp));
    TRY(ensure_complex(c, &c->eo_t));
    TRY(ensure_complex(c, &c->cg_r));
    TRY(ensure_complex(c, &c->cg_d));
    TRY(ensure_complex(c, &c->cg_Ad));
    const int n_elems = 2 * c->V;
    const double tol = c->tol;
    const int max_iter = c->max_iter;
    CgState* st = c->cg;
    const int* done = &st->done;
    k_cg_reset<<<1, 1, 0, c->stream>>>(st, tol, max_iter);
    c->launches++;
    // A phi, then x = phi ; r = phi - A phi ; d = r ; |phi|^2, |r|^2
    TRY((dev_Dhat<true>(c, U, phi, c->tmp, c->eo_t, m0)));
    TRY((dev_Dhat<false>(c, U, c->eo_t, c->tmp, c->cg_Ad, m0)));
    k_cg_start<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(phi, c->cg_Ad, x, c->cg_r, c->cg_d, n_elems, c->partials,
                                                           c->tickets + TK_DOT, &st->phi_norm2);
    KCHECK();
    c->launches++;
    // one iteration (6 kernels); the iteration index lives in the device state, so a batch of 8 starting on an odd k
    // is one CUDA graph whatever k is (as in cg_fused_loop)
    auto iteration = [&](int k) -> int {
        const int cur = k & 1;
        if (k > 0) {
            k_cg_dir_dev<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, cur, c->cg_r, c->cg_d, n_elems);
            KCHECK();
            c->launches++;
        }
        TRY((dev_Dhat<true>(c, U, c->cg_d, c->tmp, c->eo_t, m0, nullptr, nullptr, done)));
        TRY((dev_Dhat<false>(c, U, c->eo_t, c->tmp, c->cg_Ad, m0, c->cg_d, st->dAd, done)));
        k_cg_update<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(st, cur, x, c->cg_d, c->cg_r, c->cg_Ad, n_elems, c->partials,
                                                                c->tickets + TK_UPDATE, &st->rr[cur ^ 1]);
        KCHECK();
        c->launches++;
        return SM_OK;
    };
    const int batch = 8;
    cudaGraphExec_t exec = nullptr;
    int graph_kernels = 0;
    const bool graphs = c->use_graphs && max_iter > batch;
    if (graphs) {
        for (auto& g : c->eo_graphs)
            if (g.U == (const void*)U && g.x == (const void*)x && g.m0 == m0) {
                exec = g.exec;
                graph_kernels = g.kernels;
            }
    }
    int k = 0, slot = 0, prev = -1;
    TRY(iteration(k++));
    if (graphs && exec == nullptr) {
        const long long l0 = c->launches;
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = SM_OK;
        for (int i = 0; i < batch && rc == SM_OK; i++) rc = iteration(1 + i);
        if (rc == SM_OK) {
            k_cg_check_dev<<<1, 1, 0, c->stream>>>(st, DistLink{});
            c->launches++;
        }
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        if (rc != SM_OK) return rc;
        if (e != cudaSuccess) return fail(SM_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        graph_kernels = (int)(c->launches - l0);
        c->launches = l0;
        CU(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        if (c->eo_graphs.size() >= 16) {
            cudaGraphExecDestroy(c->eo_graphs.front().exec);
            c->eo_graphs.erase(c->eo_graphs.begin());
        }
        c->eo_graphs.push_back({(const void*)U, (const void*)x, m0, exec, graph_kernels});
    }
    for (;;) {
        if (exec != nullptr && k + batch <= max_iter) {
            CU(cudaGraphLaunch(exec, c->stream));
            c->launches += graph_kernels;
            k += batch;
        } else {
            const int k_end = std::min(max_iter, k + batch);
            for (; k < k_end; k++) TRY(iteration(k));
            k_cg_check_dev<<<1, 1, 0, c->stream>>>(st, DistLink{});
            KCHECK();
            c->launches++;
        }
        CU(cudaMemcpyAsync(&c->h->cg[slot], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaEventRecord(c->ev_poll[slot], c->stream));
        if (prev >= 0) {
            CU(cudaEventSynchronize(c->ev_poll[prev]));
            if (c->h->cg[prev].done) break;
        }
        if (k >= max_iter) break;
        prev = slot;
        slot ^= 1;
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpyAsync(&c->h->cg[0], st, sizeof(CgState), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (converged) *converged = c->h->cg[0].converged;
    if (iterations) *iterations = c->h->cg[0].iters;
    return SM_OK;
}

static int dev_cg_route(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    if (c->use_cluster && !c->dist()) {
        if (cols_plan(c) >= 0) return dev_cg_cols(c, U, phi, x, m0, converged, iterations);
        if (c->V <= kClusterMaxCtas * kClusterThreads && cluster_ok(c)) return dev_cg_cluster(c, U, phi, x, m0, converged, iterations);
        if (c->V <= coop_capacity(c)) return dev_cg_coop(c, U, phi, x, m0, converged, iterations);   // width_t < 32 or SM_COLS=0
    }
    if (c->solver == SM_SOLVER_MIXED && fused_ok(c) && !c->dist()) return dev_cg_mixed(c, U, phi, x, m0, converged, iterations);
    if (fused_ok(c)) return dev_cg_fused(c, U, phi, x, m0, converged, iterations);
    return dev_cg_twopass(c, U, phi, x, m0, converged, iterations);
}

// x0: start vector (null: x_0 = phi, the reference's choice, src/conjugate_gradient.cpp:16).  Everything else -- the
// recurrences, the stopping rule |r| < tol |phi| -- is the same; only the opt-in chronological solver passes one.
static int dev_cg(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations,
                  const cplx* x0 = nullptr) {
    NvtxRange nvtx("sm:conjugate_gradient");
    c->cg_x0 = (c->solver == SM_SOLVER_MIXED) ? nullptr : x0;
    const int rc = dev_cg_route(c, U, phi, x, m0, converged, iterations);
    c->cg_x0 = nullptr;
    return rc;
}
