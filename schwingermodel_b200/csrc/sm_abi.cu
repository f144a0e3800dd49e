// sm_abi.cu -- the extern "C" surface of libschwinger_b200.so (include/schwinger_b200.h).
// One translation unit: sm_context.cuh -> sm_dist.cuh -> sm_ops.cuh -> sm_cg.cuh -> sm_hmc.cuh hold the
// context and the launch logic; the kernels are in sm_kernels.cuh, sm_fused.cuh, sm_cluster_cg.cuh.
#include "sm_hmc.cuh"

// ================================================================================================
// extern "C"
// ================================================================================================
// ---- peer-memory windows (halo rows and CG sums over NVLink) -----------------------------------------
static int p2p_make_window(sm_ctx* c, cudaIpcMemHandle_t* h) {
    if (!c->dist() || (c->rt != 1 && c->rx != 1) || c->wx < 4 || c->wt < 4)
        return fail(SM_ERR_STATE, "peer-memory halos need a lattice split along one axis only");
    if (!c->win) {
        c->win_bytes = win_total_bytes(c);
        CU(cudaMalloc((void**)&c->win, c->win_bytes));
        CU(cudaMemset(c->win, 0, c->win_bytes));
        c->win_flags = win_flag(c, c->win, 0, 0);
        TRY(dev_alloc(&c->push_ticket, (size_t)1));
        CU(cudaMemset(c->push_ticket, 0, sizeof(unsigned int)));
    }
    CU(cudaIpcGetMemHandle(h, c->win));
    return SM_OK;
}

// map the neighbours' windows (halo rows) and, on up to kMaxPeers ranks, every rank's window (CG sums)
static int p2p_open_windows(sm_ctx* c, const void* all_handles) {
    if (c->p2p) return SM_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) != cudaSuccess || fn == nullptr)
        return fail(SM_ERR_CUDA, "cuStreamWaitValue32 is not available");
    c->wait_value32 = (CUresult(*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int))fn;
    const bool along_t = c->rt > 1;                      // split along t: neighbours -t / +t, halos only (sums stay with NCCL)
    const int nb_lo = along_t ? c->nb_tm : c->nb_xm, nb_hi = along_t ? c->nb_tp : c->nb_xp;
    const bool all = c->nranks <= kMaxPeers && !along_t;
    std::vector<void*> opened(c->nranks, nullptr);
    opened[c->rank] = c->win;
    auto open_rank = [&](int r) -> int {
        if (opened[r]) return SM_OK;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + (size_t)r * SM_P2P_HANDLE_BYTES, sizeof(h));
        CU(cudaIpcOpenMemHandle(&opened[r], h, cudaIpcMemLazyEnablePeerAccess));
        return SM_OK;
    };
    int rc = SM_OK;
    for (int r = 0; r < c->nranks && rc == SM_OK; r++)
        if (all || r == nb_lo || r == nb_hi) rc = open_rank(r);
    if (rc != SM_OK) {
        for (int r = 0; r < c->nranks; r++)
            if (opened[r] && r != c->rank) cudaIpcCloseMemHandle(opened[r]);
        cudaGetLastError();
        return rc;
    }
    c->peer_win[0] = opened[nb_lo];
    c->peer_win[1] = opened[nb_hi];
    if (all)
        for (int r = 0; r < c->nranks; r++) c->peer_all[r] = opened[r];
    c->p2p = true;
    c->peer_sums = all && c->use_fused;     // the kernels that do their own collectives belong to the one-pass CG
    return SM_OK;
}

// Called at the end of sm_create_dist on lattices split along x: every rank exports its window, the handles travel
// by ncclAllGather, every rank maps its peers.  All ranks then agree (ncclAllReduce of a flag) on whether the windows
// are usable -- ranks in one process, or GPUs without peer access, fall back to NCCL send/recv + all-reduce together.
static int p2p_auto_connect(sm_ctx* c) {
    if ((c->rt != 1 && c->rx != 1) || c->wx < 4 || c->wt < 4) return SM_OK;
    if (c->rt > 1 && !(c->tsplit_onepass && c->fused_tma)) return SM_OK;
    if (const char* e = getenv("SM_P2P"))
        if (atoi(e) == 0) return SM_OK;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    int ok = (p2p_make_window(c, &h) == SM_OK) ? 1 : 0;
    struct DevBuf {          // released on every return path
        void* p = nullptr;
        ~DevBuf() {
            if (p) cudaFree(p);
        }
    } b_handles, b_ok;
    CU(cudaMalloc(&b_handles.p, (size_t)c->nranks * SM_P2P_HANDLE_BYTES));
    CU(cudaMalloc(&b_ok.p, sizeof(int)));
    unsigned char* d_handles = (unsigned char*)b_handles.p;
    int* d_ok = (int*)b_ok.p;
    CU(cudaMemcpyAsync(d_handles + (size_t)c->rank * SM_P2P_HANDLE_BYTES, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(d_handles + (size_t)c->rank * SM_P2P_HANDLE_BYTES, d_handles, SM_P2P_HANDLE_BYTES, ncclChar, c->comm, c->stream));
    std::vector<unsigned char> handles((size_t)c->nranks * SM_P2P_HANDLE_BYTES);
    CU(cudaMemcpyAsync(handles.data(), d_handles, handles.size(), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (ok && p2p_open_windows(c, handles.data()) != SM_OK) ok = 0;
    CU(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c->comm, c->stream));
    int all_ok = 0;
    CU(cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!all_ok) {       // together or not at all
        c->p2p = false;
        c->peer_sums = false;
    }
    return SM_OK;
}

extern "C" {

const char* sm_last_error(void) { return g_err.c_str(); }

int sm_create(int Nx, int Nt, int device, sm_ctx** out) {
    NEED(out);
    *out = nullptr;
    if (Nx < 2 || Nt < 2) return fail(SM_ERR_ARG, "lattice must be at least 2x2");
    if ((long long)Nx * Nt > (1LL << 29)) return fail(SM_ERR_ARG, "tile too large for 32-bit site indices");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SM_ERR_CUDA, std::string("no CUDA device (libschwinger_b200 has no CPU fallback): ") +
                                     cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SM_ERR_ARG, "device index out of range");
    sm_ctx* c = new sm_ctx();
    c->Nx = Nx;
    c->Nt = Nt;
    c->wx = Nx;
    c->wt = Nt;
    c->V = Nx * Nt;
    c->device = device;
    int rc = ctx_common_init(c);
    if (rc != SM_OK) {
        delete c;
        return rc;
    }
    *out = c;
    return SM_OK;
}

int sm_nccl_unique_id(void* out_id) {
    NEED(out_id);
    TRY(nccl_load());
    static_assert(sizeof(ncclUniqueId) <= SM_NCCL_ID_BYTES, "id size");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    memset(out_id, 0, SM_NCCL_ID_BYTES);
    memcpy(out_id, &id, sizeof(id));
    return SM_OK;
}

int sm_create_dist(int Nx, int Nt, int ranks_x, int ranks_t, int rank, int device, const void* nccl_id, sm_ctx** out) {
    NEED(out);
    *out = nullptr;
    if (ranks_x < 1 || ranks_t < 1) return fail(SM_ERR_ARG, "ranks_x and ranks_t must be >= 1");
    const int nranks = ranks_x * ranks_t;
    if (nranks == 1) return sm_create(Nx, Nt, device, out);
    NEED(nccl_id);
    if (rank < 0 || rank >= nranks) return fail(SM_ERR_ARG, "rank out of range");
    // equal tiles only, as the reference enforces (include/mpi_setup.h:9-20)
    if (Nx % ranks_x != 0) return fail(SM_ERR_ARG, "Nx is not divisible by ranks_x");
    if (Nt % ranks_t != 0) return fail(SM_ERR_ARG, "Nt is not divisible by ranks_t");
    if (Nx / ranks_x < 2 || Nt / ranks_t < 2) return fail(SM_ERR_ARG, "tiles must be at least 2x2");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SM_ERR_CUDA, std::string("no CUDA device (libschwinger_b200 has no CPU fallback): ") +
                                     cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(SM_ERR_ARG, "device index out of range");
    TRY(nccl_load());
    sm_ctx* c = new sm_ctx();
    c->Nx = Nx;
    c->Nt = Nt;
    c->rx = ranks_x;
    c->rt = ranks_t;
    c->rank = rank;
    c->nranks = nranks;
    c->cx = rank / ranks_t;
    c->ct = rank % ranks_t;
    c->wx = Nx / ranks_x;
    c->wt = Nt / ranks_t;
    c->V = c->wx * c->wt;
    c->device = device;
    // Cartesian neighbours, periodic (include/mpi_setup.h:39-71): x-1 "top", x+1 "bot", t-1 "left", t+1 "right"
    auto rk = [&](int cx, int ct) { return ((cx + ranks_x) % ranks_x) * ranks_t + (ct + ranks_t) % ranks_t; };
    c->nb_xm = rk(c->cx - 1, c->ct);
    c->nb_xp = rk(c->cx + 1, c->ct);
    c->nb_tm = rk(c->cx, c->ct - 1);
    c->nb_tp = rk(c->cx, c->ct + 1);
    int rc = ctx_common_init(c);
    if (rc != SM_OK) {
        delete c;
        return rc;
    }
    auto body = [&]() -> int {
        ncclUniqueId id;
        memcpy(&id, nccl_id, sizeof(id));
        NC(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
        const size_t wx = c->wx, wt = c->wt, W = wt + 2;
        TRY(dev_alloc(&c->send_tm, wx));
        TRY(dev_alloc(&c->send_tp, wx));
        TRY(dev_alloc(&c->send_xm, wt));
        TRY(dev_alloc(&c->send_xp, wt));
        TRY(dev_alloc(&c->g_tp, wx));
        TRY(dev_alloc(&c->g_tm, wx));
        TRY(dev_alloc(&c->g_xp, wt));
        TRY(dev_alloc(&c->g_xm, wt));
        TRY(dev_alloc(&c->gg_tm, 2 * wx));
        TRY(dev_alloc(&c->gg_tp, 2 * wx));
        TRY(dev_alloc(&c->gg_xm, 2 * W));
        TRY(dev_alloc(&c->gg_xp, 2 * W));
        TRY(dev_alloc(&c->gg_send, 4 * wx + 4 * W));
        TRY(dev_alloc(&c->fg_t, 2 * wx));
        TRY(dev_alloc(&c->fg_x, 2 * wt));
        TRY(dev_alloc(&c->fg_send, 2 * wx + 2 * wt));
        for (int side = 0; side < 2; side++) {
            TRY(dev_alloc(&c->f2_U[side], 4 * wt));
            TRY(dev_alloc(&c->f2_in[side], 4 * wt));
            TRY(dev_alloc(&c->f2_r[side], 4 * wt));
            TRY(dev_alloc(&c->f2_d[0][side], 4 * wt));
            TRY(dev_alloc(&c->f2_d[1][side], 4 * wt));
            CU(cudaMemsetAsync(c->f2_d[0][side], 0, sizeof(cplx) * 4 * wt, c->stream));
            CU(cudaMemsetAsync(c->f2_d[1][side], 0, sizeof(cplx) * 4 * wt, c->stream));
        }
        return p2p_auto_connect(c);
    };
    rc = body();
    if (rc != SM_OK) {
        delete c;
        return rc;
    }
    *out = c;
    return SM_OK;
}

int sm_destroy(sm_ctx* c) {
    if (!c) return SM_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->p2p) {
        std::vector<void*> closed;
        auto close_once = [&](void* p) {
            if (p && p != (void*)c->win && std::find(closed.begin(), closed.end(), p) == closed.end()) {
                cudaIpcCloseMemHandle(p);
                closed.push_back(p);
            }
        };
        close_once(c->peer_win[0]);
        close_once(c->peer_win[1]);
        for (void* p : c->peer_all) close_once(p);
    }
    if (c->win) cudaFree(c->win);
    if (c->push_ticket) cudaFree(c->push_ticket);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    for (int kind = 0; kind < 5; kind++)
        for (int side = 0; side < 2; side++) {
            if (c->tg_col[kind][side]) cudaFree(c->tg_col[kind][side]);
            if (c->tg_row[kind][side]) cudaFree(c->tg_row[kind][side]);
        }
    if (c->tg_sendc) cudaFree(c->tg_sendc);
    if (c->tg_sendr) cudaFree(c->tg_sendr);
    if (c->eo_t) cudaFree(c->eo_t);
    if (c->eo_wsum) cudaFree(c->eo_wsum);
    if (c->chrono_prev) cudaFree(c->chrono_prev);
    if (c->chrono_guess) cudaFree(c->chrono_guess);
    void* ptrs[] = {c->partials, c->tickets, c->cg,      c->sums,    c->sums_loc, c->tmp,     c->cg_r,   c->cg_d,
                    c->cg_Ad,    c->cg_d2, c->sU,      c->sA,      c->sB,      c->sC,       c->sF,      c->U,      c->Up,
                    c->chi,      c->phi,     c->psi,     c->xi,      c->pi,       c->pip,     c->F,      c->send_tm,
                    c->send_tp,  c->send_xm, c->send_xp, c->g_tp,    c->g_tm,     c->g_xp,    c->g_xm,   c->gg_xm,
                    c->gg_xp,    c->gg_tm,   c->gg_tp,   c->gg_send, c->fg_t,     c->fg_x,    c->fg_send,
                    c->f2_U[0],  c->f2_U[1], c->f2_in[0], c->f2_in[1], c->f2_r[0], c->f2_r[1],
                    c->f2_d[0][0], c->f2_d[0][1], c->f2_d[1][0], c->f2_d[1][1]};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (void* p : c->user_fields) cudaFree(p);
    for (void* p : {(void*)c->mx_U, (void*)c->mx_r, (void*)c->mx_e, (void*)c->mx_d0, (void*)c->mx_d1, (void*)c->mx_Ad})
        if (p) cudaFree(p);
    if (c->coop_hop) cudaFree(c->coop_hop);
    if (c->cols_hop) cudaFree(c->cols_hop);
    if (c->cols_wsum) cudaFree(c->cols_wsum);
    if (c->coop_wsum) cudaFree(c->coop_wsum);
    if (c->coop_bar) cudaFree(c->coop_bar);
    for (auto& g : c->cg_graphs) cudaGraphExecDestroy(g.exec);
    for (auto& g : c->eo_graphs) cudaGraphExecDestroy(g.exec);
    if (c->h) cudaFreeHost(c->h);
    cudaEventDestroy(c->ev_a);
    cudaEventDestroy(c->ev_b);
    cudaEventDestroy(c->ev_poll[0]);
    cudaEventDestroy(c->ev_poll[1]);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->comm_stream);
    cudaEventDestroy(c->ev_ready);
    cudaEventDestroy(c->ev_t0);
    cudaEventDestroy(c->ev_t1);
    cudaEventDestroy(c->ev_ghost);
    cudaEventDestroy(c->ev_packed);
    delete c;
    return SM_OK;
}

int sm_local_dims(const sm_ctx* c, int dims[4]) {
    NEED(c);
    NEED(dims);
    dims[0] = c->wx;
    dims[1] = c->wt;
    dims[2] = c->rank;
    dims[3] = c->nranks;
    return SM_OK;
}

int sm_set_cg(sm_ctx* c, double tol, int max_iter) {
    NEED(c);
    if (!(tol > 0) || max_iter < 1 || max_iter > 60000) return fail(SM_ERR_ARG, "tol must be > 0 and 1 <= max_iter <= 60000");
    c->tol = tol;
    c->max_iter = max_iter;
    return SM_OK;
}

int sm_set_solver(sm_ctx* c, int solver) {
    NEED(c);
    if (solver != SM_SOLVER_REFERENCE && solver != SM_SOLVER_MIXED && solver != SM_SOLVER_CHRONO && solver != SM_SOLVER_EVENODD)
        return fail(SM_ERR_ARG, "unknown solver");
    if (solver == SM_SOLVER_EVENODD && (c->dist() || (c->Nx & 1) || (c->Nt & 1)))
        return fail(SM_ERR_ARG, "the even-odd solver needs a single tile with even Nx and Nt");
    c->solver = solver;
    return SM_OK;
}

int sm_last_kernel_ms(const sm_ctx* c, double* ms) {
    NEED(c);
    NEED(ms);
    *ms = c->last_ms;
    return SM_OK;
}

int sm_one_pass_dd(const sm_ctx* c, int* one_pass) {
    NEED(c);
    NEED(one_pass);
    *one_pass = fused_ok(c) ? 1 : 0;
    return SM_OK;
}

int sm_host_register(int enable) {
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        g_pin_enabled = enable != 0;
    }
    if (!enable) host_forget(nullptr);
    return SM_OK;
}

int sm_host_forget(const void* ptr) {
    if (ptr != nullptr) host_forget(ptr);
    return SM_OK;
}

int sm_device_count(int* n) {
    NEED(n);
    *n = 0;
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) return fail(SM_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return SM_OK;
}

int sm_peer_mode(const sm_ctx* c, int* mode) {
    NEED(c);
    NEED(mode);
    *mode = c->peer_sums ? 2 : (c->p2p ? 1 : 0);
    return SM_OK;
}

int sm_launch_count(const sm_ctx* c, long long* n) {
    NEED(c);
    NEED(n);
    *n = c->launches;
    return SM_OK;
}

int sm_tables(sm_ctx* c, int ranks_x, int ranks_t, int rank, int* RightPB, int* LeftPB, double* SignR, double* SignL,
              int* x_1_t1, int* x1_t_1) {
    TRY(set_device(c));
    NEED(RightPB); NEED(LeftPB); NEED(SignR); NEED(SignL); NEED(x_1_t1); NEED(x1_t_1);
    if (ranks_x < 1 || ranks_t < 1 || c->Nx % ranks_x || c->Nt % ranks_t || rank < 0 || rank >= ranks_x * ranks_t)
        return fail(SM_ERR_ARG, "bad decomposition");
    const int wx = c->Nx / ranks_x, wt = c->Nt / ranks_t, m = wx * wt;
    // antiperiodic seam keyed on the rank exactly as the reference (include/dirac_operator.h:53-58)
    const double sR = ((rank + 1) % ranks_t == 0) ? -1.0 : 1.0;
    const double sL = (rank % ranks_t == 0) ? -1.0 : 1.0;
    int *dR = nullptr, *dL = nullptr, *dA = nullptr, *dB = nullptr;
    double *dsR = nullptr, *dsL = nullptr;
    struct Release {       // the scratch tables go on every return path
        void** p[6];
        ~Release() {
            for (void** q : p)
                if (*q) cudaFree(*q);
        }
    } release{{(void**)&dR, (void**)&dL, (void**)&dA, (void**)&dB, (void**)&dsR, (void**)&dsL}};
    TRY(dev_alloc(&dR, (size_t)2 * m));
    TRY(dev_alloc(&dL, (size_t)2 * m));
    TRY(dev_alloc(&dA, (size_t)m));
    TRY(dev_alloc(&dB, (size_t)m));
    TRY(dev_alloc(&dsR, (size_t)4 * m));
    TRY(dev_alloc(&dsL, (size_t)4 * m));
    k_tables<<<(m + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(wx, wt, sR, sL, dR, dL, dsR, dsL, dA, dB);
    KCHECK();
    c->launches++;
    CU(cudaMemcpyAsync(RightPB, dR, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(LeftPB, dL, sizeof(int) * 2 * m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(x_1_t1, dA, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(x1_t_1, dB, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(SignR, dsR, sizeof(double) * 4 * m, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(SignL, dsL, sizeof(double) * 4 * m, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}

// ---- peer-memory windows: the entry points ------------------------------------------------------------
int sm_p2p_handle(sm_ctx* c, void* handle_out) {
    TRY(set_device(c));
    NEED(handle_out);
    static_assert(sizeof(cudaIpcMemHandle_t) == SM_P2P_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    TRY(p2p_make_window(c, &h));
    memcpy(handle_out, &h, sizeof(h));
    return SM_OK;
}

int sm_p2p_connect(sm_ctx* c, const void* all_handles) {
    TRY(set_device(c));
    NEED(all_handles);
    if (!c->win) return fail(SM_ERR_STATE, "sm_p2p_handle first");
    return p2p_open_windows(c, all_handles);
}

// ---- host-buffer operators ---------------------------------------------------------------------
static int host_stencil(sm_ctx* c, const double* U0, const double* U1, const double* p0, const double* p1, double* o0,
                        double* o1, double m0, int which) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(p0); NEED(p1); NEED(o0); NEED(o1);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    TRY(h2d_c(c, c->sA, p0, p1));
    tick(c);
    if (which == 0) TRY(dev_D(c, c->sU, c->sA, c->sB, m0, false));
    if (which == 1) TRY(dev_D(c, c->sU, c->sA, c->sB, m0, true));
    if (which == 2) TRY(dev_DDdag(c, c->sU, c->sA, c->sB, m0));
    TRY(tock(c));
    TRY(d2h_c(c, c->sB, o0, o1));
    return sync(c);
}

int sm_D_phi(sm_ctx* c, const double* U0, const double* U1, const double* p0, const double* p1, double* o0, double* o1,
             double m0) {
    return host_stencil(c, U0, U1, p0, p1, o0, o1, m0, 0);
}
int sm_D_dagger_phi(sm_ctx* c, const double* U0, const double* U1, const double* p0, const double* p1, double* o0,
                    double* o1, double m0) {
    return host_stencil(c, U0, U1, p0, p1, o0, o1, m0, 1);
}
int sm_D_D_dagger_phi(sm_ctx* c, const double* U0, const double* U1, const double* p0, const double* p1, double* o0,
                      double* o1, double m0) {
    return host_stencil(c, U0, U1, p0, p1, o0, o1, m0, 2);
}

int sm_dot(sm_ctx* c, const double* x0, const double* x1, const double* y0, const double* y1, double out[2]) {
    TRY(set_device(c));
    NEED(x0); NEED(x1); NEED(y0); NEED(y1); NEED(out);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sA, x0, x1));
    TRY(h2d_c(c, c->sB, y0, y1));
    tick(c);
    TRY(dev_dot_async(c, c->sA, c->sB, c->sums));
    TRY(tock(c));
    TRY(fetch_sums(c, 2));
    out[0] = c->h->sums[0];
    out[1] = c->h->sums[1];
    return SM_OK;
}

int sm_conjugate_gradient(sm_ctx* c, const double* U0, const double* U1, const double* p0, const double* p1, double* x0,
                          double* x1, double m0, int* converged, int* iterations) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(p0); NEED(p1); NEED(x0); NEED(x1);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    TRY(h2d_c(c, c->sA, p0, p1));
    tick(c);
    TRY(dev_cg(c, c->sU, c->sA, c->sB, m0, converged, iterations));
    TRY(tock(c));
    TRY(d2h_c(c, c->sB, x0, x1));
    return sync(c);
}

int sm_evenodd_solve(sm_ctx* c, const double* U0, const double* U1, const double* p0, const double* p1, double* x0,
                     double* x1, double m0, int* converged, int* iterations) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(p0); NEED(p1); NEED(x0); NEED(x1);
    if (c->dist() || (c->Nx & 1) || (c->Nt & 1)) return fail(SM_ERR_ARG, "the even-odd solver needs a single tile with even Nx and Nt");
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    TRY(h2d_c(c, c->sA, p0, p1));
    k_mask_parity<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(c->sA, c->wt, c->V, 0);
    KCHECK();
    c->launches++;
    tick(c);
    TRY(dev_cg_eo(c, c->sU, c->sA, c->sB, m0, converged, iterations));
    TRY(tock(c));
    TRY(d2h_c(c, c->sB, x0, x1));
    return sync(c);
}

int sm_phi_dag_partialD_phi(sm_ctx* c, const double* U0, const double* U1, const double* l0, const double* l1,
                            const double* r0, const double* r1, double* F0, double* F1) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(l0); NEED(l1); NEED(r0); NEED(r1); NEED(F0); NEED(F1);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    TRY(h2d_c(c, c->sA, l0, l1));
    TRY(h2d_c(c, c->sB, r0, r1));
    c->ghost_valid_for = c->f2_U_valid_for = c->tg_U_valid_for = nullptr;
    tick(c);
    TRY(dev_force(c, c->sU, c->sA, c->sB, c->sF, 0.0, true, false));
    TRY(tock(c));
    TRY(d2h_r(c, c->sF, F0, F1));
    return sync(c);
}

int sm_compute_staple(sm_ctx* c, const double* U0, const double* U1, double* K0, double* K1) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(K0); NEED(K1);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    c->ghost_valid_for = c->f2_U_valid_for = c->tg_U_valid_for = nullptr;
    tick(c);
    TRY(refresh_gauge_ghosts(c, c->sU));
    k_staple<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(gauge_view(c, c->sU), c->sB);
    KCHECK();
    c->launches++;
    TRY(tock(c));
    TRY(d2h_c(c, c->sB, K0, K1));
    return sync(c);
}

int sm_compute_plaquette(sm_ctx* c, const double* U0, const double* U1, double beta, double* P, double sums[2]) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(sums);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    c->ghost_valid_for = c->f2_U_valid_for = c->tg_U_valid_for = nullptr;
    tick(c);
    TRY(dev_plaquette(c, c->sU, beta, P ? c->sA : nullptr, c->sums));
    TRY(tock(c));
    if (P) CU(cudaMemcpyAsync(P, c->sA, sizeof(cplx) * c->V, cudaMemcpyDeviceToHost, c->stream));
    TRY(fetch_sums(c, 2));
    sums[0] = c->h->sums[0];
    sums[1] = c->h->sums[1];
    return SM_OK;
}

// ---- device-resident fields ----------------------------------------------------------------------
int sm_field_alloc(sm_ctx* c, int complex_field, double** d_field) {
    TRY(set_device(c));
    NEED(d_field);
    void* p = nullptr;
    const size_t bytes = (complex_field ? sizeof(cplx) : sizeof(double)) * 2 * (size_t)c->V;
    CU(cudaMalloc(&p, bytes));
    CU(cudaMemsetAsync(p, 0, bytes, c->stream));
    c->user_fields.push_back(p);
    *d_field = (double*)p;
    return SM_OK;
}

int sm_field_free(sm_ctx* c, double* d_field) {
    TRY(set_device(c));
    auto it = std::find(c->user_fields.begin(), c->user_fields.end(), (void*)d_field);
    if (it == c->user_fields.end()) return fail(SM_ERR_ARG, "not a field of this context");
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(*it);
    c->user_fields.erase(it);
    return SM_OK;
}

int sm_field_upload(sm_ctx* c, double* d, const double* h0, const double* h1, int complex_field) {
    TRY(set_device(c));
    NEED(d); NEED(h0); NEED(h1);
    invalidate_gauge_ghosts(c, (const cplx*)d);
    if (complex_field) TRY(h2d_c(c, (cplx*)d, h0, h1));
    else TRY(h2d_r(c, d, h0, h1));
    return sync(c);
}

int sm_field_download(sm_ctx* c, const double* d, double* h0, double* h1, int complex_field) {
    TRY(set_device(c));
    NEED(d); NEED(h0); NEED(h1);
    if (complex_field) TRY(d2h_c(c, (const cplx*)d, h0, h1));
    else TRY(d2h_r(c, d, h0, h1));
    return sync(c);
}

int sm_dev_D(sm_ctx* c, const double* dU, const double* din, double* dout, double m0, int dagger) {
    TRY(set_device(c));
    NEED(dU); NEED(din); NEED(dout);
    tick(c);
    TRY(dev_D(c, (const cplx*)dU, (const cplx*)din, (cplx*)dout, m0, dagger != 0));
    return tock(c);
}

int sm_dev_DDdag(sm_ctx* c, const double* dU, const double* din, double* dout, double m0) {
    TRY(set_device(c));
    NEED(dU); NEED(din); NEED(dout);
    tick(c);
    TRY(dev_DDdag(c, (const cplx*)dU, (const cplx*)din, (cplx*)dout, m0));
    return tock(c);
}

int sm_dev_DDdag_loop(sm_ctx* c, const double* dU, const double* din, double* dout, double m0, int reps,
                      double* ms_total) {
    TRY(set_device(c));
    NEED(dU); NEED(din); NEED(dout);
    if (reps < 1) return fail(SM_ERR_ARG, "reps must be >= 1");
    TRY(ensure_complex(c, &c->tmp));
    tick(c);
    for (int i = 0; i < reps; i++) TRY(dev_DDdag(c, (const cplx*)dU, (const cplx*)din, (cplx*)dout, m0));
    TRY(tock(c));
    if (ms_total) *ms_total = c->last_ms;
    return SM_OK;
}

int sm_dev_dot(sm_ctx* c, const double* dx, const double* dy, double out[2]) {
    TRY(set_device(c));
    NEED(dx); NEED(dy); NEED(out);
    tick(c);
    TRY(dev_dot_async(c, (const cplx*)dx, (const cplx*)dy, c->sums));
    TRY(tock(c));
    TRY(fetch_sums(c, 2));
    out[0] = c->h->sums[0];
    out[1] = c->h->sums[1];
    return SM_OK;
}

int sm_dev_cg(sm_ctx* c, const double* dU, const double* dphi, double* dx, double m0, int* converged, int* iterations) {
    TRY(set_device(c));
    NEED(dU); NEED(dphi); NEED(dx);
    tick(c);
    TRY(dev_cg(c, (const cplx*)dU, (const cplx*)dphi, (cplx*)dx, m0, converged, iterations));
    return tock(c);
}

// ---- HMC ------------------------------------------------------------------------------------------
int sm_hmc_configure(sm_ctx* c, const sm_hmc_params* p) {
    TRY(set_device(c));
    NEED(p);
    if (p->md_steps < 1) return fail(SM_ERR_ARG, "md_steps must be >= 1");
    c->hp = *p;
    return hmc_alloc(c);
}

int sm_hmc_set_gauge(sm_ctx* c, const double* U0, const double* U1) {
    TRY(set_device(c));
    NEED(U0); NEED(U1);
    TRY(hmc_alloc(c));
    c->ghost_valid_for = c->f2_U_valid_for = c->tg_U_valid_for = nullptr;
    TRY(h2d_c(c, c->U, U0, U1));
    c->hmc_has_gauge = true;
    return sync(c);
}

int sm_hmc_get_gauge(sm_ctx* c, double* U0, double* U1, int proposal) {
    TRY(set_device(c));
    NEED(U0); NEED(U1);
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "no gauge field set");
    TRY(d2h_c(c, proposal ? c->Up : c->U, U0, U1));
    return sync(c);
}

int sm_hmc_get_momenta(sm_ctx* c, double* p0, double* p1, int proposal) {
    TRY(set_device(c));
    NEED(p0); NEED(p1);
    if (!c->hmc_ready) return fail(SM_ERR_STATE, "HMC not configured");
    TRY(d2h_r(c, proposal ? c->pip : c->pi, p0, p1));
    return sync(c);
}

int sm_hmc_get_phi(sm_ctx* c, double* p0, double* p1) {
    TRY(set_device(c));
    NEED(p0); NEED(p1);
    if (!c->hmc_ready) return fail(SM_ERR_STATE, "HMC not configured");
    TRY(d2h_c(c, c->phi, p0, p1));
    return sync(c);
}

int sm_hmc_get_chi(sm_ctx* c, double* p0, double* p1) {
    TRY(set_device(c));
    NEED(p0); NEED(p1);
    if (!c->hmc_ready) return fail(SM_ERR_STATE, "HMC not configured");
    TRY(d2h_c(c, c->chi, p0, p1));
    return sync(c);
}

int sm_hmc_refresh(sm_ctx* c, uint64_t seed, uint64_t trajectory_index) {
    TRY(set_device(c));
    TRY(hmc_alloc(c));
    TileMap m{c->wx, c->wt, c->cx * c->wx, c->ct * c->wt, c->Nt, (long long)c->Nx * c->Nt};
    k_refresh<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(c->pi, c->chi, m, seed, trajectory_index);
    KCHECK();
    c->launches++;
    c->hmc_has_fields = true;
    return SM_OK;
}

int sm_hmc_inject(sm_ctx* c, const double* pi0, const double* pi1, const double* chi0, const double* chi1) {
    TRY(set_device(c));
    NEED(pi0); NEED(pi1); NEED(chi0); NEED(chi1);
    TRY(hmc_alloc(c));
    TRY(h2d_r(c, c->pi, pi0, pi1));
    TRY(h2d_c(c, c->chi, chi0, chi1));
    c->hmc_has_fields = true;
    return sync(c);
}

int sm_hmc_trajectory(sm_ctx* c, sm_traj_result* out) {
    TRY(set_device(c));
    NEED(out);
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "sm_hmc_set_gauge first");
    if (!c->hmc_has_fields) return fail(SM_ERR_STATE, "sm_hmc_refresh or sm_hmc_inject first");
    if (c->hp.md_steps < 1) return fail(SM_ERR_STATE, "sm_hmc_configure first");
    NvtxRange nvtx("sm:HMC_Update trajectory");
    TrajAcc acc;
    CU(cudaEventRecord(c->ev_t0, c->stream));
    TRY(hmc_pseudofermion(c));                                            // hmc.cpp:160
    TRY(hmc_leapfrog(c, &acc));                                           // hmc.cpp:161
    TRY(hmc_hamiltonian_async(c, c->Up, c->pip, c->phi, 0, &acc));        // hmc.cpp:162 (new)
    TRY(hmc_hamiltonian_async(c, c->U, c->pi, c->phi, 5, &acc));          //             (old)
    CU(cudaEventRecord(c->ev_t1, c->stream));
    TRY(fetch_sums(c, 10));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1));
    const double* s = c->h->sums;
    out->H_new = hamiltonian_from(s);
    out->H_old = hamiltonian_from(s + 5);
    out->dH = out->H_new - out->H_old;
    out->sum_re_plaq_new = s[1];
    out->gauge_action_new = s[2];
    out->sum_re_plaq_old = s[6];
    out->gauge_action_old = s[7];
    out->dd_applications = acc.dd_apps;
    out->cg_solves = acc.solves;
    out->cg_all_converged = acc.all_ok;
    out->cg_force_failures = acc.force_fail;
    out->reserved_ = 0;
    out->kernel_ms = ms;
    c->last_ms = ms;
    c->hmc_has_fields = false;   // pi, chi are consumed: refresh per trajectory (hmc.cpp:154-157)
    return SM_OK;
}

int sm_hmc_accept(sm_ctx* c, int accept) {
    TRY(set_device(c));
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "no gauge field set");
    if (accept) {   // GConf = GConf_copy (hmc.cpp:173) as a pointer swap
        std::swap(c->U, c->Up);
        c->ghost_valid_for = c->f2_U_valid_for = c->tg_U_valid_for = nullptr;
    }
    return SM_OK;
}

int sm_hmc_force(sm_ctx* c, const double* p0, const double* p1, double* F0, double* F1, int* converged) {
    TRY(set_device(c));
    NEED(p0); NEED(p1); NEED(F0); NEED(F1);
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "sm_hmc_set_gauge first");
    TRY(h2d_c(c, c->phi, p0, p1));
    TRY(hmc_adopt_phi(c));
    TrajAcc acc;
    tick(c);
    TRY(hmc_force(c, c->U, c->phi, c->F, &acc));
    TRY(tock(c));
    if (converged) *converged = acc.all_ok;
    TRY(d2h_r(c, c->F, F0, F1));
    return sync(c);
}

int sm_hmc_hamiltonian(sm_ctx* c, const double* pi0, const double* pi1, const double* p0, const double* p1, double* H) {
    TRY(set_device(c));
    NEED(pi0); NEED(pi1); NEED(p0); NEED(p1); NEED(H);
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "sm_hmc_set_gauge first");
    TRY(h2d_r(c, c->pi, pi0, pi1));
    TRY(h2d_c(c, c->phi, p0, p1));
    TRY(hmc_adopt_phi(c));
    tick(c);
    TRY(hmc_hamiltonian_async(c, c->U, c->pi, c->phi, 0, nullptr));
    TRY(tock(c));
    TRY(fetch_sums(c, 5));
    *H = hamiltonian_from(c->h->sums);
    return SM_OK;
}

int sm_hmc_leapfrog(sm_ctx* c, const double* pi0, const double* pi1, const double* p0, const double* p1,
                    int* all_converged) {
    TRY(set_device(c));
    NEED(pi0); NEED(pi1); NEED(p0); NEED(p1);
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "sm_hmc_set_gauge first");
    TRY(h2d_r(c, c->pi, pi0, pi1));
    TRY(h2d_c(c, c->phi, p0, p1));
    TRY(hmc_adopt_phi(c));
    TrajAcc acc;
    tick(c);
    TRY(hmc_leapfrog(c, &acc));
    TRY(tock(c));
    if (all_converged) *all_converged = acc.all_ok;
    return SM_OK;
}

// ---- configuration files ---------------------------------------------------------------------------
int sm_save_conf(int Nx, int Nt, const double* U0, const double* U1, const char* path) {
    NEED(U0); NEED(U1); NEED(path);
    FILE* f = fopen(path, "wb");
    if (!f) return fail(SM_ERR_IO, std::string("cannot open ") + path);
    // one 28-byte record per link, x -> t -> mu (src/gauge_conf.cpp:404-419); buffered by row
    std::vector<unsigned char> row((size_t)Nt * 2 * 28);
    for (int x = 0; x < Nx; x++) {
        unsigned char* w = row.data();
        for (int t = 0; t < Nt; t++) {
            const size_t n = (size_t)x * Nt + t;
            for (int mu = 0; mu < 2; mu++) {
                const int32_t hdr[3] = {x, t, mu};
                const double* src = (mu == 0 ? U0 : U1) + 2 * n;
                memcpy(w, hdr, 12);
                memcpy(w + 12, src, 16);
                w += 28;
            }
        }
        if (fwrite(row.data(), 1, row.size(), f) != row.size()) {
            fclose(f);
            return fail(SM_ERR_IO, std::string("short write to ") + path);
        }
    }
    if (fclose(f) != 0) return fail(SM_ERR_IO, std::string("cannot finish writing ") + path);   // buffered data that did not fit
    return SM_OK;
}

int sm_read_conf(int Nx, int Nt, const char* path, double* U0, double* U1) {
    NEED(U0); NEED(U1); NEED(path);
    FILE* f = fopen(path, "rb");
    if (!f) return fail(SM_ERR_IO, std::string("cannot open ") + path);
    std::vector<unsigned char> row((size_t)Nt * 2 * 28);
    for (int x = 0; x < Nx; x++) {
        if (fread(row.data(), 1, row.size(), f) != row.size()) {
            fclose(f);
            return fail(SM_ERR_IO, std::string("short read from ") + path);
        }
        const unsigned char* r = row.data();
        for (int t = 0; t < Nt; t++) {
            const size_t n = (size_t)x * Nt + t;
            for (int mu = 0; mu < 2; mu++) {
                // like the reference reader (gauge_conf.cpp:515-532) the position in the file, not the
                // stored (x,t,mu), decides where a link goes
                memcpy((mu == 0 ? U0 : U1) + 2 * n, r + 12, 16);
                r += 28;
            }
        }
    }
    fclose(f);
    return SM_OK;
}

}  // extern "C"
