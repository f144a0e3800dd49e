// sm_abi.cu -- context, launch logic and the extern "C" surface of libschwinger_b200.so.
// See include/schwinger_b200.h for the contract; DESIGN.md for the layout and kernel list.
#include "../../include/schwinger_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "sm_kernels.cuh"
#include "sm_fused.cuh"
#include "sm_cluster_cg.cuh"

using namespace sm;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(SM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));               \
    } while (0)

#define TRY(call)                   \
    do {                            \
        int rc_ = (call);           \
        if (rc_ != SM_OK) return rc_; \
    } while (0)

#define NEED(p)                                                                  \
    do {                                                                         \
        if ((p) == nullptr) return fail(SM_ERR_ARG, std::string("null argument: ") + #p); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, bound lazily so that single-GPU use never loads it (and a process that already holds
// torch's libnccl.so.2 shares that copy).
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return SM_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(SM_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + dlerror());
#define BIND(name)                                                             \
    g_nccl.name = (decltype(g_nccl.name))dlsym(h, "nccl" #name);               \
    if (!g_nccl.name) return fail(SM_ERR_NCCL, "libnccl lacks nccl" #name);
    BIND(GetUniqueId) BIND(CommInitRank) BIND(CommDestroy) BIND(Send) BIND(Recv) BIND(AllReduce) BIND(GroupStart)
    BIND(GroupEnd) BIND(GetErrorString)
#undef BIND
    g_nccl.handle = h;
    return SM_OK;
}

#define NC(call)                                                                                            \
    do {                                                                                                    \
        ncclResult_t r_ = (call);                                                                           \
        if (r_ != ncclSuccess) return fail(SM_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct HostScalars {   // pinned mirror of what comes back per call
    CgState cg[2];
    double sums[16];
};

struct sm_ctx {
    int Nx = 0, Nt = 0, rx = 1, rt = 1, rank = 0, nranks = 1, cx = 0, ct = 0;
    int wx = 0, wt = 0, V = 0;
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t comm_stream = nullptr;   // halo exchanges that overlap the interior blocks
    cudaEvent_t ev_ready = nullptr, ev_ghost = nullptr;
    bool overlap = true;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_poll[2] = {nullptr, nullptr};
    double last_ms = 0.0;
    long long launches = 0;

    // launch geometry
    dim3 wil_block, wil_grid;
    int rows_per_block = 0;
    dim3 fus_block, fus_grid;   // one-pass D D^dagger (sm_fused.cuh)
    int fus_rows = 0, fus_cols = 0;
    int fus_rb = 8, fus_split_rows = 0, fus_split_chunks = 0;   // interior/boundary launch split (split lattice)
    bool use_fused = true;      // SM_DD_PATH=twopass selects the two-pass form
    int flat_blocks_c = 0;   // grid for flat passes over 2V elements
    int flat_blocks_s = 0;   // grid for passes over V sites

    // reductions and scalars
    double* partials = nullptr;
    unsigned int* tickets = nullptr;   // one per reducing kernel type
    CgState* cg = nullptr;
    double* sums = nullptr;            // 16 doubles
    double* sums_loc = nullptr;        // staging of local sums before an all-reduce (split lattice)
    HostScalars* h = nullptr;

    double tol = 1e-10;
    int max_iter = 10000;

    // work fields (2V complex each)
    cplx *tmp = nullptr, *cg_r = nullptr, *cg_d = nullptr, *cg_Ad = nullptr, *cg_d2 = nullptr;
    // staging for the host-buffer API
    cplx *sU = nullptr, *sA = nullptr, *sB = nullptr, *sC = nullptr;
    double* sF = nullptr;
    // HMC state
    bool hmc_ready = false, hmc_has_gauge = false, hmc_has_fields = false;
    sm_hmc_params hp{};
    cplx *U = nullptr, *Up = nullptr, *chi = nullptr, *phi = nullptr, *psi = nullptr, *xi = nullptr;
    double *pi = nullptr, *pip = nullptr, *F = nullptr;

    // split lattice
    ncclComm_t comm = nullptr;
    int nb_xm = 0, nb_xp = 0, nb_tm = 0, nb_tp = 0;   // neighbour ranks
    cplx *send_tm = nullptr, *send_tp = nullptr, *send_xm = nullptr, *send_xp = nullptr;
    cplx *g_tp = nullptr, *g_tm = nullptr, *g_xp = nullptr, *g_xm = nullptr;
    // gauge ghost ring and force ghosts
    cplx *gg_xm = nullptr, *gg_xp = nullptr, *gg_tm = nullptr, *gg_tp = nullptr, *gg_send = nullptr;
    cplx *fg_t = nullptr, *fg_x = nullptr, *fg_send = nullptr;
    const cplx* ghost_valid_for = nullptr;   // gauge field whose ghost ring is current
    // 2-row ghosts for the one-pass D D^dagger on a lattice split along x ([comp][2 rows][wt] each)
    cplx *f2_U[2] = {nullptr, nullptr}, *f2_in[2] = {nullptr, nullptr}, *f2_r[2] = {nullptr, nullptr};
    cplx *f2_d[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [ping-pong][lo/hi]
    const cplx* f2_U_valid_for = nullptr;
    // peer-memory halo push (sm_p2p_connect): one window per rank, [kind: psi, r][parity][side: lo, hi][4 wt]
    // complex + 4 epoch flags; neighbours store into it over NVLink
    cplx* win = nullptr;
    unsigned int* win_flags = nullptr;
    size_t win_bytes = 0;
    void* peer_win[2] = {nullptr, nullptr};   // -x, +x neighbour's window (peer pointers)
    bool p2p = false;
    unsigned int p2p_epoch[2] = {0, 0};
    unsigned int* push_ticket = nullptr;
    CUresult (*wait_value32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;

    std::vector<void*> user_fields;

    // CUDA graphs of one batch of one-pass CG iterations, keyed on what the kernels bake in
    struct CgGraph {
        const void* U;
        const void* x;
        double m0;
        int max_iter;
        cudaGraphExec_t exec;
        int kernels;
    };
    std::vector<CgGraph> cg_graphs;
    bool use_graphs = true;
    unsigned int attr_done = 0;   // kernel attributes already set on this context's device
    int solver = SM_SOLVER_REFERENCE;
    cplxf *mx_U = nullptr, *mx_r = nullptr, *mx_e = nullptr, *mx_d0 = nullptr, *mx_d1 = nullptr, *mx_Ad = nullptr;
    bool use_cluster = true;   // whole-solve resident kernels for small lattices (SM_CLUSTER_CG=0 disables)
    int coop_sites = -1;
    cplx* coop_hop = nullptr;
    double* coop_wsum = nullptr;
    unsigned int* coop_bar = nullptr;

    bool dist() const { return nranks > 1; }
    double sR_edge() const { return (ct == rt - 1) ? -1.0 : 1.0; }
    double sL_edge() const { return (ct == 0) ? -1.0 : 1.0; }
};

enum { TK_WILSON = 0, TK_UPDATE, TK_DOT, TK_PLAQ, TK_KIN, TK_COUNT };

template <typename T>
static int dev_alloc(T** p, size_t n) {
    CU(cudaMalloc((void**)p, n * sizeof(T)));
    return SM_OK;
}

static int ensure_complex(sm_ctx* c, cplx** p) {
    if (*p) return SM_OK;
    TRY(dev_alloc(p, (size_t)2 * c->V));
    CU(cudaMemsetAsync(*p, 0, sizeof(cplx) * 2 * c->V, c->stream));
    return SM_OK;
}
static int ensure_real(sm_ctx* c, double** p) {
    if (*p) return SM_OK;
    TRY(dev_alloc(p, (size_t)2 * c->V));
    CU(cudaMemsetAsync(*p, 0, sizeof(double) * 2 * c->V, c->stream));
    return SM_OK;
}

static int ctx_common_init(sm_ctx* c) {
    CU(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major < 10)
        return fail(SM_ERR_CUDA, "libschwinger_b200 is built for sm_100a only; device is sm_" +
                                     std::to_string(prop.major) + std::to_string(prop.minor));
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;   // the comm stream outranks the compute stream: its few blocks go first when slots free up
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
    }
    CU(cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_ghost, cudaEventDisableTiming));
    if (const char* e = getenv("SM_OVERLAP")) c->overlap = atoi(e) != 0;
    if (const char* e = getenv("SM_GRAPHS")) c->use_graphs = atoi(e) != 0;
    if (const char* e = getenv("SM_CLUSTER_CG")) c->use_cluster = atoi(e) != 0;
    CU(cudaEventCreate(&c->ev_a));
    CU(cudaEventCreate(&c->ev_b));
    CU(cudaEventCreateWithFlags(&c->ev_poll[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_poll[1], cudaEventDisableTiming));

    // stencil tiles: TT sites along t (coalesced 16-byte accesses), TX rows per step
    const int TT = c->wt >= 128 ? 128 : (c->wt >= 64 ? 64 : 32);
    const int TX = kBlock / TT;
    c->wil_block = dim3(TT, TX, 1);
    const int nT = (c->wt + TT - 1) / TT;
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_wilson<false, WILSON_DOT>, kBlock, 0));
    if (occ < 1) occ = 1;
    const int target = c->sm_count * occ;
    const int steps = (c->wx + TX - 1) / TX;
    int GY = std::max(1, std::min(steps, target / nT));
    int rows = ((c->wx + GY - 1) / GY + TX - 1) / TX * TX;
    GY = (c->wx + rows - 1) / rows;
    c->rows_per_block = rows;
    c->wil_grid = dim3(nT, GY, 1);

    // one-pass D D^dagger: strips of <= BT-4 columns, chunks of rows.  Large lattices: ~8 waves of
    // blocks with >= 64 rows each (4 warm-up rows per chunk); mid-size: one resident wave.
    {
        const long long V = (long long)c->wx * c->wt;
        int BT = (c->wt + 4 <= 128 || V <= (1LL << 21)) ? 128 : 256;
        if (const char* e = getenv("SM_FUSED_BT")) BT = atoi(e) == 128 ? 128 : 256;
        const int strips = (c->wt + (BT - 4) - 1) / (BT - 4);
        c->fus_cols = (c->wt + strips - 1) / strips;     // equal strips
        const int capacity = c->sm_count * (BT == 128 ? 4 : 2);
        // rows per chunk: minimise  waves x (rows + 4 warm-up rows)  with waves = ceil(blocks / resident blocks);
        // this model reproduces the measured sweep (profiles/r01_sweep_rows.txt) to a few per cent
        auto rows_for = [&](int nrows) {
            if (const char* r = getenv("SM_FUSED_ROWS")) return std::max(1, std::min(nrows, atoi(r)));
            int best = std::min(nrows, 8);
            long long best_cost = -1;
            for (int r = std::min(nrows, 8); r <= std::min(nrows, 512); r++) {
                const long long blocks = (long long)strips * ((nrows + r - 1) / r);
                const long long cost = ((blocks + capacity - 1) / capacity) * (r + 4);
                if (best_cost < 0 || cost <= best_cost) {
                    best_cost = cost;
                    best = r;
                }
            }
            return best;
        };
        const int rows = rows_for(c->wx);
        c->fus_block = dim3(BT, 1, 1);
        c->fus_grid = dim3(strips, (c->wx + rows - 1) / rows, 1);
        c->fus_rows = rows;
        // split lattice: two thin boundary bands (the only rows that read ghost rows) + interior chunks
        c->fus_rb = 8;
        if (const char* r = getenv("SM_FUSED_RB")) c->fus_rb = std::max(2, atoi(r));
        c->fus_split_rows = c->fus_split_chunks = 0;
        if (c->wx >= 4 * c->fus_rb) {
            const int inner = c->wx - 2 * c->fus_rb;
            c->fus_split_rows = rows_for(inner);
            if (const char* r = getenv("SM_FUSED_SPLIT_ROWS")) c->fus_split_rows = std::max(1, std::min(inner, atoi(r)));
            c->fus_split_chunks = (inner + c->fus_split_rows - 1) / c->fus_split_rows;
        }
        long long min_sites = 0;                         // measured: never slower than two passes (profiles/r01_sweep_sizes_*)
        if (const char* m = getenv("SM_FUSED_MIN_SITES")) min_sites = atoll(m);
        const char* e = getenv("SM_DD_PATH");
        c->use_fused = !(e && std::string(e) == "twopass") && V >= min_sites;
        if (e && std::string(e) == "onepass") c->use_fused = true;
    }

    const int cap = c->sm_count * 8;
    c->flat_blocks_c = std::max(1, std::min(cap, (2 * c->V + kBlock - 1) / kBlock));
    c->flat_blocks_s = std::max(1, std::min(cap, (c->V + kBlock - 1) / kBlock));

    const size_t max_blocks = std::max<size_t>(std::max<size_t>((size_t)nT * GY, (size_t)cap),
                                               (size_t)c->fus_grid.x * (std::max<size_t>(c->fus_grid.y, c->fus_split_chunks) + 2));
    TRY(dev_alloc(&c->partials, max_blocks * kMaxSums));
    TRY(dev_alloc(&c->tickets, (size_t)TK_COUNT));
    CU(cudaMemsetAsync(c->tickets, 0, sizeof(unsigned int) * TK_COUNT, c->stream));
    TRY(dev_alloc(&c->cg, 1));
    CU(cudaMemsetAsync(c->cg, 0, sizeof(CgState), c->stream));
    TRY(dev_alloc(&c->sums, 16));
    TRY(dev_alloc(&c->sums_loc, 16));
    CU(cudaMemsetAsync(c->sums, 0, sizeof(double) * 16, c->stream));
    CU(cudaMallocHost((void**)&c->h, sizeof(HostScalars)));
    memset(c->h, 0, sizeof(HostScalars));
    CU(cudaStreamSynchronize(c->stream));
    return SM_OK;
}

static void tick(sm_ctx* c) { cudaEventRecord(c->ev_a, c->stream); }
static int tock(sm_ctx* c) {
    CU(cudaEventRecord(c->ev_b, c->stream));
    CU(cudaEventSynchronize(c->ev_b));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, c->ev_a, c->ev_b));
    c->last_ms = ms;
    return SM_OK;
}

#define KCHECK()                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess) return fail(SM_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)

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

template <bool DAG>
static int exchange_spinor_halo(sm_ctx* c, const cplx* U, const cplx* in, const int* done) {
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
    NC(g_nccl.GroupStart());
    if (c->rt > 1) {
        NC(g_nccl.Send(c->send_tm, 2 * (size_t)c->wx, ncclDouble, c->nb_tm, c->comm, c->stream));
        NC(g_nccl.Send(c->send_tp, 2 * (size_t)c->wx, ncclDouble, c->nb_tp, c->comm, c->stream));
        NC(g_nccl.Recv(c->g_tp, 2 * (size_t)c->wx, ncclDouble, c->nb_tp, c->comm, c->stream));
        NC(g_nccl.Recv(c->g_tm, 2 * (size_t)c->wx, ncclDouble, c->nb_tm, c->comm, c->stream));
    }
    if (c->rx > 1) {
        NC(g_nccl.Send(c->send_xm, 2 * (size_t)c->wt, ncclDouble, c->nb_xm, c->comm, c->stream));
        NC(g_nccl.Send(c->send_xp, 2 * (size_t)c->wt, ncclDouble, c->nb_xp, c->comm, c->stream));
        NC(g_nccl.Recv(c->g_xp, 2 * (size_t)c->wt, ncclDouble, c->nb_xp, c->comm, c->stream));
        NC(g_nccl.Recv(c->g_xm, 2 * (size_t)c->wt, ncclDouble, c->nb_xm, c->comm, c->stream));
    }
    NC(g_nccl.GroupEnd());
    return SM_OK;
}

// ------------------------------------------------------------------------------------------------
// operator launches on device fields
// ------------------------------------------------------------------------------------------------
template <bool DAG, int MODE>
static int launch_wilson(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, const cplx* aux = nullptr,
                         cplx* r = nullptr, cplx* d = nullptr, cplx* x = nullptr, double* sums_out = nullptr,
                         const int* done = nullptr) {
    if (c->dist()) TRY((exchange_spinor_halo<DAG>(c, U, in, done)));
    WilsonArgs a{};
    a.U = U;
    a.in = in;
    a.out = out;
    a.aux = aux;
    a.r = r;
    a.d = d;
    a.x = x;
    a.wx = c->wx;
    a.wt = c->wt;
    a.V = c->V;
    a.rows_per_block = c->rows_per_block;
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
    k_wilson<DAG, MODE><<<c->wil_grid, c->wil_block, 0, c->stream>>>(a);
    KCHECK();
    c->launches++;
    return SM_OK;
}

static int dev_D(sm_ctx* c, const cplx* U, const cplx* in, cplx* out, double m0, bool dagger) {
    if (in == out) return fail(SM_ERR_ARG, "D: in and out must not alias");
    if (dagger) return launch_wilson<true, WILSON_PLAIN>(c, U, in, out, m0);
    return launch_wilson<false, WILSON_PLAIN>(c, U, in, out, m0);
}

// D D^dagger via the context's scratch field (the reference's global DTEMP, dirac_operator.cpp:477-480)
// ---- peer-memory window ------------------------------------------------------------------------
static size_t win_ghost_elems(const sm_ctx* c) { return 4 * (size_t)c->wt; }
static cplx* win_ghost(const sm_ctx* c, void* base, int kind, int parity, int side) {
    return (cplx*)base + (size_t)((kind * 2 + parity) * 2 + side) * win_ghost_elems(c);
}
static unsigned int* win_flag(const sm_ctx* c, void* base, int kind, int side) {
    return (unsigned int*)((char*)base + sizeof(cplx) * 8 * win_ghost_elems(c)) + kind * 2 + side;
}

// push my boundary rows of `field` into both neighbours' ghosts (epoch parity) and raise their flags
static int p2p_push(sm_ctx* c, const cplx* field, int kind, cudaStream_t st) {
    const unsigned int epoch = ++c->p2p_epoch[kind];
    const int parity = epoch & 1;
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

// one-pass D D^dagger (sm_fused.cuh): a single tile, or tiles split along x only (ranks_t == 1,
// 2-row ghosts); a split along t keeps the two-pass kernels.  C = cplx (double) everywhere except in the
// inner solve of the opt-in mixed-precision CG (C = cplxf, single tile only).
template <typename C, int MODE>
static int launch_fused(sm_ctx* c, const C* U, const C* in, C* out, double m0, double* sums_out = nullptr,
                        const C* r = nullptr, C* x = nullptr, C* d_new = nullptr, int k = 0) {
    constexpr bool kDouble = std::is_same<C, cplx>::value;
    if (!kDouble && c->dist()) return fail(SM_ERR_STATE, "single-precision passes run on a single tile only");
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
    const size_t smem = fused_smem_bytes(MODE, STAGES, c->fus_block.x, sizeof(C));
    const unsigned int attr_bit = 1u << (MODE + (kDouble ? 0 : 4));
    if (!(c->attr_done & attr_bit)) {   // function attributes are per device: once per context and instantiation
        CU(cudaFuncSetAttribute(k_dd_fused<C, MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
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
            if (c->p2p) TRY(p2p_push(c, moving, kind, c->stream));   // stores into the neighbours' windows
            if (split_launch) {
                CU(cudaEventRecord(c->ev_ready, c->stream));
                CU(cudaStreamWaitEvent(c->comm_stream, c->ev_ready, 0));
            }
            if (c->p2p) {
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
        k_dd_fused<C, MODE, STAGES><<<dim3(c->fus_grid.x, 2, 1), c->fus_block, smem, c->comm_stream>>>(a);
        KCHECK();
        CU(cudaEventRecord(c->ev_ghost, c->comm_stream));
        a.chunk_mode = 1;
        k_dd_fused<C, MODE, STAGES><<<dim3(c->fus_grid.x, c->fus_split_chunks, 1), c->fus_block, smem, c->stream>>>(a);
        KCHECK();
        CU(cudaStreamWaitEvent(c->stream, c->ev_ghost, 0));
        c->launches += 2;
    } else {
        k_dd_fused<C, MODE, STAGES><<<c->fus_grid, c->fus_block, smem, c->stream>>>(a);
        KCHECK();
        c->launches++;
    }
    return SM_OK;
}

static bool fused_ok(const sm_ctx* c) { return c->use_fused && (!c->dist() || (c->rt == 1 && c->wx >= 4)); }

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

    k_cg_reset<<<1, 1, 0, c->stream>>>(st, tol);
    c->launches++;
    // Ad = DD^dagger phi ; r = phi - Ad ; d = r ; x = phi ; |phi|^2, |r|^2
    TRY((launch_wilson<true, WILSON_PLAIN>(c, U, phi, c->tmp, m0)));
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

    // one iteration: A(k) then B(k) (+ the all-reduces of their sums on a split lattice)
    auto iteration = [&](int k) -> int {
        const int cur = k & 1;
        TRY((launch_fused<C, FUSED_CG>(c, U, dbuf[cur ^ 1], Ad, m0, sum_target(c, st->dAd), r, x, dbuf[cur], k)));
        TRY(sum_finish(c, st->dAd, 2));
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
    const bool graphs = c->use_graphs && !c->dist() && max_iter > batch;
    if (graphs) {
        for (auto& g : c->cg_graphs)
            if (g.U == (const void*)U && g.x == (const void*)x && g.m0 == m0 && g.max_iter == max_iter) {
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
            k_cg_check_dev<<<1, 1, 0, c->stream>>>(st, max_iter);
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
        c->cg_graphs.push_back({(const void*)U, (const void*)x, m0, max_iter, exec, graph_kernels});
    }
    for (;;) {
        if (exec != nullptr && k + batch <= max_iter) {
            CU(cudaGraphLaunch(exec, c->stream));
            c->launches += graph_kernels;
            k += batch;
        } else {
            const int k_end = std::min(max_iter, k + batch);
            for (; k < k_end; k++) TRY(iteration(k));
            k_cg_check_dev<<<1, 1, 0, c->stream>>>(st, max_iter);
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
    k_cg_reset<<<1, 1, 0, c->stream>>>(st, c->tol);
    c->launches++;
    // x = phi ; r = phi - D D^dagger phi ; |phi|^2, |r|^2   (d_0 = r_0 is formed by the first pass)
    TRY((launch_wilson<true, WILSON_PLAIN>(c, U, phi, c->tmp, m0)));
    TRY((launch_wilson<false, WILSON_CGINIT>(c, U, c->tmp, nullptr, m0, phi, c->cg_r, c->cg_d2, x,
                                             sum_target(c, &st->phi_norm2))));
    TRY(sum_finish(c, &st->phi_norm2, 2));
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
        k_mixed_begin<<<1, 1, 0, c->stream>>>(c->cg, c->sums + 12, delta);
        KCHECK();
        c->launches++;
        TRY(cg_fused_loop<cplxf>(c, c->mx_U, c->mx_r, c->mx_e, c->mx_d0, c->mx_d1, c->mx_Ad, m0,
                                 std::max(1, c->max_iter - total)));
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

static int dev_cg_cluster(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    ResidentCgArgs a;
    TRY(resident_args(c, U, phi, x, m0, &a));
    int ctas = 1;
    while (ctas * kClusterThreads < c->V) ctas *= 2;
    if (!(c->attr_done & (1u << 8))) {
        CU(cudaFuncSetAttribute(k_cg_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        c->attr_done |= 1u << 8;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas, 1, 1);
    cfg.blockDim = dim3(kClusterThreads, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
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
        c->coop_sites = per_sm * c->sm_count * kCoopThreads;
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
        TRY(dev_alloc(&c->coop_bar, (size_t)32));
    }
    CU(cudaMemsetAsync(c->coop_bar, 0, sizeof(unsigned int) * 32, c->stream));
    a.hop = c->coop_hop;
    a.wsum = c->coop_wsum;
    a.bar = c->coop_bar;
    void* params[] = {&a};
    CU(cudaLaunchCooperativeKernel((const void*)k_cg_coop, dim3(blocks, 1, 1), dim3(kCoopThreads, 1, 1), params, 0,
                                   c->stream));
    return resident_finish(c, converged, iterations);
}

static int dev_cg(sm_ctx* c, const cplx* U, const cplx* phi, cplx* x, double m0, int* converged, int* iterations) {
    if (c->use_cluster && !c->dist()) {
        if (c->V <= kClusterMaxCtas * kClusterThreads) return dev_cg_cluster(c, U, phi, x, m0, converged, iterations);
        if (c->V <= coop_capacity(c)) return dev_cg_coop(c, U, phi, x, m0, converged, iterations);
    }
    if (c->solver == SM_SOLVER_MIXED && fused_ok(c) && !c->dist()) return dev_cg_mixed(c, U, phi, x, m0, converged, iterations);
    if (fused_ok(c)) return dev_cg_fused(c, U, phi, x, m0, converged, iterations);
    return dev_cg_twopass(c, U, phi, x, m0, converged, iterations);
}

// ------------------------------------------------------------------------------------------------
// gauge ghost ring / force ghosts for a split lattice
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_gauge_cols(const cplx* U, int wx, int wt, int V, cplx* send) {
    // send[0..2wx): column t=0 (mu0, mu1) ; send[2wx..4wx): column t=wt-1
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wx) return;
    for (int mu = 0; mu < 2; mu++) {
        send[mu * wx + i] = U[mu * V + i * wt];
        send[2 * wx + mu * wx + i] = U[mu * V + i * wt + wt - 1];
    }
}

__global__ void k_pack_gauge_rows(const cplx* U, int wx, int wt, int V, const cplx* gt_m, const cplx* gt_p,
                                  cplx* send) {
    // rows x=0 and x=wx-1 with their t=-1 / t=wt ghosts: (wt+2) per mu
    // send[0 .. 2(wt+2)): row 0 ; send[2(wt+2) .. 4(wt+2)): row wx-1
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // 0..wt+1  <->  t = j-1
    const int W = wt + 2;
    if (j >= W) return;
    for (int mu = 0; mu < 2; mu++) {
        for (int which = 0; which < 2; which++) {
            const int x = which ? wx - 1 : 0;
            cplx v;
            if (j == 0)
                v = gt_m ? gt_m[mu * wx + x] : U[mu * V + x * wt + wt - 1];
            else if (j == W - 1)
                v = gt_p ? gt_p[mu * wx + x] : U[mu * V + x * wt];
            else
                v = U[mu * V + x * wt + (j - 1)];
            send[which * 2 * W + mu * W + j] = v;
        }
    }
}

static int refresh_gauge_ghosts(sm_ctx* c, const cplx* U) {
    if (!c->dist() || c->ghost_valid_for == U) return SM_OK;
    const int wx = c->wx, wt = c->wt, W = wt + 2;
    if (c->rt > 1) {
        k_pack_gauge_cols<<<(wx + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(U, wx, wt, c->V, c->gg_send);
        KCHECK();
        c->launches++;
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(c->gg_send, 2 * (size_t)2 * wx, ncclDouble, c->nb_tm, c->comm, c->stream));            // my t=0 column
        NC(g_nccl.Send(c->gg_send + 2 * wx, 2 * (size_t)2 * wx, ncclDouble, c->nb_tp, c->comm, c->stream));   // my t=wt-1
        NC(g_nccl.Recv(c->gg_tp, 2 * (size_t)2 * wx, ncclDouble, c->nb_tp, c->comm, c->stream));
        NC(g_nccl.Recv(c->gg_tm, 2 * (size_t)2 * wx, ncclDouble, c->nb_tm, c->comm, c->stream));
        NC(g_nccl.GroupEnd());
    }
    if (c->rx > 1) {
        cplx* send = c->gg_send + 4 * wx;
        k_pack_gauge_rows<<<(W + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(
            U, wx, wt, c->V, c->rt > 1 ? c->gg_tm : nullptr, c->rt > 1 ? c->gg_tp : nullptr, send);
        KCHECK();
        c->launches++;
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(send, 2 * (size_t)2 * W, ncclDouble, c->nb_xm, c->comm, c->stream));             // my row 0
        NC(g_nccl.Send(send + 2 * W, 2 * (size_t)2 * W, ncclDouble, c->nb_xp, c->comm, c->stream));     // my row wx-1
        NC(g_nccl.Recv(c->gg_xp, 2 * (size_t)2 * W, ncclDouble, c->nb_xp, c->comm, c->stream));
        NC(g_nccl.Recv(c->gg_xm, 2 * (size_t)2 * W, ncclDouble, c->nb_xm, c->comm, c->stream));
        NC(g_nccl.GroupEnd());
    }
    c->ghost_valid_for = U;
    return SM_OK;
}

static GaugeView gauge_view(sm_ctx* c, const cplx* U) {
    GaugeView g{};
    g.U = U;
    g.wx = c->wx;
    g.wt = c->wt;
    g.V = c->V;
    g.gx_m = c->rx > 1 ? c->gg_xm : nullptr;
    g.gx_p = c->rx > 1 ? c->gg_xp : nullptr;
    g.gt_m = c->rt > 1 ? c->gg_tm : nullptr;
    g.gt_p = c->rt > 1 ? c->gg_tp : nullptr;
    return g;
}

// force ghosts: projected forward neighbours of psi and chi (src/dirac_operator.cpp:511-530)
__global__ void k_pack_force(const cplx* psi, const cplx* chi, int wx, int wt, int V, cplx* to_tm, cplx* to_xm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (to_tm != nullptr && i < wx) {
        const int n = i * wt;   // my column t = 0
        to_tm[i] = csub(chi[n], chi[V + n]);
        to_tm[wx + i] = cadd(psi[n], psi[V + n]);
    }
    if (to_xm != nullptr && i < wt) {
        const int n = i;        // my row x = 0
        const cplx c0 = chi[n], c1 = chi[V + n], p0 = psi[n], p1 = psi[V + n];
        to_xm[i] = make_double2(c0.x - c1.y, c0.y + c1.x);
        to_xm[wt + i] = make_double2(p0.x + p1.y, p0.y - p1.x);
    }
}

static int dev_force(sm_ctx* c, const cplx* U, const cplx* psi, const cplx* chi, double* F, double beta, bool fermion,
                     bool gauge) {
    if (gauge) TRY(refresh_gauge_ghosts(c, U));
    ForceArgs a{};
    a.g = gauge_view(c, U);
    if (!gauge) {   // the fermion part needs U(n) only
        a.g.gx_m = a.g.gx_p = a.g.gt_m = a.g.gt_p = nullptr;
    }
    a.psi = psi;
    a.chi = chi;
    a.F = F;
    a.beta = beta;
    a.sR_edge = c->sR_edge();
    a.fermion = fermion;
    a.gauge = gauge;
    if (c->dist() && fermion) {
        cplx* to_tm = c->rt > 1 ? c->fg_send : nullptr;
        cplx* to_xm = c->rx > 1 ? c->fg_send + 2 * c->wx : nullptr;
        const int n = std::max(c->wx, c->wt);
        k_pack_force<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(psi, chi, c->wx, c->wt, c->V, to_tm, to_xm);
        KCHECK();
        c->launches++;
        NC(g_nccl.GroupStart());
        if (c->rt > 1) {
            NC(g_nccl.Send(to_tm, 2 * (size_t)2 * c->wx, ncclDouble, c->nb_tm, c->comm, c->stream));
            NC(g_nccl.Recv(c->fg_t, 2 * (size_t)2 * c->wx, ncclDouble, c->nb_tp, c->comm, c->stream));
        }
        if (c->rx > 1) {
            NC(g_nccl.Send(to_xm, 2 * (size_t)2 * c->wt, ncclDouble, c->nb_xm, c->comm, c->stream));
            NC(g_nccl.Recv(c->fg_x, 2 * (size_t)2 * c->wt, ncclDouble, c->nb_xp, c->comm, c->stream));
        }
        NC(g_nccl.GroupEnd());
        a.fg_t = c->rt > 1 ? c->fg_t : nullptr;
        a.fg_x = c->rx > 1 ? c->fg_x : nullptr;
    }
    k_force<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(a);
    KCHECK();
    c->launches++;
    return SM_OK;
}

static int dev_plaquette(sm_ctx* c, const cplx* U, double beta, cplx* P, double* d_out2) {
    TRY(refresh_gauge_ghosts(c, U));
    k_plaquette<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(gauge_view(c, U), beta, P, c->partials,
                                                            c->tickets + TK_PLAQ, sum_target(c, d_out2));
    KCHECK();
    c->launches++;
    return sum_finish(c, d_out2, 2);
}

static int dev_kinetic(sm_ctx* c, const double* pi, double* d_out1) {
    k_kinetic<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(pi, 2 * c->V, c->partials, c->tickets + TK_KIN,
                                                          sum_target(c, d_out1));
    KCHECK();
    c->launches++;
    return sum_finish(c, d_out1, 1);
}

static int dev_leap_update(sm_ctx* c, cplx* U, double* pi, const double* F, double eps_pi, double eps_u) {
    k_leap_update<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(U, pi, F, eps_pi, eps_u, 2 * c->V);
    KCHECK();
    c->launches++;
    invalidate_gauge_ghosts(c, U);
    return SM_OK;
}

// ------------------------------------------------------------------------------------------------
// host <-> device field copies (component arrays of the reference's spinor / re_field)
// ------------------------------------------------------------------------------------------------
static int h2d_c(sm_ctx* c, cplx* d, const double* h0, const double* h1) {
    invalidate_gauge_ghosts(c, d);
    CU(cudaMemcpyAsync(d, h0, sizeof(cplx) * c->V, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + c->V, h1, sizeof(cplx) * c->V, cudaMemcpyHostToDevice, c->stream));
    return SM_OK;
}
static int d2h_c(sm_ctx* c, const cplx* d, double* h0, double* h1) {
    CU(cudaMemcpyAsync(h0, d, sizeof(cplx) * c->V, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h1, d + c->V, sizeof(cplx) * c->V, cudaMemcpyDeviceToHost, c->stream));
    return SM_OK;
}
static int h2d_r(sm_ctx* c, double* d, const double* h0, const double* h1) {
    CU(cudaMemcpyAsync(d, h0, sizeof(double) * c->V, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + c->V, h1, sizeof(double) * c->V, cudaMemcpyHostToDevice, c->stream));
    return SM_OK;
}
static int d2h_r(sm_ctx* c, const double* d, double* h0, double* h1) {
    CU(cudaMemcpyAsync(h0, d, sizeof(double) * c->V, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h1, d + c->V, sizeof(double) * c->V, cudaMemcpyDeviceToHost, c->stream));
    return SM_OK;
}
static int sync(sm_ctx* c) {
    CU(cudaStreamSynchronize(c->stream));
    return SM_OK;
}
static int fetch_sums(sm_ctx* c, int n) {
    CU(cudaMemcpyAsync(c->h->sums, c->sums, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    return sync(c);
}

static int ensure_staging(sm_ctx* c) {
    TRY(ensure_complex(c, &c->sU));
    TRY(ensure_complex(c, &c->sA));
    TRY(ensure_complex(c, &c->sB));
    TRY(ensure_complex(c, &c->sC));
    TRY(ensure_real(c, &c->sF));
    return SM_OK;
}

static int set_device(sm_ctx* c) {
    NEED(c);
    CU(cudaSetDevice(c->device));
    return SM_OK;
}

// ------------------------------------------------------------------------------------------------
// HMC pieces on device fields (src/hmc.cpp)
// ------------------------------------------------------------------------------------------------
struct TrajAcc {
    long long dd_apps = 0;
    int solves = 0;
    int all_ok = 1;
};

static int hmc_alloc(sm_ctx* c) {
    if (c->hmc_ready) return SM_OK;
    TRY(ensure_complex(c, &c->U));
    TRY(ensure_complex(c, &c->Up));
    TRY(ensure_complex(c, &c->chi));
    TRY(ensure_complex(c, &c->phi));
    TRY(ensure_complex(c, &c->psi));
    TRY(ensure_complex(c, &c->xi));
    TRY(ensure_real(c, &c->pi));
    TRY(ensure_real(c, &c->pip));
    TRY(ensure_real(c, &c->F));
    c->hmc_ready = true;
    return SM_OK;
}

// HMC::Force (hmc.cpp:44-60): psi = (DD^dagger)^-1 phi ; chi' = D^dagger psi ; fermion + gauge force
static int hmc_force(sm_ctx* c, const cplx* U, const cplx* phi, double* F, TrajAcc* acc) {
    int ok = 0, its = 0;
    TRY(dev_cg(c, U, phi, c->psi, c->hp.m0, &ok, &its));
    if (acc) {
        acc->dd_apps += ok ? its + 2 : its + 1;
        acc->solves++;
        acc->all_ok &= ok;
    }
    TRY(dev_D(c, U, c->psi, c->xi, c->hp.m0, true));
    return dev_force(c, U, c->psi, c->xi, F, c->hp.beta, true, true);
}

// HMC::Leapfrog (hmc.cpp:63-103): position first, MD_steps-1 force evaluations
static int hmc_leapfrog(sm_ctx* c, TrajAcc* acc) {
    const int md = c->hp.md_steps;
    const double eps = c->hp.trajectory_length / (md * 1.0);
    CU(cudaMemcpyAsync(c->pip, c->pi, sizeof(double) * 2 * c->V, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->Up, c->U, sizeof(cplx) * 2 * c->V, cudaMemcpyDeviceToDevice, c->stream));
    invalidate_gauge_ghosts(c, c->Up);
    TRY(dev_leap_update(c, c->Up, c->pip, nullptr, 0.0, 0.5 * eps));
    TRY(hmc_force(c, c->Up, c->phi, c->F, acc));
    for (int step = 1; step < md - 1; step++) {
        TRY(dev_leap_update(c, c->Up, c->pip, c->F, eps, eps));
        TRY(hmc_force(c, c->Up, c->phi, c->F, acc));
    }
    return dev_leap_update(c, c->Up, c->pip, c->F, eps, 0.5 * eps);
}

// HMC::Hamiltonian (hmc.cpp:135-149) = sum 1/2 pi^2 + [ beta sum Re(1-P) + Re dot((DD^dagger)^-1 phi, phi) ]
// device sums land in c->sums[base .. base+5): kinetic, sum Re P, gauge action, Re dot, Im dot
static int hmc_hamiltonian_async(sm_ctx* c, const cplx* U, const double* pi, const cplx* phi, int base, TrajAcc* acc) {
    TRY(dev_kinetic(c, pi, c->sums + base));
    TRY(dev_plaquette(c, U, c->hp.beta, nullptr, c->sums + base + 1));
    int ok = 0, its = 0;
    TRY(dev_cg(c, U, phi, c->xi, c->hp.m0, &ok, &its));
    if (acc) {
        acc->dd_apps += ok ? its + 2 : its + 1;
        acc->solves++;
        acc->all_ok &= ok;
    }
    return dev_dot_async(c, c->xi, phi, c->sums + base + 3);
}

static double hamiltonian_from(const double* s) {
    double action = s[2];
    action += s[3];
    double H = s[0];
    H += action;
    return H;
}

// ================================================================================================
// extern "C"
// ================================================================================================
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
        return SM_OK;
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
        cudaIpcCloseMemHandle(c->peer_win[0]);
        if (c->peer_win[1] != c->peer_win[0]) cudaIpcCloseMemHandle(c->peer_win[1]);
    }
    if (c->win) cudaFree(c->win);
    if (c->push_ticket) cudaFree(c->push_ticket);
    if (c->comm) g_nccl.CommDestroy(c->comm);
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
    if (c->coop_wsum) cudaFree(c->coop_wsum);
    if (c->coop_bar) cudaFree(c->coop_bar);
    for (auto& g : c->cg_graphs) cudaGraphExecDestroy(g.exec);
    if (c->h) cudaFreeHost(c->h);
    cudaEventDestroy(c->ev_a);
    cudaEventDestroy(c->ev_b);
    cudaEventDestroy(c->ev_poll[0]);
    cudaEventDestroy(c->ev_poll[1]);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->comm_stream);
    cudaEventDestroy(c->ev_ready);
    cudaEventDestroy(c->ev_ghost);
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
    if (!(tol > 0) || max_iter < 1) return fail(SM_ERR_ARG, "tol must be > 0 and max_iter >= 1");
    c->tol = tol;
    c->max_iter = max_iter;
    return SM_OK;
}

int sm_set_solver(sm_ctx* c, int solver) {
    NEED(c);
    if (solver != SM_SOLVER_REFERENCE && solver != SM_SOLVER_MIXED) return fail(SM_ERR_ARG, "unknown solver");
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
    int *dR, *dL, *dA, *dB;
    double *dsR, *dsL;
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
    TRY(sync(c));
    cudaFree(dR); cudaFree(dL); cudaFree(dA); cudaFree(dB); cudaFree(dsR); cudaFree(dsL);
    return SM_OK;
}

// ---- peer-memory halo push ------------------------------------------------------------------------
int sm_p2p_handle(sm_ctx* c, void* handle_out) {
    TRY(set_device(c));
    NEED(handle_out);
    if (!c->dist() || c->rt != 1 || c->wx < 4) return fail(SM_ERR_STATE, "peer-memory halos need a lattice split along x only");
    if (!c->win) {
        c->win_bytes = sizeof(cplx) * 8 * win_ghost_elems(c) + 256;
        CU(cudaMalloc((void**)&c->win, c->win_bytes));
        CU(cudaMemset(c->win, 0, c->win_bytes));
        c->win_flags = win_flag(c, c->win, 0, 0);
        TRY(dev_alloc(&c->push_ticket, (size_t)1));
        CU(cudaMemset(c->push_ticket, 0, sizeof(unsigned int)));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == SM_P2P_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->win));
    memcpy(handle_out, &h, sizeof(h));
    return SM_OK;
}

int sm_p2p_connect(sm_ctx* c, const void* all_handles) {
    TRY(set_device(c));
    NEED(all_handles);
    if (!c->win) return fail(SM_ERR_STATE, "sm_p2p_handle first");
    if (c->p2p) return SM_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr) != cudaSuccess || fn == nullptr)
        return fail(SM_ERR_CUDA, "cuStreamWaitValue32 is not available");
    c->wait_value32 = (CUresult(*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int))fn;
    const int nb[2] = {c->nb_xm, c->nb_xp};
    for (int s = 0; s < 2; s++) {
        if (s == 1 && nb[1] == nb[0]) {
            c->peer_win[1] = c->peer_win[0];
            break;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + (size_t)nb[s] * SM_P2P_HANDLE_BYTES, sizeof(h));
        CU(cudaIpcOpenMemHandle(&c->peer_win[s], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->p2p = true;
    return SM_OK;
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

int sm_phi_dag_partialD_phi(sm_ctx* c, const double* U0, const double* U1, const double* l0, const double* l1,
                            const double* r0, const double* r1, double* F0, double* F1) {
    TRY(set_device(c));
    NEED(U0); NEED(U1); NEED(l0); NEED(l1); NEED(r0); NEED(r1); NEED(F0); NEED(F1);
    TRY(ensure_staging(c));
    TRY(h2d_c(c, c->sU, U0, U1));
    TRY(h2d_c(c, c->sA, l0, l1));
    TRY(h2d_c(c, c->sB, r0, r1));
    c->ghost_valid_for = c->f2_U_valid_for = nullptr;
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
    c->ghost_valid_for = c->f2_U_valid_for = nullptr;
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
    c->ghost_valid_for = c->f2_U_valid_for = nullptr;
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
    c->ghost_valid_for = c->f2_U_valid_for = nullptr;
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
    TrajAcc acc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, c->stream));
    TRY(dev_D(c, c->U, c->chi, c->phi, c->hp.m0, false));                 // hmc.cpp:160
    TRY(hmc_leapfrog(c, &acc));                                           // hmc.cpp:161
    TRY(hmc_hamiltonian_async(c, c->Up, c->pip, c->phi, 0, &acc));        // hmc.cpp:162 (new)
    TRY(hmc_hamiltonian_async(c, c->U, c->pi, c->phi, 5, &acc));          //             (old)
    CU(cudaEventRecord(e1, c->stream));
    TRY(fetch_sums(c, 10));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
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
        c->ghost_valid_for = c->f2_U_valid_for = nullptr;
    }
    return SM_OK;
}

int sm_hmc_force(sm_ctx* c, const double* p0, const double* p1, double* F0, double* F1, int* converged) {
    TRY(set_device(c));
    NEED(p0); NEED(p1); NEED(F0); NEED(F1);
    if (!c->hmc_has_gauge) return fail(SM_ERR_STATE, "sm_hmc_set_gauge first");
    TRY(h2d_c(c, c->phi, p0, p1));
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
    fclose(f);
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
