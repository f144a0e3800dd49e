// sm_context.cuh -- error reporting, lazily bound NCCL, the device context (sm_ctx) and its launch geometry.
// Part of the single translation unit sm_abi.cu (static functions, included in dependency order).
#pragma once
#include "../../include/schwinger_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>

#include "sm_kernels.cuh"
#include "sm_fused.cuh"
#include "sm_fused_tma.cuh"
#include "sm_cluster_cg.cuh"
#include "sm_evenodd_cg.cuh"

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

// NVTX ranges around the phases of the path (CG solve, force, link/momentum update, Hamiltonian, trajectory): visible in
// Nsight Systems / Compute timelines; header-only (nvtx3), a few nanoseconds when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

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
    decltype(&ncclAllGather) AllGather = nullptr;
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
    BIND(GetUniqueId) BIND(CommInitRank) BIND(CommDestroy) BIND(Send) BIND(Recv) BIND(AllReduce) BIND(AllGather) BIND(GroupStart)
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
    cudaEvent_t ev_ready = nullptr, ev_ghost = nullptr, ev_packed = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;   // trajectory timing
    bool overlap = true;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_poll[2] = {nullptr, nullptr};
    double last_ms = 0.0;
    long long launches = 0;

    // launch geometry
    dim3 wil_block, wil_grid, wil_grid_plain;
    int rows_per_block = 0, rows_per_block_plain = 0;
    dim3 fus_block, fus_grid;   // one-pass D D^dagger (sm_fused.cuh)
    int fus_rows = 0, fus_cols = 0;
    int fus_rb = 8, fus_split_rows = 0, fus_split_chunks = 0;   // interior/boundary launch split (split lattice)
    bool use_fused = true;      // SM_DD_PATH=twopass selects the two-pass form
    bool fused_tma = true;      // rows staged by TMA bulk copies (k_dd_tma); SM_FUSED_TMA=0: per-thread cp.async (k_dd_fused)
    int fused_stages = 4;       // rows of shared-memory staging per block in k_dd_tma<PLAIN/DOT> (SM_FUSED_STAGES=3|4)
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
    cplx* eo_t = nullptr;   // even-odd solver: Dhat^dagger d
    double* eo_wsum = nullptr;
    int eo_coop = -1;       // cooperative even-odd CG usable on this device / lattice (-1: not asked yet; SM_EO_COOP=0 disables)
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
    // one-pass D D^dagger on lattices split along t (sm_ops.cuh: launch_fused_tsplit): 2-deep ghost columns [comp][wx][2]
    // and ghost rows widened to wt + 4 (corner entries); kinds: 0 U, 1 psi, 2 r, 3/4 the d ping-pong
    bool tsplit_onepass = true;      // SM_TSPLIT_ONEPASS=0: keep the two-pass kernels on t-splits
    bool self_t = false, self_x = false;   // SM_SELF_GHOSTS=t|x|xt (tests, one GPU): a single tile that takes its own opposite edges as ghosts
    cplx* tg_col[5][2] = {};         // [kind][lo, hi]
    cplx* tg_row[5][2] = {};
    cplx *tg_sendc = nullptr, *tg_sendr = nullptr;
    const cplx* tg_U_valid_for = nullptr;
    // peer-memory halo push (sm_p2p_connect): one window per rank, [kind: psi, r][parity][side: lo, hi][4 wt]
    // complex + 4 epoch flags; neighbours store into it over NVLink
    cplx* win = nullptr;
    unsigned int* win_flags = nullptr;
    size_t win_bytes = 0;
    void* peer_win[2] = {nullptr, nullptr};   // -x, +x neighbour's window (peer pointers)
    void* peer_all[kMaxPeers] = {nullptr};    // every rank's window as seen from here ([rank] = win); peer sums only
    bool p2p = false;
    bool peer_sums = false;                   // CG sums gathered through the windows (all ranks mapped): no NCCL in the loop
    unsigned int solve_seq = 0;               // CG solves so far: epochs of a solve are (solve_seq << 16) + k
    unsigned int p2p_epoch[2] = {0, 0};
    unsigned int* push_ticket = nullptr;
    CUresult (*wait_value32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;

    std::vector<void*> user_fields;

    // CUDA graphs of one batch of one-pass CG iterations, keyed on what the kernels bake in
    struct CgGraph {
        const void* U;
        const void* x;
        double m0;
        cudaGraphExec_t exec;
        int kernels;
    };
    std::vector<CgGraph> cg_graphs;
    std::vector<CgGraph> eo_graphs;   // ... and of the even-odd solver
    bool use_graphs = true;
    bool pdl = false;             // SM_PDL=1: programmatic dependent launch of the one-pass CG kernels (single tile; sm_common.cuh)
    unsigned int attr_done = 0;   // kernel attributes already set on this context's device
    int solver = SM_SOLVER_REFERENCE;
    const cplx* cg_x0 = nullptr;      // start vector of the running solve (null: phi, as the reference)
    cplx *chrono_prev = nullptr, *chrono_guess = nullptr;   // SM_SOLVER_CHRONO: previous Force solution, extrapolated guess
    int chrono_have = 0;              // solutions of this trajectory's Force solves kept so far (0, 1, 2)
    cplxf *mx_U = nullptr, *mx_r = nullptr, *mx_e = nullptr, *mx_d0 = nullptr, *mx_d1 = nullptr, *mx_Ad = nullptr;
    bool use_cluster = true;   // whole-solve resident kernels for small lattices (SM_CLUSTER_CG=0 disables)
    int coop_sites = -1;
    int cluster_ok = -1;       // -1 not asked yet; can the device schedule k_cg_cluster's cluster (cudaOccupancyMaxActiveClusters)
    // several rows per thread (k_cg_cols): variant chosen once per context (cols_plan)
    bool cols_planned = false, cols_enabled = true;
    int cols = -1, cols_force_S = 0, cols_force_T = 0;
    cplx* cols_hop = nullptr;
    double* cols_wsum = nullptr;
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
    CU(cudaEventCreate(&c->ev_t0));
    CU(cudaEventCreate(&c->ev_t1));
    CU(cudaEventCreateWithFlags(&c->ev_ghost, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
    if (const char* e = getenv("SM_OVERLAP")) c->overlap = atoi(e) != 0;
    if (const char* e = getenv("SM_GRAPHS")) c->use_graphs = atoi(e) != 0;
    if (const char* e = getenv("SM_FUSED_TMA")) c->fused_tma = atoi(e) != 0;
    if (const char* e = getenv("SM_PDL")) c->pdl = atoi(e) != 0;
    if (const char* e = getenv("SM_TSPLIT_ONEPASS")) c->tsplit_onepass = atoi(e) != 0;
    if (const char* e = getenv("SM_SELF_GHOSTS")) {
        c->self_t = strchr(e, 't') != nullptr && c->nranks == 1;
        c->self_x = strchr(e, 'x') != nullptr && c->nranks == 1;
    }
    if (const char* e = getenv("SM_FUSED_STAGES")) c->fused_stages = atoi(e) == 3 ? 3 : 4;
    if (const char* e = getenv("SM_CLUSTER_CG")) c->use_cluster = atoi(e) != 0;
    if (const char* e = getenv("SM_COLS")) {      // "0": off; "S,T": force a variant
        int S = 0, T = 512;
        const int got = sscanf(e, "%d%*[,x]%d", &S, &T);
        c->cols_enabled = got >= 1 && S > 0;
        if (c->cols_enabled) {
            c->cols_force_S = S;
            c->cols_force_T = T;
        }
    }
    CU(cudaEventCreate(&c->ev_a));
    CU(cudaEventCreate(&c->ev_b));
    CU(cudaEventCreateWithFlags(&c->ev_poll[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_poll[1], cudaEventDisableTiming));

    // stencil tiles: TT sites along t (coalesced 16-byte accesses), TX rows per step
    const int TT = c->wt >= 128 ? 128 : (c->wt >= 64 ? 64 : 32);
    const int TX = kBlock / TT;
    c->wil_block = dim3(TT, TX, 1);
    const int nT = (c->wt + TT - 1) / TT;
    // grids sized per kernel variant (the plain stencil needs 40 registers and fits 6 blocks per SM, the
    // variants with fused sums 58-60 and fit 4): one resident wave on lattices that live in L2, 8 waves of
    // shorter row runs (~70 rows at 8192^2) on large ones -- measured 6.41 vs 5.77 TB/s there
    // (profiles/r01_sweep_wilson.txt)
    const int waves = ((long long)c->wx * c->wt >= (1LL << 22)) ? 8 : 1;
    auto grid_for = [&](int occ, dim3* grid, int* rows_out) {
        if (occ < 1) occ = 1;
        occ *= waves;
        if (const char* e = getenv("SM_WILSON_BLOCKS_PER_SM")) occ = std::max(1, atoi(e));
        const int target = c->sm_count * occ;
        const int steps = (c->wx + TX - 1) / TX;
        int GY = std::max(1, std::min(steps, target / nT));
        int rows = ((c->wx + GY - 1) / GY + TX - 1) / TX * TX;
        GY = (c->wx + rows - 1) / rows;
        *rows_out = rows;
        *grid = dim3(nT, GY, 1);
    };
    int occ_plain = 0, occ_sum = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_plain, k_wilson<false, WILSON_PLAIN>, kBlock, 0));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sum, k_wilson<false, WILSON_DOT>, kBlock, 0));
    grid_for(occ_plain, &c->wil_grid_plain, &c->rows_per_block_plain);
    grid_for(occ_sum, &c->wil_grid, &c->rows_per_block);
    const int GY = std::max(c->wil_grid.y, c->wil_grid_plain.y);

    // one-pass D D^dagger: strips of <= BT-4 columns, chunks of rows.  Large lattices: ~8 waves of
    // blocks with >= 64 rows each (4 warm-up rows per chunk); mid-size: one resident wave.
    {
        const long long V = (long long)c->wx * c->wt;
        int BT = (c->wt + 4 <= 128 || V <= (1LL << 21)) ? 128 : 256;
        if (const char* e = getenv("SM_FUSED_BT")) BT = atoi(e) == 128 ? 128 : 256;
        int strips = (c->wt + (BT - 4) - 1) / (BT - 4);
        if (const char* e = getenv("SM_FUSED_STRIPS")) strips = std::max(strips, atoi(e));   // more, narrower strips
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

    // (+ kWilsonBoundaryBlocks: on a split lattice the boundary launch of a stencil pass joins its reduction)
    const size_t max_blocks = std::max<size_t>(std::max<size_t>((size_t)nT * GY + kWilsonBoundaryBlocks, (size_t)cap),
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

// A kernel launch that may start while its predecessor on the stream drains (programmatic dependent launch): the kernel
// itself waits (pdl_wait, sm_common.cuh) before it touches anything the predecessor wrote.  Also valid under stream capture
// (the edge becomes a programmatic dependency of the graph).
static int launch_pdl(sm_ctx* c, const void* kernel, dim3 grid, dim3 block, size_t smem, void** args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelExC(&cfg, kernel, args));
    return SM_OK;
}

#define KCHECK()                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                      \
        if (e_ != cudaSuccess) return fail(SM_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
    } while (0)
