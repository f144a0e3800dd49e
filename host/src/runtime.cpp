// runtime.cpp -- process tree + GPU context of the host shell (see b200_runtime.h).
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b200_runtime.h"
#include "variables.h"

namespace b200 {

static sm_ctx* g_ctx = nullptr;
static std::vector<int> g_child_fd;    // rank 0: socket to rank i (index i-1)
static std::vector<pid_t> g_child_pid;
static int g_parent_fd = -1;           // rank > 0: socket to rank 0

static void write_all(int fd, const void* p, std::size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n > 0) {
        ssize_t w = ::write(fd, c, n);
        if (w <= 0) {
            perror("b200: write");
            _exit(1);
        }
        c += w;
        n -= (std::size_t)w;
    }
}

static void read_all(int fd, void* p, std::size_t n) {
    char* c = static_cast<char*>(p);
    while (n > 0) {
        ssize_t r = ::read(fd, c, n);
        if (r <= 0) {
            if (r < 0) perror("b200: read");
            _exit(1);   // the peer died
        }
        c += r;
        n -= (std::size_t)r;
    }
}

void check(int rc, const char* what) {
    if (rc == SM_OK) return;
    std::cerr << "[rank " << mpi::rank << "] " << what << " failed: " << sm_last_error() << std::endl;
    exit(1);
}

// CUDA devices visible to this process, asked from a throw-away child so that the parent reaches fork() without a
// CUDA context (a context does not survive fork).  -1: could not ask.
static int visible_devices() {
    int fd[2];
    if (pipe(fd) != 0) return -1;
    fflush(nullptr);
    pid_t pid = fork();
    if (pid < 0) return -1;
    if (pid == 0) {
        close(fd[0]);
        int n = 0;
        if (sm_device_count(&n) != SM_OK) n = 0;
        if (::write(fd[1], &n, sizeof(n)) != (ssize_t)sizeof(n)) _exit(1);
        _exit(0);
    }
    close(fd[1]);
    int n = -1;
    if (::read(fd[0], &n, sizeof(n)) != (ssize_t)sizeof(n)) n = -1;
    close(fd[0]);
    int st = 0;
    waitpid(pid, &st, 0);
    return n;
}

void spawn_ranks(int n_ranks) {
    mpi::size = n_ranks;
    mpi::rank = 0;
    if (n_ranks > 1) {
        // one GPU per rank: refuse here, before any worker exists, instead of leaving the other ranks stuck in the
        // NCCL rendezvous when one of them finds no device
        int first = 0;
        if (const char* d = std::getenv("SM_DEVICE")) first = std::atoi(d);
        const int have = visible_devices();
        if (have >= 0 && first + n_ranks > have) {
            std::cerr << "ranks_x*ranks_t = " << n_ranks << " ranks need GPUs " << first << ".." << first + n_ranks - 1
                      << " but only " << have << " CUDA device(s) are visible (one GPU per rank, no CPU fallback)" << std::endl;
            exit(1);
        }
    }
    for (int r = 1; r < n_ranks; r++) {
        int sv[2];
        if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) {
            perror("socketpair");
            exit(1);
        }
        fflush(nullptr);
        pid_t pid = fork();
        if (pid < 0) {
            perror("fork");
            exit(1);
        }
        if (pid == 0) {   // worker
            close(sv[0]);
            for (int fd : g_child_fd) close(fd);
            g_child_fd.clear();
            g_child_pid.clear();
            g_parent_fd = sv[1];
            mpi::rank = r;
            return;
        }
        close(sv[1]);
        g_child_fd.push_back(sv[0]);
        g_child_pid.push_back(pid);
    }
}

void bcast(void* buf, std::size_t bytes) {
    if (mpi::size == 1) return;
    if (mpi::rank == 0)
        for (int fd : g_child_fd) write_all(fd, buf, bytes);
    else
        read_all(g_parent_fd, buf, bytes);
}

void gather(const void* mine, std::size_t bytes, void* all_on_root) {
    if (mpi::rank == 0) {
        std::memcpy(all_on_root, mine, bytes);
        for (std::size_t i = 0; i < g_child_fd.size(); i++)
            read_all(g_child_fd[i], static_cast<char*>(all_on_root) + (i + 1) * bytes, bytes);
    } else {
        write_all(g_parent_fd, mine, bytes);
    }
}

void create_context() {
    int device = mpi::rank;
    if (const char* d = std::getenv("SM_DEVICE")) device = std::atoi(d) + mpi::rank;
    if (mpi::size == 1) {
        check(sm_create(LV::Nx, LV::Nt, device, &g_ctx), "sm_create");
    } else {
        unsigned char id[SM_NCCL_ID_BYTES];
        if (mpi::rank == 0) check(sm_nccl_unique_id(id), "sm_nccl_unique_id");
        bcast(id, sizeof(id));
        check(sm_create_dist(LV::Nx, LV::Nt, mpi::ranks_x, mpi::ranks_t, mpi::rank, device, id, &g_ctx),
              "sm_create_dist");
    }
    check(sm_set_cg(g_ctx, CG::tol, CG::max_iter), "sm_set_cg");
    // the shell's fields are long-lived `new[]` arrays whose destructor tells the library (variables.h): let it
    // page-lock them so that D_phi(...), conjugate_gradient(...) on host fields copy at the full link rate
    if (!std::getenv("SM_HOST_REGISTER") || std::atoi(std::getenv("SM_HOST_REGISTER")) != 0) sm_host_register(1);
}

sm_ctx* ctx() {
    if (!g_ctx) {
        std::cerr << "b200: no device context (call initializeMPI() first)" << std::endl;
        exit(1);
    }
    return g_ctx;
}

void shutdown() {
    if (g_ctx) sm_destroy(g_ctx);
    g_ctx = nullptr;
    if (mpi::rank == 0) {
        for (int fd : g_child_fd) close(fd);
        for (pid_t p : g_child_pid) {
            int st = 0;
            waitpid(p, &st, 0);
        }
    } else if (g_parent_fd >= 0) {
        close(g_parent_fd);
    }
}

double wtime() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace b200
