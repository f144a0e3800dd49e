// minimpi: a header-only stand-in for <mpi.h>, TEST INFRASTRUCTURE ONLY.
//
// The image has no MPI.  The reference (Fabian2598/SchwingerModel) needs exactly the
// 18 MPI entry points listed in SURVEY.md App. A; this header provides them so that the
// UNMODIFIED reference sources under /root/reference compile and run here:
//   * np == 1 : everything is a local copy (reference takes its `size==1` serial branches)
//   * np  > 1 : ranks are fork()ed processes that talk through one MAP_SHARED arena
//               (per-pair FIFO mailboxes for Send/Recv, slot arrays for Allreduce/Bcast,
//               a sense-reversing barrier).  Send is always eager (copies into the
//               mailbox), which is what the reference's Send-then-Recv pattern assumes
//               (dirac_operator.cpp:66-88).
// Rank <-> Cartesian coordinates are row-major with no reordering, i.e. what every MPI
// implementation does for MPI_Cart_create on a fresh communicator.
//
// Nothing here is part of the product; it exists so oracle/_ref can be the real reference.
#ifndef MINIMPI_MPI_H
#define MINIMPI_MPI_H

#include <atomic>
#include <new>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

typedef int MPI_Comm;
typedef int MPI_Datatype;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
typedef int MPI_Op;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_SUM 1
// basic datatypes: id == size class below
#define MPI_CHAR 1
#define MPI_INT 2
#define MPI_DOUBLE 3
#define MPI_DOUBLE_COMPLEX 4

namespace minimpi {

constexpr int kMaxRanks = 64;
constexpr int kSlots = 8;          // mailbox depth per ordered pair
constexpr int kMaxTypes = 32;

struct DerivedType {               // MPI_Type_vector (+ create_resized) description
    int count, blocklen, stride;   // in units of the base type
    int base;                      // base datatype id
    long extent_bytes;             // extent used for displacements
};

struct Mail {                      // one message slot
    std::atomic<int> full;
    int tag;
    long bytes;
};

struct Shared {
    int np;
    long max_msg;
    std::atomic<int> bar_count;
    std::atomic<int> bar_sense;
    // allreduce / bcast scratch: kMaxRanks * 64 bytes
    alignas(64) unsigned char red[kMaxRanks][64];
    // mailbox headers [src][dst][slot]; payload follows the struct in the arena
    Mail mail[kMaxRanks][kMaxRanks][kSlots];
    std::atomic<long> head[kMaxRanks][kMaxRanks];   // next slot to read
    std::atomic<long> tail[kMaxRanks][kMaxRanks];   // next slot to write
};

inline Shared* g_shared = nullptr;
inline unsigned char* g_payload = nullptr;
inline int g_rank = 0;
inline int g_np = 1;
inline int g_local_sense = 0;
inline pid_t g_children[kMaxRanks];
inline int g_dims[2] = {1, 1};
inline DerivedType g_types[kMaxTypes];
inline int g_ntypes = 0;
inline long g_allreduce_calls = 0;   // lets a harness recover CG iteration counts
inline long g_sendrecv_calls = 0;

inline long type_size(MPI_Datatype t) {
    switch (t) {
        case MPI_CHAR: return 1;
        case MPI_INT: return 4;
        case MPI_DOUBLE: return 8;
        case MPI_DOUBLE_COMPLEX: return 16;
        default: return -1;
    }
}

inline unsigned char* slot_ptr(int src, int dst, int slot) {
    long idx = ((long)src * g_np + dst) * kSlots + slot;
    return g_payload + idx * g_shared->max_msg;
}

// Allocate the shared arena and fork np-1 children.  Returns this process's rank.
// The caller (rank 0 = the original process) must call join() at the end; children
// must call child_exit().
inline int spawn(int np, long max_msg_bytes) {
    g_np = np;
    g_rank = 0;
    g_local_sense = 0;
    if (np > kMaxRanks) { std::fprintf(stderr, "minimpi: too many ranks\n"); std::exit(2); }
    long payload = (long)np * np * kSlots * max_msg_bytes;
    long total = (long)sizeof(Shared) + payload + 4096;
    void* mem = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (mem == MAP_FAILED) { std::perror("minimpi mmap"); std::exit(2); }
    g_shared = new (mem) Shared();
    g_shared->np = np;
    g_shared->max_msg = max_msg_bytes;
    g_shared->bar_count.store(0);
    g_shared->bar_sense.store(0);
    g_payload = reinterpret_cast<unsigned char*>(mem) + ((sizeof(Shared) + 4095) / 4096) * 4096;
    for (int r = 1; r < np; r++) {
        std::fflush(stdout);
        std::fflush(stderr);
        pid_t pid = fork();
        if (pid < 0) { std::perror("minimpi fork"); std::exit(2); }
        if (pid == 0) { g_rank = r; return r; }
        g_children[r] = pid;
    }
    return 0;
}

inline void relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}

inline void barrier() {
    if (g_np == 1) return;
    g_local_sense ^= 1;
    if (g_shared->bar_count.fetch_add(1) == g_np - 1) {
        g_shared->bar_count.store(0);
        g_shared->bar_sense.store(g_local_sense);
    } else {
        int spins = 0;
        while (g_shared->bar_sense.load() != g_local_sense) {
            relax();
            if (++spins > 2000) { sched_yield(); spins = 0; }
        }
    }
}

[[noreturn]] inline void child_exit() {
    std::fflush(stdout);
    std::fflush(stderr);
    _exit(0);
}

inline int join() {   // rank 0 only
    int bad = 0;
    for (int r = 1; r < g_np; r++) {
        int st = 0;
        waitpid(g_children[r], &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++;
    }
    g_np = 1;
    g_rank = 0;
    return bad;
}

inline void send_bytes(const void* buf, long bytes, int dst, int tag) {
    g_sendrecv_calls++;
    if (bytes > g_shared->max_msg) {
        std::fprintf(stderr, "minimpi: message of %ld B exceeds max_msg %ld\n", bytes, g_shared->max_msg);
        std::exit(2);
    }
    long t = g_shared->tail[g_rank][dst].load();
    int slot = (int)(t % kSlots);
    Mail& m = g_shared->mail[g_rank][dst][slot];
    int spins = 0;
    while (m.full.load(std::memory_order_acquire)) {
        relax();
        if (++spins > 2000) { sched_yield(); spins = 0; }
    }
    std::memcpy(slot_ptr(g_rank, dst, slot), buf, bytes);
    m.tag = tag;
    m.bytes = bytes;
    m.full.store(1, std::memory_order_release);
    g_shared->tail[g_rank][dst].store(t + 1);
}

inline void recv_bytes(void* buf, long bytes, int src, int tag) {
    long h = g_shared->head[src][g_rank].load();
    int slot = (int)(h % kSlots);
    Mail& m = g_shared->mail[src][g_rank][slot];
    int spins = 0;
    while (!m.full.load(std::memory_order_acquire)) {
        relax();
        if (++spins > 2000) { sched_yield(); spins = 0; }
    }
    if (m.tag != tag || m.bytes != bytes) {
        std::fprintf(stderr, "minimpi: rank %d expected tag %d/%ld B from %d, got tag %d/%ld B\n",
                     g_rank, tag, bytes, src, m.tag, m.bytes);
        std::exit(2);
    }
    std::memcpy(buf, slot_ptr(src, g_rank, slot), bytes);
    m.full.store(0, std::memory_order_release);
    g_shared->head[src][g_rank].store(h + 1);
}

// copy `n` items of datatype `t` between a strided (derived) view and a packed buffer
inline void pack_or_unpack(unsigned char* strided, unsigned char* packed, MPI_Datatype t, bool pack) {
    const DerivedType& d = g_types[t - 100];
    long bs = type_size(d.base);
    for (int c = 0; c < d.count; c++) {
        unsigned char* s = strided + (long)c * d.stride * bs;
        unsigned char* p = packed + (long)c * d.blocklen * bs;
        if (pack) std::memcpy(p, s, (long)d.blocklen * bs);
        else std::memcpy(s, p, (long)d.blocklen * bs);
    }
}

}  // namespace minimpi

// ------------------------------------------------------------------ the MPI surface

inline int MPI_Init(int*, char***) {
    const char* e = std::getenv("MINIMPI_NP");
    int np = e ? std::atoi(e) : 1;
    if (np > 1 && minimpi::g_shared == nullptr) {
        const char* m = std::getenv("MINIMPI_MAXMSG");
        minimpi::spawn(np, m ? std::atol(m) : (1L << 20));
    }
    return MPI_SUCCESS;
}

inline int MPI_Finalize() {
    if (minimpi::g_np > 1) {
        minimpi::barrier();
        if (minimpi::g_rank != 0) minimpi::child_exit();
        minimpi::join();
    }
    return MPI_SUCCESS;
}

inline int MPI_Comm_size(MPI_Comm, int* size) { *size = minimpi::g_np; return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = minimpi::g_rank; return MPI_SUCCESS; }

inline double MPI_Wtime() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

inline int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm) {
    using namespace minimpi;
    if (g_np == 1) return MPI_SUCCESS;
    long bytes = (long)count * type_size(t);
    // stream through the 64-byte-per-rank scratch in chunks
    long chunk = (long)sizeof(g_shared->red);
    unsigned char* scratch = &g_shared->red[0][0];
    for (long off = 0; off < bytes; off += chunk) {
        long n = bytes - off < chunk ? bytes - off : chunk;
        if (g_rank == root) std::memcpy(scratch, (unsigned char*)buf + off, n);
        barrier();
        if (g_rank != root) std::memcpy((unsigned char*)buf + off, scratch, n);
        barrier();
    }
    return MPI_SUCCESS;
}

inline int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op, MPI_Comm) {
    using namespace minimpi;
    g_allreduce_calls++;
    int nd = (t == MPI_DOUBLE_COMPLEX) ? 2 * count : count;   // doubles
    if (t != MPI_DOUBLE && t != MPI_DOUBLE_COMPLEX) {
        std::fprintf(stderr, "minimpi: Allreduce only on doubles\n");
        std::exit(2);
    }
    if (g_np == 1) { std::memcpy(recv, send, sizeof(double) * nd); return MPI_SUCCESS; }
    if (nd > 8) { std::fprintf(stderr, "minimpi: Allreduce payload too large\n"); std::exit(2); }
    std::memcpy(g_shared->red[g_rank], send, sizeof(double) * nd);
    barrier();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < g_np; r++) {          // fixed rank order: every rank gets the same bits
        double v[8];
        std::memcpy(v, g_shared->red[r], sizeof(double) * nd);
        for (int i = 0; i < nd; i++) acc[i] += v[i];
    }
    barrier();
    std::memcpy(recv, acc, sizeof(double) * nd);
    return MPI_SUCCESS;
}

inline int MPI_Send(const void* buf, int count, MPI_Datatype t, int dst, int tag, MPI_Comm) {
    minimpi::send_bytes(buf, (long)count * minimpi::type_size(t), dst, tag);
    return MPI_SUCCESS;
}

inline int MPI_Recv(void* buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm, MPI_Status*) {
    minimpi::recv_bytes(buf, (long)count * minimpi::type_size(t), src, tag);
    return MPI_SUCCESS;
}

inline int MPI_Cart_create(MPI_Comm, int ndims, const int* dims, const int*, int, MPI_Comm* out) {
    if (ndims != 2) { std::fprintf(stderr, "minimpi: 2-D only\n"); std::exit(2); }
    minimpi::g_dims[0] = dims[0];
    minimpi::g_dims[1] = dims[1];
    *out = 1;
    return MPI_SUCCESS;
}

inline int MPI_Cart_coords(MPI_Comm, int rank, int, int* coords) {
    coords[0] = rank / minimpi::g_dims[1];
    coords[1] = rank % minimpi::g_dims[1];
    return MPI_SUCCESS;
}

inline int MPI_Cart_rank(MPI_Comm, const int* coords, int* rank) {
    int d0 = minimpi::g_dims[0], d1 = minimpi::g_dims[1];
    int c0 = ((coords[0] % d0) + d0) % d0;
    int c1 = ((coords[1] % d1) + d1) % d1;
    *rank = c0 * d1 + c1;
    return MPI_SUCCESS;
}

inline int MPI_Cart_shift(MPI_Comm c, int dir, int disp, int* src, int* dst) {
    int me[2];
    MPI_Cart_coords(c, minimpi::g_rank, 2, me);
    int a[2] = {me[0], me[1]}, b[2] = {me[0], me[1]};
    a[dir] -= disp;
    b[dir] += disp;
    MPI_Cart_rank(c, a, src);
    MPI_Cart_rank(c, b, dst);
    return MPI_SUCCESS;
}

inline int MPI_Type_vector(int count, int blocklen, int stride, MPI_Datatype base, MPI_Datatype* out) {
    using namespace minimpi;
    if (g_ntypes >= kMaxTypes) g_ntypes = 0;   // harnesses re-initialise many times
    g_types[g_ntypes] = {count, blocklen, stride, base,
                         ((long)(count - 1) * stride + blocklen) * type_size(base)};
    *out = 100 + g_ntypes++;
    return MPI_SUCCESS;
}

inline int MPI_Type_commit(MPI_Datatype*) { return MPI_SUCCESS; }

inline int MPI_Type_create_resized(MPI_Datatype in, long, long extent, MPI_Datatype* out) {
    using namespace minimpi;
    if (g_ntypes >= kMaxTypes) g_ntypes = 0;
    g_types[g_ntypes] = g_types[in - 100];
    g_types[g_ntypes].extent_bytes = extent;
    *out = 100 + g_ntypes++;
    return MPI_SUCCESS;
}

// Gatherv/Scatterv as the reference uses them (gauge_conf.cpp:390-395, 537-541): the
// per-rank side is a packed run of basic elements, the root side one derived-type item
// per rank placed at displs[r] * extent.
inline int MPI_Gatherv(const void* send, int sendcount, MPI_Datatype st, void* recv, const int* counts,
                       const int* displs, MPI_Datatype rt, int root, MPI_Comm) {
    using namespace minimpi;
    long bytes = (long)sendcount * type_size(st);
    const DerivedType& d = g_types[rt - 100];
    if (g_rank == root) {
        for (int r = 0; r < g_np; r++) {
            if (counts[r] != 1) { std::fprintf(stderr, "minimpi: Gatherv count != 1\n"); std::exit(2); }
            unsigned char* dstp = (unsigned char*)recv + (long)displs[r] * d.extent_bytes;
            if (r == root) {
                pack_or_unpack(dstp, (unsigned char*)const_cast<void*>(send), rt, false);
            } else {
                unsigned char* tmp = (unsigned char*)std::malloc(bytes);
                // large tiles travel in max_msg-sized pieces
                for (long off = 0; off < bytes; off += g_shared->max_msg) {
                    long n = bytes - off < g_shared->max_msg ? bytes - off : g_shared->max_msg;
                    recv_bytes(tmp + off, n, r, 9000);
                }
                pack_or_unpack(dstp, tmp, rt, false);
                std::free(tmp);
            }
        }
    } else {
        for (long off = 0; off < bytes; off += g_shared->max_msg) {
            long n = bytes - off < g_shared->max_msg ? bytes - off : g_shared->max_msg;
            send_bytes((const unsigned char*)send + off, n, root, 9000);
        }
    }
    return MPI_SUCCESS;
}

inline int MPI_Scatterv(const void* send, const int* counts, const int* displs, MPI_Datatype st, void* recv,
                        int recvcount, MPI_Datatype rt, int root, MPI_Comm) {
    using namespace minimpi;
    long bytes = (long)recvcount * type_size(rt);
    const DerivedType& d = g_types[st - 100];
    if (g_rank == root) {
        for (int r = 0; r < g_np; r++) {
            if (counts[r] != 1) { std::fprintf(stderr, "minimpi: Scatterv count != 1\n"); std::exit(2); }
            unsigned char* srcp = (unsigned char*)const_cast<void*>(send) + (long)displs[r] * d.extent_bytes;
            if (r == root) {
                pack_or_unpack(srcp, (unsigned char*)recv, st, true);
            } else {
                unsigned char* tmp = (unsigned char*)std::malloc(bytes);
                pack_or_unpack(srcp, tmp, st, true);
                for (long off = 0; off < bytes; off += g_shared->max_msg) {
                    long n = bytes - off < g_shared->max_msg ? bytes - off : g_shared->max_msg;
                    send_bytes(tmp + off, n, r, 9001);
                }
                std::free(tmp);
            }
        }
    } else {
        for (long off = 0; off < bytes; off += g_shared->max_msg) {
            long n = bytes - off < g_shared->max_msg ? bytes - off : g_shared->max_msg;
            recv_bytes((unsigned char*)recv + off, n, root, 9001);
        }
    }
    return MPI_SUCCESS;
}

#endif  // MINIMPI_MPI_H
