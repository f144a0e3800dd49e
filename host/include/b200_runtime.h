// b200_runtime.h -- what replaces MPI in the host shell: one process per GPU, a libschwinger_b200
// context per process, NCCL between the GPUs (inside the library) and socket pairs between the
// processes for the few host-side collectives the reference does with MPI_Bcast / MPI_Gatherv.
#ifndef SM_HOST_B200_RUNTIME_H
#define SM_HOST_B200_RUNTIME_H

#include <cstddef>

#include "schwinger_b200.h"

namespace b200 {

// Fork ranks_x*ranks_t - 1 worker processes (before any CUDA call); the caller continues as rank 0.
// Sets mpi::rank / mpi::size.  Replaces `mpirun -n N` (README.md:49 of the reference).
void spawn_ranks(int n_ranks);
// Create this process's device context (GPU = rank, or SM_DEVICE for a single rank).
// Replaces initializeMPI() (include/mpi_setup.h:96-100).
void create_context();
sm_ctx* ctx();
void shutdown();   // destroy the context, reap the workers (MPI_Finalize)
void check(int rc, const char* what);   // abort with sm_last_error() on failure, like the reference's exit(1)

// host-side collectives over the process tree
void bcast(void* buf, std::size_t bytes);                                  // MPI_Bcast from rank 0
void gather(const void* mine, std::size_t bytes, void* all_on_root);       // MPI_Gather to rank 0
double wtime();                                                            // MPI_Wtime

}  // namespace b200
#endif
