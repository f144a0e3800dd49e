/*
 * schwinger_b200.h -- C ABI of libschwinger_b200.so, the B200 (sm_100a) implementation of the
 * Schwinger-model HMC fermion hot path.
 *
 * The reference (Fabian2598/SchwingerModel) has no FFI layer: its boundary is the C++ surface of
 * include/dirac_operator.h, include/conjugate_gradient.h, include/gauge_conf.h and include/hmc.h.
 * Every entry point below names the reference function (file:line under /root/reference) it
 * replaces.  The C++ shell in host/ keeps the reference signatures and forwards here; Python
 * binds the same symbols with ctypes (schwingermodel_b200/_abi.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every call returns an int status (SM_OK == 0) and never
 *     throws or prints.  sm_last_error() gives the text of the last failure on this thread.
 *   - lattice site n = x*Nt + t (t fastest); mu=0 is time, mu=1 is space (src/variables.cpp:10-12).
 *   - a complex field is TWO arrays (mu0, mu1) of V complex doubles stored (re,im) interleaved,
 *     exactly the reference's `spinor{c_double* mu0; c_double* mu1;}` (include/variables.h:54-100);
 *     a real field is two arrays of V doubles (`re_field`, include/variables.h:102-141).
 *     std::complex<double>* may be passed as double* (layout-compatible).
 *   - "h_" arguments are HOST pointers (copied in/out inside the call); "d_" arguments are DEVICE
 *     fields obtained from sm_field_alloc (one allocation: mu0 at [0,V), mu1 at [V,2V) elements).
 *   - in a distributed context (ranks_x*ranks_t > 1, one process per GPU) every call that touches
 *     neighbours or global sums is collective, like the reference's MPI calls; host buffers hold
 *     this rank's width_x*width_t tile (n = x_local*width_t + t_local), as in the reference.
 *   - like the reference (single-threaded per rank, global scratch), a context is not re-entrant: one thread
 *     drives one context at a time; different contexts (GPUs) may be driven by different threads.
 *   - there is no CPU fallback: without a usable CUDA device sm_create fails with SM_ERR_CUDA.
 */
#ifndef SCHWINGER_B200_H
#define SCHWINGER_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SM_API __attribute__((visibility("default")))
#else
#define SM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sm_ctx sm_ctx;

enum {
    SM_OK = 0,
    SM_ERR_ARG = 1,     /* bad argument (null pointer, non-divisible decomposition ...) */
    SM_ERR_CUDA = 2,    /* CUDA runtime error (text in sm_last_error) */
    SM_ERR_NCCL = 3,    /* NCCL error */
    SM_ERR_IO = 4,      /* file could not be opened / short read */
    SM_ERR_STATE = 5    /* call sequence error (e.g. trajectory before set_gauge) */
};

#define SM_NCCL_ID_BYTES 128
#define SM_P2P_HANDLE_BYTES 64

/* ---- life cycle ------------------------------------------------------------------------------
 * replaces initializeMPI() + allocate_lattice_arrays() + periodic_boundary()
 * (include/mpi_setup.h:96-100, src/variables.cpp:47-64, include/dirac_operator.h:35-62). */
SM_API int sm_create(int Nx, int Nt, int device, sm_ctx** out);
/* rank = coords_x*ranks_t + coords_t (mpi_setup.h:39-47, row-major Cartesian, no reordering).
 * nccl_id: SM_NCCL_ID_BYTES bytes from sm_nccl_unique_id() on rank 0, broadcast by the caller. */
SM_API int sm_create_dist(int Nx, int Nt, int ranks_x, int ranks_t, int rank, int device, const void* nccl_id, sm_ctx** out);
SM_API int sm_nccl_unique_id(void* out_id /* SM_NCCL_ID_BYTES */);
/* Optional, lattices split along x only: halo rows are stored straight into the neighbour's memory over
 * NVLink (CUDA IPC) instead of ncclSend/ncclRecv.  Every rank exports a handle, the caller gathers the
 * ranks' handles in rank order (MPI_Allgather in the reference's world) and every rank connects. */
SM_API int sm_p2p_handle(sm_ctx* ctx, void* handle_out /* SM_P2P_HANDLE_BYTES */);
SM_API int sm_p2p_connect(sm_ctx* ctx, const void* all_handles /* nranks * SM_P2P_HANDLE_BYTES, rank order */);
/* Host-buffer calls copy the caller's arrays (the reference's `new[]`-allocated spinor::mu0/mu1, include/variables.h:54-100)
   to and from the device.  sm_host_register(1): page-lock every caller buffer the first time it is passed and keep it
   locked (pointer-keyed cache), so later copies run at the full host-link rate instead of through the driver's pageable
   staging; the owner must call sm_host_forget(ptr) before freeing such a buffer.  sm_host_register(0) (default) unlocks
   everything and stops. */
SM_API int sm_host_register(int enable);
SM_API int sm_host_forget(const void* ptr);
/* number of CUDA devices visible to this process; initialises the CUDA runtime, so a launcher that forks one process
   per GPU (the stand-in for `mpirun -n ranks_x*ranks_t`, README.md:49 of the reference) asks from a throw-away child */
SM_API int sm_device_count(int* n);
/* how this context exchanges halos and sums (the reference: MPI_Send/Recv, src/dirac_operator.cpp:66-88, and
   MPI_Allreduce, include/variables.h:190): 0 = single tile or NCCL send/recv + all-reduce; 1 = halo rows by peer-memory
   stores; 2 = halo rows and the CG iteration's sums by the kernels themselves over peer memory (no NCCL in the loop).
   sm_create_dist connects the windows itself on lattices split along x (SM_P2P=0 in the environment disables). */
SM_API int sm_peer_mode(const sm_ctx* ctx, int* mode);
SM_API int sm_destroy(sm_ctx* ctx);
SM_API const char* sm_last_error(void);
/* local tile: dims[0]=width_x, dims[1]=width_t, dims[2]=rank, dims[3]=nranks */
SM_API int sm_local_dims(const sm_ctx* ctx, int dims[4]);
/* CG controls = the reference's mutable globals CG::tol / CG::max_iter (src/variables.cpp:35-38; defaults 1e-10 and 10000).
 * max_iter <= 60000 (the iteration index shares a 32-bit epoch with the solve number on split lattices). */
SM_API int sm_set_cg(sm_ctx* ctx, double tol, int max_iter);
/* Solver used by every CG of the context (conjugate_gradient, HMC::Force, HMC::Action):
 *   SM_SOLVER_REFERENCE  the reference's algorithm in double precision (default; dH parity <= 1e-8)
 *   SM_SOLVER_MIXED      opt-in: single-precision inner CG inside a double-precision defect correction
 *                        (lattices above ~75k sites on a single tile); same stopping criterion, checked on the
 *                        TRUE residual, but a different iterate than the reference's
 *   SM_SOLVER_CHRONO     opt-in: the reference's CG arithmetic and stopping rule, but inside a trajectory the solves of
 *                        HMC::Force (and the proposal's HMC::Action) start from the previous solution / the linear
 *                        extrapolation of the last two instead of from phi (conjugate_gradient.cpp:16): fewer iterations,
 *                        a different iterate, dH differs at the CG-truncation level
 *   SM_SOLVER_EVENODD    opt-in: even-odd preconditioned HMC.  The pseudofermion lives on the even sites with the action
 *                        phi_e^dagger (Dhat Dhat^dagger)^-1 phi_e, Dhat = m - (1/4m) H_eo H_oe the Schur complement of D
 *                        (same determinant as D D^dagger up to a constant, hence the same gauge-field distribution);
 *                        every solve is a CG on half the sites with a ~4x smaller condition number near the critical
 *                        mass.  A different Markov chain than the reference's: dH is not comparable trajectory by
 *                        trajectory; plaquette agrees statistically, acceptance is higher in equilibrium.  From a hot start
 *                        its leapfrog error is the larger one: thermalise with the reference solver (or a finer step)
 *                        first.  Single tile, even Nx and Nt. */
enum { SM_SOLVER_REFERENCE = 0, SM_SOLVER_MIXED = 1, SM_SOLVER_CHRONO = 2, SM_SOLVER_EVENODD = 3 };
SM_API int sm_set_solver(sm_ctx* ctx, int solver);
/* device time (ms) of the last sm_* call on this context, measured with CUDA events on the
 * context's stream around the kernels only (no copies) */
SM_API int sm_last_kernel_ms(const sm_ctx* ctx, double* ms);
/* 1 if D D^dagger runs as the one-pass kernel on this context (single tile or x-only split), 0 if as two stencil passes */
SM_API int sm_one_pass_dd(const sm_ctx* ctx, int* one_pass);
/* number of kernels this library launched on the context since creation */
SM_API int sm_launch_count(const sm_ctx* ctx, long long* n);

/* ---- geometry --------------------------------------------------------------------------------
 * periodic_boundary() tables of `rank` in a ranks_x*ranks_t decomposition, produced by the SAME
 * device index arithmetic the kernels use (include/dirac_operator.h:35-62).
 * RightPB/LeftPB: int32[2*m] indexed [2n+mu]; SignR/SignL: complex[2*m] as (re,im);
 * x_1_t1 / x1_t_1: int32[m]; m = width_x*width_t. */
SM_API int sm_tables(sm_ctx* ctx, int ranks_x, int ranks_t, int rank, int* RightPB, int* LeftPB, double* SignR, double* SignL,
              int* x_1_t1, int* x1_t_1);

/* ---- drop-in operators on HOST buffers (copies inside the call) ------------------------------ */
/* D_phi          src/dirac_operator.cpp:24-244  (include/dirac_operator.h:71) */
SM_API int sm_D_phi(sm_ctx* ctx, const double* h_U0, const double* h_U1, const double* h_phi0, const double* h_phi1,
             double* h_out0, double* h_out1, double m0);
/* D_dagger_phi   src/dirac_operator.cpp:247-473 (include/dirac_operator.h:80) */
SM_API int sm_D_dagger_phi(sm_ctx* ctx, const double* h_U0, const double* h_U1, const double* h_phi0, const double* h_phi1,
                    double* h_out0, double* h_out1, double m0);
/* D_D_dagger_phi src/dirac_operator.cpp:477-480 (include/dirac_operator.h:87) */
SM_API int sm_D_D_dagger_phi(sm_ctx* ctx, const double* h_U0, const double* h_U1, const double* h_phi0, const double* h_phi1,
                      double* h_out0, double* h_out1, double m0);
/* dot            include/variables.h:181-192: sum_n x conj(y) over both components, global */
SM_API int sm_dot(sm_ctx* ctx, const double* h_x0, const double* h_x1, const double* h_y0, const double* h_y1,
           double out_re_im[2]);
/* conjugate_gradient  src/conjugate_gradient.cpp:4-67 (include/conjugate_gradient.h:16).
 * x0 = phi, stop when ||r|| < tol*||phi||; *converged = the reference's return value (1/0);
 * *iterations = the reference's k at exit; tol / max_iter from sm_set_cg. */
SM_API int sm_conjugate_gradient(sm_ctx* ctx, const double* h_U0, const double* h_U1, const double* h_phi0,
                          const double* h_phi1, double* h_x0, double* h_x1, double m0, int* converged,
                          int* iterations);
/* phi_dag_partialD_phi  src/dirac_operator.cpp:486-580 (include/dirac_operator.h:93) */
/* x_e = (Dhat Dhat^dagger)^-1 phi_e on the even sites (phi's odd sites are ignored, x's are zero): the solve of the
 * opt-in even-odd HMC (SM_SOLVER_EVENODD), with the reference CG's start vector and stopping rule
 * (src/conjugate_gradient.cpp:16,45) on the Schur complement of D */
SM_API int sm_evenodd_solve(sm_ctx* ctx, const double* h_U0, const double* h_U1, const double* h_phi0, const double* h_phi1,
                            double* h_x0, double* h_x1, double m0, int* converged, int* iterations);
SM_API int sm_phi_dag_partialD_phi(sm_ctx* ctx, const double* h_U0, const double* h_U1, const double* h_left0,
                            const double* h_left1, const double* h_right0, const double* h_right1, double* h_F0,
                            double* h_F1);
/* GaugeConf::Compute_Staple  src/gauge_conf.cpp:89-373 */
SM_API int sm_compute_staple(sm_ctx* ctx, const double* h_U0, const double* h_U1, double* h_K0, double* h_K1);
/* GaugeConf::Compute_Plaquette01 src/gauge_conf.cpp:41-85 (+ MeasureSp_HMC :427-437,
 * Compute_gaugeAction :441-449).  h_P may be NULL.  sums[0] = sum Re P, sums[1] = beta*sum Re(1-P) */
SM_API int sm_compute_plaquette(sm_ctx* ctx, const double* h_U0, const double* h_U1, double beta, double* h_P,
                         double sums[2]);

/* ---- device-resident fields and operators ---------------------------------------------------- */
/* complex!=0: 2*V complex doubles, else 2*V doubles */
SM_API int sm_field_alloc(sm_ctx* ctx, int complex_field, double** d_field);
SM_API int sm_field_free(sm_ctx* ctx, double* d_field);
SM_API int sm_field_upload(sm_ctx* ctx, double* d_field, const double* h_mu0, const double* h_mu1, int complex_field);
SM_API int sm_field_download(sm_ctx* ctx, const double* d_field, double* h_mu0, double* h_mu1, int complex_field);
SM_API int sm_dev_D(sm_ctx* ctx, const double* d_U, const double* d_in, double* d_out, double m0, int dagger);
SM_API int sm_dev_DDdag(sm_ctx* ctx, const double* d_U, const double* d_in, double* d_out, double m0);
SM_API int sm_dev_dot(sm_ctx* ctx, const double* d_x, const double* d_y, double out_re_im[2]);
SM_API int sm_dev_cg(sm_ctx* ctx, const double* d_U, const double* d_phi, double* d_x, double m0, int* converged,
              int* iterations);
/* `reps` back-to-back D D^dagger applications (in -> out), one host synchronisation: for benches */
SM_API int sm_dev_DDdag_loop(sm_ctx* ctx, const double* d_U, const double* d_in, double* d_out, double m0, int reps,
                      double* ms_total);

/* ---- device-resident HMC (src/hmc.cpp) --------------------------------------------------------
 * The context owns U, U', pi, pi', F, chi, phi and the CG work vectors; one trajectory is
 * RandomPI/RandomCHI (or injected fields) -> phi = D chi -> Leapfrog -> dH; only scalars return. */
typedef struct {
    double beta, m0, trajectory_length;
    int md_steps;
} sm_hmc_params;

typedef struct {
    double dH, H_old, H_new;          /* Hamiltonian(U',pi') - Hamiltonian(U,pi)  (hmc.cpp:162) */
    double sum_re_plaq_new, gauge_action_new;   /* MeasureSp_HMC / Compute_gaugeAction of U' */
    double sum_re_plaq_old, gauge_action_old;   /* ... of U */
    long long dd_applications;        /* D D^dagger applications in all CG solves of the trajectory */
    int cg_solves, cg_all_converged;
    double kernel_ms;                 /* device time of the trajectory */
    int cg_force_failures;            /* CG solves inside HMC::Force that did not converge (hmc.cpp:46-56: one illConf dump each) */
    int reserved_;
} sm_traj_result;

SM_API int sm_hmc_configure(sm_ctx* ctx, const sm_hmc_params* p);
SM_API int sm_hmc_set_gauge(sm_ctx* ctx, const double* h_U0, const double* h_U1);          /* upload U */
SM_API int sm_hmc_get_gauge(sm_ctx* ctx, double* h_U0, double* h_U1, int proposal);        /* download U (0) or U' (1) */
SM_API int sm_hmc_get_momenta(sm_ctx* ctx, double* h_pi0, double* h_pi1, int proposal);
SM_API int sm_hmc_get_phi(sm_ctx* ctx, double* h_phi0, double* h_phi1);
/* the Gaussian chi of HMC::RandomCHI (hmc.cpp:19-28) as refreshed or injected (statistical tests of the generator) */
SM_API int sm_hmc_get_chi(sm_ctx* ctx, double* h_chi0, double* h_chi1);
/* HMC::RandomPI + HMC::RandomCHI (hmc.cpp:5-28) from a counter-based device generator */
SM_API int sm_hmc_refresh(sm_ctx* ctx, uint64_t seed, uint64_t trajectory_index);
/* same fields supplied by the caller (parity tests: identical pi, chi on both sides) */
SM_API int sm_hmc_inject(sm_ctx* ctx, const double* h_pi0, const double* h_pi1, const double* h_chi0, const double* h_chi1);
/* phi = D chi; Leapfrog; dH (hmc.cpp:160-162).  U is untouched, the proposal stays in U'. */
SM_API int sm_hmc_trajectory(sm_ctx* ctx, sm_traj_result* out);
/* Metropolis outcome decided by the host (hmc.cpp:166-177): accept!=0 makes U' the current U */
SM_API int sm_hmc_accept(sm_ctx* ctx, int accept);
/* HMC::Force on the current U and phi (hmc.cpp:44-60), result to host: for parity tests */
SM_API int sm_hmc_force(sm_ctx* ctx, const double* h_phi0, const double* h_phi1, double* h_F0, double* h_F1, int* converged);
/* HMC::Hamiltonian(U, pi, phi) with host-supplied pi and phi on the current U (hmc.cpp:135-149) */
SM_API int sm_hmc_hamiltonian(sm_ctx* ctx, const double* h_pi0, const double* h_pi1, const double* h_phi0,
                       const double* h_phi1, double* H);
/* HMC::Leapfrog from host-supplied pi, phi on the current U; results stay in U', pi' */
SM_API int sm_hmc_leapfrog(sm_ctx* ctx, const double* h_pi0, const double* h_pi1, const double* h_phi0,
                    const double* h_phi1, int* all_converged);

/* ---- binary configuration files ---------------------------------------------------------------
 * SaveConf (src/gauge_conf.cpp:378-423) and GaugeConf::readBinary (:495-546): headerless 28-byte
 * records (int32 x, int32 t, int32 mu, double re, double im), loops x -> t -> mu, GLOBAL lattice.
 * Host-side helpers over host buffers of the global field (no device work). */
SM_API int sm_save_conf(int Nx, int Nt, const double* h_U0, const double* h_U1, const char* path);
SM_API int sm_read_conf(int Nx, int Nt, const char* path, double* h_U0, double* h_U1);

#ifdef __cplusplus
}
#endif
#endif
