// C-ABI harness around the UNMODIFIED reference sources (TEST INFRASTRUCTURE ONLY).
//
// Built by oracle/ref_build/build_ref.sh together with
//   /root/reference/src/{statistics,variables,gauge_conf,dirac_operator,conjugate_gradient,hmc}.cpp
// (compiled where they lie; nothing is copied) into oracle/_ref/libref_<NS>x<NT>.so.
// The lattice size is fixed per library, exactly like the reference executable
// (CMakeLists.txt:17-20): -DCONFIG_H -DNS=.. -DNT=.. bypasses include/config.h.
//
// Field convention on this ABI: a complex field is `double[2][V][2]` = (mu, site, re/im),
// i.e. the reference's spinor{mu0,mu1} concatenated; a real field is `double[2][V]`;
// sites are n = x*Nt + t (variables.cpp:10-12).
//
// Private members of class HMC (Leapfrog, Hamiltonian, PConf, chi ...) are reached with the
// `#define private public` trick, which does not change the class layout.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "mpi.h"

#define private public
#include "hmc.h"
#undef private
#include "mpi_setup.h"

namespace {

int g_cfg_rx = 0, g_cfg_rt = 0, g_cfg_np = 0;

// Mirrors main.cpp:72-74 for the given decomposition (this process = minimpi::g_rank).
void setup(int ranks_x, int ranks_t) {
    int np = ranks_x * ranks_t;
    if (g_cfg_rx == ranks_x && g_cfg_rt == ranks_t && g_cfg_np == np && mpi::rank == minimpi::g_rank) return;
    if (LeftPB != nullptr) free_lattice_arrays();
    mpi::size = np;
    mpi::rank = minimpi::g_rank;
    mpi::ranks_x = ranks_x;
    mpi::ranks_t = ranks_t;
    initializeMPI();
    allocate_lattice_arrays();
    periodic_boundary();
    g_cfg_rx = ranks_x;
    g_cfg_rt = ranks_t;
    g_cfg_np = np;
}

double now() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// global (mu, n, re/im) array  ->  this rank's tile as a reference `spinor`
void load_c(spinor& s, const double* g) {
    const long V = LV::Ntot;
    const int wx = mpi::width_x, wt = mpi::width_t;
    const int cx = mpi::rank / mpi::ranks_t, ct = mpi::rank % mpi::ranks_t;
    for (int x = 0; x < wx; x++)
        for (int t = 0; t < wt; t++) {
            long gn = (long)(cx * wx + x) * LV::Nt + (ct * wt + t);
            int n = x * wt + t;
            s.mu0[n] = c_double(g[2 * gn], g[2 * gn + 1]);
            s.mu1[n] = c_double(g[2 * (V + gn)], g[2 * (V + gn) + 1]);
        }
}

void store_c(const spinor& s, double* g) {
    const long V = LV::Ntot;
    const int wx = mpi::width_x, wt = mpi::width_t;
    const int cx = mpi::rank / mpi::ranks_t, ct = mpi::rank % mpi::ranks_t;
    for (int x = 0; x < wx; x++)
        for (int t = 0; t < wt; t++) {
            long gn = (long)(cx * wx + x) * LV::Nt + (ct * wt + t);
            int n = x * wt + t;
            g[2 * gn] = s.mu0[n].real();
            g[2 * gn + 1] = s.mu0[n].imag();
            g[2 * (V + gn)] = s.mu1[n].real();
            g[2 * (V + gn) + 1] = s.mu1[n].imag();
        }
}

void load_r(re_field& f, const double* g) {
    const long V = LV::Ntot;
    const int wx = mpi::width_x, wt = mpi::width_t;
    const int cx = mpi::rank / mpi::ranks_t, ct = mpi::rank % mpi::ranks_t;
    for (int x = 0; x < wx; x++)
        for (int t = 0; t < wt; t++) {
            long gn = (long)(cx * wx + x) * LV::Nt + (ct * wt + t);
            int n = x * wt + t;
            f.mu0[n] = g[gn];
            f.mu1[n] = g[V + gn];
        }
}

void store_r(const re_field& f, double* g) {
    const long V = LV::Ntot;
    const int wx = mpi::width_x, wt = mpi::width_t;
    const int cx = mpi::rank / mpi::ranks_t, ct = mpi::rank % mpi::ranks_t;
    for (int x = 0; x < wx; x++)
        for (int t = 0; t < wt; t++) {
            long gn = (long)(cx * wx + x) * LV::Nt + (ct * wt + t);
            int n = x * wt + t;
            g[gn] = f.mu0[n];
            g[V + gn] = f.mu1[n];
        }
}

HMC make_hmc(GaugeConf& G, int md, double tau, double beta, double m0) {
    return HMC(G, md, tau, /*Ntherm*/ 0, /*Nmeas*/ 0, /*Nsteps*/ 0, beta, LV::Nx, LV::Nt, m0, /*saveconf*/ 0);
}

void* shared_alloc(size_t bytes) {
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) { std::perror("mmap"); std::exit(2); }
    return p;
}

}  // namespace

extern "C" {

int ref_nx() { return LV::Nx; }
int ref_nt() { return LV::Nt; }

// ---- geometry (dirac_operator.h:35-62) -------------------------------------------------
// Tables of rank `rank` in a ranks_x x ranks_t decomposition (local-wrap indices, world-rank
// keyed antiperiodic signs).  Sizes: RightPB/LeftPB 2*maxSize ints, SignR/SignL 2*maxSize
// complex (as re,im pairs), x_1_t1/x1_t_1 maxSize ints.
void ref_tables(int ranks_x, int ranks_t, int rank, int* rpb, int* lpb, double* sr, double* sl, int* xm1tp1,
                int* xp1tm1) {
    int keep = minimpi::g_rank;
    minimpi::g_rank = rank;
    g_cfg_np = 0;   // force re-setup
    setup(ranks_x, ranks_t);
    int m = mpi::maxSize;
    std::memcpy(rpb, RightPB, sizeof(int) * 2 * m);
    std::memcpy(lpb, LeftPB, sizeof(int) * 2 * m);
    std::memcpy(sr, SignR, sizeof(double) * 4 * m);
    std::memcpy(sl, SignL, sizeof(double) * 4 * m);
    std::memcpy(xm1tp1, x_1_t1, sizeof(int) * m);
    std::memcpy(xp1tm1, x1_t_1, sizeof(int) * m);
    minimpi::g_rank = keep;
    g_cfg_np = 0;
}

// ---- hot start (gauge_conf.cpp:23-36) with a fixed srand seed ---------------------------
void ref_hot_start(unsigned seed, double* U) {
    setup(1, 1);
    srand(seed);
    GaugeConf G;
    G.initialization();
    store_c(G.Conf, U);
}

// ---- operators (dirac_operator.cpp) ------------------------------------------------------
void ref_D(const double* U, const double* phi, double* out, double m0, int dagger) {
    setup(1, 1);
    spinor u, p, o;
    load_c(u, U);
    load_c(p, phi);
    if (dagger) D_dagger_phi(u, p, o, m0);
    else D_phi(u, p, o, m0);
    store_c(o, out);
}

void ref_DDdag(const double* U, const double* phi, double* out, double m0) {
    setup(1, 1);
    spinor u, p, o;
    load_c(u, U);
    load_c(p, phi);
    D_D_dagger_phi(u, p, o, m0);
    store_c(o, out);
}

void ref_dot(const double* x, const double* y, double* out2) {
    setup(1, 1);
    spinor a, b;
    load_c(a, x);
    load_c(b, y);
    c_double z = dot(a, b);
    out2[0] = z.real();
    out2[1] = z.imag();
}

// conjugate_gradient.cpp:4-67.  *dd_apps = number of D_D_dagger_phi applications, recovered
// from the number of MPI_Allreduce calls (2 before the loop + 2 per iteration).
int ref_cg(const double* U, const double* phi, double* x, double m0, double tol, int max_iter, int* dd_apps,
           double* seconds) {
    setup(1, 1);
    spinor u, p, o;
    load_c(u, U);
    load_c(p, phi);
    double tol0 = CG::tol;
    int mi0 = CG::max_iter;
    CG::tol = tol;
    CG::max_iter = max_iter;
    long a0 = minimpi::g_allreduce_calls;
    double t0 = now();
    int ok = conjugate_gradient(u, p, o, m0);
    double t1 = now();
    long calls = minimpi::g_allreduce_calls - a0;
    if (dd_apps) *dd_apps = (int)(1 + (calls - 2) / 2);
    if (seconds) *seconds = t1 - t0;
    CG::tol = tol0;
    CG::max_iter = mi0;
    store_c(o, x);
    return ok;
}

void ref_fermion_force(const double* U, const double* left, const double* right, double* F) {
    setup(1, 1);
    spinor u, l, r;
    load_c(u, U);
    load_c(l, left);
    load_c(r, right);
    re_field f = phi_dag_partialD_phi(u, l, r);
    store_r(f, F);
}

// ---- gauge observables (gauge_conf.cpp) --------------------------------------------------
void ref_staple(const double* U, double* K) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    G.Compute_Staple();
    store_c(G.Staples, K);
}

// P: double[V][2]; sums[0] = MeasureSp_HMC, sums[1] = Compute_gaugeAction(beta)
void ref_plaquette(const double* U, double beta, double* P, double* sums) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    G.Compute_Plaquette01();
    for (long n = 0; n < LV::Ntot; n++) {
        P[2 * n] = G.Plaquette01[n].real();
        P[2 * n + 1] = G.Plaquette01[n].imag();
    }
    sums[0] = G.MeasureSp_HMC();
    sums[1] = G.Compute_gaugeAction(beta);
}

// ---- HMC internals (hmc.cpp) ---------------------------------------------------------------
// HMC::Force (hmc.cpp:44-60): CG -> D^dagger psi -> fermion force -> + gauge force
int ref_force(const double* U, const double* phi, double beta, double m0, double* F) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    spinor p;
    load_c(p, phi);
    HMC h = make_hmc(G, 2, 1.0, beta, m0);
    h.Force(h.GConf, p);
    store_r(h.Forces, F);
    return h.CG_convergence;
}

double ref_action(const double* U, const double* phi, double beta, double m0) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    spinor p;
    load_c(p, phi);
    HMC h = make_hmc(G, 2, 1.0, beta, m0);
    return h.Action(h.GConf, p);
}

double ref_hamiltonian(const double* U, const double* pi, const double* phi, double beta, double m0) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    spinor p;
    load_c(p, phi);
    re_field P;
    load_r(P, pi);
    HMC h = make_hmc(G, 2, 1.0, beta, m0);
    return h.Hamiltonian(h.GConf, P, p);
}

// HMC::Leapfrog (hmc.cpp:63-103) from injected (U, pi) with pseudofermion phi
void ref_leapfrog(const double* U, const double* pi, const double* phi, int md, double tau, double beta, double m0,
                  double* U_out, double* pi_out) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    spinor p;
    load_c(p, phi);
    HMC h = make_hmc(G, md, tau, beta, m0);
    load_r(h.PConf, pi);
    h.Leapfrog(p);
    store_c(h.GConf_copy.Conf, U_out);
    store_r(h.PConf_copy, pi_out);
}

// One HMC_Update (hmc.cpp:151-181) with injected pi and chi instead of the irreproducible
// generators; Metropolis is left to the caller.  out: phi = D chi, evolved (U', pi'),
// H[0] = Hamiltonian(U,pi), H[1] = Hamiltonian(U',pi'), aux[0] = sum Re P of U' (MeasureSp_HMC),
// aux[1] = gauge action of U'.
int ref_trajectory(const double* U, const double* pi, const double* chi, int md, double tau, double beta, double m0,
                   double tol, double* phi_out, double* U_out, double* pi_out, double* H, double* aux,
                   double* seconds) {
    setup(1, 1);
    double tol0 = CG::tol;
    CG::tol = tol;
    GaugeConf G;
    load_c(G.Conf, U);
    HMC h = make_hmc(G, md, tau, beta, m0);
    load_r(h.PConf, pi);
    load_c(h.chi, chi);
    double t0 = now();
    spinor phi(mpi::maxSize);
    D_phi(h.GConf.Conf, h.chi, phi, m0);
    h.Leapfrog(phi);
    double Hn = h.Hamiltonian(h.GConf_copy, h.PConf_copy, phi);
    double Ho = h.Hamiltonian(h.GConf, h.PConf, phi);
    double t1 = now();
    if (seconds) *seconds = t1 - t0;
    H[0] = Ho;
    H[1] = Hn;
    aux[0] = h.GConf_copy.MeasureSp_HMC();
    aux[1] = h.GConf_copy.Compute_gaugeAction(beta);
    if (phi_out) store_c(phi, phi_out);
    store_c(h.GConf_copy.Conf, U_out);
    store_r(h.PConf_copy, pi_out);
    CG::tol = tol0;
    return h.CG_convergence;
}

// ---- configuration files (gauge_conf.cpp:378-423, 495-546) -------------------------------
void ref_save_conf(const double* U, const char* name) {
    setup(1, 1);
    GaugeConf G;
    load_c(G.Conf, U);
    SaveConf(G, name);
}

void ref_read_binary(const char* name, double* U) {
    setup(1, 1);
    GaugeConf G;
    G.readBinary(name);
    store_c(G.Conf, U);
}

// format() of variables.h:197-203 (file-name mangling of beta and m0)
void ref_format(double v, char* out, int cap) {
    std::string s = format(v);
    std::strncpy(out, s.c_str(), cap - 1);
    out[cap - 1] = 0;
}

double ref_jackknife(const double* dat, int n, int bins) {
    std::vector<double> v(dat, dat + n);
    return Jackknife_error(v, bins);
}

// ---- timing on host cores (the CPU baseline) ------------------------------------------------
// Runs the reference on ranks_x*ranks_t forked ranks (1x1 = in-process, serial branches).
//   op 0: `reps` applications of D_D_dagger_phi      -> seconds for all reps
//   op 1: one conjugate_gradient solve                -> seconds, *iters = DD^dagger applications
//   op 2: `reps` x { D_D_dagger_phi, dot, axpy-class CG updates } = reps CG iterations worth of
//         work without the convergence logic (bounded sample of a long solve)
// `out` (may be NULL) receives the last result in global layout.  Returns 0 on success.
int ref_timed(int ranks_x, int ranks_t, int op, const double* U, const double* phi, double m0, int reps, double tol,
              int max_iter, double* out, double* seconds, int* iters) {
    const int np = ranks_x * ranks_t;
    const long V = LV::Ntot;
    double* shared_out = nullptr;
    double* shared_scal = (double*)shared_alloc(4096);
    if (out && np > 1) shared_out = (double*)shared_alloc(sizeof(double) * 4 * V);
    if (np > 1) {
        long maxmsg = 16L * std::max(LV::Nx / ranks_x, LV::Nt / ranks_t);
        minimpi::spawn(np, std::max(maxmsg, 4096L));
    }
    g_cfg_np = 0;
    setup(ranks_x, ranks_t);
    {
        spinor u(mpi::maxSize), p(mpi::maxSize), o(mpi::maxSize);
        load_c(u, U);
        load_c(p, phi);
        double tol0 = CG::tol;
        int mi0 = CG::max_iter;
        CG::tol = tol;
        CG::max_iter = max_iter;
        int its = 0;
        minimpi::barrier();
        long a0 = minimpi::g_allreduce_calls;
        double t0 = now();
        if (op == 0) {
            for (int r = 0; r < reps; r++) D_D_dagger_phi(u, p, o, m0);
            its = reps;
        } else if (op == 1) {
            conjugate_gradient(u, p, o, m0);
            its = (int)(1 + (minimpi::g_allreduce_calls - a0 - 2) / 2);
        } else {
            // the loop body of conjugate_gradient.cpp:31-62 with alpha, beta frozen
            spinor x(p), r(p), d(p);
            c_double alpha(1e-3, 0), beta(0.5, 0);
            for (int k = 0; k < reps; k++) {
                D_D_dagger_phi(u, d, o, m0);
                c_double dAd = dot(d, o);
                for (int n = 0; n < mpi::maxSize; n++) {
                    x.mu0[n] += alpha * d.mu0[n];
                    x.mu1[n] += alpha * d.mu1[n];
                    r.mu0[n] -= alpha * o.mu0[n];
                    r.mu1[n] -= alpha * o.mu1[n];
                }
                c_double rr = dot(r, r);
                for (int n = 0; n < mpi::maxSize; n++) {
                    d.mu0[n] *= beta;
                    d.mu1[n] *= beta;
                    d.mu0[n] += r.mu0[n];
                    d.mu1[n] += r.mu1[n];
                }
                shared_scal[8] = dAd.real() + rr.real();
            }
            its = reps;
        }
        minimpi::barrier();
        double t1 = now();
        CG::tol = tol0;
        CG::max_iter = mi0;
        if (mpi::rank == 0) {
            shared_scal[0] = t1 - t0;
            shared_scal[1] = its;
        }
        if (out) store_c(o, np > 1 ? shared_out : out);
        minimpi::barrier();
    }
    if (np > 1) {
        if (minimpi::g_rank != 0) minimpi::child_exit();
        int bad = minimpi::join();
        if (out) std::memcpy(out, shared_out, sizeof(double) * 4 * V);
        if (shared_out) munmap(shared_out, sizeof(double) * 4 * V);
        if (bad) return 1;
    }
    *seconds = shared_scal[0];
    if (iters) *iters = (int)shared_scal[1];
    munmap(shared_scal, 4096);
    g_cfg_np = 0;
    return 0;
}

}  // extern "C"
