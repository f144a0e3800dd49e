/*
 * schwinger_oracle.c -- CPU restatement (plain C99) of the reference's HMC fermion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for schwingermodel_b200: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.  The product
 * never calls it and has no CPU fallback.
 *
 * PARITY PIN: every function here is checked against the UNMODIFIED reference compiled from
 * /root/reference into oracle/_ref (tests/test_oracle_vs_reference.py) and against the golden
 * vectors that build produced (tests/golden, made by tests/golden/make_golden.py).  The
 * reference itself ships no tests or known-answer vectors.
 *
 * Each function cites the reference file:line it restates.  Unlike the reference, the lattice
 * size is a run-time argument, so one library serves every size.
 *
 * Conventions (reference: src/variables.cpp:10-12, include/variables.h:54-141):
 *   site n = x*Nt + t (t fastest); mu=0 is time, mu=1 is space.
 *   complex field  = double[2][V][2]  (mu | spin component, site, re/im)
 *   real field     = double[2][V]
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cd;

static inline int wrap(int a, int b) { int r = a % b; return r < 0 ? r + b : r; }   /* variables.h:171-174 */

/* ------------------------------------------------------------------------------------------
 * Geometry: include/dirac_operator.h:35-62 (periodic_boundary), per-rank local-wrap tables.
 * rank -> antiperiodic seam keyed on the WORLD rank exactly as the reference does.
 * ---------------------------------------------------------------------------------------- */
void so_tables(int Nx, int Nt, int ranks_x, int ranks_t, int rank, int* RightPB, int* LeftPB, double* SignR,
               double* SignL, int* x_1_t1, int* x1_t_1) {
    const int wx = Nx / ranks_x, wt = Nt / ranks_t;
    for (int x = 0; x < wx; x++)
        for (int t = 0; t < wt; t++) {
            const int n = x * wt + t;
            x_1_t1[n] = wrap(x - 1, wx) * wt + wrap(t + 1, wt);
            x1_t_1[n] = wrap(x + 1, wx) * wt + wrap(t - 1, wt);
            for (int mu = 0; mu < 2; mu++) {
                const int dt = (mu == 0), dx = (mu == 1);
                RightPB[2 * n + mu] = wrap(x + dx, wx) * wt + wrap(t + dt, wt);
                LeftPB[2 * n + mu] = wrap(x - dx, wx) * wt + wrap(t - dt, wt);
                double sr = 1.0, sl = 1.0;
                if ((rank + 1) % ranks_t == 0 && mu == 0 && t == wt - 1) sr = -1.0;
                if (rank % ranks_t == 0 && mu == 0 && t == 0) sl = -1.0;
                SignR[2 * (2 * n + mu)] = sr;
                SignR[2 * (2 * n + mu) + 1] = 0.0;
                SignL[2 * (2 * n + mu)] = sl;
                SignL[2 * (2 * n + mu) + 1] = 0.0;
            }
        }
}

typedef struct {
    int Nx, Nt, V;
    int *R, *L;        /* RightPB / LeftPB, [2n+mu] */
    double *sR, *sL;   /* time-direction signs per site (mu=1 signs are always +1) */
} geom;

static geom make_geom(int Nx, int Nt) {
    geom g;
    g.Nx = Nx; g.Nt = Nt; g.V = Nx * Nt;
    g.R = (int*)malloc(sizeof(int) * 2 * g.V);
    g.L = (int*)malloc(sizeof(int) * 2 * g.V);
    g.sR = (double*)malloc(sizeof(double) * g.V);
    g.sL = (double*)malloc(sizeof(double) * g.V);
    for (int x = 0; x < Nx; x++)
        for (int t = 0; t < Nt; t++) {
            int n = x * Nt + t;
            g.R[2 * n] = x * Nt + wrap(t + 1, Nt);
            g.R[2 * n + 1] = wrap(x + 1, Nx) * Nt + t;
            g.L[2 * n] = x * Nt + wrap(t - 1, Nt);
            g.L[2 * n + 1] = wrap(x - 1, Nx) * Nt + t;
            g.sR[n] = (t == Nt - 1) ? -1.0 : 1.0;
            g.sL[n] = (t == 0) ? -1.0 : 1.0;
        }
    return g;
}

static void free_geom(geom* g) { free(g->R); free(g->L); free(g->sR); free(g->sL); }

/* ------------------------------------------------------------------------------------------
 * Hot start: src/gauge_conf.cpp:23-36 with the truncated pi of src/variables.cpp:3.
 * ---------------------------------------------------------------------------------------- */
void so_hot_start(unsigned seed, int Nx, int Nt, double* U) {
    const int V = Nx * Nt;
    const double pi_ref = 3.14159265359;
    cd* u0 = (cd*)U;
    cd* u1 = u0 + V;
    srand(seed);
    for (int n = 0; n < V; n++) {
        double th0 = 2.0 * pi_ref * ((double)rand() / (RAND_MAX));
        u0[n] = cos(th0) + I * sin(th0);
        double th1 = 2.0 * pi_ref * ((double)rand() / (RAND_MAX));
        u1[n] = cos(th1) + I * sin(th1);
    }
}

/* ------------------------------------------------------------------------------------------
 * D and D^dagger: src/dirac_operator.cpp:29-44 and :253-268 (serial branches).
 * The expression shape (products associated left to right, real scalings, i as a complex
 * constant) mirrors the reference so the results are bit-identical to it.
 * ---------------------------------------------------------------------------------------- */
static void apply_D(const geom* g, const cd* U0, const cd* U1, const cd* p0, const cd* p1, cd* o0, cd* o1, double m0,
                    int dagger) {
    const cd i_ = I;
    for (int n = 0; n < g->V; n++) {
        const int r0 = g->R[2 * n], r1 = g->R[2 * n + 1], l0 = g->L[2 * n], l1 = g->L[2 * n + 1];
        const cd sR0 = g->sR[n], sL0 = g->sL[n], one = 1.0;
        if (!dagger) {
            o0[n] = (m0 + 2) * p0[n] - 0.5 * (U0[n] * sR0 * (p0[r0] - p1[r0]) + U1[n] * one * (p0[r1] + i_ * p1[r1]) +
                                              conj(U0[l0]) * sL0 * (p0[l0] + p1[l0]) +
                                              conj(U1[l1]) * one * (p0[l1] - i_ * p1[l1]));
            o1[n] = (m0 + 2) * p1[n] - 0.5 * (U0[n] * sR0 * (-p0[r0] + p1[r0]) + U1[n] * one * (-i_ * p0[r1] + p1[r1]) +
                                              conj(U0[l0]) * sL0 * (p0[l0] + p1[l0]) +
                                              conj(U1[l1]) * one * (i_ * p0[l1] + p1[l1]));
        } else {
            o0[n] = (m0 + 2) * p0[n] - 0.5 * (conj(U0[l0]) * sL0 * (p0[l0] - p1[l0]) +
                                              conj(U1[l1]) * one * (p0[l1] + i_ * p1[l1]) +
                                              U0[n] * sR0 * (p0[r0] + p1[r0]) + U1[n] * one * (p0[r1] - i_ * p1[r1]));
            o1[n] = (m0 + 2) * p1[n] - 0.5 * (conj(U0[l0]) * sL0 * (-p0[l0] + p1[l0]) +
                                              conj(U1[l1]) * one * (-i_ * p0[l1] + p1[l1]) +
                                              U0[n] * sR0 * (p0[r0] + p1[r0]) + U1[n] * one * (i_ * p0[r1] + p1[r1]));
        }
    }
}

void so_D(int Nx, int Nt, const double* U, const double* phi, double* out, double m0, int dagger) {
    geom g = make_geom(Nx, Nt);
    const int V = g.V;
    apply_D(&g, (const cd*)U, (const cd*)U + V, (const cd*)phi, (const cd*)phi + V, (cd*)out, (cd*)out + V, m0, dagger);
    free_geom(&g);
}

/* D D^dagger: src/dirac_operator.cpp:477-480 (D^dagger first, then D) */
static void apply_DDdag(const geom* g, const cd* U, const cd* in, cd* out, cd* tmp, double m0) {
    const int V = g->V;
    apply_D(g, U, U + V, in, in + V, tmp, tmp + V, m0, 1);
    apply_D(g, U, U + V, tmp, tmp + V, out, out + V, m0, 0);
}

void so_DDdag(int Nx, int Nt, const double* U, const double* phi, double* out, double m0) {
    geom g = make_geom(Nx, Nt);
    cd* tmp = (cd*)malloc(sizeof(cd) * 2 * g.V);
    apply_DDdag(&g, (const cd*)U, (const cd*)phi, (cd*)out, tmp, m0);
    free(tmp);
    free_geom(&g);
}

/* dot: include/variables.h:181-192 -- sum_n x conj(y), mu0 then mu1 per site, sequential */
static cd dot_c(int V, const cd* x, const cd* y) {
    cd z = 0;
    for (int n = 0; n < V; n++) {
        z += x[n] * conj(y[n]);
        z += x[V + n] * conj(y[V + n]);
    }
    return z;
}

void so_dot(int Nx, int Nt, const double* x, const double* y, double* out2) {
    cd z = dot_c(Nx * Nt, (const cd*)x, (const cd*)y);
    out2[0] = creal(z);
    out2[1] = cimag(z);
}

/* ------------------------------------------------------------------------------------------
 * Conjugate gradient on D D^dagger: src/conjugate_gradient.cpp:4-67.
 * x0 = phi, complex alpha/beta, recursive residual, stop when ||r|| < tol*||phi||.
 * returns 1 (converged) / 0; *dd_apps = number of D D^dagger applications.
 * ---------------------------------------------------------------------------------------- */
static int cg_solve(const geom* g, const cd* U, const cd* phi, cd* x, double m0, double tol, int max_iter, int* dd_apps) {
    const int V = g->V, N = 2 * V;
    cd* r = (cd*)calloc(N, sizeof(cd));
    cd* d = (cd*)calloc(N, sizeof(cd));
    cd* Ad = (cd*)calloc(N, sizeof(cd));
    cd* tmp = (cd*)calloc(N, sizeof(cd));
    int k = 0, apps = 0, ok = 0;
    memcpy(x, phi, sizeof(cd) * N);
    apply_DDdag(g, U, x, Ad, tmp, m0);
    apps++;
    for (int n = 0; n < N; n++) r[n] = phi[n] - Ad[n];
    memcpy(d, r, sizeof(cd) * N);
    cd r_norm2 = dot_c(V, r, r);
    double phi_norm = sqrt(creal(dot_c(V, phi, phi)));
    while (k < max_iter) {
        apply_DDdag(g, U, d, Ad, tmp, m0);
        apps++;
        cd alpha = r_norm2 / dot_c(V, d, Ad);
        for (int n = 0; n < V; n++) {           /* same site-major order as the reference loop */
            x[n] += alpha * d[n];
            x[V + n] += alpha * d[V + n];
            r[n] -= alpha * Ad[n];
            r[V + n] -= alpha * Ad[V + n];
        }
        double err_sqr = creal(dot_c(V, r, r));
        double err = sqrt(err_sqr);
        if (err < tol * phi_norm) { ok = 1; break; }
        cd beta = err_sqr / r_norm2;
        for (int n = 0; n < N; n++) {
            d[n] *= beta;
            d[n] += r[n];
        }
        r_norm2 = err_sqr;
        k++;
    }
    if (dd_apps) *dd_apps = apps;
    free(r); free(d); free(Ad); free(tmp);
    return ok;
}

int so_cg(int Nx, int Nt, const double* U, const double* phi, double* x, double m0, double tol, int max_iter,
          int* dd_apps) {
    geom g = make_geom(Nx, Nt);
    int ok = cg_solve(&g, (const cd*)U, (const cd*)phi, (cd*)x, m0, tol, max_iter, dd_apps);
    free_geom(&g);
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * Fermion-force derivative: src/dirac_operator.cpp:493-507 (eq. 37-38), forward hops only.
 * ---------------------------------------------------------------------------------------- */
static void fermion_force(const geom* g, const cd* U, const cd* left, const cd* right, double* F) {
    const int V = g->V;
    const cd *U0 = U, *U1 = U + V, *l0 = left, *l1 = left + V, *q0 = right, *q1 = right + V;
    const cd i_ = I;
    for (int n = 0; n < V; n++) {
        const int rt = g->R[2 * n], rx = g->R[2 * n + 1];
        const cd s = g->sR[n], one = 1.0;
        F[n] = cimag(U0[n] * s * (conj(l0[n] - l1[n])) * (q0[rt] - q1[rt]) -
                     conj(U0[n]) * s * (conj(l0[rt] + l1[rt])) * (q0[n] + q1[n]));
        F[V + n] = cimag(U1[n] * one * (conj(l0[n]) - i_ * conj(l1[n])) * (q0[rx] + i_ * q1[rx]) +
                         conj(U1[n]) * one * (conj(l0[rx]) + i_ * conj(l1[rx])) * (-q0[n] + i_ * q1[n]));
    }
}

void so_fermion_force(int Nx, int Nt, const double* U, const double* left, const double* right, double* F) {
    geom g = make_geom(Nx, Nt);
    fermion_force(&g, (const cd*)U, (const cd*)left, (const cd*)right, F);
    free_geom(&g);
}

/* ------------------------------------------------------------------------------------------
 * Staples: src/gauge_conf.cpp:96-127; plaquette :44-48; sums :427-449.
 * ---------------------------------------------------------------------------------------- */
static void staples(const geom* g, const cd* U, cd* K) {
    const int V = g->V, Nx = g->Nx, Nt = g->Nt;
    const cd *U0 = U, *U1 = U + V;
    for (int n = 0; n < V; n++) {
        const int x = n / Nt, t = n % Nt;
        const int x1 = g->R[2 * n + 1], xm = g->L[2 * n + 1], t1 = g->R[2 * n], tm = g->L[2 * n];
        const int xm_t1 = wrap(x - 1, Nx) * Nt + wrap(t + 1, Nt);
        const int x1_tm = wrap(x + 1, Nx) * Nt + wrap(t - 1, Nt);
        K[n] = U1[n] * U0[x1] * conj(U1[t1]) + conj(U1[xm]) * U0[xm] * U1[xm_t1];
        K[V + n] = U0[n] * U1[t1] * conj(U0[x1]) + conj(U0[tm]) * U1[tm] * U0[x1_tm];
    }
}

void so_staple(int Nx, int Nt, const double* U, double* K) {
    geom g = make_geom(Nx, Nt);
    staples(&g, (const cd*)U, (cd*)K);
    free_geom(&g);
}

static void plaquettes(const geom* g, const cd* U, cd* P) {
    const int V = g->V;
    const cd *U0 = U, *U1 = U + V;
    for (int n = 0; n < V; n++) P[n] = U0[n] * U1[g->R[2 * n]] * conj(U0[g->R[2 * n + 1]]) * conj(U1[n]);
}

/* sums[0] = sum Re P (MeasureSp_HMC), sums[1] = beta * sum Re(1-P) (Compute_gaugeAction) */
void so_plaquette(int Nx, int Nt, const double* U, double beta, double* P, double* sums) {
    geom g = make_geom(Nx, Nt);
    cd* p = (cd*)P;
    plaquettes(&g, (const cd*)U, p);
    double sp = 0.0, sg = 0.0;
    for (int n = 0; n < g.V; n++) sp += creal(p[n]);
    for (int n = 0; n < g.V; n++) sg += beta * creal(1.0 - p[n]);
    sums[0] = sp;
    sums[1] = sg;
    free_geom(&g);
}

/* ------------------------------------------------------------------------------------------
 * HMC pieces: src/hmc.cpp.
 * ---------------------------------------------------------------------------------------- */
/* HMC::Force (hmc.cpp:44-60) + Force_G (:32-40): F = fermion term, then += -beta Im(U conj K) */
static int total_force(const geom* g, const cd* U, const cd* phi, double beta, double m0, double tol, int max_iter,
                       double* F, int* its_acc) {
    const int V = g->V, N = 2 * V;
    cd* psi = (cd*)calloc(N, sizeof(cd));
    cd* chi = (cd*)calloc(N, sizeof(cd));
    cd* K = (cd*)calloc(N, sizeof(cd));
    int apps = 0;
    int ok = cg_solve(g, U, phi, psi, m0, tol, max_iter, &apps);
    if (its_acc) *its_acc += apps;
    apply_D(g, U, U + V, psi, psi + V, chi, chi + V, m0, 1);
    fermion_force(g, U, psi, chi, F);
    staples(g, U, K);
    for (int n = 0; n < V; n++) {
        F[n] += -beta * cimag(U[n] * conj(K[n]));
        F[V + n] += -beta * cimag(U[V + n] * conj(K[V + n]));
    }
    free(psi); free(chi); free(K);
    return ok;
}

int so_force(int Nx, int Nt, const double* U, const double* phi, double beta, double m0, double tol, int max_iter,
             double* F) {
    geom g = make_geom(Nx, Nt);
    int ok = total_force(&g, (const cd*)U, (const cd*)phi, beta, m0, tol, max_iter, F, NULL);
    free_geom(&g);
    return ok;
}

/* HMC::Action (hmc.cpp:105-133) */
static double action(const geom* g, const cd* U, const cd* phi, double beta, double m0, double tol, int max_iter,
                     int* ok, int* its_acc, double* plaq_sums) {
    const int V = g->V, N = 2 * V;
    cd* P = (cd*)calloc(V, sizeof(cd));
    cd* x = (cd*)calloc(N, sizeof(cd));
    plaquettes(g, U, P);
    double a = 0.0;
    for (int n = 0; n < V; n++) a += beta * creal(1.0 - P[n]);
    if (plaq_sums) {
        double sp = 0.0;
        for (int n = 0; n < V; n++) sp += creal(P[n]);
        plaq_sums[0] = sp;
        plaq_sums[1] = a;
    }
    int apps = 0;
    int c = cg_solve(g, U, phi, x, m0, tol, max_iter, &apps);
    if (ok) *ok = c;
    if (its_acc) *its_acc += apps;
    a += creal(dot_c(V, x, phi));
    free(P); free(x);
    return a;
}

/* HMC::Hamiltonian (hmc.cpp:135-149) */
static double hamiltonian(const geom* g, const cd* U, const double* pi, const cd* phi, double beta, double m0,
                          double tol, int max_iter, int* ok, int* its_acc, double* plaq_sums) {
    const int V = g->V;
    double h = 0.0;
    for (int n = 0; n < V; n++) {
        h += 0.5 * pi[n] * pi[n];
        h += 0.5 * pi[V + n] * pi[V + n];
    }
    h += action(g, U, phi, beta, m0, tol, max_iter, ok, its_acc, plaq_sums);
    return h;
}

double so_action(int Nx, int Nt, const double* U, const double* phi, double beta, double m0, double tol, int max_iter) {
    geom g = make_geom(Nx, Nt);
    double a = action(&g, (const cd*)U, (const cd*)phi, beta, m0, tol, max_iter, NULL, NULL, NULL);
    free_geom(&g);
    return a;
}

double so_hamiltonian(int Nx, int Nt, const double* U, const double* pi, const double* phi, double beta, double m0,
                      double tol, int max_iter) {
    geom g = make_geom(Nx, Nt);
    double h = hamiltonian(&g, (const cd*)U, pi, (const cd*)phi, beta, m0, tol, max_iter, NULL, NULL, NULL);
    free_geom(&g);
    return h;
}

/* link update U <- U * exp(i c eps pi)  (hmc.cpp:70-71, 82-86, 96-100) */
static void update_links(int N, cd* U, const double* pi, double coef) {
    for (int n = 0; n < N; n++) U[n] = U[n] * cexp(coef * I * pi[n]);
}

/* HMC::Leapfrog (hmc.cpp:63-103): position-first, MD_steps-1 force evaluations */
static int leapfrog(const geom* g, const cd* U, const double* pi, const cd* phi, int md, double tau, double beta,
                    double m0, double tol, int max_iter, cd* Uo, double* pio, int* its_acc) {
    const int V = g->V, N = 2 * V;
    const double eps = tau / (md * 1.0);
    double* F = (double*)calloc(N, sizeof(double));
    int ok = 1;
    memcpy(pio, pi, sizeof(double) * N);
    memcpy(Uo, U, sizeof(cd) * N);
    update_links(N, Uo, pio, 0.5 * eps);
    ok &= total_force(g, Uo, phi, beta, m0, tol, max_iter, F, its_acc);
    for (int step = 1; step < md - 1; step++) {
        for (int n = 0; n < N; n++) pio[n] += eps * F[n];
        update_links(N, Uo, pio, eps);
        ok &= total_force(g, Uo, phi, beta, m0, tol, max_iter, F, its_acc);
    }
    for (int n = 0; n < N; n++) pio[n] += eps * F[n];
    update_links(N, Uo, pio, 0.5 * eps);
    free(F);
    return ok;
}

int so_leapfrog(int Nx, int Nt, const double* U, const double* pi, const double* phi, int md, double tau, double beta,
                double m0, double tol, int max_iter, double* U_out, double* pi_out) {
    geom g = make_geom(Nx, Nt);
    int ok = leapfrog(&g, (const cd*)U, pi, (const cd*)phi, md, tau, beta, m0, tol, max_iter, (cd*)U_out, pi_out, NULL);
    free_geom(&g);
    return ok;
}

/* One HMC_Update (hmc.cpp:151-181) with injected pi, chi; Metropolis left to the caller.
 * H[0]=H(U,pi), H[1]=H(U',pi'); aux[0]=sum Re P(U'), aux[1]=gauge action(U'); aux[2]=total
 * D D^dagger applications.  returns AND of CG convergence flags. */
int so_trajectory(int Nx, int Nt, const double* U, const double* pi, const double* chi, int md, double tau, double beta,
                  double m0, double tol, int max_iter, double* phi_out, double* U_out, double* pi_out, double* H,
                  double* aux) {
    geom g = make_geom(Nx, Nt);
    const int V = g.V, N = 2 * V;
    cd* phi = (cd*)calloc(N, sizeof(cd));
    const cd* u = (const cd*)U;
    const cd* c = (const cd*)chi;
    int its = 0, ok = 1, o1 = 1, o2 = 1;
    double ps[2];
    apply_D(&g, u, u + V, c, c + V, phi, phi + V, m0, 0);   /* hmc.cpp:160 */
    ok &= leapfrog(&g, u, pi, phi, md, tau, beta, m0, tol, max_iter, (cd*)U_out, pi_out, &its);
    H[1] = hamiltonian(&g, (const cd*)U_out, pi_out, phi, beta, m0, tol, max_iter, &o1, &its, ps);
    H[0] = hamiltonian(&g, u, pi, phi, beta, m0, tol, max_iter, &o2, &its, NULL);
    aux[0] = ps[0];
    aux[1] = ps[1];
    aux[2] = (double)its;
    if (phi_out) memcpy(phi_out, phi, sizeof(cd) * N);
    free(phi);
    free_geom(&g);
    return ok & o1 & o2;
}

/* ------------------------------------------------------------------------------------------
 * Binary configuration files: src/gauge_conf.cpp:404-419 (writer), :515-532 (reader).
 * 28-byte records (int32 x, int32 t, int32 mu, double re, double im), x -> t -> mu.
 * ---------------------------------------------------------------------------------------- */
int so_save_conf(int Nx, int Nt, const double* U, const char* name) {
    FILE* f = fopen(name, "wb");
    if (!f) return 1;
    const int V = Nx * Nt;
    for (int x = 0; x < Nx; x++)
        for (int t = 0; t < Nt; t++)
            for (int mu = 0; mu < 2; mu++) {
                const int n = x * Nt + t;
                int32_t hdr[3] = {x, t, mu};
                double v[2] = {U[2 * (mu * V + n)], U[2 * (mu * V + n) + 1]};
                fwrite(hdr, sizeof(int32_t), 3, f);
                fwrite(v, sizeof(double), 2, f);
            }
    fclose(f);
    return 0;
}

int so_read_binary(int Nx, int Nt, const char* name, double* U) {
    FILE* f = fopen(name, "rb");
    if (!f) return 1;
    const int V = Nx * Nt;
    for (int x = 0; x < Nx; x++)
        for (int t = 0; t < Nt; t++)
            for (int mu = 0; mu < 2; mu++) {
                const int n = x * Nt + t;   /* the reader trusts loop position, not the stored x,t,mu */
                int32_t hdr[3];
                double v[2];
                if (fread(hdr, sizeof(int32_t), 3, f) != 3 || fread(v, sizeof(double), 2, f) != 2) {
                    fclose(f);
                    return 2;
                }
                U[2 * (mu * V + n)] = v[0];
                U[2 * (mu * V + n) + 1] = v[1];
            }
    fclose(f);
    return 0;
}

/* Jackknife error: src/statistics.cpp:6-34 (leave-one-bin-out means, sqrt((b-1)/b * sum dev^2)) */
double so_jackknife(const double* dat, int n, int bins) {
    const int per = n / bins;
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += dat[i] * 1.0;
    mean = mean / n;
    double err = 0.0;
    for (int i = 0; i < bins; i++) {
        double s = 0.0;
        for (int k = 0; k < bins; k++)
            for (int j = k * per; j < k * per + per; j++)
                if (k != i) s += dat[j];
        s = s / (n - per);
        err += pow(s - mean, 2);
    }
    return sqrt(err * (bins - 1) / bins);
}
