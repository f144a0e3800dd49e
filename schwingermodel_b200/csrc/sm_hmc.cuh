// sm_hmc.cuh -- gauge ghosts, force / plaquette / kinetic / leapfrog launches, host<->device copies, HMC pieces (src/hmc.cpp).
// Part of the single translation unit sm_abi.cu (static functions, included in dependency order).
#pragma once
#include "sm_cg.cuh"

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
    NvtxRange nvtx("sm:leapfrog update");
    k_leap_update<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(U, pi, F, eps_pi, eps_u, 2 * c->V);
    KCHECK();
    c->launches++;
    invalidate_gauge_ghosts(c, U);
    return SM_OK;
}

// ------------------------------------------------------------------------------------------------
// host <-> device field copies (component arrays of the reference's spinor / re_field)
// ------------------------------------------------------------------------------------------------
// Caller buffers of the host-buffer entry points are pageable (`new[]` in the reference's spinor, include/variables.h:
// 54-100): a copy from pageable memory is staged by the driver and runs at a fraction of the link rate.  Opt-in
// (sm_host_register(1); the C++ shell in host/ switches it on): page-lock a caller buffer the first time it is seen
// and remember it by address; the owner calls sm_host_forget(ptr) before it frees the buffer (the shell's field
// destructor does).  Off by default because a stale entry for memory that was freed and mapped again is a hazard only
// the owner can rule out.
struct HostPin {
    const void* p;
    size_t bytes;   // 0: registration refused (e.g. already page-locked by its owner): do not try again
};
static std::mutex g_pin_mu;
static std::vector<HostPin> g_pins;
static bool g_pin_enabled = false;

static void host_pin(const void* p, size_t bytes) {
    if (!g_pin_enabled || bytes < ((size_t)1 << 16)) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (auto& e : g_pins)
        if (e.p == p) {
            if (e.bytes == 0 || e.bytes >= bytes) return;
            cudaHostUnregister((void*)p);          // grew: pin again below
            e = g_pins.back();
            g_pins.pop_back();
            break;
        }
    if (g_pins.size() >= 256) return;
    const cudaError_t err = cudaHostRegister((void*)p, bytes, cudaHostRegisterDefault);
    if (err != cudaSuccess) cudaGetLastError();
    g_pins.push_back({p, err == cudaSuccess ? bytes : 0});
}

static void host_forget(const void* p) {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (size_t i = 0; i < g_pins.size(); i++)
        if (g_pins[i].p == p || p == nullptr) {
            if (g_pins[i].bytes) {
                if (cudaHostUnregister((void*)g_pins[i].p) != cudaSuccess) cudaGetLastError();
            }
            g_pins[i] = g_pins.back();
            g_pins.pop_back();
            if (p != nullptr) return;
            i--;
        }
}

static int h2d_c(sm_ctx* c, cplx* d, const double* h0, const double* h1) {
    invalidate_gauge_ghosts(c, d);
    host_pin(h0, sizeof(cplx) * c->V);
    host_pin(h1, sizeof(cplx) * c->V);
    CU(cudaMemcpyAsync(d, h0, sizeof(cplx) * c->V, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + c->V, h1, sizeof(cplx) * c->V, cudaMemcpyHostToDevice, c->stream));
    return SM_OK;
}
static int d2h_c(sm_ctx* c, const cplx* d, double* h0, double* h1) {
    host_pin(h0, sizeof(cplx) * c->V);
    host_pin(h1, sizeof(cplx) * c->V);
    CU(cudaMemcpyAsync(h0, d, sizeof(cplx) * c->V, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(h1, d + c->V, sizeof(cplx) * c->V, cudaMemcpyDeviceToHost, c->stream));
    return SM_OK;
}
static int h2d_r(sm_ctx* c, double* d, const double* h0, const double* h1) {
    host_pin(h0, sizeof(double) * c->V);
    host_pin(h1, sizeof(double) * c->V);
    CU(cudaMemcpyAsync(d, h0, sizeof(double) * c->V, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d + c->V, h1, sizeof(double) * c->V, cudaMemcpyHostToDevice, c->stream));
    return SM_OK;
}
static int d2h_r(sm_ctx* c, const double* d, double* h0, double* h1) {
    host_pin(h0, sizeof(double) * c->V);
    host_pin(h1, sizeof(double) * c->V);
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
    int force_fail = 0;   // non-converged solves inside HMC::Force (the reference dumps an illConf for each, hmc.cpp:48-56)
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

// HMC::Force (hmc.cpp:44-60): psi = (DD^dagger)^-1 phi ; chi' = D^dagger psi ; fermion + gauge force.
// Opt-in SM_SOLVER_CHRONO: the solve starts from the solution of the previous force evaluation of the
// trajectory (second one) or from the linear extrapolation 2 psi_{-1} - psi_{-2} of the last two (later ones) instead
// of from phi; the CG itself and its stopping rule are unchanged.
// Opt-in even-odd HMC (SM_SOLVER_EVENODD): the pseudofermion lives on the even sites and its action is
//   S_pf = phi_e^dagger (Dhat Dhat^dagger)^-1 phi_e,   Dhat = m - (1/4m) H_eo H_oe  the Schur complement of D,
// which has the same determinant as D D^dagger up to a constant (det D = m^{V/2} det Dhat), i.e. the same distribution of
// gauge fields.  With X_e = (Dhat Dhat^dagger)^-1 phi_e and Y_e = Dhat^dagger X_e the variation is
//   dS = -2 Re( X^dagger dD Y )  on the full lattice with  X_o = -(1/m) (D^dagger X_e)_o ,  Y_o = -(1/m) (D Y_e)_o ,
// so the force is the reference's bilinear phi_dag_partialD_phi(U, X, Y) (src/dirac_operator.cpp:486-580) of the
// completed fields, plus the unchanged gauge force.
static int hmc_force_eo(sm_ctx* c, const cplx* U, const cplx* phi, double* F, TrajAcc* acc) {
    int ok = 0, its = 0;
    const double m0 = c->hp.m0, m = m0 + 2;
    TRY(dev_cg_eo(c, U, phi, c->psi, m0, &ok, &its));
    if (acc) {
        acc->dd_apps += ok ? its + 2 : its + 1;
        acc->solves++;
        acc->all_ok &= ok;
        acc->force_fail += ok ? 0 : 1;
    }
    TRY((dev_Dhat<true>(c, U, c->psi, c->tmp, c->xi, m0)));                                           // Y_e
    TRY((launch_wilson_eo<true, WILSON_EO>(c, U, c->psi, c->tmp, m0, 1, nullptr, 0.0, -1.0 / m)));     // X_o
    k_add_into<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(c->psi, c->tmp, 2 * c->V);
    KCHECK();
    TRY((launch_wilson_eo<false, WILSON_EO>(c, U, c->xi, c->tmp, m0, 1, nullptr, 0.0, -1.0 / m)));    // Y_o
    k_add_into<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(c->xi, c->tmp, 2 * c->V);
    KCHECK();
    c->launches += 2;
    return dev_force(c, U, c->psi, c->xi, F, c->hp.beta, true, true);
}

static int hmc_force(sm_ctx* c, const cplx* U, const cplx* phi, double* F, TrajAcc* acc) {
    NvtxRange nvtx("sm:HMC::Force");
    if (c->solver == SM_SOLVER_EVENODD) return hmc_force_eo(c, U, phi, F, acc);
    int ok = 0, its = 0;
    const cplx* x0 = nullptr;
    if (c->solver == SM_SOLVER_CHRONO && c->chrono_have > 0) {
        TRY(ensure_complex(c, &c->chrono_guess));
        if (c->chrono_have >= 2) {
            k_extrapolate<<<c->flat_blocks_c, kBlock, 0, c->stream>>>(c->psi, c->chrono_prev, c->chrono_guess, 2 * c->V);
            KCHECK();
            c->launches++;
        } else {
            CU(cudaMemcpyAsync(c->chrono_guess, c->psi, sizeof(cplx) * 2 * c->V, cudaMemcpyDeviceToDevice, c->stream));
        }
        x0 = c->chrono_guess;
    }
    if (c->solver == SM_SOLVER_CHRONO) {          // psi is about to be overwritten: keep it as psi_{-2} of the next solve
        TRY(ensure_complex(c, &c->chrono_prev));
        if (c->chrono_have > 0)
            CU(cudaMemcpyAsync(c->chrono_prev, c->psi, sizeof(cplx) * 2 * c->V, cudaMemcpyDeviceToDevice, c->stream));
    }
    TRY(dev_cg(c, U, phi, c->psi, c->hp.m0, &ok, &its, x0));
    if (c->solver == SM_SOLVER_CHRONO) c->chrono_have = std::min(2, c->chrono_have + 1);
    if (acc) {
        acc->dd_apps += ok ? its + 2 : its + 1;
        acc->solves++;
        acc->all_ok &= ok;
        acc->force_fail += ok ? 0 : 1;
    }
    TRY(dev_D(c, U, c->psi, c->xi, c->hp.m0, true));
    return dev_force(c, U, c->psi, c->xi, F, c->hp.beta, true, true);
}

// phi = D chi (hmc.cpp:160); even-odd HMC: phi_e = Dhat chi_e, chi's odd sites dropped
static int hmc_pseudofermion(sm_ctx* c) {
    if (c->solver != SM_SOLVER_EVENODD) return dev_D(c, c->U, c->chi, c->phi, c->hp.m0, false);
    TRY(ensure_complex(c, &c->tmp));
    k_mask_parity<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(c->chi, c->wt, c->V, 0);
    KCHECK();
    c->launches++;
    return dev_Dhat<false>(c, c->U, c->chi, c->tmp, c->phi, c->hp.m0);
}

// a pseudofermion supplied by the caller (test entry points): the even-odd action sees its even sites only
static int hmc_adopt_phi(sm_ctx* c) {
    if (c->solver != SM_SOLVER_EVENODD) return SM_OK;
    k_mask_parity<<<c->flat_blocks_s, kBlock, 0, c->stream>>>(c->phi, c->wt, c->V, 0);
    KCHECK();
    c->launches++;
    return SM_OK;
}

// HMC::Leapfrog (hmc.cpp:63-103): position first, MD_steps-1 force evaluations
static int hmc_leapfrog(sm_ctx* c, TrajAcc* acc) {
    NvtxRange nvtx("sm:HMC::Leapfrog");
    const int md = c->hp.md_steps;
    const double eps = c->hp.trajectory_length / (md * 1.0);
    CU(cudaMemcpyAsync(c->pip, c->pi, sizeof(double) * 2 * c->V, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->Up, c->U, sizeof(cplx) * 2 * c->V, cudaMemcpyDeviceToDevice, c->stream));
    invalidate_gauge_ghosts(c, c->Up);
    c->chrono_have = 0;     // a new trajectory: U' jumps back to U, earlier solutions are no guide
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
    NvtxRange nvtx("sm:HMC::Hamiltonian");
    TRY(dev_kinetic(c, pi, c->sums + base));
    TRY(dev_plaquette(c, U, c->hp.beta, nullptr, c->sums + base + 1));
    int ok = 0, its = 0;
    // opt-in chronological start: the proposal's action solve sits half a link step after the last force solve
    const cplx* x0 = (c->solver == SM_SOLVER_CHRONO && c->chrono_have > 0 && U == c->Up) ? c->psi : nullptr;
    if (c->solver == SM_SOLVER_EVENODD) TRY(dev_cg_eo(c, U, phi, c->xi, c->hp.m0, &ok, &its));   // phi lives on the even sites
    else TRY(dev_cg(c, U, phi, c->xi, c->hp.m0, &ok, &its, x0));
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
