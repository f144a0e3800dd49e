// hmc.cpp -- HMC driver: the control flow of the reference's src/hmc.cpp:151-215 around the
// device-resident trajectory of libschwinger_b200.so.
#include "hmc.h"

#include <cmath>
#include <random>
#include <sstream>

#include "b200_runtime.h"

HMC::HMC(GaugeConf& GConf, const int& MD_steps, const double& trajectory_length, const int& Ntherm, const int& Nmeas,
         const int& Nsteps, const double& beta, const int& Nspace, const int& Ntime, const double& m0,
         const int& saveconf)
    : Nx(Nspace), Nt(Ntime), Ntot(Nspace * Ntime), MD_steps(MD_steps), Ntherm(Ntherm), Nmeas(Nmeas), Nsteps(Nsteps),
      saveconf(saveconf), conf_i(0), trajectory_length(trajectory_length), beta(beta), m0(m0), Ep(0), dEp(0), gS(0),
      dgS(0), acceptance_rate(0), CG_convergence(1), illConfId(0), therm(false), start_from_conf(false), GConf(GConf),
      sum_re_plaq(0), gauge_action(0), seed(0), dd_apps_total(0), device_ms_total(0), traj_count(0) {
    // one seed for the device generator, shared by all ranks (the fields are indexed by global site);
    // SM_SEED makes a run reproducible, otherwise std::random_device like the reference (hmc.cpp:7-8)
    if (mpi::rank == 0) {
        if (const char* s = std::getenv("SM_SEED")) seed = std::strtoull(s, nullptr, 10);
        else {
            std::random_device rd;
            seed = ((std::uint64_t)rd() << 32) ^ rd();
        }
    }
    b200::bcast(&seed, sizeof(seed));
}

void HMC::pull_conf() {
    b200::check(sm_hmc_get_gauge(b200::ctx(), raw(GConf.Conf.mu0), raw(GConf.Conf.mu1), 0), "sm_hmc_get_gauge");
}

void HMC::HMC_Update() {
    sm_ctx* c = b200::ctx();
    b200::check(sm_hmc_refresh(c, seed, (std::uint64_t)traj_count), "sm_hmc_refresh");   // RandomPI, RandomCHI
    sm_traj_result r;
    b200::check(sm_hmc_trajectory(c, &r), "sm_hmc_trajectory");   // phi = D chi, Leapfrog, deltaH
    traj_count++;
    dd_apps_total += r.dd_applications;
    device_ms_total += r.kernel_ms;
    CG_convergence = r.cg_all_converged;
    // hmc.cpp:46-56: every Force whose CG did not converge dumps the current (accepted) configuration GConf -- not the
    // proposal that failed -- and bumps illConfId; the two Action solves do not (their dump is commented out, :121-131).
    // The trajectory has already run on the device, so the dumps come after it: same files, same contents.
    for (int f = 0; f < r.cg_force_failures; f++) {
        std::ostringstream name;
        name << "2D_U1_" << Nx << "x" << Nt << "_b" << format(beta) << "_m" << format(m0) << "_illConf" << illConfId
             << ".ctxt";
        if (f == 0) pull_conf();
        SaveConf(GConf, name.str());
        illConfId += 1;
    }
    double u = 0.0;
    if (mpi::rank == 0) u = rand_range(0, 1);   // same number on all ranks (hmc.cpp:166-169)
    b200::bcast(&u, sizeof(double));
    const bool accept = u <= exp(-r.dH);
    b200::check(sm_hmc_accept(c, accept ? 1 : 0), "sm_hmc_accept");
    if (accept) {
        sum_re_plaq = r.sum_re_plaq_new;
        gauge_action = r.gauge_action_new;
        if (therm) acceptance_rate += 1.0;
    } else {
        sum_re_plaq = r.sum_re_plaq_old;
        gauge_action = r.gauge_action_old;
    }
}

void HMC::HMC_algorithm() {
    sm_ctx* c = b200::ctx();
    std::vector<double> SpVector(Nmeas), gAction(Nmeas);
    if (!start_from_conf) GConf.initialization();   // hot start (hmc.cpp:186)
    sm_hmc_params p{beta, m0, trajectory_length, MD_steps};
    b200::check(sm_set_cg(c, CG::tol, CG::max_iter), "sm_set_cg");
    b200::check(sm_hmc_configure(c, &p), "sm_hmc_configure");
    b200::check(sm_hmc_set_gauge(c, raw(GConf.Conf.mu0), raw(GConf.Conf.mu1)), "sm_hmc_set_gauge");
    for (int i = 0; i < Ntherm; i++) {
        HMC_Update();
        if (i % 100 == 0 && mpi::rank == 0)
            std::cout << "Conf " << i << " out of " << Ntherm << " for thermalization" << std::endl;
    }
    therm = true;
    if (mpi::rank == 0) std::cout << "Thermalization done" << std::endl;
    conf_i = 0;
    for (int i = 0; i < Nmeas; i++) {
        conf_i += 1;
        HMC_Update();
        SpVector[i] = sum_re_plaq;    // MeasureSp_HMC of the current configuration
        gAction[i] = gauge_action;    // Compute_gaugeAction(beta)
        if (saveconf == 1) {
            std::ostringstream name;
            name << "2D_U1_Ns" << Nx << "_Nt" << Nt << "_b" << format(beta) << "_m" << format(m0) << "_" << i << ".ctxt";
            pull_conf();
            SaveConf(GConf, name.str());
        }
        if (i != Nmeas - 1)
            for (int j = 0; j < Nsteps; j++) {
                conf_i += 1;
                HMC_Update();
            }
    }
    pull_conf();
    Ep = mean(SpVector) / (Ntot * 1.0);
    dEp = Jackknife_error(SpVector, 20) / (Ntot * 1.0);
    gS = mean(gAction) / (Ntot * 1.0);
    dgS = Jackknife_error(gAction, 20) / (Ntot * 1.0);
}
