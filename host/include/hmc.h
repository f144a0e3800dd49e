// hmc.h -- Hybrid Monte Carlo driver with the reference's public interface (include/hmc.h:8-66).
// The whole trajectory (refresh, phi = D chi, leapfrog with its CG solves, both Hamiltonians)
// runs on the GPU inside libschwinger_b200.so; only dH and the plaquette sums come back, and the
// Metropolis decision, the measurement history and the file output stay here.
#ifndef SM_HOST_HMC_H
#define SM_HOST_HMC_H
#include <cstdint>

#include "conjugate_gradient.h"
#include "gauge_conf.h"

class HMC {
public:
    HMC(GaugeConf& GConf, const int& MD_steps, const double& trajectory_length, const int& Ntherm, const int& Nmeas,
        const int& Nsteps, const double& beta, const int& Nspace, const int& Ntime, const double& m0,
        const int& saveconf);
    ~HMC() {}

    void HMC_algorithm();
    double getEp() { return Ep; }
    double getdEp() { return dEp; }
    double getgS() { return gS; }
    double getdgS() { return dgS; }
    double getacceptance_rate(int conf_number) { return acceptance_rate / ((conf_number)*1.0); }

    // extras of the B200 build (not in the reference): start from the configuration held in the
    // GaugeConf passed to the constructor instead of a hot start; work counters
    void set_start_from_conf(bool on) { start_from_conf = on; }
    long long dd_applications() const { return dd_apps_total; }
    double device_seconds() const { return device_ms_total * 1e-3; }
    int trajectories() const { return traj_count; }

private:
    int Nx, Nt, Ntot;
    int MD_steps, Ntherm, Nmeas, Nsteps;
    int saveconf;
    int conf_i;
    double trajectory_length, beta, m0;
    double Ep, dEp, gS, dgS;
    double acceptance_rate;
    int CG_convergence;
    int illConfId;
    bool therm;
    bool start_from_conf;
    GaugeConf GConf;
    double sum_re_plaq, gauge_action;   // of the current configuration
    std::uint64_t seed;
    long long dd_apps_total;
    double device_ms_total;
    int traj_count;

    void HMC_Update();
    void pull_conf();   // device U -> GConf.Conf (for SaveConf)
};

#endif
