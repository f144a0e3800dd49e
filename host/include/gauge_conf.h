// gauge_conf.h -- gauge configuration container (reference: include/gauge_conf.h:18-114).
#ifndef SM_HOST_GAUGE_CONF_H
#define SM_HOST_GAUGE_CONF_H
#include <fstream>
#include <string>

#include "statistics.h"
#include "variables.h"

c_double RandomU1();   // uniform phase from rand() (reference: src/gauge_conf.cpp:23-29)

class GaugeConf {
public:
    GaugeConf() : Conf(mpi::maxSize), Staples(mpi::maxSize), Plaquette01(new c_double[mpi::maxSize]()) {}
    GaugeConf(const GaugeConf& o) : Conf(o.Conf), Staples(o.Staples), Plaquette01(new c_double[mpi::maxSize]) {
        std::copy(o.Plaquette01, o.Plaquette01 + mpi::maxSize, Plaquette01);
    }
    GaugeConf& operator=(const GaugeConf& o) {
        if (this != &o) {
            Conf = o.Conf;
            Staples = o.Staples;
            std::copy(o.Plaquette01, o.Plaquette01 + mpi::maxSize, Plaquette01);
        }
        return *this;
    }
    ~GaugeConf() { delete[] Plaquette01; }

    void initialization();        // hot start: RandomU1() per link, mu0 then mu1 per site

    spinor Conf;                  // links U_mu(n): mu0 = time, mu1 = space
    spinor Staples;
    c_double* Plaquette01;        // U_01(n)

    void Compute_Staple();        // GPU: sm_compute_staple
    void Compute_Plaquette01();   // GPU: sm_compute_plaquette (fills Plaquette01 and caches both sums)
    double MeasureSp_HMC();                           // sum Re U_01 (plaquettes must be current)
    double Compute_gaugeAction(const double& beta);   // beta * sum Re(1 - U_01)

    void readBinary(const std::string& name);   // the .ctxt format SaveConf writes
};

// gathers the tiles on rank 0 and writes 28-byte records (x, t, mu, re, im), x -> t -> mu
void SaveConf(const GaugeConf& GConf, const std::string& Name);

#endif
