// api_probe.cpp -- calls the host shell through the REFERENCE's C++ interface (spinor / re_field
// containers, D_phi, D_dagger_phi, D_D_dagger_phi, dot, conjugate_gradient, phi_dag_partialD_phi,
// GaugeConf::{readBinary, Compute_Staple, Compute_Plaquette01, MeasureSp_HMC, Compute_gaugeAction},
// SaveConf) exactly as reference code would, and dumps the results as raw doubles for
// tests/test_host_shell.py to compare with the oracle.
//   usage: api_probe <conf.ctxt> <phi.bin> <m0> <beta> <out.bin> <saved.ctxt>
#include <cstdio>
#include <fstream>

#include "b200_runtime.h"
#include "hmc.h"
#include "mpi_setup.h"

static void dump(std::ofstream& o, const c_double* p, int n) { o.write(reinterpret_cast<const char*>(p), sizeof(c_double) * n); }
static void dump(std::ofstream& o, const double* p, int n) { o.write(reinterpret_cast<const char*>(p), sizeof(double) * n); }

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const double m0 = atof(argv[3]), beta = atof(argv[4]);
    mpi::ranks_x = mpi::ranks_t = 1;
    b200::spawn_ranks(1);
    initializeMPI();
    allocate_lattice_arrays();
    periodic_boundary();

    GaugeConf G;
    G.readBinary(argv[1]);
    spinor phi(mpi::maxSize), out(mpi::maxSize), x(mpi::maxSize);
    {
        std::ifstream in(argv[2], std::ios::binary);
        in.read(reinterpret_cast<char*>(phi.mu0), sizeof(c_double) * mpi::maxSize);
        in.read(reinterpret_cast<char*>(phi.mu1), sizeof(c_double) * mpi::maxSize);
        if (!in) return 3;
    }
    std::ofstream o(argv[5], std::ios::binary);
    const int V = mpi::maxSize;
    o.write(reinterpret_cast<const char*>(RightPB), sizeof(int) * 2 * V);
    o.write(reinterpret_cast<const char*>(LeftPB), sizeof(int) * 2 * V);
    dump(o, SignR, 2 * V);
    dump(o, SignL, 2 * V);
    D_phi(G.Conf, phi, out, m0);
    dump(o, out.mu0, V); dump(o, out.mu1, V);
    D_dagger_phi(G.Conf, phi, out, m0);
    dump(o, out.mu0, V); dump(o, out.mu1, V);
    D_D_dagger_phi(G.Conf, phi, out, m0);
    dump(o, out.mu0, V); dump(o, out.mu1, V);
    const c_double z = dot(phi, out);
    dump(o, &z, 1);
    const double ok = conjugate_gradient(G.Conf, phi, x, m0);
    const double its = conjugate_gradient_last_iterations();
    dump(o, &ok, 1); dump(o, &its, 1);
    dump(o, x.mu0, V); dump(o, x.mu1, V);
    D_dagger_phi(G.Conf, x, out, m0);
    re_field F = phi_dag_partialD_phi(G.Conf, x, out);
    dump(o, F.mu0, V); dump(o, F.mu1, V);
    G.Compute_Staple();
    dump(o, G.Staples.mu0, V); dump(o, G.Staples.mu1, V);
    G.Compute_Plaquette01();
    dump(o, G.Plaquette01, V);
    const double sp = G.MeasureSp_HMC(), sg = G.Compute_gaugeAction(beta);
    dump(o, &sp, 1); dump(o, &sg, 1);
    o.close();
    GaugeConf H = G;          // deep copy, then write it back out
    SaveConf(H, argv[6]);
    free_lattice_arrays();
    b200::shutdown();
    return 0;
}
