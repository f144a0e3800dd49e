// gauge_conf.cpp -- gauge observables (GPU) and configuration I/O (host).
#include "gauge_conf.h"

#include <cmath>
#include <vector>

#include "b200_runtime.h"

c_double RandomU1() {
    const double theta = 2.0 * pi * ((double)rand() / (RAND_MAX));
    return c_double(cos(theta), sin(theta));
}

void GaugeConf::initialization() {
    for (int n = 0; n < mpi::maxSize; n++) {
        Conf.mu0[n] = RandomU1();
        Conf.mu1[n] = RandomU1();
    }
}

void GaugeConf::Compute_Staple() {
    b200::check(sm_compute_staple(b200::ctx(), raw(Conf.mu0), raw(Conf.mu1), raw(Staples.mu0), raw(Staples.mu1)),
                "Compute_Staple");
}

void GaugeConf::Compute_Plaquette01() {
    double sums[2];
    b200::check(sm_compute_plaquette(b200::ctx(), raw(Conf.mu0), raw(Conf.mu1), 1.0, raw(Plaquette01), sums),
                "Compute_Plaquette01");
}

static double global_sum_re_plaquette(const GaugeConf* g) {
    // sum over this rank's Plaquette01, then over the ranks
    double local = 0.0;
    for (int n = 0; n < mpi::maxSize; n++) local += std::real(g->Plaquette01[n]);
    if (mpi::size == 1) return local;
    std::vector<double> all(mpi::size);
    b200::gather(&local, sizeof(double), all.data());
    double s = 0.0;
    if (mpi::rank == 0)
        for (double v : all) s += v;
    b200::bcast(&s, sizeof(double));
    return s;
}

double GaugeConf::MeasureSp_HMC() { return global_sum_re_plaquette(this); }

double GaugeConf::Compute_gaugeAction(const double& beta) {
    double local = 0.0;
    for (int n = 0; n < mpi::maxSize; n++) local += beta * std::real(1.0 - Plaquette01[n]);
    if (mpi::size == 1) return local;
    std::vector<double> all(mpi::size);
    b200::gather(&local, sizeof(double), all.data());
    double s = 0.0;
    if (mpi::rank == 0)
        for (double v : all) s += v;
    b200::bcast(&s, sizeof(double));
    return s;
}

// ---- tiles <-> global field ---------------------------------------------------------------------
// rank r = (cx, ct) owns rows [cx*width_x, ...) x columns [ct*width_t, ...) of the global lattice,
// the same placement the reference builds with MPI_Type_vector + displacements (gauge_conf.cpp:383-387)
static void place_tile(const c_double* tile, int r, c_double* global) {
    const int cx = r / mpi::ranks_t, ct = r % mpi::ranks_t;
    for (int x = 0; x < mpi::width_x; x++)
        std::copy(tile + (size_t)x * mpi::width_t, tile + (size_t)(x + 1) * mpi::width_t,
                  global + (size_t)(cx * mpi::width_x + x) * LV::Nt + ct * mpi::width_t);
}

static void take_tile(const c_double* global, int r, c_double* tile) {
    const int cx = r / mpi::ranks_t, ct = r % mpi::ranks_t;
    for (int x = 0; x < mpi::width_x; x++) {
        const c_double* src = global + (size_t)(cx * mpi::width_x + x) * LV::Nt + ct * mpi::width_t;
        std::copy(src, src + mpi::width_t, tile + (size_t)x * mpi::width_t);
    }
}

static void gather_global(const spinor& local, spinor& global) {
    const size_t bytes = sizeof(c_double) * mpi::maxSize;
    std::vector<c_double> all((size_t)(mpi::rank == 0 ? mpi::size : 1) * mpi::maxSize);
    for (int comp = 0; comp < 2; comp++) {
        b200::gather(comp == 0 ? local.mu0 : local.mu1, bytes, all.data());
        if (mpi::rank == 0)
            for (int r = 0; r < mpi::size; r++)
                place_tile(all.data() + (size_t)r * mpi::maxSize, r, comp == 0 ? global.mu0 : global.mu1);
    }
}

static void scatter_global(spinor& global, spinor& local) {
    // rank 0 holds `global`; everybody receives it and keeps its own tile (setup path, not hot)
    b200::bcast(global.mu0, sizeof(c_double) * LV::Ntot);
    b200::bcast(global.mu1, sizeof(c_double) * LV::Ntot);
    take_tile(global.mu0, mpi::rank, local.mu0);
    take_tile(global.mu1, mpi::rank, local.mu1);
}

void SaveConf(const GaugeConf& GConf, const std::string& Name) {
    spinor GlobalConf(mpi::rank == 0 ? LV::Ntot : 1);
    gather_global(GConf.Conf, GlobalConf);
    if (mpi::rank != 0) return;
    if (sm_save_conf(LV::Nx, LV::Nt, raw(GlobalConf.mu0), raw(GlobalConf.mu1), Name.c_str()) != SM_OK)
        std::cerr << "Error opening file: " << Name << std::endl;
}

void GaugeConf::readBinary(const std::string& name) {
    spinor GlobalConf(LV::Ntot);
    int ok = 1;
    if (mpi::rank == 0) ok = sm_read_conf(LV::Nx, LV::Nt, name.c_str(), raw(GlobalConf.mu0), raw(GlobalConf.mu1)) == SM_OK;
    b200::bcast(&ok, sizeof(int));
    if (!ok) {
        if (mpi::rank == 0) std::cerr << "File " << name << " not found " << std::endl;
        exit(1);
    }
    scatter_global(GlobalConf, Conf);
}
