// variables.cpp -- globals of the host shell (reference: src/variables.cpp).
#include "variables.h"

#include "b200_runtime.h"
#include "mpi_setup.h"

double pi = 3.14159265359;   // the reference's truncated value (src/variables.cpp:3), used by the hot start

namespace mpi {
int rank = 0, size = 1, maxSize = LV::Ntot;
int ranks_x = 1, ranks_t = 1, width_x = LV::Nx, width_t = LV::Nt;
int rank2d = 0;
int coords[2] = {0, 0};
int top = 0, bot = 0, right = 0, left = 0;
int bot_left = 0, bot_right = 0, top_left = 0, top_right = 0;
}  // namespace mpi

namespace CG {
int max_iter = 10000;
double tol = 1e-10;
}  // namespace CG

int* LeftPB = nullptr;
int* RightPB = nullptr;
c_double* SignL = nullptr;
c_double* SignR = nullptr;
int* x_1_t1 = nullptr;
int* x1_t_1 = nullptr;

spinor DTEMP(1);
spinor TEMP(1);

int Coords(const int& x, const int& t) { return x * mpi::width_t + t; }

void allocate_lattice_arrays() {
    const int m = mpi::maxSize;
    LeftPB = new int[2 * m];
    RightPB = new int[2 * m];
    SignL = new c_double[2 * m];
    SignR = new c_double[2 * m];
    x_1_t1 = new int[m];
    x1_t_1 = new int[m];
    DTEMP = spinor(m);
    TEMP = spinor(m);
}

void free_lattice_arrays() {
    delete[] LeftPB;
    delete[] RightPB;
    delete[] SignL;
    delete[] SignR;
    delete[] x_1_t1;
    delete[] x1_t_1;
    LeftPB = RightPB = x_1_t1 = x1_t_1 = nullptr;
    SignL = SignR = nullptr;
}

c_double dot(const spinor& x, const spinor& y) {
    double z[2];
    b200::check(sm_dot(b200::ctx(), raw(x.mu0), raw(x.mu1), raw(y.mu0), raw(y.mu1), z), "dot");
    return c_double(z[0], z[1]);
}

// ---- topology (mpi_setup.h) -------------------------------------------------------------------
void assignWidth() {
    if (mpi::ranks_t * mpi::ranks_x != mpi::size) {
        if (mpi::rank == 0) {
            std::cout << "ranks_t * ranks_x != total number of ranks" << std::endl;
            std::cout << mpi::ranks_t * mpi::ranks_x << " != " << mpi::size << std::endl;
        }
        exit(1);
    }
    if (LV::Nx % mpi::ranks_x != 0 || LV::Nt % mpi::ranks_t != 0) {
        if (mpi::rank == 0) std::cout << "Nx (Nt) is not exactly divisible by rank_x (rank_t)" << std::endl;
        exit(1);
    }
    mpi::width_x = LV::Nx / mpi::ranks_x;
    mpi::width_t = LV::Nt / mpi::ranks_t;
    mpi::maxSize = mpi::width_t * mpi::width_x;
}

void buildCartesianTopology() {
    // row-major ranks, periodic in both directions: coords = (rank / ranks_t, rank % ranks_t)
    auto at = [](int cx, int ct) { return mod(cx, mpi::ranks_x) * mpi::ranks_t + mod(ct, mpi::ranks_t); };
    mpi::rank2d = mpi::rank;
    const int cx = mpi::coords[0] = mpi::rank / mpi::ranks_t;
    const int ct = mpi::coords[1] = mpi::rank % mpi::ranks_t;
    mpi::left = at(cx, ct - 1);
    mpi::right = at(cx, ct + 1);
    mpi::top = at(cx - 1, ct);
    mpi::bot = at(cx + 1, ct);
    mpi::bot_left = at(cx + 1, ct - 1);
    mpi::bot_right = at(cx + 1, ct + 1);
    mpi::top_left = at(cx - 1, ct - 1);
    mpi::top_right = at(cx - 1, ct + 1);
}

void initializeMPI() {
    assignWidth();
    buildCartesianTopology();
    b200::create_context();
}
