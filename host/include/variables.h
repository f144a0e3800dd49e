// variables.h -- data containers and globals of the host shell.
//
// Same public names and meaning as the reference's include/variables.h (c_double, spinor,
// re_field, namespaces mpi / LV / CG, the periodic-boundary tables, dot(), mod(), format()),
// so code written against the reference compiles against this header.  The arithmetic behind
// every function lives on the GPU in libschwinger_b200.so; "ranks" are GPUs, one process each.
#ifndef SM_HOST_VARIABLES_H
#define SM_HOST_VARIABLES_H

#include <algorithm>
#include <complex>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "config.h"
#include "schwinger_b200.h"

typedef std::complex<double> c_double;
extern double pi;

namespace mpi {   // reference: include/variables.h:17-37 (MPI world -> one process per GPU)
extern int rank, size, maxSize;
extern int ranks_x, ranks_t, width_x, width_t;
extern int rank2d;
extern int coords[2];
extern int top, bot, right, left;
extern int bot_left, bot_right, top_left, top_right;
}  // namespace mpi

namespace LV {    // reference: include/variables.h:41-46
constexpr int Nx = NS;
constexpr int Nt = NT;
constexpr int Ntot = Nx * Nt;
}  // namespace LV

namespace CG {    // reference: include/variables.h:48-51 (mutable, read by conjugate_gradient)
extern int max_iter;
extern double tol;
}  // namespace CG

// Two-component field: mu0/mu1 are the two spin components of a spinor or the two link
// directions of a gauge field (reference: spinor, include/variables.h:54-100) or of a real
// field (re_field, :102-141).  Owning, zero-initialised, deep-copying.
template <typename T>
struct field2 {
    T* mu0;
    T* mu1;
    int size;

    explicit field2(int N = LV::Ntot) : mu0(new T[N]()), mu1(new T[N]()), size(N) {}
    field2(const field2& o) : mu0(new T[o.size]), mu1(new T[o.size]), size(o.size) { copy_from(o); }
    field2& operator=(const field2& o) {
        if (this == &o) return *this;
        if (size != o.size) {
            release();
            size = o.size;
            mu0 = new T[size];
            mu1 = new T[size];
        }
        copy_from(o);
        return *this;
    }
    ~field2() { release(); }
    void clearBuffer() {
        std::fill(mu0, mu0 + size, T());
        std::fill(mu1, mu1 + size, T());
    }

private:
    // the library may have page-locked these arrays (sm_host_register): tell it before they go
    void release() {
        sm_host_forget(mu0);
        sm_host_forget(mu1);
        delete[] mu0;
        delete[] mu1;
    }
    void copy_from(const field2& o) {
        std::copy(o.mu0, o.mu0 + size, mu0);
        std::copy(o.mu1, o.mu1 + size, mu1);
    }
};

typedef field2<c_double> spinor;
typedef field2<double> re_field;
typedef spinor c_matrix;

// raw views for the C ABI (std::complex<double> is layout-compatible with double[2])
inline const double* raw(const c_double* p) { return reinterpret_cast<const double*>(p); }
inline double* raw(c_double* p) { return reinterpret_cast<double*>(p); }

// geometry tables of this rank's tile (reference: include/variables.h:143-154)
int Coords(const int& x, const int& t);
extern int* LeftPB;
extern int* RightPB;
extern c_double* SignL;
extern c_double* SignR;
extern int* x_1_t1;
extern int* x1_t_1;
void allocate_lattice_arrays();
void free_lattice_arrays();

// host scratch spinors kept for source compatibility (reference: variables.h:157-165)
extern spinor DTEMP;
extern spinor TEMP;

inline int mod(int a, int b) {
    const int r = a % b;
    return r < 0 ? r + b : r;
}

// A.B = sum_n A_n conj(B_n) over both components and all ranks (reference: variables.h:181-192)
c_double dot(const spinor& x, const spinor& y);

// fixed, 4 decimals, decimal point removed: 2 -> "20000", -0.18 -> "-01800" (reference: variables.h:197-203)
inline std::string format(const double& number) {
    std::ostringstream s;
    s << std::fixed << std::setprecision(4) << number;
    std::string out = s.str();
    out.erase(out.find('.'), 1);
    return out;
}

#endif
