// dirac_operator.cpp -- forwards the reference's operator calls to the GPU library.
#include "dirac_operator.h"

#include "b200_runtime.h"

c_double I_number(0, 1);

void periodic_boundary() {
    b200::check(sm_tables(b200::ctx(), mpi::ranks_x, mpi::ranks_t, mpi::rank, RightPB, LeftPB, raw(SignR), raw(SignL),
                          x_1_t1, x1_t_1),
                "periodic_boundary");
}

void D_phi(const spinor& U, const spinor& phi, spinor& Dphi, const double& m0) {
    b200::check(sm_D_phi(b200::ctx(), raw(U.mu0), raw(U.mu1), raw(phi.mu0), raw(phi.mu1), raw(Dphi.mu0), raw(Dphi.mu1), m0),
                "D_phi");
}

void D_dagger_phi(const spinor& U, const spinor& phi, spinor& Dphi, const double& m0) {
    b200::check(sm_D_dagger_phi(b200::ctx(), raw(U.mu0), raw(U.mu1), raw(phi.mu0), raw(phi.mu1), raw(Dphi.mu0),
                                raw(Dphi.mu1), m0),
                "D_dagger_phi");
}

void D_D_dagger_phi(const spinor& U, const spinor& phi, spinor& Dphi, const double& m0) {
    b200::check(sm_D_D_dagger_phi(b200::ctx(), raw(U.mu0), raw(U.mu1), raw(phi.mu0), raw(phi.mu1), raw(Dphi.mu0),
                                  raw(Dphi.mu1), m0),
                "D_D_dagger_phi");
}

re_field phi_dag_partialD_phi(const spinor& U, const spinor& left, const spinor& right) {
    re_field F(mpi::maxSize);
    b200::check(sm_phi_dag_partialD_phi(b200::ctx(), raw(U.mu0), raw(U.mu1), raw(left.mu0), raw(left.mu1), raw(right.mu0),
                                        raw(right.mu1), F.mu0, F.mu1),
                "phi_dag_partialD_phi");
    return F;
}
