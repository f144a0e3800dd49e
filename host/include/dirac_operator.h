// dirac_operator.h -- Wilson Dirac operator entry points with the reference's signatures
// (include/dirac_operator.h:35-93).  Each call forwards to libschwinger_b200.so (sm_D_phi, ...):
// the callers own all buffers, outputs are pre-sized spinors of mpi::maxSize sites, in and out
// must not alias, and every call is collective over the ranks.
#ifndef SM_HOST_DIRAC_OPERATOR_H
#define SM_HOST_DIRAC_OPERATOR_H
#include "variables.h"

extern c_double I_number;

// fills RightPB / LeftPB / SignR / SignL / x_1_t1 / x1_t_1 for this rank from the device's own
// index arithmetic (sm_tables); reference: dirac_operator.h:35-62
void periodic_boundary();

void D_phi(const spinor& U, const spinor& phi, spinor& Dphi, const double& m0);
void D_dagger_phi(const spinor& U, const spinor& phi, spinor& Dphi, const double& m0);
void D_D_dagger_phi(const spinor& U, const spinor& phi, spinor& Dphi, const double& m0);
// Im[ left^dagger (dD/domega) right ] per link (eq. 37-38 of HMC_doc.pdf)
re_field phi_dag_partialD_phi(const spinor& U, const spinor& left, const spinor& right);

#endif
