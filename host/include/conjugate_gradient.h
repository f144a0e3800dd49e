// conjugate_gradient.h -- reference: include/conjugate_gradient.h:16.
// Solves D D^dagger x = phi on the GPU (x0 = phi, stop at ||r|| < CG::tol * ||phi||,
// at most CG::max_iter iterations).  Returns 1 if converged, 0 otherwise.
#ifndef SM_HOST_CONJUGATE_GRADIENT_H
#define SM_HOST_CONJUGATE_GRADIENT_H
#include <cmath>
#include <iostream>

#include "dirac_operator.h"

int conjugate_gradient(const spinor& U, const spinor& phi, spinor& x, const double& m0);
// iterations of the most recent solve on this rank (the reference does not expose it)
int conjugate_gradient_last_iterations();

#endif
