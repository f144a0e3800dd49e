// conjugate_gradient.cpp -- reference: src/conjugate_gradient.cpp:4-67, solved on the GPU.
#include "conjugate_gradient.h"

#include "b200_runtime.h"

static int g_last_iterations = 0;

int conjugate_gradient(const spinor& U, const spinor& phi, spinor& x, const double& m0) {
    sm_ctx* c = b200::ctx();
    b200::check(sm_set_cg(c, CG::tol, CG::max_iter), "sm_set_cg");   // CG::tol / CG::max_iter are mutable globals
    int converged = 0;
    b200::check(sm_conjugate_gradient(c, raw(U.mu0), raw(U.mu1), raw(phi.mu0), raw(phi.mu1), raw(x.mu0), raw(x.mu1), m0,
                                      &converged, &g_last_iterations),
                "conjugate_gradient");
    if (!converged && mpi::rank2d == 0)
        std::cout << "CG did not converge in " << CG::max_iter << " iterations" << std::endl;
    return converged;
}

int conjugate_gradient_last_iterations() { return g_last_iterations; }
