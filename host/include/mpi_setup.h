// mpi_setup.h -- rank topology.  Same entry point as the reference (include/mpi_setup.h:96-100):
// initializeMPI() validates ranks_x*ranks_t, computes the tile widths, fills the Cartesian
// neighbour ranks and brings up this rank's GPU context.
#ifndef SM_HOST_MPI_SETUP_H
#define SM_HOST_MPI_SETUP_H
#include "variables.h"

void assignWidth();             // reference: mpi_setup.h:6-24 (exit(1) on a bad decomposition)
void buildCartesianTopology();  // reference: mpi_setup.h:39-71
void initializeMPI();

#endif
