/* <mpi.h> for programs that are built against include/FRIES without an MPI installation: the single-process stand-in
 * of include/FRIES/fries_global.hpp (one process drives the GPU(s); MPI_COMM_WORLD has one rank).  Add this directory to
 * the include path only when no real MPI is wanted. */
#pragma once
#include "../FRIES/fries_global.hpp"
