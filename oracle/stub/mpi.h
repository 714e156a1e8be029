/* oracle/stub/mpi.h -- placeholder so the reference's locate.c (which includes <mpi.h> but
 * calls no MPI routine on the L2 grid-search path) compiles without an MPI installation. */
#ifndef ORACLE_STUB_MPI_H
#define ORACLE_STUB_MPI_H
typedef int MPI_Comm;
#endif
