// TEST INFRASTRUCTURE ONLY: declaration-level stand-in for <mpi.h>, so that the reference's headers that mention MPI types
// (src/mpi.h, pulled in by src/acc/cpu/cpu_ml_optimiser.h) parse without an MPI installation.  Nothing here can be called.
#pragma once
typedef int MPI_Comm;
typedef int MPI_Group;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
