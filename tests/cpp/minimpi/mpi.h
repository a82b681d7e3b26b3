// tests/cpp/minimpi/mpi.h - TEST INFRASTRUCTURE.  A small multi-PROCESS stand-in for the MPI calls the reference's host code makes
// in its USE_MPI build (src/main.cc:88-101, src/ic_generator.cc:729-732, include/particle_container.hh, include/general.hh:95-183
// ...), so that the reference's own MPI executable can run with one rank per slab on the library in a container without MPI
// (tests/test_reference_driver.py).  Ranks are separately started processes (environment MINIMPI_RANK / MINIMPI_SIZE /
// MINIMPI_SHM); collectives go through a shared-memory file with a spin barrier, reductions in rank order (deterministic).
// Point-to-point calls - used by src/grid_fft.cc, which the library replaces, and by ghost-zone / glass code - are declared and
// abort, except the collective ring shift MPI_Sendrecv_replace of include/grid_interpolate.hh.
#pragma once
#include <cstddef>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
enum { MPI_COMM_WORLD = 0, MPI_SUCCESS = 0, MPI_INFO_NULL = 0, MPI_MAX_LIBRARY_VERSION_STRING = 64, MPI_MAX_ERROR_STRING = 64 };
enum { MPI_THREAD_SINGLE = 0, MPI_THREAD_FUNNELED = 1, MPI_THREAD_SERIALIZED = 2, MPI_THREAD_MULTIPLE = 3 };
enum { MPI_SUM = 1, MPI_MAX = 2, MPI_MIN = 3 };
enum { MPI_CHAR = 1, MPI_BYTE, MPI_INT, MPI_UNSIGNED, MPI_FLOAT, MPI_DOUBLE, MPI_LONG_DOUBLE, MPI_LONG_LONG, MPI_UNSIGNED_LONG,
       MPI_UNSIGNED_LONG_LONG, MPI_C_FLOAT_COMPLEX, MPI_C_DOUBLE_COMPLEX, MPI_C_LONG_DOUBLE_COMPLEX };
#define MPI_IN_PLACE ((void*)1)
#define MPI_STATUS_IGNORE ((MPI_Status*)0)

extern "C" {
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm, int* rank);
int MPI_Comm_size(MPI_Comm, int* size);
double MPI_Wtime(void);
int MPI_Barrier(MPI_Comm);
int MPI_Get_library_version(char* version, int* len);
int MPI_Error_string(int, char* s, int* len);
int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op, MPI_Comm);
int MPI_Allgather(const void* send, int scount, MPI_Datatype st, void* recv, int rcount, MPI_Datatype rt, MPI_Comm);
int MPI_Alltoall(const void* send, int scount, MPI_Datatype st, void* recv, int rcount, MPI_Datatype rt, MPI_Comm);
int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm);
/* declared only (abort when called) */
int MPI_Isend(const void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request*);
int MPI_Send(const void*, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void*, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status*);
int MPI_Sendrecv_replace(void*, int, MPI_Datatype, int, int, int, int, MPI_Comm, MPI_Status*);
int MPI_Test(MPI_Request*, int*, MPI_Status*);
int MPI_Wait(MPI_Request*, MPI_Status*);
int MPI_Alltoallv(const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm);
int MPI_Exscan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void*, void*, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
}
