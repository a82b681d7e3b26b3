// oracle/refshim/mpi.h - TEST INFRASTRUCTURE.  A ONE-RANK stand-in for the MPI calls the reference's Grid_FFT / convolver
// sources make in their USE_MPI build (src/grid_fft.cc:120-140, 336-492; include/grid_fft.hh:527-690;
// include/general.hh:95-183), so that the MPI code path - the one the reference's CI builds and its golden files come
// from: transposed k-layout, pad/unpad through FourierInterpolateCopyTo with the Nyquist planes dropped - can be compiled
// and run in a container without MPI.  Collectives are copies; a send to self is kept until the matching receive.
#pragma once
#include <cstddef>
#include <cstring>
#include <ctime>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;   // the value is the element size in bytes
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
enum { MPI_COMM_WORLD = 0, MPI_SUCCESS = 0, MPI_INFO_NULL = 0, MPI_THREAD_FUNNELED = 1, MPI_MAX_LIBRARY_VERSION_STRING = 64 };
enum { MPI_SUM = 1, MPI_MAX = 2, MPI_MIN = 3 };
#define MPI_IN_PLACE ((void*)1)
#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_UNSIGNED 4
#define MPI_FLOAT 4
#define MPI_DOUBLE 8
#define MPI_LONG_DOUBLE 16
#define MPI_LONG_LONG 8
#define MPI_UNSIGNED_LONG 8
#define MPI_UNSIGNED_LONG_LONG 8
#define MPI_C_FLOAT_COMPLEX 8
#define MPI_C_DOUBLE_COMPLEX 16
#define MPI_C_LONG_DOUBLE_COMPLEX 32

namespace mpi_shim {
struct Pending { const void* buf; size_t bytes; int tag; bool done; };
inline std::vector<Pending>& pending() { static std::vector<Pending> p; return p; }
}  // namespace mpi_shim

inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
inline double MPI_Wtime() { return double(std::clock()) / CLOCKS_PER_SEC; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Get_library_version(char* v, int* len) { std::strcpy(v, "one-rank shim"); *len = 13; return MPI_SUCCESS; }
inline int MPI_Allreduce(const void* s, void* r, int count, MPI_Datatype t, MPI_Op, MPI_Comm) {
    if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, size_t(count) * t);
    return MPI_SUCCESS;
}
inline int MPI_Allgather(const void* s, int scount, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm) {
    if (s != MPI_IN_PLACE && s != r) std::memcpy(r, s, size_t(scount) * st);
    return MPI_SUCCESS;
}
inline int MPI_Isend(const void* buf, int count, MPI_Datatype t, int /*dest*/, int tag, MPI_Comm, MPI_Request* req) {
    mpi_shim::pending().push_back({buf, size_t(count) * t, tag, false});
    *req = int(mpi_shim::pending().size()) - 1;
    return MPI_SUCCESS;
}
inline int MPI_Recv(void* buf, int count, MPI_Datatype t, int /*source*/, int tag, MPI_Comm, MPI_Status* st) {
    for (auto& p : mpi_shim::pending())
        if (!p.done && p.tag == tag) {
            std::memcpy(buf, p.buf, p.bytes < size_t(count) * t ? p.bytes : size_t(count) * t);
            p.done = true;
            if (st) { st->MPI_SOURCE = 0; st->MPI_TAG = tag; st->MPI_ERROR = MPI_SUCCESS; }
            return MPI_SUCCESS;
        }
    if (st) st->MPI_ERROR = 1;
    return 1;   // no matching send: a real MPI would block forever
}
// declared for template code of the reference that the one-rank driver never instantiates into a call with > 1 rank
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
inline int MPI_Error_string(int, char* s, int* len) { s[0] = 0; *len = 0; return MPI_SUCCESS; }
inline int MPI_Bcast(void*, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Sendrecv_replace(void*, int, MPI_Datatype, int, int, int, int, MPI_Comm, MPI_Status* st) {   // self: data stays
    if (st) st->MPI_ERROR = MPI_SUCCESS;
    return MPI_SUCCESS;
}
inline int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) {
    MPI_Request r;
    return MPI_Isend(buf, count, t, dest, tag, c, &r);
}
inline int MPI_Alltoall(const void* s, int scount, MPI_Datatype st, void* r, int, MPI_Datatype, MPI_Comm) {
    if (s != r) std::memcpy(r, s, size_t(scount) * st);
    return MPI_SUCCESS;
}
inline int MPI_Alltoallv(const void* s, const int* scounts, const int* sdispls, MPI_Datatype st, void* r, const int*, const int* rdispls,
                         MPI_Datatype, MPI_Comm) {
    std::memcpy((char*)r + size_t(rdispls[0]) * st, (const char*)s + size_t(sdispls[0]) * st, size_t(scounts[0]) * st);
    return MPI_SUCCESS;
}
inline int MPI_Test(MPI_Request* req, int* flag, MPI_Status*) { *flag = mpi_shim::pending()[*req].done ? 1 : 0; return MPI_SUCCESS; }
inline int MPI_Wait(MPI_Request* req, MPI_Status* st) {
    if (st) st->MPI_ERROR = mpi_shim::pending()[*req].done ? MPI_SUCCESS : 1;
    bool all = true;
    for (auto& p : mpi_shim::pending()) all = all && p.done;
    if (all && *req == int(mpi_shim::pending().size()) - 1) mpi_shim::pending().clear();   // last request of a batch
    return MPI_SUCCESS;
}
