// tests/cpp/minimpi/minimpi.cc - TEST INFRASTRUCTURE: see mpi.h.  Shared file layout: header (barrier state) + one slot per rank.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include "mpi.h"

namespace {
struct Header { std::atomic<int> arrived, generation; char pad[56]; };
Header* g_hdr = nullptr;
char* g_slots = nullptr;
size_t g_slot = 0;
int g_rank = 0, g_size = 1;

void die(const char* what) { std::fprintf(stderr, "minimpi rank %d: %s\n", g_rank, what); std::abort(); }
void barrier() {
    if (g_size == 1) return;
    const int gen = g_hdr->generation.load();
    if (g_hdr->arrived.fetch_add(1) + 1 == g_size) { g_hdr->arrived.store(0); g_hdr->generation.fetch_add(1); }
    else while (g_hdr->generation.load() == gen) usleep(20);
}
size_t tsize(MPI_Datatype t) {
    switch (t) {
        case MPI_CHAR: case MPI_BYTE: return 1;
        case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
        case MPI_DOUBLE: case MPI_LONG_LONG: case MPI_UNSIGNED_LONG: case MPI_UNSIGNED_LONG_LONG: case MPI_C_FLOAT_COMPLEX: return 8;
        case MPI_LONG_DOUBLE: case MPI_C_DOUBLE_COMPLEX: return 16;
        case MPI_C_LONG_DOUBLE_COMPLEX: return 32;
    }
    die("unknown datatype");
    return 0;
}
char* slot(int r) { return g_slots + size_t(r) * g_slot; }
template <typename T> void reduce(T* out, int count, MPI_Op op) {
    for (int i = 0; i < count; ++i) {
        T acc = reinterpret_cast<const T*>(slot(0))[i];
        for (int r = 1; r < g_size; ++r) {
            const T v = reinterpret_cast<const T*>(slot(r))[i];
            acc = op == MPI_SUM ? T(acc + v) : (op == MPI_MAX ? std::max(acc, v) : std::min(acc, v));
        }
        out[i] = acc;
    }
}
}  // namespace

extern "C" {
int MPI_Init_thread(int*, char***, int required, int* provided) {
    const char *r = std::getenv("MINIMPI_RANK"), *s = std::getenv("MINIMPI_SIZE"), *f = std::getenv("MINIMPI_SHM"), *m = std::getenv("MINIMPI_SLOT_MB");
    g_rank = r ? std::atoi(r) : 0; g_size = s ? std::atoi(s) : 1;
    g_slot = size_t(m ? std::atoi(m) : 64) << 20;
    *provided = required;
    if (g_size == 1) return MPI_SUCCESS;
    if (!f) die("MINIMPI_SHM is not set");
    const int fd = open(f, O_RDWR);
    const size_t bytes = sizeof(Header) + size_t(g_size) * g_slot;
    if (fd < 0 || size_t(lseek(fd, 0, SEEK_END)) < bytes) die("shared file missing or too small");
    void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (p == MAP_FAILED) die("mmap failed");
    g_hdr = static_cast<Header*>(p);
    g_slots = static_cast<char*>(p) + sizeof(Header);
    barrier();
    return MPI_SUCCESS;
}
int MPI_Finalize(void) { barrier(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = g_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int* size) { *size = g_size; return MPI_SUCCESS; }
double MPI_Wtime(void) { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int MPI_Barrier(MPI_Comm) { barrier(); return MPI_SUCCESS; }
int MPI_Get_library_version(char* v, int* len) { std::strcpy(v, "minimpi (shared-memory test stand-in)"); *len = int(std::strlen(v)); return MPI_SUCCESS; }
int MPI_Error_string(int, char* s, int* len) { s[0] = 0; *len = 0; return MPI_SUCCESS; }

int MPI_Allreduce(const void* send, void* recv, int count, MPI_Datatype t, MPI_Op op, MPI_Comm) {
    const size_t bytes = size_t(count) * tsize(t);
    const void* src = (send == MPI_IN_PLACE) ? recv : send;
    if (g_size == 1) { if (src != recv) std::memcpy(recv, src, bytes); return MPI_SUCCESS; }
    if (bytes > g_slot) die("Allreduce larger than a slot");
    std::memcpy(slot(g_rank), src, bytes);
    barrier();
    switch (t) {
        case MPI_DOUBLE: reduce(static_cast<double*>(recv), count, op); break;
        case MPI_FLOAT: reduce(static_cast<float*>(recv), count, op); break;
        case MPI_INT: reduce(static_cast<int*>(recv), count, op); break;
        case MPI_UNSIGNED: reduce(static_cast<unsigned*>(recv), count, op); break;
        case MPI_LONG_LONG: reduce(static_cast<long long*>(recv), count, op); break;
        case MPI_UNSIGNED_LONG: case MPI_UNSIGNED_LONG_LONG: reduce(static_cast<unsigned long long*>(recv), count, op); break;
        case MPI_LONG_DOUBLE: reduce(static_cast<long double*>(recv), count, op); break;
        default: die("Allreduce: datatype not supported");
    }
    barrier();
    return MPI_SUCCESS;
}
int MPI_Allgather(const void* send, int scount, MPI_Datatype st, void* recv, int, MPI_Datatype, MPI_Comm) {
    const size_t bytes = size_t(scount) * tsize(st);
    const void* src = (send == MPI_IN_PLACE) ? static_cast<char*>(recv) + size_t(g_rank) * bytes : send;
    if (bytes > g_slot) die("Allgather larger than a slot");
    if (g_size == 1) { if (src != recv) std::memcpy(recv, src, bytes); return MPI_SUCCESS; }
    std::memcpy(slot(g_rank), src, bytes);
    barrier();
    for (int r = 0; r < g_size; ++r) std::memcpy(static_cast<char*>(recv) + size_t(r) * bytes, slot(r), bytes);
    barrier();
    return MPI_SUCCESS;
}
int MPI_Alltoall(const void* send, int scount, MPI_Datatype st, void* recv, int, MPI_Datatype, MPI_Comm) {
    const size_t chunk = size_t(scount) * tsize(st);
    if (chunk * g_size > g_slot) die("Alltoall larger than a slot (raise MINIMPI_SLOT_MB)");
    if (g_size == 1) { if (send != recv) std::memcpy(recv, send, chunk); return MPI_SUCCESS; }
    std::memcpy(slot(g_rank), send, chunk * g_size);
    barrier();
    for (int r = 0; r < g_size; ++r) std::memcpy(static_cast<char*>(recv) + size_t(r) * chunk, slot(r) + size_t(g_rank) * chunk, chunk);
    barrier();
    return MPI_SUCCESS;
}
int MPI_Bcast(void* buf, int count, MPI_Datatype t, int root, MPI_Comm) {
    const size_t bytes = size_t(count) * tsize(t);
    if (g_size == 1) return MPI_SUCCESS;
    if (bytes > g_slot) die("Bcast larger than a slot");
    if (g_rank == root) std::memcpy(slot(root), buf, bytes);
    barrier();
    if (g_rank != root) std::memcpy(buf, slot(root), bytes);
    barrier();
    return MPI_SUCCESS;
}
/* point-to-point: only a rank's messages to ITSELF (what src/grid_fft.cc's FourierInterpolateCopyTo and grid_ghosts.hh do on one
   rank - the pure-CPU reference executable of tests/cpp/Makefile); a send is kept until the matching receive */
struct Pending { const void* buf; size_t bytes; int tag; bool done; };
static std::vector<Pending> g_pending;
int MPI_Isend(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm, MPI_Request* req) {
    if (dest != g_rank) die("point-to-point between different ranks is not part of this stand-in");
    g_pending.push_back({buf, size_t(count) * tsize(t), tag, false});
    *req = int(g_pending.size()) - 1;
    return MPI_SUCCESS;
}
int MPI_Send(const void* buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c) { MPI_Request r; return MPI_Isend(buf, count, t, dest, tag, c, &r); }
int MPI_Recv(void* buf, int count, MPI_Datatype t, int source, int tag, MPI_Comm, MPI_Status* st) {
    if (source != g_rank) die("point-to-point between different ranks is not part of this stand-in");
    for (auto& p : g_pending)
        if (!p.done && p.tag == tag) {
            std::memcpy(buf, p.buf, std::min(p.bytes, size_t(count) * tsize(t)));
            p.done = true;
            if (st) { st->MPI_SOURCE = g_rank; st->MPI_TAG = tag; st->MPI_ERROR = MPI_SUCCESS; }
            return MPI_SUCCESS;
        }
    die("MPI_Recv without a matching send");
    return 1;
}
/* the one point-to-point pattern of the host code outside grid_fft.cc: a ring shift every rank takes part in (include/grid_interpolate.hh:78-85) */
int MPI_Sendrecv_replace(void* buf, int count, MPI_Datatype t, int, int, int source, int, MPI_Comm, MPI_Status* st) {
    const size_t bytes = size_t(count) * tsize(t);
    if (bytes > g_slot) die("Sendrecv_replace larger than a slot");
    if (g_size > 1) {
        std::memcpy(slot(g_rank), buf, bytes);
        barrier();
        std::memcpy(buf, slot(source), bytes);
        barrier();
    }
    if (st) { st->MPI_SOURCE = source; st->MPI_TAG = 0; st->MPI_ERROR = MPI_SUCCESS; }
    return MPI_SUCCESS;
}
int MPI_Test(MPI_Request* req, int* flag, MPI_Status*) { *flag = g_pending[*req].done ? 1 : 0; return MPI_SUCCESS; }
int MPI_Wait(MPI_Request* req, MPI_Status* st) {
    if (st) st->MPI_ERROR = g_pending[*req].done ? MPI_SUCCESS : 1;
    bool all = true;
    for (auto& p : g_pending) all = all && p.done;
    if (all && *req == int(g_pending.size()) - 1) g_pending.clear();   // last request of a batch
    return MPI_SUCCESS;
}
int MPI_Alltoallv(const void*, const int*, const int*, MPI_Datatype, void*, const int*, const int*, MPI_Datatype, MPI_Comm) { die("MPI_Alltoallv is not part of this stand-in"); return 1; }
int MPI_Exscan(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm) { die("MPI_Exscan is not part of this stand-in"); return 1; }
int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op op, int, MPI_Comm c) { return MPI_Allreduce(s, r, n, t, op, c); }   /* every rank gets the result */
}
