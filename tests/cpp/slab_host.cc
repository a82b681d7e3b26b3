// tests/cpp/slab_host.cc - a C++ host with ONE PROCESS PER RANK driving slab-decomposed grids through the drop-in headers, the
// way the reference's USE_MPI build does (src/grid_fft.cc:115-192: real space split along axis 0, k-space transposed and
// split along axis 1; Grid_FFT<real_t, true>).  Built twice from this source (tests/test_cpp_slab_host.py):
//   slab_host_ref   ref  <out.bin>                              whole grid in one process: the values every slab must reproduce
//   slab_host_rank  rank <r> <nranks> <shm file> <ref.bin>      -DGRID_FFT_DISTRIBUTED=true: this process is rank r
// Both run the same host code - a "plugin" filling the noise through relem on the local planes, phi(1) by a host lambda over
// the local k-space range (get_k / is_nyquist_mode in the transposed layout), the phi(2) source with the call sequence of
// src/ic_generator.cc:482-486, apply_InverseLaplacian, a gradient by host lambda, the reductions mean()/std() - and the ranks
// compare their planes with the reference's.  The exchange of the slab transposes goes through the library's staged mode
// (mfb_dist_init_staged): a host callback that here moves the chunks through a shared-memory file with a spin barrier -
// what MPI_Alltoall would do in an MPI host.  CPU thread-emulation build only (the callback sees host pointers).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <vector>

#include "../../include/mfb/convolution.hh"

#define CHECK(c)                                                                              \
    do {                                                                                      \
        if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); std::exit(1); }     \
    } while (0)

static const size_t n = 32;
static double noise(size_t i, size_t j, size_t k) {   // a fixed function of the GLOBAL cell index
    unsigned long long x = ((i * n + j) * n + k + 1) * 6364136223846793005ull + 1442695040888963407ull;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29;
    return double(int(x & 0xfffff) - 524288) / 524288.0;
}

// ---- exchange through a shared-memory file (ranks are separately started processes)
struct Shm {
    std::atomic<int> arrived, generation;
    char pad[56];
    char data[1];
};
static Shm* g_shm = nullptr;
static int g_rank = 0, g_nranks = 1;
static size_t g_slot_bytes = 0;
static void barrier() {
    const int gen = g_shm->generation.load();
    if (g_shm->arrived.fetch_add(1) + 1 == g_nranks) { g_shm->arrived.store(0); g_shm->generation.fetch_add(1); }
    else while (g_shm->generation.load() == gen) usleep(50);
}
static void exchange(const void* send, void* recv, size_t chunk_bytes, void*) {
    CHECK(chunk_bytes * g_nranks <= g_slot_bytes);
    std::memcpy(g_shm->data + size_t(g_rank) * g_slot_bytes, send, chunk_bytes * g_nranks);
    barrier();
    for (int r = 0; r < g_nranks; ++r)   // chunk g_rank of rank r's send buffer is chunk r of my receive buffer
        std::memcpy(static_cast<char*>(recv) + size_t(r) * chunk_bytes, g_shm->data + size_t(r) * g_slot_bytes + size_t(g_rank) * chunk_bytes, chunk_bytes);
    barrier();
}

using grid = Grid_FFT<real_t>;   // <real_t, GRID_FFT_DISTRIBUTED>

// the host code under test; returns [dphi2/dx planes | phi2-source planes] of the local slab and two reductions
static std::vector<double> run_host(double stats[2]) {
    std::array<size_t, 3> N{n, n, n};
    std::array<real_t, 3> L{100.0, 100.0, 100.0};
    grid w(N, L), phi(N, L, true, kspace_id), phi2(N, L, true, kspace_id), gx(N, L, true, kspace_id);
    const size_t l0 = size_t(w.local_0_size_), o0 = size_t(w.local_0_start_);
    CHECK(w.size(0) == l0 && w.is_distributed() == GRID_FFT_DISTRIBUTED);
    for (size_t i = 0; i < l0; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) w.relem(i, j, k) = noise(o0 + i, j, k);
    stats[0] = w.mean(); stats[1] = w.std();
    w.FourierTransformForward();
    // phi = -w / k^2, Nyquist planes zeroed: a host lambda over the LOCAL k-space range (transposed layout on slabs)
    phi.FourierTransformForward(false);
    for (size_t i = 0; i < phi.size(0); ++i) for (size_t j = 0; j < phi.size(1); ++j) for (size_t k = 0; k < phi.size(2); ++k) {
        const auto kk = phi.get_k<real_t>(i, j, k);
        const double k2 = kk.norm_squared();
        phi.kelem(i, j, k) = (k2 > 0 && !phi.is_nyquist_mode(i, j, k)) ? -w.kelem(i, j, k) / k2 : 0.0;
    }
    OrszagConvolver<real_t> Conv(N, L);
    Conv.convolve_SumOfHessians(phi, {0, 0}, phi, {1, 1}, {2, 2}, op::assign_to(phi2));   // src/ic_generator.cc:482-486
    Conv.convolve_Hessians(phi, {1, 1}, phi, {2, 2}, op::add_to(phi2));
    Conv.convolve_Hessians(phi, {0, 1}, phi, {0, 1}, op::subtract_from(phi2));
    Conv.convolve_Hessians(phi, {0, 2}, phi, {0, 2}, op::subtract_from(phi2));
    Conv.convolve_Hessians(phi, {1, 2}, phi, {1, 2}, op::subtract_from(phi2));
    gx.FourierTransformForward(false);
    gx.copy_from(phi2);
    phi2.apply_InverseLaplacian();
    for (size_t i = 0; i < phi2.size(0); ++i) for (size_t j = 0; j < phi2.size(1); ++j) for (size_t k = 0; k < phi2.size(2); ++k)
        phi2.kelem(i, j, k) *= phi2.gradient(0, {i, j, k});                                 // d/dx by the host gradient operator
    phi2.FourierTransformBackward();
    gx.FourierTransformBackward();
    std::vector<double> out(2 * l0 * n * n);
    for (size_t i = 0; i < l0; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) {
        out[(i * n + j) * n + k] = phi2.relem(i, j, k);
        out[l0 * n * n + (i * n + j) * n + k] = gx.relem(i, j, k);
    }
    return out;
}

int main(int argc, char** argv) {
    if (argc >= 3 && std::strcmp(argv[1], "ref") == 0) {
        CHECK(!GRID_FFT_DISTRIBUTED);
        CHECK(mfb_init(0) == 0);
        double st[2];
        std::vector<double> v = run_host(st);
        FILE* f = std::fopen(argv[2], "wb");
        CHECK(f && std::fwrite(st, 8, 2, f) == 2 && std::fwrite(v.data(), 8, v.size(), f) == v.size());
        std::fclose(f);
        std::printf("slab host reference ok (%zu values)\n", v.size());
        return 0;
    }
    CHECK(argc >= 6 && std::strcmp(argv[1], "rank") == 0 && GRID_FFT_DISTRIBUTED);
    g_rank = std::atoi(argv[2]); g_nranks = std::atoi(argv[3]);
    CHECK(mfb_init(0) == 0);
    int nn[3] = {int(n), int(n), int(n)};
    const size_t xbuf = mfb_dist_xbuf_bytes(nn, 1, g_nranks);
    g_slot_bytes = xbuf;
    const size_t shm_bytes = sizeof(Shm) + size_t(g_nranks) * g_slot_bytes;
    const int fd = open(argv[4], O_RDWR);
    CHECK(fd >= 0 && size_t(lseek(fd, 0, SEEK_END)) >= shm_bytes);   // the launcher made the (sparse, zero-filled) file large enough
    g_shm = static_cast<Shm*>(mmap(nullptr, shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0));
    CHECK(g_shm != MAP_FAILED);
    if (mfb_dist_init_staged(g_rank, g_nranks, xbuf, exchange, nullptr) != 0) { std::printf("mfb_dist_init_staged: %s\n", mfb_last_error()); return 2; }
    double st[2];
    std::vector<double> v = run_host(st);
    // compare with the whole-grid run: this rank's planes
    FILE* f = std::fopen(argv[5], "rb");
    CHECK(f);
    double rst[2];
    std::vector<double> ref(2 * n * n * n);
    CHECK(std::fread(rst, 8, 2, f) == 2 && std::fread(ref.data(), 8, ref.size(), f) == ref.size());
    std::fclose(f);
    const size_t l0 = n / g_nranks, o0 = g_rank * l0, slab = l0 * n * n, full = n * n * n;
    double emax[2] = {0, 0}, rmax[2] = {0, 0};
    for (int part = 0; part < 2; ++part)
        for (size_t q = 0; q < slab; ++q) {
            const double r = ref[part * full + o0 * n * n + q];
            emax[part] = std::max(emax[part], std::fabs(v[part * slab + q] - r));
            rmax[part] = std::max(rmax[part], std::fabs(r));
        }
    std::printf("rank %d: grad phi2 err %.2e of %.2e, phi2 source err %.2e of %.2e, mean %.3e / %.3e, std %.6f / %.6f, exchanges %lld\n", g_rank,
                emax[0], rmax[0], emax[1], rmax[1], st[0], rst[0], st[1], rst[1], mfb_dist_exchanges());
    CHECK(rmax[0] > 0 && rmax[1] > 0 && emax[0] < 1e-12 * rmax[0] && emax[1] < 1e-12 * rmax[1]);
    CHECK(std::fabs(st[0] - rst[0]) < 1e-14 && std::fabs(st[1] - rst[1]) < 1e-13);   // reductions are global on every rank
    CHECK(mfb_dist_exchanges() > 0 && mfb_dist_check() == 0);
    barrier();
    std::printf("slab host rank %d ok\n", g_rank);
    return 0;
}
