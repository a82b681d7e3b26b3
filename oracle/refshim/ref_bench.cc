// oracle/refshim/ref_bench.cc - BASELINE INFRASTRUCTURE.  Times the reference's OWN hot-path sources on the host cores:
// Grid_FFT<real_t> (src/grid_fft.cc), OrszagConvolver<real_t> (include/convolution.hh) and op::* compiled where they lie
// under /root/reference (non-MPI build, OpenMP on) with the FFTW calls served by fftw_fast.cc.  One step = the numerical
// body of ic_generator::run for a DM-only sc run: the source stage src/ic_generator.cc:475-574 term by term through the
// reference's convolver (5 products at 2LPT, 33 at 3LPT), the growth-factor scaling :561-574, and the six displacement /
// velocity assemblies + backward transforms :832-942 (written here as the same per-mode loops, with Grid_FFT::gradient).
// RNG, cosmology, transfer functions, particle read-out and output are outside (they need GSL / HDF5).
//
//   ref_bench <n> <order> <steps> <warmup> [<dumpdir>]     prints "seconds_per_step <t>"; with dumpdir also writes the input
//   phi (k-space) and the resulting phi2 / phi3 / A3x as complex128 [n][n][n/2+1] for the check against the oracle.
#include <chrono>
#include <cstdio>
#include <string>
#include <vector>

#include <general.hh>
#include <grid_fft.hh>
#include <convolution.hh>
#include <operators.hh>

std::map<cosmo_species, std::string> cosmo_species_name = {{cosmo_species::dm, "dm"}, {cosmo_species::baryon, "baryon"},
                                                           {cosmo_species::neutrino, "neutrino"}};
namespace CONFIG {
int MPI_thread_support = -1;
int MPI_task_rank = 0;
int MPI_task_size = 1;
bool MPI_ok = false;
bool MPI_threads_ok = false;
bool FFTW_threads_ok = true;
int num_threads = 1;
}  // namespace CONFIG
size_t global_mem_high_mark, local_mem_high_mark;
extern "C" {
const char* GIT_TAG = "oracle-shim";
const char* GIT_REV = "oracle-shim";
const char* GIT_BRANCH = "oracle-shim";
}
using grid_t = Grid_FFT<real_t>;

static void dump(grid_t& g, const std::string& dir, const std::string& name) {
    FILE* f = std::fopen((dir + "/" + name + ".bin").c_str(), "wb");
    const size_t n0 = g.n_[0], n1 = g.n_[1], n2 = g.n_[2];
    for (size_t i = 0; i < n0; ++i)
        for (size_t j = 0; j < n1; ++j)
            for (size_t k = 0; k < n2 / 2 + 1; ++k) { ccomplex_t v = g.kelem(i, j, k); std::fwrite(&v, sizeof(v), 1, f); }
    std::fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 5) { std::printf("usage: ref_bench n order steps warmup [dumpdir]\n"); return 1; }
    const size_t n = std::strtoul(argv[1], nullptr, 10);
    const int order = std::atoi(argv[2]), steps = std::atoi(argv[3]), warmup = std::atoi(argv[4]);
    const std::string dir = argc > 5 ? argv[5] : "";
    const real_t L = 100.0, D = 0.0257, vfac = 396.69;
    const std::array<size_t, 3> N{n, n, n};
    const std::array<real_t, 3> B{L, L, L};
    grid_t phi0(N, B), phi(N, B), phi2(N, B), phi3(N, B), A3x(N, B), A3y(N, B), A3z(N, B), tmp(N, B);
    grid_t* A3[3] = {&A3x, &A3y, &A3z};
    // a smooth random first-order potential: white noise -> k-space -> / (1 + k^2)^2, Nyquist planes and DC zero (NGenIC-like)
    phi0.FourierTransformBackward(false);
    unsigned long long s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        phi0.relem(i, j, k) = real_t((s >> 11) * (1.0 / 9007199254740992.0) - 0.5);
    }
    phi0.FourierTransformForward();
    phi0.apply_function_k_dep([&](auto x, auto k) -> ccomplex_t { const real_t k2 = k.norm_squared(); return x / ((1.0 + k2) * (1.0 + k2)); });
    for (size_t i = 0; i < phi0.size(0); ++i) for (size_t j = 0; j < phi0.size(1); ++j) for (size_t k = 0; k < phi0.size(2); ++k)
        if (i == n / 2 || j == n / 2 || k == n / 2) phi0.kelem(i, j, k) = 0.0;
    phi0.zero_DC_mode();

    OrszagConvolver<real_t> Conv(N, B);
    struct H2 { int a, b; };
    double total = 0.0;
    for (int it = 0; it < warmup + steps; ++it) {
        const auto t0 = std::chrono::steady_clock::now();
        phi.FourierTransformForward(false);
        phi.copy_from(phi0);
        phi2.FourierTransformForward(false);
        phi3.FourierTransformForward(false);
        for (auto* a : A3) a->FourierTransformForward(false);
        if (order > 1) {   // src/ic_generator.cc:482-499
            Conv.convolve_SumOfHessians(phi, {0, 0}, phi, {1, 1}, {2, 2}, op::assign_to(phi2));
            Conv.convolve_Hessians(phi, {1, 1}, phi, {2, 2}, op::add_to(phi2));
            for (H2 h : {H2{0, 1}, H2{0, 2}, H2{1, 2}}) Conv.convolve_Hessians(phi, {h.a, h.b}, phi, {h.a, h.b}, op::subtract_from(phi2));
            phi2.apply_InverseLaplacian();
        }
        if (order > 2) {   // :522-556
            Conv.convolve_Hessians(phi, {0, 0}, phi, {1, 1}, phi, {2, 2}, op::assign_to(phi3));
            Conv.convolve_Hessians(phi, {0, 1}, phi, {0, 2}, phi, {1, 2}, op::multiply_add_to(phi3, 2.0));
            for (int d = 0; d < 3; ++d) {
                const int p = (d + 1) % 3, q = (d + 2) % 3;
                Conv.convolve_Hessians(phi, {p, q}, phi, {p, q}, phi, {d, d}, op::subtract_from(phi3));
            }
            for (int d = 0; d < 3; ++d) {
                const int p = (d + 1) % 3, q = (d + 2) % 3;
                Conv.convolve_SumOfHessians(phi, {d, d}, phi2, {p, p}, {q, q}, op::multiply_add_to(phi3, -5.0 / 7.0));
            }
            for (H2 h : {H2{0, 1}, H2{0, 2}, H2{1, 2}}) Conv.convolve_Hessians(phi, {h.a, h.b}, phi2, {h.a, h.b}, op::multiply_add_to(phi3, 10.0 / 7.0));
            phi3.apply_InverseLaplacian();
            for (int d = 0; d < 3; ++d) {
                const int p = (d + 1) % 3, q = (d + 2) % 3;
                Conv.convolve_Hessians(phi2, {d, p}, phi, {d, q}, op::assign_to(*A3[d]));
                Conv.convolve_Hessians(phi2, {d, q}, phi, {d, p}, op::subtract_from(*A3[d]));
                Conv.convolve_DifferenceOfHessians(phi, {p, q}, phi2, {p, p}, {q, q}, op::add_to(*A3[d]));
                Conv.convolve_DifferenceOfHessians(phi2, {p, q}, phi, {p, p}, {q, q}, op::subtract_from(*A3[d]));
                A3[d]->apply_InverseLaplacian();
            }
        }
        if (it == 0 && !dir.empty()) {   // unscaled potentials for the check against the oracle
            dump(phi, dir, "phi");
            if (order > 1) dump(phi2, dir, "phi2");
            if (order > 2) { dump(phi3, dir, "phi3"); dump(A3x, dir, "A3x"); }
        }
        phi *= -D;                                   // :561-574
        if (order > 1) phi2 *= -3.0 / 7.0 * D * D;
        if (order > 2) { phi3 *= 1.0 / 3.0 * D * D * D; for (auto* a : A3) (*a) *= 1.0 / 7.0 * D * D * D; }
        // the six component fields :832-942 (sc lattice: no shift FFTs)
        for (int kind = 0; kind < 2; ++kind) {
            const real_t w1 = kind ? vfac : 1.0, w2 = kind ? 2 * vfac : 1.0, w3 = kind ? 3 * vfac : 1.0;
            for (int d = 0; d < 3; ++d) {
                const int p = (d + 1) % 3, q = (d + 2) % 3;
                tmp.FourierTransformForward(false);
#pragma omp parallel for
                for (size_t i = 0; i < phi.size(0); ++i)
                    for (size_t j = 0; j < phi.size(1); ++j)
                        for (size_t k = 0; k < phi.size(2); ++k) {
                            const size_t idx = phi.get_idx(i, j, k);
                            auto tot = w1 * phi.kelem(idx);
                            if (order > 1) tot += w2 * phi2.kelem(idx);
                            if (order > 2) tot += w3 * phi3.kelem(idx);
                            tmp.kelem(idx) = phi.gradient(d, {i, j, k}) * tot;
                            if (order > 2)
                                tmp.kelem(idx) += w3 * (phi.gradient(p, {i, j, k}) * A3[q]->kelem(idx) - phi.gradient(q, {i, j, k}) * A3[p]->kelem(idx));
                        }
                tmp.zero_DC_mode();
                tmp.FourierTransformBackward();
            }
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (it >= warmup) total += dt;
    }
    std::printf("seconds_per_step %.6f\n", total / steps);
    std::printf("checksum %.17g\n", (double)tmp.std());
    return 0;
}
