// oracle/refshim/ref_driver.cc - TEST INFRASTRUCTURE.  Runs the reference's OWN Grid_FFT<real_t>, OrszagConvolver<real_t> and
// op::* (compiled from /root/reference/include/{grid_fft,convolution,operators}.hh and src/grid_fft.cc against the shims in
// this directory) on a given first-order potential and dumps what they produce, so that the CPU oracle - and through the
// committed fixtures the CUDA path - can be pinned against the reference itself for the quantities its goldens do not cover
// (phi2, phi3, A3, every named product, FFT normalisation, Poisson solve, shift_field, FourierInterpolateCopyTo).
// Two builds (oracle/Makefile): ref_driver is the reference's non-MPI build, where pad_insert keeps the +Nyquist planes
// (include/convolution.hh:575-592) - it agrees with the MPI build for Nyquist-free input only; ref_driver_mpi is the
// USE_MPI build on one rank (mpi.h / fftw3-mpi.h shims): transposed k-layout, pad / unpad through FourierInterpolateCopyTo
// with the Nyquist planes dropped (:594-608, :657-665) - the code path the reference's CI builds and its goldens come from,
// and the semantics the CUDA path implements for Nyquist-carrying noise (MUSIC1, Panphasia).
//
//   ref_driver <n> <boxlen> <Dplus0> <phi_k.bin> <psi_k.bin> <outdir> [<phi_nyq_k.bin> <psi_nyq_k.bin> <cplx_r.bin>]
//   phi_k / psi_k: complex128 [n][n][n/2+1] half-spectra (kelem layout), Nyquist-free;
//   optional: two Nyquist-CARRYING half-spectra for the NaiveConvolver (which has no MPI/non-MPI ambiguity: it never pads)
//   and a complex128 [n][n][n] real-space field for Grid_FFT<ccomplex_t> (the c2c transforms of the PPT branch)
#include <cstdio>
#include <string>
#include <vector>

#include <general.hh>
#include <grid_fft.hh>
#include <convolution.hh>
#include <operators.hh>
#include <grid_interpolate.hh>
#if defined(USE_MPI)   // the reference's particle generator only compiles in its MPI build (the non-MPI branches of
#define USE_HDF5       // grid_ghosts.hh / glass_loader.hh use MPI unguarded), and its glass loader wants HDF_IO.hh
#include <particle_container.hh>
#include <particle_generator.hh>
#endif

std::map<cosmo_species, std::string> cosmo_species_name = {{cosmo_species::dm, "dm"}, {cosmo_species::baryon, "baryon"},
                                                           {cosmo_species::neutrino, "neutrino"}};
// globals the reference defines in src/main.cc (which is not part of this build)
namespace CONFIG {
int MPI_thread_support = -1;
int MPI_task_rank = 0;
int MPI_task_size = 1;
bool MPI_ok = false;
bool MPI_threads_ok = false;
bool FFTW_threads_ok = false;
int num_threads = 1;
}  // namespace CONFIG
size_t global_mem_high_mark, local_mem_high_mark;
extern "C" {
const char* GIT_TAG = "oracle-shim";
const char* GIT_REV = "oracle-shim";
const char* GIT_BRANCH = "oracle-shim";
}

using grid_t = Grid_FFT<real_t>;

// Files hold half-spectra as [kx][ky][kz] and real fields as [x][y][z].  In the USE_MPI build k-space is stored transposed
// ([ky_loc][kx][kz], src/grid_fft.cc:125-128): kelem(i,j,k) then takes i = ky, j = kx.
#if defined(USE_MPI)
static constexpr bool k_transposed = true;
#else
static constexpr bool k_transposed = false;
#endif
template <class G> static ccomplex_t& kref(G& g, size_t kx, size_t ky, size_t kz) {
    return k_transposed ? g.kelem(ky, kx, kz) : g.kelem(kx, ky, kz);
}
static bool load_k(grid_t& g, const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    g.FourierTransformForward(false);
    const size_t n0 = g.n_[0], n1 = g.n_[1], nzc = g.n_[2] / 2 + 1;
    std::vector<ccomplex_t> row(nzc);
    bool ok = true;
    for (size_t i = 0; i < n0 && ok; ++i)
        for (size_t j = 0; j < n1 && ok; ++j) {
            ok = std::fread(row.data(), sizeof(ccomplex_t), row.size(), f) == row.size();
            for (size_t k = 0; k < nzc; ++k) kref(g, i, j, k) = row[k];
        }
    std::fclose(f);
    return ok;
}
static void dump(grid_t& g, const std::string& dir, const std::string& name) {
    FILE* f = std::fopen((dir + "/" + name + ".bin").c_str(), "wb");
    const size_t n0 = g.n_[0], n1 = g.n_[1], n2 = g.n_[2];
    if (g.space_ == kspace_id) {
        for (size_t i = 0; i < n0; ++i)
            for (size_t j = 0; j < n1; ++j)
                for (size_t k = 0; k < n2 / 2 + 1; ++k) { ccomplex_t v = kref(g, i, j, k); std::fwrite(&v, sizeof(v), 1, f); }
    } else {
        for (size_t i = 0; i < n0; ++i)
            for (size_t j = 0; j < n1; ++j)
                for (size_t k = 0; k < n2; ++k) { real_t v = g.relem(i, j, k); std::fwrite(&v, sizeof(v), 1, f); }
    }
    std::fclose(f);
}

#if defined(USE_MPI)
// particle::lattice_generator of the reference (include/particle_generator.hh:129-170, 292-350, 422-467, 520-562) on the
// fields phi (positions, masses) and psi (velocities) in real space:
//   ref_driver_mpi particles <n> <boxlen> <phi_k.bin> <psi_k.bin> <outdir> <lattice_type> <lattice_index> <zero_field 0|1>
//                            [<DoBaryons yes|no> <ParticleMaskType>]      (lattice_type -2: the masked sc lattice, :171-260)
static int run_particles(int argc, char** argv) {
    if (argc < 10) return 1;
    const size_t n = std::strtoul(argv[2], nullptr, 10);
    const real_t L = std::atof(argv[3]);
    const std::string dir = argv[6];
    const auto lt = particle::lattice(std::atoi(argv[7]));
    const int lidx = std::atoi(argv[8]);
    const bool zero = std::atoi(argv[9]) != 0;
    const std::array<size_t, 3> N{n, n, n};
    const std::array<real_t, 3> B{L, L, L};
    grid_t phi(N, B), psi(N, B), tmp(N, B);
    if (!load_k(phi, argv[4]) || !load_k(psi, argv[5])) return 2;
    {
        FILE* f = std::fopen((dir + "/empty.conf").c_str(), "w");
        std::fprintf(f, "[setup]\nDoBaryons = %s\nParticleMaskType = %s\n", argc > 10 ? argv[10] : "no", argc > 11 ? argv[11] : "3");
        std::fclose(f);
    }
    config_file cf(dir + "/empty.conf");
    const real_t lunit = L, munit = 2.5;
    particle::lattice_generator<grid_t> lg(lt, lidx, true, true, true, size_t(lidx), tmp, cf);
    auto fresh = [&](grid_t& src) {
        tmp.FourierTransformForward(false);
        tmp.copy_from(src);
        if (zero) tmp.zero();
        tmp.FourierTransformBackward();
    };
    for (int d = 0; d < 3; ++d) { fresh(phi); lg.set_positions(lt, lidx, d, lunit, true, tmp, cf); }
    for (int d = 0; d < 3; ++d) { fresh(psi); lg.set_velocities(lt, lidx, d, true, tmp, cf); }
    fresh(phi);
    lg.set_masses(lt, lidx, munit, true, tmp, cf);
    const particle::container& pc = lg.get_particles();
    const uint64_t np = pc.get_local_num_particles();
    FILE* f = std::fopen((dir + "/particles.bin").c_str(), "wb");
    std::fwrite(&np, sizeof(np), 1, f);
    std::fwrite(pc.get_pos64_ptr(), sizeof(double), 3 * np, f);
    std::fwrite(pc.get_vel64_ptr(), sizeof(double), 3 * np, f);
    std::fwrite(pc.get_mass64_ptr(), sizeof(double), np, f);
    std::fwrite(pc.get_ids64_ptr(), sizeof(uint64_t), np, f);
    std::fclose(f);
    std::printf("ref_driver particles ok\n");
    return 0;
}
#endif

#if !defined(USE_MPI)
// The interpolation building blocks of the PLT gradient and of glass loads, from the reference's own headers:
//   ref_driver cic <n> <boxlen> <phi_k.bin> <points.bin> <npts> <out.bin>
// points: npts x 3 doubles in index units.  Per point: Re/Im Grid_FFT::get_cic_kspace (include/grid_fft.hh:347-373; z is
// clamped into the half-spectrum by the caller), the three components of get_k<real_t>(x,y,z) (:299-313), then - on the
// backward-transformed field - grid_interpolate<1>::get_cic_at (include/grid_interpolate.hh:103-148) and
// compensation_kernel(get_k(floor(x),floor(y),floor(z))) (:195-208): 7 doubles.
static int run_cic(int argc, char** argv) {
    if (argc < 8) return 1;
    const size_t n = std::strtoul(argv[2], nullptr, 10);
    const real_t L = std::atof(argv[3]);
    const size_t npts = std::strtoul(argv[6], nullptr, 10);
    grid_t phi({n, n, n}, {L, L, L});
    if (!load_k(phi, argv[4])) return 2;
    std::vector<double> pts(3 * npts), out(7 * npts);
    { FILE* f = std::fopen(argv[5], "rb"); if (!f || std::fread(pts.data(), 8, pts.size(), f) != pts.size()) return 2; std::fclose(f); }
    for (size_t p = 0; p < npts; ++p) {
        const real_t x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
        const real_t zk = z * real_t(n / 2) / real_t(n);   // into [0, n/2)
        const ccomplex_t c = phi.get_cic_kspace({x, y, zk});
        const auto kv = phi.get_k<real_t>(x, y, zk);
        out[7 * p] = std::real(c); out[7 * p + 1] = std::imag(c);
        out[7 * p + 2] = kv[0]; out[7 * p + 3] = kv[1]; out[7 * p + 4] = kv[2];
    }
    phi.FourierTransformBackward();
    grid_interpolate<1, grid_t> interp(phi);
    for (size_t p = 0; p < npts; ++p) {
        const real_t x = pts[3 * p], y = pts[3 * p + 1], z = pts[3 * p + 2];
        out[7 * p + 5] = interp.get_cic_at({x, y, z});
        out[7 * p + 6] = std::real(interp.compensation_kernel(phi.get_k<real_t>(size_t(x), size_t(y), size_t(z) % (n / 2 + 1))));
    }
    FILE* f = std::fopen(argv[7], "wb");
    std::fwrite(out.data(), 8, out.size(), f);
    std::fclose(f);
    std::printf("ref_driver cic ok\n");
    return 0;
}
#endif

int main(int argc, char** argv) {
#if defined(USE_MPI)
    if (argc > 1 && std::string(argv[1]) == "particles") return run_particles(argc, argv);
#else
    if (argc > 1 && std::string(argv[1]) == "cic") return run_cic(argc, argv);
#endif
    if (argc < 7) { std::printf("usage: ref_driver n boxlen Dplus0 phi_k.bin psi_k.bin outdir\n"); return 1; }
    const size_t n = std::strtoul(argv[1], nullptr, 10);
    const real_t L = std::atof(argv[2]), D = std::atof(argv[3]);
    const std::string dir = argv[6];
    const std::array<size_t, 3> N{n, n, n};
    const std::array<real_t, 3> B{L, L, L};
    grid_t phi(N, B), psi(N, B), phi2(N, B), phi3(N, B), A3x(N, B), A3y(N, B), A3z(N, B), tmp(N, B);
    if (!load_k(phi, argv[4]) || !load_k(psi, argv[5])) { std::printf("cannot read the input spectra\n"); return 2; }

    // ---- transforms and k-space operators of Grid_FFT
    tmp.FourierTransformForward(false);
    tmp.copy_from(phi);
    tmp.FourierTransformBackward();
    dump(tmp, dir, "phi_real");
    std::printf("mean %.17g std %.17g\n", (double)tmp.mean(), (double)tmp.std());
    tmp.shift_field(vec3_t<real_t>({0.5, 0.25, 0.5}));
    dump(tmp, dir, "phi_real_shifted");
    tmp.FourierTransformForward();
    tmp.apply_InverseLaplacian();
    dump(tmp, dir, "invlap_shifted_phi");
    {
        const std::array<size_t, 3> Np{3 * n / 2, 3 * n / 2, 3 * n / 2};
        grid_t big(Np, B);
        big.FourierTransformForward(false);
        psi.FourierInterpolateCopyTo(big);
        dump(big, dir, "psi_interp_3half");
    }

    // ---- every named product of the convolver, on two different fields
    OrszagConvolver<real_t> Conv(N, B);
    tmp.FourierTransformForward(false);
    Conv.convolve_Hessians(phi, {0, 1}, psi, {1, 2}, op::assign_to(tmp));
    dump(tmp, dir, "hess_01_12");
    Conv.convolve_SumOfHessians(phi, {0, 0}, psi, {1, 1}, {2, 2}, op::assign_to(tmp));
    dump(tmp, dir, "sumhess_00_11_22");
    Conv.convolve_DifferenceOfHessians(phi, {1, 2}, psi, {1, 1}, {2, 2}, op::assign_to(tmp));
    dump(tmp, dir, "diffhess_12_11_22");
    Conv.convolve_Hessians(phi, {0, 0}, psi, {0, 2}, phi, {2, 2}, op::assign_to(tmp));
    dump(tmp, dir, "hess3_00_02_22");
    Conv.convolve_Gradients(phi, {0}, psi, {2}, op::assign_to(tmp));
    dump(tmp, dir, "grad_0_2");
    Conv.convolve_Gradient_and_Hessian(phi, {1}, psi, {0, 2}, op::assign_to(tmp));
    dump(tmp, dir, "gradhess_1_02");
    Conv.multiply_field(phi, psi, op::assign_to(tmp));
    dump(tmp, dir, "multiply");
    Conv.multiply_field(phi, psi, op::multiply_add_to(tmp, 0.5));     // the accumulating output operators
    Conv.multiply_field(phi, phi, op::subtract_from(tmp));
    Conv.multiply_field(psi, psi, op::add_to(tmp));
    dump(tmp, dir, "multiply_ops");

    // ---- the LPT source stage, src/ic_generator.cc:475-574, term by term through the reference's convolver
    phi2.FourierTransformForward(false);
    phi3.FourierTransformForward(false);
    grid_t* A3[3] = {&A3x, &A3y, &A3z};
    for (auto* a : A3) a->FourierTransformForward(false);
    struct H2 { int a, b; };
    Conv.convolve_SumOfHessians(phi, {0, 0}, phi, {1, 1}, {2, 2}, op::assign_to(phi2));
    Conv.convolve_Hessians(phi, {1, 1}, phi, {2, 2}, op::add_to(phi2));
    for (H2 h : {H2{0, 1}, H2{0, 2}, H2{1, 2}}) Conv.convolve_Hessians(phi, {h.a, h.b}, phi, {h.a, h.b}, op::subtract_from(phi2));
    phi2.apply_InverseLaplacian();
    // 3a: determinant of the Hessian
    Conv.convolve_Hessians(phi, {0, 0}, phi, {1, 1}, phi, {2, 2}, op::assign_to(phi3));
    Conv.convolve_Hessians(phi, {0, 1}, phi, {0, 2}, phi, {1, 2}, op::multiply_add_to(phi3, 2.0));
    for (int d = 0; d < 3; ++d) {
        const int p = (d + 1) % 3, q = (d + 2) % 3;   // (yz, yz | xx), (zx, zx | yy), (xy, xy | zz)
        Conv.convolve_Hessians(phi, {p, q}, phi, {p, q}, phi, {d, d}, op::subtract_from(phi3));
    }
    // 3b
    for (int d = 0; d < 3; ++d) {
        const int p = (d + 1) % 3, q = (d + 2) % 3;
        Conv.convolve_SumOfHessians(phi, {d, d}, phi2, {p, p}, {q, q}, op::multiply_add_to(phi3, -5.0 / 7.0));
    }
    for (H2 h : {H2{0, 1}, H2{0, 2}, H2{1, 2}}) Conv.convolve_Hessians(phi, {h.a, h.b}, phi2, {h.a, h.b}, op::multiply_add_to(phi3, 10.0 / 7.0));
    phi3.apply_InverseLaplacian();
    // transverse part
    for (int d = 0; d < 3; ++d) {
        const int p = (d + 1) % 3, q = (d + 2) % 3;
        Conv.convolve_Hessians(phi2, {d, p}, phi, {d, q}, op::assign_to(*A3[d]));
        Conv.convolve_Hessians(phi2, {d, q}, phi, {d, p}, op::subtract_from(*A3[d]));
        Conv.convolve_DifferenceOfHessians(phi, {p, q}, phi2, {p, p}, {q, q}, op::add_to(*A3[d]));
        Conv.convolve_DifferenceOfHessians(phi2, {p, q}, phi, {p, p}, {q, q}, op::subtract_from(*A3[d]));
        A3[d]->apply_InverseLaplacian();
    }
    // ---- NaiveConvolver (include/convolution.hh:296-426) on fields whose Nyquist planes are populated
    if (argc >= 10) {
        grid_t a(N, B), b(N, B);
        if (!load_k(a, argv[7]) || !load_k(b, argv[8])) { std::printf("cannot read the Nyquist-carrying spectra\n"); return 2; }
        NaiveConvolver<real_t> NConv(N, B);
        tmp.FourierTransformForward(false);
        NConv.convolve_Hessians(a, {0, 1}, b, {1, 2}, op::assign_to(tmp));
        dump(tmp, dir, "naive_hess_01_12");
        NConv.convolve_SumOfHessians(a, {0, 0}, b, {1, 1}, {2, 2}, op::assign_to(tmp));
        dump(tmp, dir, "naive_sumhess_00_11_22");
        NConv.convolve_Hessians(a, {0, 0}, b, {0, 2}, a, {2, 2}, op::assign_to(tmp));
        dump(tmp, dir, "naive_hess3_00_02_22");
        NConv.convolve_Gradients(a, {0}, b, {2}, op::assign_to(tmp));
        dump(tmp, dir, "naive_grad_0_2");
        NConv.convolve_Gradient_and_Hessian(a, {1}, b, {0, 2}, op::assign_to(tmp));
        dump(tmp, dir, "naive_gradhess_1_02");
        NConv.multiply_field(a, b, op::assign_to(tmp));
        dump(tmp, dir, "naive_multiply");
        // Grid_FFT<ccomplex_t>: forward c2c with the symmetric norm, and back
        Grid_FFT<ccomplex_t> c(N, B);
        FILE* f = std::fopen(argv[9], "rb");
        if (!f) { std::printf("cannot read %s\n", argv[9]); return 2; }
        c.FourierTransformBackward(false);
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) {
            ccomplex_t v;
            if (std::fread(&v, sizeof(v), 1, f) != 1) return 2;
            c.relem(i, j, k) = v;
        }
        std::fclose(f);
        c.FourierTransformForward();
        FILE* o = std::fopen((dir + "/cplx_k.bin").c_str(), "wb");
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) { ccomplex_t v = kref(c, i, j, k); std::fwrite(&v, sizeof(v), 1, o); }
        std::fclose(o);
        c.FourierTransformBackward();
        o = std::fopen((dir + "/cplx_roundtrip.bin").c_str(), "wb");
        for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) for (size_t k = 0; k < n; ++k) { ccomplex_t v = c.relem(i, j, k); std::fwrite(&v, sizeof(v), 1, o); }
        std::fclose(o);
    }
    phi *= -D;
    phi2 *= -3.0 / 7.0 * D * D;
    phi3 *= 1.0 / 3.0 * D * D * D;
    for (auto* a : A3) (*a) *= 1.0 / 7.0 * D * D * D;
    dump(phi, dir, "phi_scaled");
    dump(phi2, dir, "phi2");
    dump(phi3, dir, "phi3");
    dump(A3x, dir, "A3x");
    dump(A3y, dir, "A3y");
    dump(A3z, dir, "A3z");
    std::printf("ref_driver ok\n");
    return 0;
}
