// C++ host driver in the shape of ic_generator::run (src/ic_generator.cc:90-982), built ONLY from the drop-in headers of
// include/mfb/ and the C ABI: an RNG "plugin" fills the white-noise grid through relem, a transfer-function "plugin" is
// tabulated once per distinct |k|, the LPT source stage and the species output loop run on the device, and an output
// "plugin" reads the particle container's std::vectors (like output_gadget_hdf5.cc:277-300) and dumps them.
// tests/test_cpp_driver.py compares the dump with the CPU oracle run on the same noise.
//
//   ic_driver <GridRes> <LPTorder> <lattice 0..3> <noise.bin> <out.bin>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/mfb/convolution.hh"
#include "../../include/mfb/particle_generator.hh"

struct mock_config {};

// --- "RNG plugin": RNG_plugin::Fill_Grid(Grid_FFT<real_t>&) (include/random_plugin.hh:26-38), here from a file
static bool Fill_Grid(Grid_FFT<real_t>& g, const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    g.FourierTransformBackward(false);
    std::vector<double> row(g.size(2));
    for (size_t i = 0; i < g.size(0); ++i)
        for (size_t j = 0; j < g.size(1); ++j) {
            if (std::fread(row.data(), sizeof(double), row.size(), f) != row.size()) { std::fclose(f); return false; }
            for (size_t k = 0; k < g.size(2); ++k) g.relem(i, j, k) = row[k];
        }
    std::fclose(f);
    return true;
}
// --- "transfer-function plugin": TransferFunction_plugin::compute(k, type) (include/transfer_function_plugin.hh:44-100),
//     folded with k^{ns/2} like cosmology::calculator::get_amplitude (include/cosmology_calculator.hh:405-408)
static double get_amplitude(double k) {
    if (k <= 0.0) return 0.0;
    const double ns = 0.9649, q = k / 0.5;
    const double T = std::log1p(2.34 * q) / (2.34 * q) * std::pow(1 + 3.89 * q + std::pow(16.1 * q, 2) + std::pow(5.46 * q, 3) + std::pow(6.71 * q, 4), -0.25);
    return std::pow(k, 0.5 * ns) * T;
}
// --- "output plugin": write_particle_data(const particle::container&, ...) (include/output_plugin.hh:33-87)
static bool write_particle_data(const particle::container& pc, const char* path) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    const uint64_t np = pc.get_local_num_particles();
    std::fwrite(&np, sizeof(np), 1, f);
    std::fwrite(pc.get_pos64_ptr(), sizeof(double), 3 * np, f);
    std::fwrite(pc.get_vel64_ptr(), sizeof(double), 3 * np, f);
    std::fwrite(pc.get_ids64_ptr(), sizeof(uint64_t), np, f);
    std::fclose(f);
    return true;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        if ((call) != 0) { std::printf("%s: %s\n", #call, mfb_last_error()); return 3; } \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 6) { std::printf("usage: ic_driver GridRes LPTorder lattice noise.bin out.bin\n"); return 1; }
    const size_t ngrid = std::strtoul(argv[1], nullptr, 10);
    const int LPTorder = std::atoi(argv[2]);
    const auto lattice_type = particle::lattice(std::atoi(argv[3]));
    if (mfb_init(0) != 0) { std::printf("mfb_init: %s\n", mfb_last_error()); return 2; }
    try {
        const real_t boxlen = 100.0, Dplus0 = 0.0257, vfac = 396.69, lunit = boxlen, vunit = boxlen / std::sqrt(1.0 / 51.0);
        std::array<size_t, 3> N{ngrid, ngrid, ngrid};
        std::array<real_t, 3> L{boxlen, boxlen, boxlen};
        mock_config cf;
        Grid_FFT<real_t> wnoise(N, L), phi(N, L), phi2(N, L), phi3(N, L), A3x(N, L), A3y(N, L), A3z(N, L), tmp(N, L);
        Grid_FFT<real_t>* A3[3] = {&A3x, &A3y, &A3z};

        // :313-358  white noise -> k-space, volume factor
        if (!Fill_Grid(wnoise, argv[4])) { std::printf("cannot read %s\n", argv[4]); return 1; }
        wnoise.FourierTransformForward();
        const real_t volfac = std::pow(boxlen / ngrid / 2.0 / M_PI, 1.5);
        CK(mfb_noise_condition(wnoise.device_handle(), volfac, 0, 0, 0));

        // :461-468  phi = - wnoise * A(k) / k^2, the plugin evaluated once per distinct |k|
        const size_t h = ngrid / 2, ntab = 3 * h * h + 1;
        std::vector<double> table(ntab);
        for (size_t m = 0; m < ntab; ++m) table[m] = get_amplitude(2.0 * M_PI / boxlen * std::sqrt(double(m)));
        phi.FourierTransformForward(false);
        CK(mfb_apply_radial_table(phi.device_handle(), wnoise.device_handle(), table.data(), ntab, 1));

        // :475-574  sources and growth scaling
        phi2.FourierTransformForward(false); phi3.FourierTransformForward(false);
        for (auto* a : A3) a->FourierTransformForward(false);
        if (LPTorder > 1) {
            OrszagConvolver<real_t> Conv(N, L);
            mfb_grid* a3h[3] = {A3x.device_handle(), A3y.device_handle(), A3z.device_handle()};
            CK(mfb_lpt_sources(Conv.handle(), LPTorder, Dplus0, phi.device_handle(), phi2.device_handle(),
                               LPTorder > 2 ? phi3.device_handle() : nullptr, LPTorder > 2 ? a3h : nullptr, nullptr));
        } else {
            phi *= -Dplus0;
        }

        // :832-955  species output loop (one species, particles)
        particle::lattice_generator<Grid_FFT<real_t>> lg(lattice_type, 0, true, true, false, 0, tmp, cf);
        mfb_grid* a3h[3] = {A3x.device_handle(), A3y.device_handle(), A3z.device_handle()};
        mfb_grid* p2 = LPTorder > 1 ? phi2.device_handle() : nullptr;
        mfb_grid* p3 = LPTorder > 2 ? phi3.device_handle() : nullptr;
        for (int idim = 0; idim < 3; ++idim) {
            tmp.FourierTransformForward(false);
            CK(mfb_lpt_assemble(tmp.device_handle(), idim, phi.device_handle(), p2, p3, LPTorder > 2 ? a3h : nullptr, 1.0, 1.0, 1.0, 1.0,
                                lunit / boxlen, nullptr, nullptr, 0, 0.0));
            tmp.FourierTransformBackward();
            lg.set_positions(lattice_type, 0, idim, lunit, true, tmp, cf);
        }
        for (int idim = 0; idim < 3; ++idim) {
            tmp.FourierTransformForward(false);
            CK(mfb_lpt_assemble(tmp.device_handle(), idim, phi.device_handle(), p2, p3, LPTorder > 2 ? a3h : nullptr, vfac, 2 * vfac, 3 * vfac,
                                3 * vfac, vunit / boxlen, nullptr, nullptr, 0, 0.0));
            tmp.FourierTransformBackward();
            lg.set_velocities(lattice_type, 0, idim, true, tmp, cf);
        }
        if (!write_particle_data(lg.get_particles(), argv[5])) { std::printf("cannot write %s\n", argv[5]); return 1; }
        std::printf("ic_driver ok (%s): %zu particles\n", mfb_build_info(), lg.get_particles().get_local_num_particles());
    } catch (const std::exception& e) {
        std::printf("exception: %s\n", e.what());
        return 4;
    }
    mfb_finalize();
    return 0;
}
