// include/mfb/particle_plt.hh - particle::lattice_gradient as the device path sees it (include/particle_plt.hh:50-587,
// build option ENABLE_PLT).  The reference's class computes, on the host and once per run, the eigen-decomposition of the
// lattice's dynamical matrix on an ng^3 helper mesh (Ewald sums, GSL eigen-solver) and keeps the results in the k-space
// grids D_xx_, D_yy_, D_zz_ (eigenvalue in the radial / theta / phi direction) and D_xy_ (growth-rate correction); its
// gradient() / vfac_corr() read them with Grid_FFT::get_cic_kspace.  That set-up stays host code.  This mirror takes the
// real parts of the four finished tables ([ng][ng][ng/2+1]) and hands them to the device, where the assemble kernel
// evaluates gradient() / vfac_corr() per mode (mfb_lpt_assemble_lg) in place of op::fourier_gradient
// (src/ic_generator.cc:377-381, 855-858, 911-929).
#pragma once
#include <vector>

#include "grid_fft.hh"

namespace particle {
class lattice_gradient {
    mfb_gradient* h_ = nullptr;

public:
    lattice_gradient(int ngridself, int GridRes, double BoxLength, const std::vector<double>& D_r, const std::vector<double>& D_theta,
                     const std::vector<double>& D_phi, const std::vector<double>& mu) {
        const size_t want = size_t(ngridself) * ngridself * (ngridself / 2 + 1);
        if (D_r.size() != want || D_theta.size() != want || D_phi.size() != want || mu.size() != want)
            throw std::runtime_error("lattice_gradient: tables must be [ng][ng][ng/2+1]");
        const double* tabs[4] = {D_r.data(), D_theta.data(), D_phi.data(), mu.data()};
        mfb_detail::check(mfb_gradient_create_plt(&h_, ngridself, GridRes, BoxLength, tabs));
    }
    lattice_gradient(const lattice_gradient&) = delete;
    ~lattice_gradient() { if (h_) mfb_gradient_destroy(h_); }
    const mfb_gradient* device_handle() const { return h_; }
};
}  // namespace particle
