// include/mfb/grid_fft_complex.hh - the Grid_FFT<ccomplex_t> of the semiclassical (PPT / Eulerian) branch of
// ic_generator::run (src/ic_generator.cc:685-803) on top of mfb_cgrid_* / mfb_ppt_*.
//
// Only the host driver ever sees a complex grid (no plugin does), and every use the reference makes of it is a host
// lambda handed to apply_function_* / assign_function_of_grids_r.  A lambda cannot run on the device, so each of those
// call sites becomes one named operation below; the transforms keep their names.  `ppt::run` is the whole branch.
#pragma once
#include <cmath>
#include <functional>
#include <vector>

#include "grid_fft.hh"

class Grid_FFT_complex {
    mfb_cgrid* h_ = nullptr;

public:
    std::array<size_t, 3> n_;
    std::array<real_t, 3> length_;

    Grid_FFT_complex(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L) : n_(N), length_(L) {
        int n[3] = {int(N[0]), int(N[1]), int(N[2])};
        double l[3] = {L[0], L[1], L[2]};
        mfb_detail::check(mfb_cgrid_create(&h_, n, l));
    }
    Grid_FFT_complex(const Grid_FFT_complex&) = delete;
    Grid_FFT_complex& operator=(const Grid_FFT_complex&) = delete;
    ~Grid_FFT_complex() { if (h_) mfb_cgrid_destroy(h_); }

    space_t space() const { return mfb_cgrid_space(h_) == MFB_KSPACE ? kspace_id : rspace_id; }
    size_t size(size_t i) const { return n_[i]; }
    void FourierTransformForward(bool do_transform = true) { mfb_detail::check(mfb_cgrid_fft_forward(h_, do_transform ? 1 : 0)); }
    void FourierTransformBackward(bool do_transform = true) { mfb_detail::check(mfb_cgrid_fft_backward(h_, do_transform ? 1 : 0)); }
    void copy_from(const Grid_FFT_complex& g) { mfb_detail::check(mfb_cgrid_copy(h_, g.h_)); }
    // grad_psi.apply_function_k_dep([&](auto x, auto k){ return x * ccomplex_t(0.0, k[idim]); })   (:790-792)
    void apply_gradient(int idim) { mfb_detail::check(mfb_cgrid_gradient(h_, idim)); }
    // host copies [n0][n1][n2], for plugins / diagnostics
    std::vector<ccomplex_t> download() const {
        std::vector<ccomplex_t> v(mfb_cgrid_nelem(h_));
        mfb_detail::check(mfb_cgrid_download(h_, reinterpret_cast<double*>(v.data())));
        return v;
    }
    void upload(const std::vector<ccomplex_t>& v, space_t sp) {
        if (v.size() != mfb_cgrid_nelem(h_)) throw std::runtime_error("Grid_FFT_complex::upload: size mismatch");
        mfb_detail::check(mfb_cgrid_upload(h_, reinterpret_cast<const double*>(v.data()), sp == kspace_id ? MFB_KSPACE : MFB_RSPACE));
    }
    mfb_cgrid* device_handle() { return h_; }
    const mfb_cgrid* device_handle() const { return h_; }
};

namespace ppt {
using rgrid = Grid_FFT<real_t>;
// rho.apply_function_r(sqrt(1 + C rho))                                                        (:703-705)
inline void sqrt_density(rgrid& rho, real_t C_species) {
    rho.FourierTransformBackward();
    mfb_detail::check(mfb_ppt_sqrt_density(rho.device_handle(), C_species));
}
// hbar = max|delta phi| / pi / D+ * 1.01 over nearest neighbours (:713-734; the x loop stops at size(0)-1 there too)
inline real_t hbar_from_potential(rgrid& phi, real_t Dplus0) {
    phi.FourierTransformBackward();
    double m = 0.0;
    mfb_detail::check(mfb_grid_max_neighbour_diff(phi.device_handle(), &m));
    return m / M_PI / Dplus0 * 1.01;
}
// psi = rho exp(i (phi [+ phi2]) / (hbar D+))                                                   (:738-751)
inline void init_psi(Grid_FFT_complex& psi, rgrid& rho, rgrid& phi, rgrid* phi2, real_t hbar, real_t Dplus0) {
    phi.FourierTransformBackward();
    if (phi2) phi2->FourierTransformBackward();
    mfb_detail::check(mfb_ppt_init_psi(psi.device_handle(), rho.device_handle(), phi.device_handle(),
                                       phi2 ? phi2->device_handle() : nullptr, hbar, Dplus0));
}
// psi.apply_function_k_dep(psi * exp(-i/2 hbar k^2 D+))                                         (:757-760)
inline void drift(Grid_FFT_complex& psi, real_t hbar, real_t Dplus0) {
    psi.FourierTransformForward();
    mfb_detail::check(mfb_ppt_drift(psi.device_handle(), hbar, Dplus0));
}
// psi *= exp(i phi2 / (hbar D+))                                                                (:763-767)
inline void kick(Grid_FFT_complex& psi, rgrid& phi2, real_t hbar, real_t Dplus0) {
    psi.FourierTransformBackward();
    phi2.FourierTransformBackward();
    mfb_detail::check(mfb_ppt_kick(psi.device_handle(), phi2.device_handle(), hbar, Dplus0));
}
// rho = |psi|^2 - 1                                                                             (:772-775)
inline void density(rgrid& rho, Grid_FFT_complex& psi) {
    psi.FourierTransformBackward();
    rho.FourierTransformBackward(false);
    mfb_detail::check(mfb_ppt_density(rho.device_handle(), psi.device_handle()));
}
// tmp = vunit * hbar * Im(conj(psi) grad_psi) / (1 + rho)                                       (:796-799)
inline void velocity(rgrid& tmp, Grid_FFT_complex& psi, Grid_FFT_complex& grad_psi, rgrid& rho, real_t hbar, real_t vunit) {
    tmp.FourierTransformBackward(false);
    mfb_detail::check(mfb_ppt_velocity(tmp.device_handle(), psi.device_handle(), grad_psi.device_handle(), rho.device_handle(), hbar, vunit));
}

// The whole branch with the reference's control flow.  `rho` enters holding delta_bc in k-space (wnoise * A_delta_bc(k),
// :697-700); `write(grid, component)` plays the_output_plugin->write_grid_data (component 0 = density, 1..3 = vx..vz).
inline real_t run(rgrid& rho, rgrid& phi, rgrid* phi2, rgrid& tmp, int LPTorder, real_t Dplus0, real_t vfac, real_t boxlen,
                  real_t C_species, real_t velocity_unit, const std::function<void(rgrid&, int)>& write) {
    rho.zero_DC_mode();
    sqrt_density(rho, C_species);
    const real_t hbar = hbar_from_potential(phi, Dplus0);
    Grid_FFT_complex psi(rho.n_, rho.length_);
    rgrid* p2 = (LPTorder >= 2) ? phi2 : nullptr;
    init_psi(psi, rho, phi, p2, hbar, Dplus0);
    drift(psi, hbar, Dplus0);
    psi.FourierTransformBackward();
    if (p2) kick(psi, *p2, hbar, Dplus0);
    density(rho, psi);
    write(rho, 0);
    Grid_FFT_complex grad_psi(rho.n_, rho.length_);
    const real_t vunit = Dplus0 * vfac / boxlen * velocity_unit;   // :779
    for (int idim = 0; idim < 3; ++idim) {
        grad_psi.copy_from(psi);
        grad_psi.FourierTransformForward();
        grad_psi.apply_gradient(idim);
        grad_psi.FourierTransformBackward();
        velocity(tmp, psi, grad_psi, rho, hbar, vunit);
        write(tmp, 1 + idim);
    }
    return hbar;
}
}  // namespace ppt
