// include/mfb/operators.hh - output operators of the reference (include/operators.hh:25-35) as tagged
// descriptors: the accumulate happens inside the last FFT pass on the device instead of a host lambda.
#pragma once
#include "grid_fft.hh"

namespace op {
template <typename field> struct out_op { field* g; int code; double alpha; };
template <typename field> inline out_op<field> assign_to(field& g) { return {&g, MFB_OP_ASSIGN, 0.0}; }
template <typename field, typename val> inline out_op<field> multiply_add_to(field& g, val x) { return {&g, MFB_OP_MULADD, double(x)}; }
template <typename field> inline out_op<field> add_to(field& g) { return {&g, MFB_OP_ADD, 0.0}; }
template <typename field> inline out_op<field> subtract_from(field& g) { return {&g, MFB_OP_SUB, 0.0}; }

// op::fourier_gradient (include/operators.hh:37-61): the plain spectral gradient as the host-side operator object that
// ic_generator::run hands to its assembly loops (src/ic_generator.cc:377-381); Nyquist component = 0, vfac_corr = 1.  On the
// device path the same rule lives in the assemble kernel (mfb_lpt_assemble).
class fourier_gradient {
    real_t boxlen_, k0_;
    size_t n_, nhalf_;

public:
    template <class config_t>
    explicit fourier_gradient(const config_t& the_config)
        : boxlen_(the_config.template get_value<double>("setup", "BoxLength")), k0_(2.0 * M_PI / boxlen_),
          n_(the_config.template get_value<size_t>("setup", "GridRes")), nhalf_(n_ / 2) {}
    fourier_gradient(real_t boxlen, size_t n) : boxlen_(boxlen), k0_(2.0 * M_PI / boxlen), n_(n), nhalf_(n / 2) {}
    ccomplex_t gradient(const int idim, std::array<size_t, 3> ijk) const {
        const real_t m = (ijk[idim] != nhalf_) ? (real_t(ijk[idim]) - real_t(ijk[idim] > nhalf_) * real_t(n_)) : 0.0;
        return ccomplex_t(0.0, m * k0_);
    }
    real_t vfac_corr(std::array<size_t, 3>) const { return 1.0; }
};
}  // namespace op
