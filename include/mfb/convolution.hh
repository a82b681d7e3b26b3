// include/mfb/convolution.hh - OrszagConvolver<real_t> and NaiveConvolver<real_t> with the named products of
// BaseConvolver (include/convolution.hh:28-293, 296-426, 429-669) on top of mfb_conv_*.  Same method names and
// argument order; the output operator is one of op::assign_to/add_to/subtract_from/multiply_add_to
// (include/mfb/operators.hh).  The reference picks the convolver at compile time (CONVOLVER_TYPE); so does a host
// built against these headers.
#pragma once
#include <array>

#include "operators.hh"

template <typename data_t, typename derived_t>
class BaseConvolver {
    mfb_conv* c_ = nullptr;
    using grid_t = Grid_FFT<data_t>;
    template <class O> static mfb_grid* out(O& o) {
        o.g->FourierTransformForward(false);
        return o.g->device_handle();
    }
    static mfb_grid* in(grid_t& g) { g.FourierTransformForward(); return g.device_handle(); }

protected:
    BaseConvolver(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L, bool naive) {
        int n[3] = {int(N[0]), int(N[1]), int(N[2])};
        double l[3] = {L[0], L[1], L[2]};
        mfb_detail::check(naive ? mfb_conv_create_naive(&c_, n, l) : mfb_conv_create(&c_, n, l));
    }

public:
    BaseConvolver(const BaseConvolver&) = delete;
    ~BaseConvolver() { if (c_) mfb_conv_destroy(c_); }

    template <typename opp>
    void convolve_Gradients(grid_t& inl, const std::array<int, 1>& d1l, grid_t& inr, const std::array<int, 1>& d1r, opp o) {
        mfb_detail::check(mfb_conv_gradients(c_, in(inl), d1l[0], in(inr), d1r[0], out(o), o.code, o.alpha));
    }
    template <typename opp>
    void convolve_Gradient_and_Hessian(grid_t& inl, const std::array<int, 1>& d1l, grid_t& inr, const std::array<int, 2>& d2r, opp o) {
        mfb_detail::check(mfb_conv_gradient_hessian(c_, in(inl), d1l[0], in(inr), d2r.data(), out(o), o.code, o.alpha));
    }
    template <typename opp>
    void convolve_Hessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inr, const std::array<int, 2>& d2r, opp o) {
        mfb_detail::check(mfb_conv_hessians(c_, in(inl), d2l.data(), in(inr), d2r.data(), out(o), o.code, o.alpha));
    }
    template <typename opp>
    void convolve_Hessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inm, const std::array<int, 2>& d2m, grid_t& inr,
                           const std::array<int, 2>& d2r, opp o) {
        mfb_detail::check(mfb_conv_hessians3(c_, in(inl), d2l.data(), in(inm), d2m.data(), in(inr), d2r.data(), out(o), o.code, o.alpha));
    }
    template <typename opp>
    void convolve_SumOfHessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inr, const std::array<int, 2>& d2r1,
                                const std::array<int, 2>& d2r2, opp o) {
        mfb_detail::check(mfb_conv_sum_hessians(c_, in(inl), d2l.data(), in(inr), d2r1.data(), d2r2.data(), out(o), o.code, o.alpha));
    }
    template <typename opp>
    void convolve_DifferenceOfHessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inr, const std::array<int, 2>& d2r1,
                                       const std::array<int, 2>& d2r2, opp o) {
        mfb_detail::check(mfb_conv_diff_hessians(c_, in(inl), d2l.data(), in(inr), d2r1.data(), d2r2.data(), out(o), o.code, o.alpha));
    }
    template <typename opp> void multiply_field(grid_t& inl, grid_t& inr, opp o) {
        mfb_detail::check(mfb_conv_multiply(c_, in(inl), in(inr), out(o), o.code, o.alpha));
    }
    mfb_conv* handle() { return c_; }
};

template <typename data_t>
class OrszagConvolver : public BaseConvolver<data_t, OrszagConvolver<data_t>> {   // include/convolution.hh:429-669
public:
    OrszagConvolver(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L) : BaseConvolver<data_t, OrszagConvolver<data_t>>(N, L, false) {}
};
template <typename data_t>
class NaiveConvolver : public BaseConvolver<data_t, NaiveConvolver<data_t>> {    // include/convolution.hh:296-426
public:
    NaiveConvolver(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L) : BaseConvolver<data_t, NaiveConvolver<data_t>>(N, L, true) {}
};
