// include/mfb/convolution.hh - OrszagConvolver<real_t> and NaiveConvolver<real_t> with the named products of
// BaseConvolver (include/convolution.hh:28-293, 296-426, 429-669) on top of mfb_conv_*.  Same method names and
// argument order; the output operator is one of op::assign_to/add_to/subtract_from/multiply_add_to
// (include/mfb/operators.hh).  The reference picks the convolver at compile time (CONVOLVER_TYPE); so does a host
// built against these headers.
#pragma once
#include <array>
#include <type_traits>

#include "operators.hh"

template <typename data_t, typename derived_t>
class BaseConvolver {
    mfb_conv* c_ = nullptr;
    using grid_t = Grid_FFT<data_t>;
    template <class O> static mfb_grid* out(O& o) {
        o.g->FourierTransformForward(false);
        return o.g->device_handle();
    }
    // operands first (transformed to k-space if they are not: include/convolution.hh:142-143), the output grid - which may BE one
    // of the operands, src/ic_generator.cc:437 - only after that: declaring it k-space must not hide a pending operand transform
    static mfb_grid* in(grid_t& g) { g.FourierTransformForward(); return g.device_handle(); }

protected:
    BaseConvolver(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L, bool naive) {
        int n[3] = {int(N[0]), int(N[1]), int(N[2])};
        double l[3] = {L[0], L[1], L[2]};
        mfb_detail::check(naive ? mfb_conv_create_naive(&c_, n, l) : mfb_conv_create(&c_, n, l));
        N_ = N; L_ = L;
    }

public:
    BaseConvolver(const BaseConvolver&) = delete;
    ~BaseConvolver() { if (c_) mfb_conv_destroy(c_); }

    template <typename opp>
    void convolve_Gradients(grid_t& inl, const std::array<int, 1>& d1l, grid_t& inr, const std::array<int, 1>& d1r, opp o) {
        mfb_grid *a = in(inl), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_gradients(c_, a, d1l[0], b, d1r[0], r, o.code, o.alpha));
    }
    template <typename opp>
    void convolve_Gradient_and_Hessian(grid_t& inl, const std::array<int, 1>& d1l, grid_t& inr, const std::array<int, 2>& d2r, opp o) {
        mfb_grid *a = in(inl), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_gradient_hessian(c_, a, d1l[0], b, d2r.data(), r, o.code, o.alpha));
    }
    template <typename opp>
    void convolve_Hessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inr, const std::array<int, 2>& d2r, opp o) {
        mfb_grid *a = in(inl), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_hessians(c_, a, d2l.data(), b, d2r.data(), r, o.code, o.alpha));
    }
    template <typename opp>
    void convolve_Hessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inm, const std::array<int, 2>& d2m, grid_t& inr,
                           const std::array<int, 2>& d2r, opp o) {
        mfb_grid *a = in(inl), *m = in(inm), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_hessians3(c_, a, d2l.data(), m, d2m.data(), b, d2r.data(), r, o.code, o.alpha));
    }
    template <typename opp>
    void convolve_SumOfHessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inr, const std::array<int, 2>& d2r1,
                                const std::array<int, 2>& d2r2, opp o) {
        mfb_grid *a = in(inl), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_sum_hessians(c_, a, d2l.data(), b, d2r1.data(), d2r2.data(), r, o.code, o.alpha));
    }
    template <typename opp>
    void convolve_DifferenceOfHessians(grid_t& inl, const std::array<int, 2>& d2l, grid_t& inr, const std::array<int, 2>& d2r1,
                                       const std::array<int, 2>& d2r2, opp o) {
        mfb_grid *a = in(inl), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_diff_hessians(c_, a, d2l.data(), b, d2r1.data(), d2r2.data(), r, o.code, o.alpha));
    }
    template <typename opp> void multiply_field(grid_t& inl, grid_t& inr, opp o) {
        mfb_grid *a = in(inl), *b = in(inr), *r = out(o);
        mfb_detail::check(mfb_conv_multiply(c_, a, b, r, o.code, o.alpha));
    }
    // convolve2 / convolve3 with the reference's host lambdas kfunc(i,j,k) -> ccomplex_t (include/convolution.hh:512-549): the
    // lambdas are evaluated on the host over the local k-space index range - exactly the loop of the reference's MPI-build
    // pad_insert (:593-603) - into temporaries, which the device then multiplies (mfb_conv_generic / mfb_conv_generic3).
    // output_op is one of the tagged operators of op:: (accumulated on the device) or, like in the reference, any callable
    // (size_t idx, data_t value) over the ntot_ elements of the result grid, run on the host over the downloaded result.
    template <typename kfunc1, typename kfunc2, typename opp> void convolve2(kfunc1 kf1, kfunc2 kf2, opp o) {
        grid_t a(N_, L_, true, kspace_id), b(N_, L_, true, kspace_id);
        fill(a, kf1); fill(b, kf2);
        finish(o, [&](mfb_grid* og, int code, double alpha) {
            return mfb_conv_generic(c_, a.device_handle(), b.device_handle(), og, code, alpha); });
    }
    template <typename kfunc1, typename kfunc2, typename kfunc3, typename opp> void convolve3(kfunc1 kf1, kfunc2 kf2, kfunc3 kf3, opp o) {
        grid_t a(N_, L_, true, kspace_id), b(N_, L_, true, kspace_id), c3(N_, L_, true, kspace_id);
        fill(a, kf1); fill(b, kf2); fill(c3, kf3);
        finish(o, [&](mfb_grid* og, int code, double alpha) {
            return mfb_conv_generic3(c_, a.device_handle(), b.device_handle(), c3.device_handle(), og, code, alpha); });
    }
    mfb_conv* handle() { return c_; }

private:
    std::array<size_t, 3> N_;
    std::array<real_t, 3> L_;
    template <typename kfunc> static void fill(grid_t& g, kfunc kf) {
        g.FourierTransformForward(false);
        for (size_t i = 0; i < g.size(0); ++i)
            for (size_t j = 0; j < g.size(1); ++j)
                for (size_t k = 0; k < g.size(2); ++k) g.kelem(i, j, k) = kf(i, j, k);
    }
    template <typename T> struct is_tagged : std::false_type {};
    template <typename F> struct is_tagged<op::out_op<F>> : std::true_type {};
    template <typename opp, typename call> void finish(opp o, call run) {
        if constexpr (is_tagged<opp>::value) {
            mfb_detail::check(run(out(o), o.code, o.alpha));
        } else {
            grid_t res(N_, L_, true, kspace_id);
            res.FourierTransformForward(false);
            mfb_detail::check(run(res.device_handle(), MFB_OP_ASSIGN, 0.0));
            for (size_t i = 0; i < res.ntot_; ++i) o(i, res[i]);   // element type data_t, like unpad's loop (:659-663)
        }
    }
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
