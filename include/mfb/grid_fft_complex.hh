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

// ------------------------------------------------------------------------------------------------------------------
// Grid_FFT<ccomplex_t> itself, so that the PPT branch of the reference's ic_generator::run (src/ic_generator.cc:690-801)
// compiles and runs as written: dense [n0][n1][n2] complex elements in both spaces (src/grid_fft.cc:56-70), the c2c transforms
// on the device (mfb_cgrid_fft_*), and a lazily synchronised host mirror that the reference's host lambdas
// (apply_function_k_dep, assign_function_of_grids_r, ... include/grid_fft.hh:698-796) run on - the same coherence rule as the
// real grid.  The named device operations of namespace ppt below are the fast spelling of the same branch.  Whole grid or slab.
template <bool bdistributed>
class Grid_FFT<ccomplex_t, bdistributed> {
public:
    using data_t = ccomplex_t;
    static constexpr bool is_distributed_trait{bdistributed};
    std::array<size_t, 3> n_, nhalf_;
    std::array<size_t, 4> sizes_;
    size_t npr_, npc_, ntot_;
    std::array<real_t, 3> length_, kfac_, kny_, dx_;
    space_t space_;
    ccomplex_t* data_;
    ccomplex_t* cdata_;
    real_t fft_norm_fac_;
    bool ballocated_;
    ptrdiff_t local_0_start_, local_1_start_, local_0_size_, local_1_size_;

    Grid_FFT(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L, bool allocate = true, space_t initialspace = rspace_id)
        : n_(N), length_(L), space_(initialspace), data_(nullptr), cdata_(nullptr), ballocated_(false) {
        if (allocate) this->allocate();
    }
    Grid_FFT(const Grid_FFT&) = delete;
    Grid_FFT& operator=(const Grid_FFT&) = delete;
    ~Grid_FFT() { reset(); }
    void reset() {
        if (h_) { mfb_cgrid_destroy(h_); h_ = nullptr; }
        host_.clear(); host_.shrink_to_fit();
        data_ = cdata_ = nullptr; ballocated_ = false;
    }
    void allocate() {
        if (ballocated_) return;
        int n[3] = {int(n_[0]), int(n_[1]), int(n_[2])};
        double L[3] = {length_[0], length_[1], length_[2]};
        mfb_detail::ensure_init(n);
        mfb_detail::check(mfb_cgrid_create(&h_, n, L));
        int loc[4];
        const bool slab = mfb_cgrid_local(h_, loc) != 0;   // {local_0_start_, local_0_size_, local_1_start_, local_1_size_}
        if (slab != bdistributed) {
            mfb_cgrid_destroy(h_); h_ = nullptr;
            throw std::runtime_error(bdistributed ? "Grid_FFT<ccomplex_t, true>: the library is not slab-decomposed (mfb_dist_init)"
                                                  : "Grid_FFT<ccomplex_t, false>: the library is slab-decomposed - use Grid_FFT<ccomplex_t, true>");
        }
        local_0_start_ = loc[0]; local_0_size_ = loc[1]; local_1_start_ = loc[2]; local_1_size_ = loc[3];
        ntot_ = mfb_cgrid_nelem(h_);   // n0 n1 n2, or the larger of the slab's two layouts ([x_loc][n1][n2] / [ky_loc][n0][n2])
        host_.assign(ntot_, ccomplex_t(0.0, 0.0));
        data_ = cdata_ = host_.data();
        fft_norm_fac_ = 1.0 / std::sqrt(real_t(n_[0]) * real_t(n_[1]) * real_t(n_[2]));
        npr_ = npc_ = n_[2];
        for (int i = 0; i < 3; ++i) {
            nhalf_[i] = n_[i] / 2; kfac_[i] = 2.0 * M_PI / length_[i]; kny_[i] = kfac_[i] * n_[i] / 2; dx_[i] = length_[i] / n_[i];
        }
        set_sizes();
        ballocated_ = true;
        host_dirty_ = true; dev_newer_ = false;
        if (space_ == kspace_id) mfb_detail::check(mfb_cgrid_fft_forward(h_, 0));
    }
    const Grid_FFT* get_grid(size_t) const { return this; }
    bool is_distributed() const noexcept { return bdistributed; }
    bool is_allocated() const noexcept { return ballocated_; }
    size_t memsize() const noexcept { return ntot_; }
    size_t size(size_t i) const noexcept { return sizes_[i]; }
    size_t rsize(size_t i) const noexcept { return n_[i]; }
    size_t local_size() const noexcept { return size_t(local_0_size_) * n_[1] * n_[2]; }
    const std::array<size_t, 3>& get_global_size() const noexcept { return n_; }
    const std::array<real_t, 3>& get_dx() const noexcept { return dx_; }
    size_t get_idx(size_t i, size_t j, size_t k) const noexcept { return (i * sizes_[1] + j) * sizes_[2] + k; }
    template <typename ft> vec3_t<ft> get_k(size_t i, size_t j, size_t k) const noexcept {   // include/grid_fft.hh:281-296
        if (bdistributed) {   // transposed slab: the first index runs over the local ky rows
            const size_t ip = i + size_t(local_1_start_);
            return vec3_t<ft>(ft((real_t(j) - real_t(j > nhalf_[0]) * n_[0]) * kfac_[0]), ft((real_t(ip) - real_t(ip > nhalf_[1]) * n_[1]) * kfac_[1]),
                              ft((real_t(k) - real_t(k > nhalf_[2]) * n_[2]) * kfac_[2]));
        }
        return vec3_t<ft>(ft((real_t(i) - real_t(i > nhalf_[0]) * n_[0]) * kfac_[0]), ft((real_t(j) - real_t(j > nhalf_[1]) * n_[1]) * kfac_[1]),
                          ft((real_t(k) - real_t(k > nhalf_[2]) * n_[2]) * kfac_[2]));
    }
    template <typename ft> vec3_t<ft> get_r(size_t i, size_t j, size_t k) const noexcept {
        return vec3_t<ft>(ft(real_t(i + local_0_start_) * dx_[0]), ft(real_t(j) * dx_[1]), ft(real_t(k) * dx_[2]));
    }
    // ---- host element access
    ccomplex_t& operator[](size_t i) { to_host(true); return data_[i]; }
    ccomplex_t& relem(size_t i, size_t j, size_t k) { to_host(true); return data_[get_idx(i, j, k)]; }
    const ccomplex_t& relem(size_t i, size_t j, size_t k) const { to_host(false); return data_[get_idx(i, j, k)]; }
    ccomplex_t& kelem(size_t i, size_t j, size_t k) { to_host(true); return data_[get_idx(i, j, k)]; }
    const ccomplex_t& kelem(size_t i, size_t j, size_t k) const { to_host(false); return data_[get_idx(i, j, k)]; }
    ccomplex_t& relem(size_t idx) { to_host(true); return data_[idx]; }
    ccomplex_t& kelem(size_t idx) { to_host(true); return data_[idx]; }
    // ---- device operations
    mfb_cgrid* device_handle() { to_device(); dev_newer_ = true; return h_; }
    void FourierTransformForward(bool do_transform = true) {
        if (space_ == kspace_id) return;
        if (do_transform) { to_device(); dev_newer_ = true; }
        mfb_detail::check(mfb_cgrid_fft_forward(h_, do_transform ? 1 : 0));
        space_ = kspace_id; set_sizes();
    }
    void FourierTransformBackward(bool do_transform = true) {
        if (space_ == rspace_id) return;
        if (do_transform) { to_device(); dev_newer_ = true; }
        mfb_detail::check(mfb_cgrid_fft_backward(h_, do_transform ? 1 : 0));
        space_ = rspace_id; set_sizes();
    }
    void zero() noexcept { to_host(true); std::fill(host_.begin(), host_.end(), ccomplex_t(0.0, 0.0)); }
    void copy_from(const Grid_FFT& g) {
        const_cast<Grid_FFT&>(g).to_device();
        mfb_detail::check(mfb_cgrid_copy(h_, g.h_));
        space_ = g.space_; set_sizes(); host_dirty_ = false; dev_newer_ = true;
    }
    // ---- the functor family on the host mirror (include/grid_fft.hh:698-796)
    template <typename F> void apply_function_k(const F& f) { to_host(true); for (size_t q = 0, m = sizes_[0] * sizes_[1] * sizes_[2]; q < m; ++q) data_[q] = f(data_[q]); }
    template <typename F> void apply_function_r(const F& f) { apply_function_k(f); }
    template <typename F> void apply_function_k_dep(const F& f) {
        to_host(true);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            auto& e = data_[get_idx(i, j, k)]; e = f(e, this->template get_k<real_t>(i, j, k));
        }
    }
    template <typename F> void apply_function_r_dep(const F& f) {
        to_host(true);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            auto& e = data_[get_idx(i, j, k)]; e = f(e, this->template get_r<real_t>(i, j, k));
        }
    }
    template <typename F, typename... Grids> void assign_function_of_grids_r(const F& f, Grids&... grids) {
        to_host(true); (grids.to_host(false), ...);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k)
            data_[get_idx(i, j, k)] = f((grids.data_[grids.get_idx(i, j, k)])...);
    }
    template <typename F, typename... Grids> void assign_function_of_grids_k(const F& f, Grids&... grids) { assign_function_of_grids_r(f, grids...); }

private:
    mfb_cgrid* h_ = nullptr;
    std::vector<ccomplex_t> host_;
    void set_sizes() {   // dense elements: [x_loc][n1][n2] in real space, [ky_loc][n0][n2] (transposed) in k-space on slabs
        if (space_ == rspace_id) sizes_ = {size_t(local_0_size_), n_[1], n_[2], n_[2]};
        else sizes_ = {size_t(local_1_size_), n_[0], n_[2], n_[2]};
    }
    mutable std::atomic<bool> host_dirty_{false}, dev_newer_{false};   // thread-safe like the real grid's (host lambdas inside OpenMP teams)
    mutable std::mutex sync_mu_;
    void to_host(bool will_write) const {
        if (dev_newer_.load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> lk(sync_mu_);
            if (dev_newer_.load(std::memory_order_relaxed)) {
                mfb_detail::check(mfb_cgrid_download(h_, reinterpret_cast<double*>(const_cast<ccomplex_t*>(host_.data()))));
                dev_newer_.store(false, std::memory_order_release);
            }
        }
        if (will_write && !host_dirty_.load(std::memory_order_relaxed)) host_dirty_.store(true, std::memory_order_relaxed);
    }
    void to_device() const {
        if (host_dirty_) {
            mfb_detail::check(mfb_cgrid_upload(h_, reinterpret_cast<const double*>(host_.data()), space_ == kspace_id ? MFB_KSPACE : MFB_RSPACE));
            host_dirty_ = false;
        }
    }
    template <typename, bool> friend class Grid_FFT;
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
