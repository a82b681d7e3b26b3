// include/mfb/grid_fft.hh - GPU-backed drop-in for the reference's Grid_FFT<real_t>
// (include/grid_fft.hh:40-914, src/grid_fft.cc:40-492) on top of the C ABI (include/mfb200.h).
//
// Same public data members and member functions as the reference class, so that the RNG, transfer and
// output plugins - which touch a grid only through n_, ntot_, sizes_, local_*_, data_, relem/kelem,
// size/rsize/global_size, zero/reset/allocate, FourierTransform*, FourierInterpolateCopyTo, mean/std -
// compile against it untouched.  The numerics run on the device; `data_` is a host mirror kept coherent
// lazily: host accessors pull the device copy when it is newer and mark the host copy dirty (non-const
// access), device operations push a dirty host copy first.
//
// Differences to the reference, all outside the plugin contract:
//  * the arbitrary-lambda functor family (apply_function_k/_r/_k_dep/_r_dep, assign_function_of_grids_*)
//    runs on the HOST mirror (it takes a host lambda); the pipeline itself uses the named device operators
//    of mfb200.h instead (mfb_noise_condition, mfb_apply_radial_table, mfb_lpt_assemble, ...).
//  * Grid_FFT<real_t, true> is a SLAB of the reference's FFTW-MPI decomposition (src/grid_fft.cc:115-192): one host process
//    per GPU, local_0_* / local_1_* taken from the library (mfb_grid_local) after mfb_dist_init + mfb_dist_connect (the
//    host carries the 64-byte IPC handles between its ranks, e.g. with MPI_Allgather - INTEGRATION.md), k-space stored
//    transposed [ky_loc][kx][kz] exactly like FFTW_MPI_TRANSPOSED_OUT.  Grid_FFT<real_t, false> is the whole grid.
#pragma once
#include <array>
#include <atomic>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../mfb200.h"
#ifdef USE_MPI
#include <mpi.h>
#endif

#ifndef MFB_HAVE_REFERENCE_GENERAL_HH
using real_t = double;                   // CODE_PRECISION=DOUBLE (include/general.hh:50-53)
using ccomplex_t = std::complex<double>;
// stand-in for the reference's vec3_t (include/math/vec3.hh) where its headers are not on the include path
template <typename T> struct vec3_t {
    T x, y, z;
    vec3_t() : x(0), y(0), z(0) {}
    vec3_t(T a, T b, T c) : x(a), y(b), z(c) {}
    T& operator[](size_t i) noexcept { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](size_t i) const noexcept { return i == 0 ? x : (i == 1 ? y : z); }
    T norm_squared() const noexcept { return x * x + y * y + z * z; }
    T norm() const noexcept { return std::sqrt(norm_squared()); }
};
#else
#include <math/vec3.hh>
#endif
#if defined(__has_include)
#if __has_include(<hdf5.h>)
#define MFB_MIRROR_HAVE_HDF5 1
#include <hdf5.h>
#endif
#endif
#include <fstream>
#include <iomanip>
#include <mutex>
#include <vector>

enum space_t { kspace_id, rspace_id };  // include/grid_fft.hh:29  (kspace_id == MFB_KSPACE == 0)

namespace mfb_detail {
inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(mfb_last_error());
}
// a host whose main() knows nothing of the library (the reference's src/main.cc as it is) gets the device selected at the first
// grid allocation: $MFB_DEVICE, else (MPI rank) % (visible devices) in a USE_MPI build, else device 0.  mfb_init is idempotent, so a host that
// called it (or mfb_dist_init) itself keeps its choice.
// USE_MPI build with more than one rank: one rank per GPU, every grid a slab (src/grid_fft.cc:115-192).  The ranks exchange the
// 64-byte handles of their peer-mapped exchange blocks once (MPI_Allgather) and the slab transposes then run inside the FFT
// passes over NVLink; with the thread-emulation build of the kernels (CPU tests) or $MFB_STAGED_EXCHANGE the exchange is handed
// to MPI_Alltoall instead (mfb_dist_init_staged; on GPUs the callback sees device pointers, i.e. that mode wants a CUDA-aware
// MPI).  `n` is the first grid of the run (GridRes^3 in ic_generator::run): the
// exchange buffers are sized for it and for its 3/2-padded products; their number follows the free device memory
// ($MFB_EXCHANGE_SLOTS overrides).
#ifdef USE_MPI
inline void mpi_exchange(const void* send, void* recv, size_t chunk_bytes, void*) {
    MPI_Alltoall(send, int(chunk_bytes / sizeof(double)), MPI_DOUBLE, recv, int(chunk_bytes / sizeof(double)), MPI_DOUBLE, MPI_COMM_WORLD);
}
#endif
inline void ensure_init(const int n[3]) {
    static bool done = false;
    if (done) return;
    const char* dev = std::getenv("MFB_DEVICE");
    int rank = 0, size = 1;
#ifdef USE_MPI
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    MPI_Comm_size(MPI_COMM_WORLD, &size);
#endif
    const int ndev = mfb_device_count();
    check(mfb_init(dev ? std::atoi(dev) : (ndev > 0 ? rank % ndev : 0)));
#ifdef USE_MPI
    if (size > 1 && mfb_dist_nranks() <= 1) {
        const size_t xbuf = mfb_dist_xbuf_bytes(n, 1, size);
        const bool staged = std::strstr(mfb_build_info(), "emulation") || std::getenv("MFB_STAGED_EXCHANGE");
        // exchange buffers in the peer-mapped block: the transforms need 2; the shared-transform source stage uses all of them as
        // slots and wants 10 at 3LPT (8, 7, 6, 5: narrower loops).  The largest count that leaves room for the 8 grids and the
        // particle arrays of an LPT run and a margin is taken - the rule of monofonic_b200/dist.py:pick_nslots - the same on every rank.
        size_t freeb = 0;
        check(mfb_device_mem_info(&freeb, nullptr));
        if (staged) freeb /= 2;   // staged mode keeps a staging copy of every buffer
        const size_t cells = size_t(n[0]) * n[1] * n[2] / size_t(size);
        const size_t fixed = 8 * 8 * cells * size_t(n[2] + 2) / size_t(n[2]) + 32 * cells + (size_t(3) << 30);
        int nslots = 2;
        const char* want = std::getenv("MFB_EXCHANGE_SLOTS");
        if (want) nslots = std::atoi(want);
        else
            for (int k : {10, 8, 7, 6, 5})
                if (size_t(k) * xbuf + fixed <= freeb) { nslots = k; break; }
        int agreed = nslots;
        MPI_Allreduce(&nslots, &agreed, 1, MPI_INT, MPI_MIN, MPI_COMM_WORLD);
        if (staged) {
            check(mfb_dist_init_staged_slots(rank, size, xbuf, agreed, mpi_exchange, nullptr));
        } else {
            char mine[64];
            std::vector<char> all(size_t(64) * size);
            check(mfb_dist_init_slots(rank, size, xbuf, agreed, mine));
            MPI_Allgather(mine, 64, MPI_CHAR, all.data(), 64, MPI_CHAR, MPI_COMM_WORLD);
            check(mfb_dist_connect(all.data()));
        }
    }
#endif
    (void)n; (void)size;
    done = true;
}
}  // namespace mfb_detail

// like the reference (include/grid_fft.hh:31-40): a host built with USE_MPI (one rank per GPU) gets slab grids by default
#ifndef GRID_FFT_DISTRIBUTED
#ifdef USE_MPI
#define GRID_FFT_DISTRIBUTED true
#else
#define GRID_FFT_DISTRIBUTED false
#endif
#endif
template <typename data_t_, bool bdistributed = GRID_FFT_DISTRIBUTED>
class Grid_FFT {
    static_assert(std::is_same<data_t_, double>::value, "FP64 real grids here; Grid_FFT<ccomplex_t> is the specialisation of grid_fft_complex.hh");

public:
    using data_t = data_t_;
    static constexpr bool is_distributed_trait{bdistributed};

    std::array<size_t, 3> n_, nhalf_;
    std::array<size_t, 4> sizes_;
    size_t npr_, npc_;
    size_t ntot_;
    std::array<real_t, 3> length_, kfac_, kny_, dx_;
    space_t space_;
    data_t* data_;        // host mirror
    ccomplex_t* cdata_;
    real_t fft_norm_fac_;
    bool ballocated_;
    ptrdiff_t local_0_start_, local_1_start_;
    ptrdiff_t local_0_size_, local_1_size_;

    Grid_FFT(const std::array<size_t, 3>& N, const std::array<real_t, 3>& L, bool allocate = true,
             space_t initialspace = rspace_id)
        : n_(N), length_(L), space_(initialspace), data_(nullptr), cdata_(nullptr), ballocated_(false), h_(nullptr),
          host_dirty_(false), dev_newer_(false) {
        if (allocate) this->allocate();
    }
    Grid_FFT(const Grid_FFT&) = delete;
    Grid_FFT& operator=(const Grid_FFT&) = delete;
    ~Grid_FFT() { reset(); }

    void reset() {
        if (h_) { mfb_grid_destroy(h_); h_ = nullptr; }
        if (data_) { std::free(data_); data_ = nullptr; cdata_ = nullptr; }
        ballocated_ = false;
    }

    void allocate() {  // src/grid_fft.cc:40-112 (non-distributed branch)
        if (ballocated_) return;
        int n[3] = {int(n_[0]), int(n_[1]), int(n_[2])};
        double L[3] = {length_[0], length_[1], length_[2]};
        mfb_detail::ensure_init(n);
        mfb_detail::check(mfb_grid_create(&h_, n, L, space_ == kspace_id ? MFB_KSPACE : MFB_RSPACE));
        int loc[4];
        const bool slab = mfb_grid_local(h_, loc) != 0;   // {local_0_start_, local_0_size_, local_1_start_, local_1_size_}
        if (slab != bdistributed) {
            mfb_grid_destroy(h_); h_ = nullptr;
            throw std::runtime_error(bdistributed ? "Grid_FFT<.., true>: the library is not slab-decomposed (mfb_dist_init)"
                                                  : "Grid_FFT<.., false>: the library is slab-decomposed - use Grid_FFT<real_t, true>");
        }
        ntot_ = mfb_grid_ntot(h_);   // (n2 + 2) n1 n0, or the slab's share (src/grid_fft.cc:45,122)
        data_ = static_cast<data_t*>(std::calloc(ntot_, sizeof(data_t)));
        if (!data_) throw std::runtime_error("Grid_FFT: host mirror allocation failed");
        cdata_ = reinterpret_cast<ccomplex_t*>(data_);
        fft_norm_fac_ = 1.0 / std::sqrt(real_t(n_[0]) * real_t(n_[1]) * real_t(n_[2]));
        npr_ = n_[2] + 2;
        npc_ = n_[2] / 2 + 1;
        for (int i = 0; i < 3; ++i) {
            nhalf_[i] = n_[i] / 2;
            kfac_[i] = 2.0 * M_PI / length_[i];
            kny_[i] = kfac_[i] * n_[i] / 2;
            dx_[i] = length_[i] / n_[i];
        }
        local_0_start_ = loc[0]; local_0_size_ = loc[1]; local_1_start_ = loc[2]; local_1_size_ = loc[3];
        set_sizes();
        ballocated_ = true;
        host_dirty_ = true;  // calloc'ed zeros are the truth until a device op runs
        dev_newer_ = false;
    }

    // ---- geometry (include/grid_fft.hh:104-138, 213-318)
    const Grid_FFT* get_grid(size_t) const { return this; }
    bool is_distributed() const noexcept { return bdistributed; }
    bool is_allocated() const noexcept { return ballocated_; }
    size_t memsize() const noexcept { return ntot_; }
    size_t size(size_t i) const noexcept { return sizes_[i]; }
    size_t local_size() const noexcept { return local_0_size_ * n_[1] * n_[2]; }
    size_t global_size() const noexcept { return n_[0] * n_[1] * n_[2]; }
    size_t global_size(size_t i) const noexcept { return n_[i]; }
    size_t rsize(size_t i) const noexcept { return (i == 0) ? local_0_size_ : n_[i]; }
    size_t get_idx(size_t i, size_t j, size_t k) const noexcept { return (i * sizes_[1] + j) * sizes_[3] + k; }
    size_t get_cell_idx_1d(size_t i, size_t j, size_t k) const noexcept { return ((i + local_0_start_) * n_[1] + j) * n_[2] + k; }
    real_t get_dx(int idim) const noexcept { return dx_[idim]; }
    const std::array<real_t, 3>& get_dx() const noexcept { return dx_; }
    bool is_nyquist_mode(size_t i, size_t j, size_t k) const {   // include/grid_fft.hh:140-147 (first index: local k_y row on slabs)
        return (i + size_t(local_1_start_) == n_[1] / 2) || (j == n_[0] / 2) || (k == n_[2] / 2);
    }
    template <typename ft> vec3_t<ft> get_k(size_t i, size_t j, size_t k) const noexcept {   // include/grid_fft.hh:281-296
        if (bdistributed) {   // transposed slab: the first index runs over the local ky rows
            const size_t ip = i + size_t(local_1_start_);
            return vec3_t<ft>(ft((real_t(j) - real_t(j > nhalf_[0]) * n_[0]) * kfac_[0]), ft((real_t(ip) - real_t(ip > nhalf_[1]) * n_[1]) * kfac_[1]),
                              ft((real_t(k) - real_t(k > nhalf_[2]) * n_[2]) * kfac_[2]));
        }
        return vec3_t<ft>(ft((real_t(i) - real_t(i > nhalf_[0]) * n_[0]) * kfac_[0]), ft((real_t(j) - real_t(j > nhalf_[1]) * n_[1]) * kfac_[1]),
                          ft((real_t(k) - real_t(k > nhalf_[2]) * n_[2]) * kfac_[2]));
    }
    template <typename ft> vec3_t<ft> get_r(size_t i, size_t j, size_t k) const noexcept {   // :213-228
        return vec3_t<ft>(ft(real_t(i + local_0_start_) * dx_[0]), ft(real_t(j) * dx_[1]), ft(real_t(k) * dx_[2]));
    }
    template <typename ft> vec3_t<ft> get_unit_r(size_t i, size_t j, size_t k) const noexcept {   // :230-240
        return vec3_t<ft>(ft(real_t(i + local_0_start_) / real_t(n_[0])), ft(real_t(j) / real_t(n_[1])), ft(real_t(k) / real_t(n_[2])));
    }
    template <typename ft> vec3_t<ft> get_unit_r_shifted(size_t i, size_t j, size_t k, const vec3_t<real_t> s) const noexcept {   // :242-252
        return vec3_t<ft>(ft((real_t(i + local_0_start_) + s.x) / real_t(n_[0])), ft((real_t(j) + s.y) / real_t(n_[1])),
                          ft((real_t(k) + s.z) / real_t(n_[2])));
    }
    vec3_t<size_t> get_cell_idx_3d(size_t i, size_t j, size_t k) const noexcept { return vec3_t<size_t>(i + local_0_start_, j, k); }
    // trilinear (CIC) read-out on the host mirror, single slab like the reference's ("doesn't work with MPI"):
    // get_cic at a position in box units, periodic in all three axes (include/grid_fft.hh:320-345); get_cic_kspace at a fractional
    // k-space index, periodic in the first two axes and clamped at the last retained k_z (:347-373; the PLT tables are read this way)
    data_t get_cic(const vec3_t<real_t>& v) const noexcept {
        real_t x[3];
        size_t lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            x[d] = real_t(std::fmod(v[d] / length_[d] + 1.0, 1.0) * n_[d]);
            lo[d] = size_t(x[d]); hi[d] = (lo[d] + 1) % n_[d]; x[d] -= real_t(lo[d]);
        }
        data_t val = 0.0;
        for (int c = 0; c < 8; ++c) {
            const bool a = c & 4, b = c & 2, e = c & 1;
            val += this->relem(a ? hi[0] : lo[0], b ? hi[1] : lo[1], e ? hi[2] : lo[2]) * (a ? x[0] : 1.0 - x[0]) * (b ? x[1] : 1.0 - x[1]) * (e ? x[2] : 1.0 - x[2]);
        }
        return val;
    }
    ccomplex_t get_cic_kspace(const vec3_t<real_t> v) const noexcept {
        real_t x[3];
        size_t lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            const int i = int(std::floor(v[d]));
            x[d] = v[d] - real_t(i); lo[d] = size_t(i);
            hi[d] = d < 2 ? size_t(i + 1) % size(d) : size_t(std::min(i + 1, int(size(2)) - 1));
        }
        ccomplex_t val = 0.0;
        for (int c = 0; c < 8; ++c) {
            const bool a = c & 4, b = c & 2, e = c & 1;
            val += this->kelem(a ? hi[0] : lo[0], b ? hi[1] : lo[1], e ? hi[2] : lo[2]) * ((a ? x[0] : 1.0 - x[0]) * (b ? x[1] : 1.0 - x[1]) * (e ? x[2] : 1.0 - x[2]));
        }
        return val;
    }
    std::array<size_t, 3> get_k3(size_t i, size_t j, size_t k) const noexcept {   // include/grid_fft.hh:315-318
        return bdistributed ? std::array<size_t, 3>({j, i + size_t(local_1_start_), k}) : std::array<size_t, 3>({i, j, k});
    }
    ccomplex_t gradient(int idim, std::array<size_t, 3> ijk) const {   // include/grid_fft.hh:376-385
        if (bdistributed) { ijk[0] += size_t(local_1_start_); std::swap(ijk[0], ijk[1]); }
        real_t rgrad = (ijk[idim] != nhalf_[idim]) ? (real_t(ijk[idim]) - real_t(ijk[idim] > nhalf_[idim]) * n_[idim]) * kfac_[idim] : 0.0;
        return ccomplex_t(0.0, rgrad);
    }

    // ---- host element access: lazily synchronised mirror (include/grid_fft.hh:179-211)
    data_t& operator[](size_t i) { to_host(true); return data_[i]; }
    data_t& relem(size_t i, size_t j, size_t k) { to_host(true); return data_[get_idx(i, j, k)]; }
    const data_t& relem(size_t i, size_t j, size_t k) const { to_host(false); return data_[get_idx(i, j, k)]; }
    ccomplex_t& kelem(size_t i, size_t j, size_t k) { to_host(true); return cdata_[get_idx(i, j, k)]; }
    const ccomplex_t& kelem(size_t i, size_t j, size_t k) const { to_host(false); return cdata_[get_idx(i, j, k)]; }
    ccomplex_t& kelem(size_t idx) { to_host(true); return cdata_[idx]; }
    const ccomplex_t& kelem(size_t idx) const { to_host(false); return cdata_[idx]; }
    data_t& relem(size_t idx) { to_host(true); return data_[idx]; }
    const data_t& relem(size_t idx) const { to_host(false); return data_[idx]; }
    //! plugins that hand data_ to C code (random_panphasia.cc:177) call this around the raw access
    data_t* host_data(bool will_write = true) { to_host(will_write); return data_; }

    // ---- device operations
    mfb_grid* device_handle() { to_device(); dev_newer_ = true; return h_; }
    void FourierTransformForward(bool do_transform = true) {
        if (space_ != kspace_id) {
            if (do_transform) { to_device(); mfb_detail::check(mfb_fft_forward(h_, 1)); dev_newer_ = true; }
            else mfb_detail::check(mfb_fft_forward(h_, 0));
            space_ = kspace_id;
            set_sizes();
        }
    }
    void FourierTransformBackward(bool do_transform = true) {
        if (space_ != rspace_id) {
            if (do_transform) { to_device(); mfb_detail::check(mfb_fft_backward(h_, 1)); dev_newer_ = true; }
            else mfb_detail::check(mfb_fft_backward(h_, 0));
            space_ = rspace_id;
            set_sizes();
        }
    }
    void FourierInterpolateCopyTo(Grid_FFT& grid_to) {
        to_device();
        grid_to.to_device();
        mfb_detail::check(mfb_fourier_resample(h_, grid_to.h_));
        space_ = kspace_id; set_sizes(); dev_newer_ = true;
        grid_to.space_ = kspace_id; grid_to.set_sizes(); grid_to.dev_newer_ = true;
    }
    void zero() noexcept { mfb_grid_zero(h_); host_dirty_ = false; dev_newer_ = true; }
    void copy_from(const Grid_FFT& g) {
        const_cast<Grid_FFT&>(g).to_device();
        mfb_detail::check(mfb_grid_copy(h_, g.h_));
        space_ = g.space_; set_sizes(); host_dirty_ = false; dev_newer_ = true;
    }
    Grid_FFT& operator*=(data_t x) { dev_op(mfb_grid_scale(prep(), x)); return *this; }
    Grid_FFT& operator/=(data_t x) { dev_op(mfb_grid_scale(prep(), 1.0 / x)); return *this; }
    Grid_FFT& apply_InverseLaplacian() { dev_op(mfb_grid_inverse_laplacian(prep())); after_k(); return *this; }
    Grid_FFT& apply_Laplacian() { dev_op(mfb_grid_laplacian(prep(), 0)); after_k(); return *this; }
    Grid_FFT& apply_negative_Laplacian() { dev_op(mfb_grid_laplacian(prep(), 1)); after_k(); return *this; }
    void zero_DC_mode() { dev_op(mfb_grid_zero_dc(prep())); }
    template <class vec3> void shift_field(const vec3& s, bool transform_back = true) {
        double ss[3] = {double(s[0]), double(s[1]), double(s[2])};
        dev_op(mfb_grid_shift_field(prep(), ss, transform_back ? 1 : 0));
        space_ = transform_back ? rspace_id : kspace_id;
        set_sizes();
    }
    real_t mean() const { double v; auto* s = const_cast<Grid_FFT*>(this); mfb_detail::check(mfb_grid_mean(s->prep(), &v)); return v; }
    real_t std() const { double v; auto* s = const_cast<Grid_FFT*>(this); mfb_detail::check(mfb_grid_std(s->prep(), &v)); return v; }
    real_t compute_2norm() const { double v; auto* s = const_cast<Grid_FFT*>(this); mfb_detail::check(mfb_grid_2norm(s->prep(), &v)); return v; }

    // ---- host-lambda functor family (runs on the mirror; include/grid_fft.hh:451-483, 698-808)
    template <typename F> void apply_function_k(const F& f) {
        to_host(true);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            auto& e = cdata_[get_idx(i, j, k)]; e = f(e);
        }
    }
    template <typename F> void apply_function_r(const F& f) {
        to_host(true);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            auto& e = data_[get_idx(i, j, k)]; e = f(e);
        }
    }

    // f(elem, k) / f(elem, r) over every element (include/grid_fft.hh:764-796)
    template <typename F> void apply_function_k_dep(const F& f) {
        to_host(true);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            auto& e = cdata_[get_idx(i, j, k)]; e = f(e, this->template get_k<real_t>(i, j, k));
        }
    }
    template <typename F> void apply_function_r_dep(const F& f) {
        to_host(true);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            auto& e = data_[get_idx(i, j, k)]; e = f(e, this->template get_r<real_t>(i, j, k));
        }
    }
    // this(x) = f(g1(x), g2(x), ...) in real space / Fourier space, optionally with the index triple or the k vector as
    // first argument (include/grid_fft.hh:698-762); the fNL / gNL terms of src/ic_generator.cc:402-455 are written with these
    template <typename F, typename... Grids> void assign_function_of_grids_r(const F& f, Grids&... grids) {
        to_host(true); (grids.to_host(false), ...);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k)
            data_[get_idx(i, j, k)] = f((grids.data_[grids.get_idx(i, j, k)])...);
    }
    template <typename F, typename... Grids> void assign_function_of_grids_k(const F& f, Grids&... grids) {
        to_host(true); (grids.to_host(false), ...);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k)
            cdata_[get_idx(i, j, k)] = f((grids.cdata_[grids.get_idx(i, j, k)])...);
    }
    template <typename F, typename... Grids> void assign_function_of_grids_ijk(const F& f, Grids&... grids) {
        to_host(true); (grids.to_host(false), ...);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k)
            cdata_[get_idx(i, j, k)] = f(std::array<size_t, 3>{i, j, k}, (grids.cdata_[grids.get_idx(i, j, k)])...);
    }
    template <typename F, typename... Grids> void assign_function_of_grids_kdep(const F& f, Grids&... grids) {
        to_host(true); (grids.to_host(false), ...);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k)
            cdata_[get_idx(i, j, k)] = f(this->template get_k<real_t>(i, j, k), (grids.cdata_[grids.get_idx(i, j, k)])...);
    }

    // ---- diagnostics used by output plugins (output_generic.cc:120, output_grafic2.cc:233-243)
    //! the real-space field as a dataset [n0][n1][n2] of `fname` (src/grid_fft.cc:806-1047, non-MPI branch).  Uses the HDF5 C
    //! API of whatever <hdf5.h> the host build provides; without HDF5 on the include path it throws.
    void Write_to_HDF5(std::string fname, std::string datasetname) const {
#ifdef MFB_MIRROR_HAVE_HDF5
        if (space_ != rspace_id) throw std::runtime_error("Write_to_HDF5: grid must be in real space");
        to_host(false);
        hid_t file = H5Fopen(fname.c_str(), H5F_ACC_RDWR, H5P_DEFAULT);
        if (file < 0) file = H5Fcreate(fname.c_str(), H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
        if (file < 0) throw std::runtime_error("Write_to_HDF5: cannot open " + fname);
        hsize_t dims[3] = {hsize_t(n_[0]), hsize_t(n_[1]), hsize_t(n_[2])};
        const hid_t fspace = H5Screate_simple(3, dims, nullptr);
        const hid_t dset = H5Dcreate2(file, datasetname.c_str(), H5T_NATIVE_DOUBLE, fspace, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        std::vector<double> plane(n_[1] * n_[2]);
        hsize_t count[3] = {1, hsize_t(n_[1]), hsize_t(n_[2])};
        const hid_t mspace = H5Screate_simple(3, count, nullptr);
        for (size_t i = 0; i < n_[0]; ++i) {   // one x-plane at a time: strip the two pad reals per row
            for (size_t j = 0; j < n_[1]; ++j) for (size_t k = 0; k < n_[2]; ++k) plane[j * n_[2] + k] = data_[(i * n_[1] + j) * npr_ + k];
            hsize_t offset[3] = {hsize_t(i), 0, 0};
            H5Sselect_hyperslab(fspace, H5S_SELECT_SET, offset, nullptr, count, nullptr);
            H5Dwrite(dset, H5T_NATIVE_DOUBLE, mspace, fspace, H5P_DEFAULT, plane.data());
        }
        H5Sclose(mspace); H5Dclose(dset); H5Sclose(fspace); H5Fclose(file);
#else
        (void)fname; (void)datasetname;
        throw std::runtime_error("Write_to_HDF5: this host build has no <hdf5.h>");
#endif
    }
    //! spherically binned power spectrum of the k-space field (src/grid_fft.cc:1133-1209): linear bins of width k_min = the smallest
    //! fundamental mode up to the largest Nyquist wave number; the vectors are resized here; modes with k_z > 0 count twice
    void Compute_PowerSpectrum(std::vector<double>& bin_k, std::vector<double>& bin_P, std::vector<double>& bin_eP, std::vector<size_t>& bin_count) {
        FourierTransformForward();
        to_host(false);
        const real_t kmax = std::max(std::max(kfac_[0] * nhalf_[0], kfac_[1] * nhalf_[1]), kfac_[2] * nhalf_[2]);
        const real_t kmin = std::min(std::min(kfac_[0], kfac_[1]), kfac_[2]), dk = kmin;
        const int nbins = int(kmax / kmin);
        bin_count.assign(nbins, 0); bin_k.assign(nbins, 0.0); bin_P.assign(nbins, 0.0); bin_eP.assign(nbins, 0.0);
        for (size_t i = 0; i < sizes_[0]; ++i) for (size_t j = 0; j < sizes_[1]; ++j) for (size_t k = 0; k < sizes_[2]; ++k) {
            const double kk = this->template get_k<double>(i, j, k).norm();
            if (!(kk >= kmin && kk < kmax)) continue;
            const int b = int(kk / dk);
            const double w = (k == 0) ? 1.0 : 2.0, p = std::norm(cdata_[get_idx(i, j, k)]);
            bin_P[b] += w * p; bin_eP[b] += w * p * p; bin_k[b] += w * kk; bin_count[b] += size_t(w);
        }
        const real_t volfac = length_[0] * length_[1] * length_[2] / std::pow(2.0 * M_PI, 3.0), fftfac = fft_norm_fac_ * fft_norm_fac_;
        for (int b = 0; b < nbins; ++b) {
            if (!bin_count[b]) continue;
            const double c = double(bin_count[b]);
            bin_k[b] /= c;
            bin_P[b] = bin_P[b] / c * volfac * fftfac;
            bin_eP[b] = std::sqrt(bin_eP[b] / c - bin_P[b] * bin_P[b]) / std::sqrt(c) * volfac * fftfac;
        }
    }
    //! the binned spectrum as a four-column text file (src/grid_fft.cc:1104-1131)
    void Write_PowerSpectrum(std::string ofname) {
        std::vector<double> bin_k, bin_P, bin_eP;
        std::vector<size_t> bin_count;
        Compute_PowerSpectrum(bin_k, bin_P, bin_eP, bin_count);
        std::ofstream ofs(ofname.c_str());
        ofs << "# " << std::setw(14) << "k" << std::setw(16) << "P(k)" << std::setw(16) << "err. P(k)" << std::setw(16) << "#modes" << "\n";
        for (size_t b = 0; b < bin_k.size(); ++b)
            if (bin_count[b] > 0)
                ofs << std::setw(16) << bin_k[b] << std::setw(16) << bin_P[b] << std::setw(16) << bin_eP[b] << std::setw(16) << bin_count[b] << std::endl;
    }
    //! external constraint field (src/grid_fft.cc:548-668; [random] ConstraintFieldFile): a cubic FP64 dataset replaces this grid
    //! - the grid takes the dataset's size - and is normalised to unit variance about zero as the reference does
    void Read_from_HDF5(const std::string Filename, const std::string ObjName) {
#ifdef MFB_MIRROR_HAVE_HDF5
        const hid_t file = H5Fopen(Filename.c_str(), H5F_ACC_RDONLY, H5P_DEFAULT);
        if (file < 0) throw std::runtime_error("Read_from_HDF5: cannot open " + Filename);
        const hid_t dset = H5Dopen2(file, ObjName.c_str(), H5P_DEFAULT);
        if (dset < 0) { H5Fclose(file); throw std::runtime_error("Read_from_HDF5: dataset '" + ObjName + "' does not exist in " + Filename); }
        const hid_t fspace = H5Dget_space(dset);
        hsize_t dims[3] = {0, 0, 0};
        if (H5Sget_simple_extent_ndims(fspace) != 3 || H5Sget_simple_extent_dims(fspace, dims, nullptr) < 0 || dims[0] != dims[1] || dims[0] != dims[2]) {
            H5Sclose(fspace); H5Dclose(dset); H5Fclose(file);
            throw std::runtime_error("Read_from_HDF5: '" + ObjName + "' is not a cubic three-dimensional dataset");
        }
        std::vector<double> buf(size_t(dims[0]) * dims[1] * dims[2]);
        const herr_t rc = H5Dread(dset, H5T_NATIVE_DOUBLE, H5S_ALL, H5S_ALL, H5P_DEFAULT, buf.data());
        H5Sclose(fspace); H5Dclose(dset); H5Fclose(file);
        if (rc < 0) throw std::runtime_error("Read_from_HDF5: reading '" + ObjName + "' failed");
        reset();
        for (int d = 0; d < 3; ++d) n_[d] = size_t(dims[d]);
        space_ = rspace_id;
        allocate();
        double s1 = 0, s2 = 0;
        for (double v : buf) { s1 += v; s2 += v * v; }
        s1 /= double(buf.size()); s2 /= double(buf.size());
        const double stdw = std::sqrt(s2 - s1 * s1);
        to_host(true);
        for (size_t i = 0; i < n_[0]; ++i) for (size_t j = 0; j < n_[1]; ++j) for (size_t k = 0; k < n_[2]; ++k)
            data_[(i * n_[1] + j) * npr_ + k] = buf[(i * n_[1] + j) * n_[2] + k] / stdw;
#else
        (void)Filename; (void)ObjName;
        throw std::runtime_error("Read_from_HDF5: this host build has no <hdf5.h>");
#endif
    }

private:
    mfb_grid* h_;
    // lazy coherence between the host mirror and the device buffer.  Host accessors may be called from inside an OpenMP team (the
    // reference's own loops are `#pragma omp parallel for` around kelem / relem): the one download a stale mirror needs is taken by
    // whichever thread comes first, the others wait for it; device operations are issued from serial code.
    mutable std::atomic<bool> host_dirty_, dev_newer_;
    mutable std::mutex sync_mu_;
    void set_sizes() {  // src/grid_fft.cc:98-111, 176-189, 230-233, 262-265
        if (space_ == rspace_id) { sizes_ = {size_t(local_0_size_), n_[1], n_[2], npr_}; }
        else { sizes_ = {size_t(local_1_size_), n_[0], npc_, npc_}; }
    }
    void to_host(bool will_write) const {
        if (dev_newer_.load(std::memory_order_acquire)) {
            std::lock_guard<std::mutex> lk(sync_mu_);
            if (dev_newer_.load(std::memory_order_relaxed)) {
                mfb_detail::check(mfb_grid_download(h_, data_));
                dev_newer_.store(false, std::memory_order_release);
            }
        }
        if (will_write && !host_dirty_.load(std::memory_order_relaxed)) host_dirty_.store(true, std::memory_order_relaxed);
    }
    void to_device() {
        if (host_dirty_) { mfb_detail::check(mfb_grid_upload(h_, data_)); host_dirty_ = false; }
    }
    mfb_grid* prep() { to_device(); return h_; }
    void dev_op(int rc) { mfb_detail::check(rc); dev_newer_ = true; }
    void after_k() { space_ = kspace_id; set_sizes(); }
    template <typename, bool> friend class Grid_FFT;
    template <typename, typename> friend class BaseConvolver;
    template <typename> friend class OrszagConvolver;
    template <typename> friend class mfb_lattice_access;
};
