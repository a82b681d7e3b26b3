"""Host-side mirrors of the reference classes on the hot path.

Names, argument meaning and error behaviour follow the reference:
  Grid_FFT            include/grid_fft.hh:40-914, src/grid_fft.cc
  OrszagConvolver     include/convolution.hh:28-293 (named products), :429-669
  op.*                include/operators.hh:25-35
  Particles           particle::lattice_generator / particle::container
                      (include/particle_generator.hh, include/particle_container.hh)
Every method is a thin call into the C ABI (include/mfb200.h); numpy is only used to move
host buffers in and out.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

kspace_id, rspace_id = 0, 1  # enum space_t, include/grid_fft.hh:29
_state = {"lib": None}


def _lib():
    if _state["lib"] is None:
        raise L.MfbError("monofonic_b200.init() has not been called")
    return _state["lib"]


def init(device: int = 0, lib_path: str | None = None):
    """Load libmfb200.so and bind it to CUDA device `device`.  Fails loudly without a GPU."""
    lib = L.load(lib_path)
    L.check(lib.mfb_init(int(device)), lib)
    _state["lib"] = lib
    return lib


def finalize():
    if _state["lib"] is not None:
        _state["lib"].mfb_finalize()
        _state["lib"] = None


def synchronize():
    L.check(_lib().mfb_synchronize(), _lib())


def kernel_launches() -> int:
    return int(_lib().mfb_kernel_launches())


def _i3(v):
    return (C.c_int * 3)(*[int(x) for x in v])


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _i2(v):
    return (C.c_int * 2)(*[int(x) for x in v])


class _OutOp:
    """op::assign_to / add_to / subtract_from / multiply_add_to  (include/operators.hh:25-35)"""

    def __init__(self, grid, code, alpha=0.0):
        self.grid, self.code, self.alpha = grid, code, float(alpha)


class op:
    @staticmethod
    def assign_to(g):
        return _OutOp(g, 0)

    @staticmethod
    def add_to(g):
        return _OutOp(g, 1)

    @staticmethod
    def subtract_from(g):
        return _OutOp(g, 2)

    @staticmethod
    def multiply_add_to(g, x):
        return _OutOp(g, 3, x)


class Grid_FFT:
    """Device-resident FFT-able grid (Grid_FFT<real_t>, non-distributed layout)."""

    def __init__(self, N, Lbox, allocate=True, initialspace=rspace_id):
        self.n_ = tuple(int(x) for x in N)
        self.length_ = tuple(float(x) for x in Lbox)
        self._h = C.c_void_p()
        self._initialspace = initialspace
        self._lib = _lib()
        if allocate:
            self.allocate()

    # -- allocation ---------------------------------------------------------------------
    def allocate(self):
        if self._h:
            return
        L.check(self._lib.mfb_grid_create(C.byref(self._h), _i3(self.n_), _d3(self.length_), self._initialspace), self._lib)
        loc = (C.c_int * 4)()
        self.distributed_ = bool(self._lib.mfb_grid_local(self._h, loc))
        # slab bookkeeping with the reference's names (include/grid_fft.hh:60-63)
        self.local_0_start_, self.local_0_size_, self.local_1_start_, self.local_1_size_ = [int(v) for v in loc]

    def is_distributed(self):
        return self.distributed_

    def is_allocated(self):
        return bool(self._h)

    def reset(self):
        if self._h:
            self._lib.mfb_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.reset()
        except Exception:
            pass

    # -- geometry -----------------------------------------------------------------------
    @property
    def space_(self):
        return self._lib.mfb_grid_space(self._h)

    @property
    def ntot_(self):
        return int(self._lib.mfb_grid_ntot(self._h))

    def size(self, i):
        """local extents; k-space of a slab is transposed: (ky_loc, kx, kz) (src/grid_fft.cc:176-189)"""
        n0, n1, n2 = self.n_
        if self.space_ == rspace_id:
            return (self.local_0_size_, n1, n2, n2 + 2)[i]
        if self.distributed_:
            return (self.local_1_size_, n0, n2 // 2 + 1, n2 // 2 + 1)[i]
        return (n0, n1, n2 // 2 + 1, n2 // 2 + 1)[i]

    def global_size(self):
        return self.n_[0] * self.n_[1] * self.n_[2]

    # -- host <-> device ----------------------------------------------------------------
    def upload_raw(self, buf: np.ndarray):
        buf = np.ascontiguousarray(buf, dtype=np.float64).reshape(-1)
        assert buf.size == self.ntot_, (buf.size, self.ntot_)
        L.check(self._lib.mfb_grid_upload(self._h, buf.ctypes.data_as(C.c_void_p)), self._lib)

    def fill_noise_ngenic_layout(self, seed: int):
        """white noise in k-space with the layout of RNG_ngenic::Fill_Grid, generated on the device (Philox4x32-10)"""
        L.check(self._lib.mfb_noise_fill_ngenic(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF), self._lib)

    def fill_noise_ngenic(self, seed: int):
        """the white noise RNG_ngenic::Fill_Grid gives for this [random] seed - the reference's own generator (RANLUX,
        gsl_rng_ranlxd1), one stream per column - generated on the device"""
        L.check(self._lib.mfb_noise_fill_ngenic_ranlxd(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF), self._lib)

    def stream_in(self, buf: np.ndarray):
        """asynchronous upload_raw on the ingest stream (`buf` should be pinned and must stay alive and unchanged until the
        copy has run); the compute stream continues with work that does not touch this grid.  Pair with stream_in_wait()."""
        if buf.dtype != np.float64 or buf.size != self.ntot_ or not buf.flags["C_CONTIGUOUS"]:
            raise ValueError(f"stream_in: host buffer must be C-contiguous float64[{self.ntot_}]")
        L.check(self._lib.mfb_grid_stream_in(self._h, buf.ctypes.data_as(C.c_void_p)), self._lib)

    def stream_in_wait(self):
        """the compute stream (not the host) waits for a pending stream_in"""
        L.check(self._lib.mfb_grid_stream_in_wait(self._h), self._lib)

    def download_raw(self) -> np.ndarray:
        buf = np.empty(self.ntot_, dtype=np.float64)
        L.check(self._lib.mfb_grid_download(self._h, buf.ctypes.data_as(C.c_void_p)), self._lib)
        return buf

    def set_real(self, x: np.ndarray):
        """fill from a [n0][n1][n2] real array (relem view); grid is declared real-space"""
        n0, n1, n2 = self.n_
        l0, o0 = self.local_0_size_, self.local_0_start_
        if x.shape[0] == n0 and l0 != n0:   # global array given to a slab: keep the local planes
            x = x[o0:o0 + l0]
        buf = np.zeros(self.ntot_).reshape(-1, n2 + 2)
        buf[:l0 * n1, :n2] = np.asarray(x).reshape(l0 * n1, n2)
        self.FourierTransformBackward(False)
        self.upload_raw(buf)

    def get_real(self) -> np.ndarray:
        """the local slab [local_0_size][n1][n2]"""
        n0, n1, n2 = self.n_
        assert self.space_ == rspace_id
        l0 = self.local_0_size_
        return self.download_raw().reshape(-1, n2 + 2)[:l0 * n1, :n2].reshape(l0, n1, n2).copy()

    def set_kspace(self, X: np.ndarray):
        """fill from a [kx][ky][kz] half-spectrum (kelem view); grid is declared k-space"""
        n0, n1, n2 = self.n_
        nzc = n2 // 2 + 1
        if self.distributed_:
            # slab k-layout is [ky_loc][kx][kz]; accept the global [kx][ky][kz] array or the local transposed one
            l1, o1 = self.local_1_size_, self.local_1_start_
            if X.shape == (n0, n1, nzc):
                X = np.transpose(X[:, o1:o1 + l1, :], (1, 0, 2))
            assert X.shape == (l1, n0, nzc)
        else:
            assert X.shape == (n0, n1, nzc)
        X = np.ascontiguousarray(X, dtype=np.complex128)
        buf = np.zeros(self.ntot_)
        buf[:X.size * 2] = X.view(np.float64).reshape(-1)
        self.FourierTransformForward(False)
        self.upload_raw(buf)

    def get_kspace(self) -> np.ndarray:
        """[kx][ky][kz]; for a slab the local part in the transposed layout [ky_loc][kx][kz]"""
        n0, n1, n2 = self.n_
        nzc = n2 // 2 + 1
        assert self.space_ == kspace_id
        raw = self.download_raw().view(np.complex128)
        if self.distributed_:
            l1 = self.local_1_size_
            return raw[:l1 * n0 * nzc].reshape(l1, n0, nzc).copy()
        return raw.reshape(n0, n1, nzc).copy()

    # -- transforms ---------------------------------------------------------------------
    def FourierTransformForward(self, do_transform=True):
        L.check(self._lib.mfb_fft_forward(self._h, int(bool(do_transform))), self._lib)

    def FourierTransformBackward(self, do_transform=True):
        L.check(self._lib.mfb_fft_backward(self._h, int(bool(do_transform))), self._lib)

    def FourierInterpolateCopyTo(self, grid_to: "Grid_FFT"):
        L.check(self._lib.mfb_fourier_resample(self._h, grid_to._h), self._lib)

    # -- element-wise -------------------------------------------------------------------
    def zero(self):
        L.check(self._lib.mfb_grid_zero(self._h), self._lib)

    def copy_from(self, g: "Grid_FFT"):
        L.check(self._lib.mfb_grid_copy(self._h, g._h), self._lib)

    def __imul__(self, x):
        L.check(self._lib.mfb_grid_scale(self._h, float(x)), self._lib)
        return self

    def __itruediv__(self, x):
        L.check(self._lib.mfb_grid_scale(self._h, 1.0 / float(x)), self._lib)
        return self

    def axpy(self, a, x: "Grid_FFT", b):
        L.check(self._lib.mfb_grid_axpy(self._h, float(a), x._h, float(b)), self._lib)

    def affine(self, a, b):
        L.check(self._lib.mfb_grid_affine(self._h, float(a), float(b)), self._lib)

    def zero_DC_mode(self):
        L.check(self._lib.mfb_grid_zero_dc(self._h), self._lib)

    def apply_InverseLaplacian(self):
        L.check(self._lib.mfb_grid_inverse_laplacian(self._h), self._lib)
        return self

    def apply_Laplacian(self):
        L.check(self._lib.mfb_grid_laplacian(self._h, 0), self._lib)
        return self

    def apply_negative_Laplacian(self):
        L.check(self._lib.mfb_grid_laplacian(self._h, 1), self._lib)
        return self

    def shift_field(self, s, transform_back=True):
        L.check(self._lib.mfb_grid_shift_field(self._h, _d3(s), int(bool(transform_back))), self._lib)

    # -- reductions ---------------------------------------------------------------------
    def _red(self, fn):
        out = C.c_double()
        L.check(fn(self._h, C.byref(out)), self._lib)
        return out.value

    def mean(self):
        return self._red(self._lib.mfb_grid_mean)

    def std(self):
        return self._red(self._lib.mfb_grid_std)

    def compute_2norm(self):
        return self._red(self._lib.mfb_grid_2norm)

    def max_neighbour_diff(self):
        return self._red(self._lib.mfb_grid_max_neighbour_diff)


class Grid_FFT_complex:
    """Grid_FFT<ccomplex_t> (PPT / Eulerian branch, src/ic_generator.cc:690,778): dense [n0][n1][n2] complex grid with
    c2c transforms of the same symmetric normalisation.  Single slab."""

    def __init__(self, N, Lbox):
        self.n_ = tuple(int(x) for x in N)
        self.length_ = tuple(float(x) for x in Lbox)
        self._lib = _lib()
        self._h = C.c_void_p()
        L.check(self._lib.mfb_cgrid_create(C.byref(self._h), _i3(self.n_), _d3(self.length_)), self._lib)
        loc = (C.c_int * 4)()
        self.distributed_ = bool(self._lib.mfb_cgrid_local(self._h, loc))
        self.local_0_start_, self.local_0_size_, self.local_1_start_, self.local_1_size_ = (int(v) for v in loc)

    def __del__(self):
        try:
            if self._h:
                self._lib.mfb_cgrid_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    @property
    def space_(self):
        return self._lib.mfb_cgrid_space(self._h)

    def set(self, x: np.ndarray, space=rspace_id):
        """fill from the global [n0][n1][n2] array (real space: [x][y][z]; k-space: [kx][ky][kz]); a slab keeps its part -
        planes of x in real space, the transposed [ky_loc][kx][kz] block in k-space"""
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert x.shape == self.n_
        if self.distributed_:
            if int(space) == rspace_id:
                loc = x[self.local_0_start_:self.local_0_start_ + self.local_0_size_]
            else:
                loc = np.transpose(x[:, self.local_1_start_:self.local_1_start_ + self.local_1_size_], (1, 0, 2))
            buf = np.zeros(int(self._lib.mfb_cgrid_nelem(self._h)), dtype=np.complex128)
            buf[:loc.size] = np.ascontiguousarray(loc).reshape(-1)
            x = buf
        L.check(self._lib.mfb_cgrid_upload(self._h, x.ctypes.data_as(C.c_void_p), int(space)), self._lib)

    def get(self) -> np.ndarray:
        """single slab: the [n0][n1][n2] array.  Slabs: the local part - [x_loc][n1][n2] in real space, [ky_loc][n0][n2] in
        k-space"""
        if not self.distributed_:
            out = np.empty(self.n_, dtype=np.complex128)
            L.check(self._lib.mfb_cgrid_download(self._h, out.ctypes.data_as(C.c_void_p)), self._lib)
            return out
        buf = np.empty(int(self._lib.mfb_cgrid_nelem(self._h)), dtype=np.complex128)
        L.check(self._lib.mfb_cgrid_download(self._h, buf.ctypes.data_as(C.c_void_p)), self._lib)
        n0, n1, n2 = self.n_
        shape = (self.local_0_size_, n1, n2) if self.space_ == rspace_id else (self.local_1_size_, n0, n2)
        return buf[:shape[0] * shape[1] * shape[2]].reshape(shape).copy()

    def copy_from(self, g: "Grid_FFT_complex"):
        L.check(self._lib.mfb_cgrid_copy(self._h, g._h), self._lib)

    def FourierTransformForward(self, do_transform=True):
        L.check(self._lib.mfb_cgrid_fft_forward(self._h, int(bool(do_transform))), self._lib)

    def FourierTransformBackward(self, do_transform=True):
        L.check(self._lib.mfb_cgrid_fft_backward(self._h, int(bool(do_transform))), self._lib)

    def apply_gradient(self, idim):
        """apply_function_k_dep(x * i k[idim]) with k from get_k (src/ic_generator.cc:790-792)"""
        L.check(self._lib.mfb_cgrid_gradient(self._h, int(idim)), self._lib)


class OrszagConvolver:
    """OrszagConvolver<real_t> with the named products of BaseConvolver."""

    _create = "mfb_conv_create"

    def __init__(self, N, Lbox):
        self._lib = _lib()
        self._h = C.c_void_p()
        self.n_ = tuple(int(x) for x in N)
        L.check(getattr(self._lib, self._create)(C.byref(self._h), _i3(N), _d3(Lbox)), self._lib)

    def __del__(self):
        try:
            if self._h:
                self._lib.mfb_conv_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def workspace_bytes(self):
        return int(self._lib.mfb_conv_workspace_bytes(self._h))

    def borrow_accumulator(self, particles):
        """lend the (idle during the source stage) position/velocity arrays of `particles` as accumulator workspace when the
        convolver could not afford its own; None returns the loan"""
        L.check(self._lib.mfb_conv_borrow_accumulator(self._h, particles._h if particles is not None else None), self._lib)

    def sources_info(self):
        """what the last mfb_lpt_sources call did: {"path": 1 (per product) | 2 (shared x passes, multi-product z pass),
        "slots", "loops", "x_passes", "y_passes", "z_launches"}"""
        info = (C.c_longlong * 6)()
        path = int(self._lib.mfb_conv_sources_info(self._h, info))
        return {"path": path, "slots": int(info[0]), "loops": int(info[1]), "x_passes": int(info[2]),
                "y_passes": int(info[3]), "z_launches": int(info[4])}

    def grouped(self):
        """True when the accumulator workspace was granted, i.e. mfb_lpt_sources sums the products of one target before a
        single forward transform"""
        return bool(self._lib.mfb_conv_grouped(self._h))

    def convolve_Hessians(self, *args):
        """(inl, d2l, inr, d2r, output_op)  or  (inl, d2l, inm, d2m, inr, d2r, output_op)"""
        if len(args) == 5:
            inl, d2l, inr, d2r, o = args
            rc = self._lib.mfb_conv_hessians(self._h, inl._h, _i2(d2l), inr._h, _i2(d2r), o.grid._h, o.code, o.alpha)
        elif len(args) == 7:
            inl, d2l, inm, d2m, inr, d2r, o = args
            rc = self._lib.mfb_conv_hessians3(self._h, inl._h, _i2(d2l), inm._h, _i2(d2m), inr._h, _i2(d2r), o.grid._h,
                                              o.code, o.alpha)
        else:
            raise TypeError("convolve_Hessians takes 5 or 7 arguments")
        L.check(rc, self._lib)

    def convolve_SumOfHessians(self, inl, d2l, inr, d2r1, d2r2, o):
        L.check(self._lib.mfb_conv_sum_hessians(self._h, inl._h, _i2(d2l), inr._h, _i2(d2r1), _i2(d2r2), o.grid._h,
                                                o.code, o.alpha), self._lib)

    def convolve_DifferenceOfHessians(self, inl, d2l, inr, d2r1, d2r2, o):
        L.check(self._lib.mfb_conv_diff_hessians(self._h, inl._h, _i2(d2l), inr._h, _i2(d2r1), _i2(d2r2), o.grid._h,
                                                 o.code, o.alpha), self._lib)

    def convolve_Gradients(self, inl, d1l, inr, d1r, o):
        L.check(self._lib.mfb_conv_gradients(self._h, inl._h, int(d1l[0]), inr._h, int(d1r[0]), o.grid._h, o.code, o.alpha),
                self._lib)

    def convolve_Gradient_and_Hessian(self, inl, d1l, inr, d2r, o):
        L.check(self._lib.mfb_conv_gradient_hessian(self._h, inl._h, int(d1l[0]), inr._h, _i2(d2r), o.grid._h, o.code,
                                                    o.alpha), self._lib)

    def multiply_field(self, inl, inr, o):
        L.check(self._lib.mfb_conv_multiply(self._h, inl._h, inr._h, o.grid._h, o.code, o.alpha), self._lib)

    def convolve2(self, spec1, spec2, o):
        """convolve2(kf1, kf2, output_op) (include/convolution.hh:512-535) with the operand spectra handed over as k-space
        grids holding kf1(i,j,k), kf2(i,j,k) instead of host lambdas"""
        L.check(self._lib.mfb_conv_generic(self._h, spec1._h, spec2._h, o.grid._h, o.code, o.alpha), self._lib)

    def convolve3(self, spec1, spec2, spec3, o):
        """convolve3(kf1, kf2, kf3, output_op) (:542-549): (spec1 * spec2) truncated to the grid, then * spec3"""
        L.check(self._lib.mfb_conv_generic3(self._h, spec1._h, spec2._h, spec3._h, o.grid._h, o.code, o.alpha), self._lib)


class NaiveConvolver(OrszagConvolver):
    """NaiveConvolver<real_t> (include/convolution.hh:296-426): the same named products without 3/2 padding."""
    _create = "mfb_conv_create_naive"


LATTICE = {"sc": 0, "bcc": 1, "fcc": 2, "rsc": 3, "glass": -1, "masked": -2}


def particle_type_mask(DoBaryons: bool, ParticleMaskType: int = 3):
    """the 2x2x2 species mask of a masked lattice as lattice_generator's ctor sets it up from [setup] DoBaryons and
    ParticleMaskType (include/particle_generator.hh:184-205): index 4a+2b+c for cell parity (a,b,c)"""
    if not DoBaryons:
        return [0] * 8
    m = [-1] * 8
    m[0] = 0
    if ParticleMaskType & 1:
        m[1] = m[2] = m[4] = 0      # edge centres
    if ParticleMaskType & 2:
        m[3] = m[5] = m[6] = 0      # face centres
    m[7] = 1                        # baryon at the cell centre
    return m


class lattice_gradient:
    """particle::lattice_gradient (include/particle_plt.hh:50-587, build option ENABLE_PLT) as the device sees it: the four
    k-space tables its gradient() / vfac_corr() read (Re D_xx_, D_yy_, D_zz_, D_xy_ on the ng^3 helper mesh, computed by the
    host's eigen-decomposition), uploaded once.  Passed to LPTPipeline as `lg` in place of op::fourier_gradient."""

    def __init__(self, ngridself, GridRes, BoxLength, D_r, D_theta, D_phi, mu):
        self._lib = _lib()
        self._h = C.c_void_p()
        ng = int(ngridself)
        tabs = [np.ascontiguousarray(t, dtype=np.float64) for t in (D_r, D_theta, D_phi, mu)]
        for t in tabs:
            if t.shape != (ng, ng, ng // 2 + 1):
                raise ValueError(f"lattice_gradient: tables must be [{ng}][{ng}][{ng // 2 + 1}]")
        arr = (C.c_void_p * 4)(*[t.ctypes.data for t in tabs])
        L.check(self._lib.mfb_gradient_create_plt(C.byref(self._h), ng, int(GridRes), float(BoxLength), arr), self._lib)

    def __del__(self):
        try:
            if self._h:
                self._lib.mfb_gradient_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass


class Particles:
    """particle::lattice_generator + particle::container: Bravais lattices, the masked sc lattice (`mask` = the 2x2x2 species
    mask, see particle_type_mask) and glass loads (`glass` = (positions in grid units [np][3], offset of the first particle))."""

    def __init__(self, lattice_type, lattice_index, b64reals, b64ids, bwithmasses, IDoffset, field: Grid_FFT, mask=None, glass=None):
        self._lib = _lib()
        self._h = C.c_void_p()
        lt = LATTICE[lattice_type] if isinstance(lattice_type, str) else int(lattice_type)
        self.lattice_type, self.lattice_index = lt, int(lattice_index)
        self.b64reals, self.b64ids, self.bwithmasses = bool(b64reals), bool(b64ids), bool(bwithmasses)
        if lt == LATTICE["masked"]:
            m = (C.c_int * 8)(*[int(v) for v in (mask if mask is not None else [0] * 8)])
            L.check(self._lib.mfb_particles_create_masked(C.byref(self._h), _i3(field.n_), self.lattice_index, m, int(self.b64reals),
                                                          int(self.b64ids), int(self.bwithmasses)), self._lib)
        elif lt == LATTICE["glass"]:
            if glass is None:
                raise RuntimeError("glass load needs the glass positions (glass_loader::glass_posr)")
            if self.bwithmasses:
                raise RuntimeError("cannot have individual particle masses for glasses")   # particle_generator.hh:399-403
            pos = np.ascontiguousarray(glass[0], dtype=np.float64).reshape(-1, 3)
            L.check(self._lib.mfb_particles_create_glass(C.byref(self._h), _i3(field.n_), pos.ctypes.data_as(C.c_void_p), pos.shape[0],
                                                         int(glass[1]), int(self.b64reals), int(self.b64ids)), self._lib)
        else:
            L.check(self._lib.mfb_particles_create(C.byref(self._h), _i3(field.n_), lt, int(self.b64reals), int(self.b64ids),
                                                   int(self.bwithmasses)), self._lib)
        L.check(self._lib.mfb_particles_set_ids(self._h, int(IDoffset)), self._lib)

    def __del__(self):
        try:
            if self._h:
                self._lib.mfb_particles_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def get_local_num_particles(self):
        return int(self._lib.mfb_particles_count(self._h))

    def reset_species(self, lattice_index, IDoffset):
        """reuse the container for the next species (the reference makes a new lattice_generator per species,
        src/ic_generator.cc:636-645: secondary lattice and ID offset 1 for the baryons)"""
        self.lattice_index = int(lattice_index)
        L.check(self._lib.mfb_particles_set_ids(self._h, int(IDoffset)), self._lib)

    def set_positions(self, idim, lunit, field: Grid_FFT):
        L.check(self._lib.mfb_particles_set_positions(self._h, self.lattice_index, int(idim), float(lunit), field._h), self._lib)

    def set_velocities(self, idim, field: Grid_FFT):
        L.check(self._lib.mfb_particles_set_velocities(self._h, self.lattice_index, int(idim), field._h), self._lib)

    def set_masses(self, munit, field: Grid_FFT):
        L.check(self._lib.mfb_particles_set_masses(self._h, self.lattice_index, float(munit), field._h), self._lib)

    _ATTR = {"pos": 0, "vel": 1, "ids": 2, "mass": 3}

    def stream_out(self, attr: str, host: np.ndarray):
        """hand one finished attribute to the host (output plugin side) on the copy stream while the compute stream goes
        on with the next component; `host` should be pinned.  Pair with stream_wait()."""
        n = self.get_local_num_particles()
        rt = np.float64 if self.b64reals else np.float32
        it = np.uint64 if self.b64ids else np.uint32
        want = (it, n) if attr == "ids" else (rt, n if attr == "mass" else 3 * n)
        if host.dtype != want[0] or host.size != want[1] or not host.flags["C_CONTIGUOUS"]:
            raise ValueError(f"stream_out({attr}): host buffer must be C-contiguous {want[0].__name__}[{want[1]}]")
        L.check(self._lib.mfb_particles_stream_out(self._h, self._ATTR[attr], host.ctypes.data_as(C.c_void_p)), self._lib)

    def stream_wait(self):
        L.check(self._lib.mfb_particles_stream_wait(self._h), self._lib)

    def download(self, pos=True, vel=True, ids=True, mass=True, out=None):
        """-> dict of host arrays (positions/velocities [np,3], ids [np], masses [np]).
        `out` may hold preallocated (e.g. pinned) arrays under the same keys."""
        n = self.get_local_num_particles()
        rt = np.float64 if self.b64reals else np.float32
        it = np.uint64 if self.b64ids else np.uint32
        out = out if out is not None else {}
        res = {}
        if pos:
            res["pos"] = out.get("pos") if "pos" in out else np.empty((n, 3), rt)
        if vel:
            res["vel"] = out.get("vel") if "vel" in out else np.empty((n, 3), rt)
        if ids:
            res["ids"] = out.get("ids") if "ids" in out else np.empty(n, it)
        if mass and self.bwithmasses:
            res["mass"] = out.get("mass") if "mass" in out else np.empty(n, rt)
        p = lambda k: res[k].ctypes.data_as(C.c_void_p) if k in res else None
        L.check(self._lib.mfb_particles_download(self._h, p("pos"), p("vel"), p("ids"), p("mass")), self._lib)
        return res
