"""CPU oracle for the monofonIC LPT field pipeline  --  TEST INFRASTRUCTURE ONLY.

This is a numpy (pocketfft, FP64) restatement of the reference algorithm for the hot path
named in BASELINE.json.  It is the *checker* used by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product path (monofonic_b200/) never
imports it; if the CUDA library is missing the product fails loudly instead.

Parity status: PINNED, twice.
  (1) tests/test_oracle_golden.py checks this module against all five golden regression files of the
      reference (tests/references/*.hdf5, converted to tests/golden/*.npz by tests/golden/make_golden.py) at
      the reference's own tolerance np.allclose(rtol=1e-9, atol=1e-14) (tests/scripts/compare_hdf5.py:66-75):
      end results (displacements, velocities, densities, masses, IDs).
  (2) tests/test_oracle_vs_reference.py checks it against the reference's OWN Grid_FFT / OrszagConvolver / op::
      sources, compiled where they lie (oracle/Makefile -> oracle/_ref/ref_driver; FFTW3 replaced by the exact-DFT
      shim of oracle/refshim/) and run in the build container: every named product with every output operator,
      the phi2 / phi3 / A3 source stage, Poisson solve, shift_field, FourierInterpolateCopyTo, the NaiveConvolver,
      c2c transforms - agreement 1e-15..2e-14 - in the reference's non-MPI build AND in its USE_MPI build on one rank
      (oracle/_ref/ref_driver_mpi; one-rank mpi.h / fftw3-mpi.h shims), the latter also on fields with populated
      Nyquist planes, whose treatment (dropped by pad / unpad) this module restates; and particle::lattice_generator
      (MPI build): IDs and bare-lattice positions bit-identical, displaced positions to 1 ulp of the box, for sc / bcc /
      fcc / rsc and both lattice indices.  The outputs are committed as
      tests/golden/ref_intermediates_32.npz for the GPU box.

Every function cites the reference file:line it follows (paths relative to /root/reference).

Conventions (include/grid_fft.hh, src/grid_fft.cc):
  * real grid  [n0][n1][n2]           (the two pad reals per z-row are not modelled)
  * k grid     [kx][ky][kz], kz = 0..n2/2   (non-transposed layout; the reference's
    MPI build stores [ky][kx][kz] - a pure storage permutation, src/grid_fft.cc:125-128)
  * symmetric normalisation: fwd = rfftn/sqrt(N^3), bwd = irfftn*sqrt(N^3)
    (src/grid_fft.cc:69,199-205,222-224,256-257)
  * Orszag pad/unpad follow the MPI-build semantics (Nyquist planes are never copied),
    because the goldens were produced by an MPI-enabled build.
"""
from __future__ import annotations

import numpy as np

try:  # multi-threaded pocketfft if scipy is around (same algorithm as numpy.fft)
    import scipy.fft as _sfft
except Exception:  # pragma: no cover
    _sfft = None


# --------------------------------------------------------------------------------------
# FFT backend (pluggable so that bench.py can time the same algorithm with MKL via torch)
# --------------------------------------------------------------------------------------
class FFTBackend:
    """numpy/scipy pocketfft.  workers>1 uses scipy's threaded pocketfft."""

    def __init__(self, workers: int = 1):
        self.workers = int(workers)

    def rfftn(self, x):
        if _sfft is not None:
            return _sfft.rfftn(x, workers=self.workers)
        return np.fft.rfftn(x)

    def irfftn(self, X, s):
        if _sfft is not None:
            return _sfft.irfftn(X, s=s, workers=self.workers)
        return np.fft.irfftn(X, s=s)

    def fftn(self, x):
        if _sfft is not None:
            return _sfft.fftn(x, workers=self.workers)
        return np.fft.fftn(x)

    def ifftn(self, x):
        if _sfft is not None:
            return _sfft.ifftn(x, workers=self.workers)
        return np.fft.ifftn(x)


class TorchFFTBackend(FFTBackend):
    """torch CPU (MKL) FFT with all host threads - only used as a *timed CPU baseline*."""

    def __init__(self, threads: int | None = None):
        import torch

        self.torch = torch
        if threads:
            torch.set_num_threads(int(threads))
        self.workers = torch.get_num_threads()

    def rfftn(self, x):
        t = self.torch
        return t.fft.rfftn(t.from_numpy(np.ascontiguousarray(x))).numpy()

    def irfftn(self, X, s):
        t = self.torch
        return t.fft.irfftn(t.from_numpy(np.ascontiguousarray(X)), s=tuple(s)).numpy()

    def fftn(self, x):
        t = self.torch
        return t.fft.fftn(t.from_numpy(np.ascontiguousarray(x))).numpy()

    def ifftn(self, x):
        t = self.torch
        return t.fft.ifftn(t.from_numpy(np.ascontiguousarray(x))).numpy()


_DEFAULT_FFT = FFTBackend(1)


# --------------------------------------------------------------------------------------
# Grid helpers  (include/grid_fft.hh)
# --------------------------------------------------------------------------------------
class GridSpec:
    """Geometry of an N^3 grid in a box of side L  (src/grid_fft.cc:82-91)."""

    def __init__(self, n: int, boxlen: float, fft: FFTBackend | None = None):
        self.n = int(n)
        self.L = float(boxlen)
        self.nh = self.n // 2
        self.nzc = self.n // 2 + 1
        self.kfac = 2.0 * np.pi / self.L
        self.dx = self.L / self.n
        self.norm = 1.0 / np.sqrt(float(self.n) ** 3)
        self.fft = fft or _DEFAULT_FFT
        i = np.arange(self.n)
        # signed integer wavenumber: n - N*[n > N/2]   (include/grid_fft.hh:287-293; note '>')
        self.nsigned = (i - self.n * (i > self.nh)).astype(np.float64)
        self.k1d = self.nsigned * self.kfac
        # gradient wavenumber: zero at the Nyquist index (include/grid_fft.hh:376-385,
        # include/operators.hh:50-55)
        self.g1d = np.where(i == self.nh, 0.0, self.k1d)

    # -- FFTs with the reference's symmetric normalisation (src/grid_fft.cc:199-270)
    def fwd(self, x):
        return self.fft.rfftn(x) * self.norm

    def bwd(self, X):
        n = self.n
        return self.fft.irfftn(X, (n, n, n)) * (float(n) ** 3 * self.norm)

    # -- broadcastable k vectors on the half-spectrum [kx][ky][kz]
    def kvec(self):
        kx = self.k1d[:, None, None]
        ky = self.k1d[None, :, None]
        kz = self.k1d[None, None, : self.nzc]
        return kx, ky, kz

    def k2(self):
        kx, ky, kz = self.kvec()
        return kx * kx + ky * ky + kz * kz

    def grad(self, d):
        """i*k_d with k_d := 0 at that axis' Nyquist index  (include/grid_fft.hh:376-385)."""
        shape = [1, 1, 1]
        if d == 2:
            g = self.g1d[: self.nzc]
        else:
            g = self.g1d
        shape[d] = g.shape[0]
        return 1j * g.reshape(shape)

    def zeros_k(self):
        return np.zeros((self.n, self.n, self.nzc), dtype=np.complex128)


def zero_dc(X):
    """Grid_FFT::zero_DC_mode in k-space  (include/grid_fft.hh:844-850)."""
    X[0, 0, 0] = 0.0
    return X


def inverse_laplacian(g: GridSpec, X):
    """Grid_FFT::apply_InverseLaplacian: x <- -x/|k|^2, then DC = 0
    (include/grid_fft.hh:440-449).  The reference divides by zero at DC and then overwrites."""
    k2 = g.k2()
    with np.errstate(divide="ignore", invalid="ignore"):
        out = -X / k2
    out[0, 0, 0] = 0.0
    return out


def shift_field_k(g: GridSpec, X, s):
    """k-space part of Grid_FFT::shift_field (include/grid_fft.hh:832-842):
    x <- x * exp(i (s.x kx dx + s.y ky dy + s.z kz dz)) for every stored mode (get_k, so the
    Nyquist index carries +N/2)."""
    kx, ky, kz = g.kvec()
    ph = (s[0] * kx + s[1] * ky + s[2] * kz) * g.dx
    return X * np.exp(1j * ph)


def shift_field_r(g: GridSpec, x, s):
    """Grid_FFT::shift_field on a real-space field: fwd, phase, bwd."""
    return g.bwd(shift_field_k(g, g.fwd(x), s))


# --------------------------------------------------------------------------------------
# Fourier-space resampling == Grid_FFT::FourierInterpolateCopyTo, MPI branch
# (src/grid_fft.cc:382-394, 432-463): only modes strictly inside both Nyquist limits are
# copied, the target is zeroed first.
# --------------------------------------------------------------------------------------
def fourier_interpolate_copy(src: GridSpec, dst: GridSpec, X):
    nf, nt = src.n, dst.n
    out = dst.zeros_k()
    if nf == nt:
        out[...] = X
        return out
    fny_left = min(nf // 2, nt // 2)
    fny_right = max(nf - nt // 2, nf // 2)
    isend = np.array([i for i in range(nf) if (i < fny_left or i > fny_right)], dtype=np.int64)
    irecv = np.where(isend < fny_left, isend, isend + nt - nf)
    kz = fny_left  # k < fny2_left
    out[np.ix_(irecv, irecv, np.arange(kz))] = X[np.ix_(isend, isend, np.arange(kz))]
    return out


# --------------------------------------------------------------------------------------
# Orszag 3/2 convolver  (include/convolution.hh:429-669, MPI-build semantics)
# --------------------------------------------------------------------------------------
class OrszagConvolver:
    def __init__(self, g: GridSpec):
        self.g = g
        self.gp = GridSpec(3 * g.n // 2, g.L, g.fft)  # include/convolution.hh:464
        self.nfft = 0

    def pad_insert(self, A):
        """pad_insert, MPI branch (include/convolution.hh:594-608): multiply by 1.5^1.5 and
        FourierInterpolateCopyTo the padded grid (Nyquist planes dropped)."""
        rfac = 1.5 ** 1.5
        return fourier_interpolate_copy(self.g, self.gp, A * rfac)

    def unpad(self, P):
        """unpad, MPI branch (include/convolution.hh:657-665): FourierInterpolateCopyTo(fbuf),
        then divide by rfac = sqrt(Np^3)/sqrt(N^3)."""
        rfac = np.sqrt(float(self.gp.n) ** 3) / np.sqrt(float(self.g.n) ** 3)
        return fourier_interpolate_copy(self.gp, self.g, P) / rfac

    def convolve2(self, A, B):
        """OrszagConvolver::convolve2 (include/convolution.hh:512-535)."""
        a = self.gp.bwd(self.pad_insert(A))
        b = self.gp.bwd(self.pad_insert(B))
        self.nfft += 3
        return self.unpad(self.gp.fwd(a * b))

    def convolve3(self, A, B, C):
        """convolve3 = convolve2(convolve2(A,B) truncated to the N grid, C)
        (include/convolution.hh:542-549)."""
        return self.convolve2(self.convolve2(A, B), C)

    # ---- named products of BaseConvolver (include/convolution.hh:75-292)
    def hess(self, F, a, b):
        g = self.g
        return g.grad(a) * g.grad(b) * F  # :147-151

    def convolve_Hessians(self, F, d2l, G, d2r):
        return self.convolve2(self.hess(F, *d2l), self.hess(G, *d2r))  # :137-160

    def convolve_Hessians3(self, F, d2l, G, d2m, H, d2r):
        return self.convolve3(self.hess(F, *d2l), self.hess(G, *d2m), self.hess(H, *d2r))  # :172-203

    def convolve_SumOfHessians(self, F, d2l, G, d2r1, d2r2):
        g = self.g
        fac = g.grad(d2r1[0]) * g.grad(d2r1[1]) + g.grad(d2r2[0]) * g.grad(d2r2[1])
        return self.convolve2(self.hess(F, *d2l), fac * G)  # :214-239

    def convolve_DifferenceOfHessians(self, F, d2l, G, d2r1, d2r2):
        g = self.g
        fac = g.grad(d2r1[0]) * g.grad(d2r1[1]) - g.grad(d2r2[0]) * g.grad(d2r2[1])
        return self.convolve2(self.hess(F, *d2l), fac * G)  # :250-274

    def convolve_Gradients(self, F, d1l, G, d1r):
        g = self.g
        return self.convolve2(g.grad(d1l) * F, g.grad(d1r) * G)  # :75-96

    def convolve_Gradient_and_Hessian(self, F, d1l, G, d2r):
        """uses get_k (no Nyquist zeroing) and sign -i k  (include/convolution.hh:106-127)."""
        kv = self.g.kvec()
        return self.convolve2((-1j * kv[d1l]) * F, (-kv[d2r[0]] * kv[d2r[1]]) * G)

    def multiply_field(self, F, G):
        return self.convolve2(F, G)  # :282-292


class NaiveConvolver(OrszagConvolver):
    """NaiveConvolver (include/convolution.hh:296-426): same API, no padding."""

    def __init__(self, g: GridSpec):
        self.g = g
        self.gp = g
        self.nfft = 0

    def convolve2(self, A, B):
        self.nfft += 3
        return self.g.fwd(self.g.bwd(A) * self.g.bwd(B))

    def convolve3(self, A, B, C):
        self.nfft += 4
        return self.g.fwd(self.g.bwd(A) * self.g.bwd(B) * self.g.bwd(C))


# --------------------------------------------------------------------------------------
# Noise conditioning and first-order potential  (src/ic_generator.cc:341-358, 461-468)
# --------------------------------------------------------------------------------------
def condition_noise(g: GridSpec, W, fix=False, invert=False, remove_corners=False):
    volfac = (g.L / g.n / 2.0 / np.pi) ** 1.5  # :210
    W = W.copy()
    if fix:
        a = np.abs(W)
        W = np.where(a != 0.0, W / np.where(a != 0.0, a, 1.0), W)  # :342-344
    W = (-1.0 if invert else 1.0) * W / volfac  # :345
    if remove_corners:
        kny = g.kfac * g.n / 2
        W = np.where(np.sqrt(g.k2()) > kny, 0.0, W)  # :352-358
    return W


def phi1_from_noise(g: GridSpec, W, amplitude):
    """phi = -W*A(|k|)/|k|^2, DC=0  (src/ic_generator.cc:461-468).
    `amplitude` is a vectorised callable A(|k|) (cosmology_calculator.hh:405-408)."""
    k2 = g.k2()
    kmod = np.sqrt(k2)
    with np.errstate(divide="ignore", invalid="ignore"):
        amp = amplitude(kmod)
        phi = -(W * amp) / (kmod * kmod)
    phi[0, 0, 0] = 0.0
    return phi


def radial_table(g: GridSpec, amplitude):
    """Tabulate A(|k|) on the distinct integer |n|^2 values of the grid:
    table[m] = A(kfac*sqrt(m)), m = 0..3(N/2)^2.  Host-side helper shared by tests."""
    m = np.arange(3 * g.nh * g.nh + 1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.asarray(amplitude(g.kfac * np.sqrt(m)), dtype=np.float64)
    t = np.where(np.isfinite(t), t, 0.0)
    return t


# --------------------------------------------------------------------------------------
# LPT source terms  (src/ic_generator.cc:475-574)
# --------------------------------------------------------------------------------------
def lpt_potentials(g: GridSpec, phi, order, Dplus0=1.0, convolver=None, tides=None):
    """Returns dict(phi, phi2, phi3, A3=[..]) *after* growth-factor scaling (:561-574).
    phi is the unscaled first-order potential in k-space."""
    conv = convolver or OrszagConvolver(g)
    g1 = -Dplus0
    g2 = -3.0 / 7.0 * Dplus0 * Dplus0 if order > 1 else 0.0
    g3 = 1.0 / 3.0 * Dplus0 ** 3 if order > 2 else 0.0
    g3c = 1.0 / 7.0 * Dplus0 ** 3 if order > 2 else 0.0
    out = {"phi": None, "phi2": None, "phi3": None, "A3": None}
    phi2 = phi3 = None
    A3 = None
    if order > 1:
        # :482-486
        phi2 = conv.convolve_SumOfHessians(phi, (0, 0), phi, (1, 1), (2, 2))
        phi2 = phi2 + conv.convolve_Hessians(phi, (1, 1), phi, (2, 2))
        phi2 = phi2 - conv.convolve_Hessians(phi, (0, 1), phi, (0, 1))
        phi2 = phi2 - conv.convolve_Hessians(phi, (0, 2), phi, (0, 2))
        phi2 = phi2 - conv.convolve_Hessians(phi, (1, 2), phi, (1, 2))
        if tides is not None:  # :488-497
            lam = np.asarray(tides, dtype=np.float64)
            kx, ky, kz = g.kvec()
            fac_aniso = kx * kx * lam[0] + ky * ky * lam[1] + kz * kz * lam[2]
            phi2 = phi2 - (lam.sum() * g.k2() + 4.0 / 3.0 * fac_aniso) * phi
        phi2 = inverse_laplacian(g, phi2)  # :499
    if order > 2:
        # 3a  :522-526
        phi3 = conv.convolve_Hessians3(phi, (0, 0), phi, (1, 1), phi, (2, 2))
        phi3 = phi3 + 2.0 * conv.convolve_Hessians3(phi, (0, 1), phi, (0, 2), phi, (1, 2))
        phi3 = phi3 - conv.convolve_Hessians3(phi, (1, 2), phi, (1, 2), phi, (0, 0))
        phi3 = phi3 - conv.convolve_Hessians3(phi, (0, 2), phi, (0, 2), phi, (1, 1))
        phi3 = phi3 - conv.convolve_Hessians3(phi, (0, 1), phi, (0, 1), phi, (2, 2))
        # 3b  :533-538  (un-scaled phi2)
        phi3 = phi3 + (-5.0 / 7.0) * conv.convolve_SumOfHessians(phi, (0, 0), phi2, (1, 1), (2, 2))
        phi3 = phi3 + (-5.0 / 7.0) * conv.convolve_SumOfHessians(phi, (1, 1), phi2, (2, 2), (0, 0))
        phi3 = phi3 + (-5.0 / 7.0) * conv.convolve_SumOfHessians(phi, (2, 2), phi2, (0, 0), (1, 1))
        phi3 = phi3 + (10.0 / 7.0) * conv.convolve_Hessians(phi, (0, 1), phi2, (0, 1))
        phi3 = phi3 + (10.0 / 7.0) * conv.convolve_Hessians(phi, (0, 2), phi2, (0, 2))
        phi3 = phi3 + (10.0 / 7.0) * conv.convolve_Hessians(phi, (1, 2), phi2, (1, 2))
        phi3 = inverse_laplacian(g, phi3)  # :539
        # transverse term  :545-556
        A3 = []
        for d in range(3):
            p, q = (d + 1) % 3, (d + 2) % 3
            a = conv.convolve_Hessians(phi2, (d, p), phi, (d, q))
            a = a - conv.convolve_Hessians(phi2, (d, q), phi, (d, p))
            a = a + conv.convolve_DifferenceOfHessians(phi, (p, q), phi2, (p, p), (q, q))
            a = a - conv.convolve_DifferenceOfHessians(phi2, (p, q), phi, (p, p), (q, q))
            A3.append(inverse_laplacian(g, a))
    out["phi"] = phi * g1
    if order > 1:
        out["phi2"] = phi2 * g2
    if order > 2:
        out["phi3"] = phi3 * g3
        out["A3"] = [a * g3c for a in A3]
    return out


# --------------------------------------------------------------------------------------
# Displacement / velocity assembly  (src/ic_generator.cc:832-871, 888-942)
# --------------------------------------------------------------------------------------
class PLTGradient:
    """particle::lattice_gradient::gradient / vfac_corr (include/particle_plt.hh:551-585, build option ENABLE_PLT), given
    the four helper-mesh tables the class holds after its eigen-decomposition (Re D_xx_ = radial, D_yy_ = theta, D_zz_ = phi
    eigenvalue, D_xy_ = growth-rate correction; [ng][ng][ng/2+1]).  The decomposition itself (Ewald sums, GSL eigen-solver)
    is host work outside the path and is not restated."""

    def __init__(self, ng, ngmapto, boxlen, D_r, D_theta, D_phi, mu):
        self.ng, self.ngmapto, self.boxlen = int(ng), int(ngmapto), float(boxlen)
        self.mapratio = float(ng) / float(ngmapto)                       # :523
        self.tab = [np.asarray(t, dtype=np.float64) for t in (D_r, D_theta, D_phi, mu)]

    def _coords(self, g: GridSpec):
        n = g.n
        idx = np.arange(n, dtype=np.float64) * self.mapratio
        return idx[:, None, None], idx[None, :, None], idx[None, None, : g.nzc]

    def cic_kspace(self, T, x, y, z):
        """Grid_FFT::get_cic_kspace (include/grid_fft.hh:347-373), real part; x, y, z broadcastable index arrays"""
        ng, nzc = self.ng, self.ng // 2 + 1
        x, y, z = np.broadcast_arrays(x, y, z)
        ix, iy, iz = np.floor(x).astype(int), np.floor(y).astype(int), np.floor(z).astype(int)
        dx, dy, dz = x - ix, y - iy, z - iz
        tx, ty, tz = 1.0 - dx, 1.0 - dy, 1.0 - dz
        ix1, iy1, iz1 = (ix + 1) % ng, (iy + 1) % ng, np.minimum(iz + 1, nzc - 1)
        v = T[ix, iy, iz] * tx * ty * tz
        v = v + T[ix, iy, iz1] * tx * ty * dz
        v = v + T[ix, iy1, iz] * tx * dy * tz
        v = v + T[ix, iy1, iz1] * tx * dy * dz
        v = v + T[ix1, iy, iz] * dx * ty * tz
        v = v + T[ix1, iy, iz1] * dx * ty * dz
        v = v + T[ix1, iy1, iz] * dx * dy * tz
        v = v + T[ix1, iy1, iz1] * dx * dy * dz
        return v

    def all(self, g: GridSpec):
        """-> ([Im gradient(0), Im gradient(1), Im gradient(2)], vfac_corr) on the mode lattice [n][n][nzc]"""
        x, y, z = self._coords(g)
        ng = float(self.ng)
        h = float(self.ng // 2)
        tau = 2.0 * np.pi
        # D_xx_.get_k<real_t>(ix, iy, iz) on the unit-length helper mesh (include/grid_fft.hh:299-313)
        kx, ky, kz = np.broadcast_arrays((x - (x > h) * ng) * tau, (y - (y > h) * ng) * tau, (z - (z > h) * ng) * tau)
        kr = np.sqrt(kx * kx + ky * ky + kz * kz)
        kmod = kr / self.mapratio / self.boxlen
        Dr, Dt, Dp = (self.cic_kspace(self.tab[t], x, y, z) for t in range(3))
        with np.errstate(divide="ignore", invalid="ignore"):
            kphi = np.where(kr > 0.0, np.arctan2(ky, kx), 0.0)
            ktheta = np.where(kr > 0.0, np.arccos(np.where(kr > 0.0, kz / np.where(kr > 0.0, kr, 1.0), 0.0)), 0.0)
        st, ct, sp, cp = np.sin(ktheta), np.cos(ktheta), np.sin(kphi), np.cos(kphi)
        G = [kmod * (Dr * st * cp + Dt * ct * cp - Dp * sp), kmod * (Dr * st * sp + Dt * ct * sp + Dp * cp),
             kmod * (Dr * ct - Dt * st)]
        return G, self.cic_kspace(self.tab[3], x, y, z)


def cic_compensation_kernel(g: GridSpec):
    """grid_interpolate<1,.>::compensation_kernel (include/grid_interpolate.hh:195-208) on the mode lattice:
    1 / (sinc(pi k_x / 2 k_ny) sinc(..) sinc(..))^(1+1), k from get_k (the Nyquist index carries +N/2)"""
    kny = g.kfac * (g.n // 2)

    def sinc(x):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(np.abs(x) > 1e-10, np.sin(x) / np.where(x == 0.0, 1.0, x), 1.0)
    kx, ky, kz = g.kvec()
    f = sinc(0.5 * np.pi * kx / kny) * sinc(0.5 * np.pi * ky / kny) * sinc(0.5 * np.pi * kz / kny)
    return 1.0 / (f * f)


def _gradients(g: GridSpec, lg):
    """(list of the three gradient factors as complex arrays, vfac_corr) for op::fourier_gradient (lg None) or PLT"""
    if lg is None:
        return [g.grad(d) for d in range(3)], 1.0
    G, vc = lg.all(g)
    return [1j * a for a in G], vc


def displacement_k(g: GridSpec, pots, order, d, scale=1.0, lg=None, glass=False):
    """k-space Psi_d * scale; scale = lunit/boxlen (:866).  lg: a PLTGradient in place of op::fourier_gradient (:377-381);
    glass: CIC compensation kernel (:861-863)."""
    p, q = (d + 1) % 3, (d + 2) % 3
    gr, _ = _gradients(g, lg)
    tot = pots["phi"].copy()
    if order > 1:
        tot = tot + pots["phi2"]
    if order > 2:
        tot = tot + pots["phi3"]
    tmp = gr[d] * tot
    if order > 2:
        tmp = tmp + gr[p] * pots["A3"][q] - gr[q] * pots["A3"][p]
    if glass:
        tmp = tmp * cic_compensation_kernel(g)
    tmp = tmp * scale
    return zero_dc(tmp)


def velocity_k(g: GridSpec, pots, order, d, vfac, scale=1.0, vbc=None, tides=None, Dplus0=None, lg=None, glass=False):
    """k-space v_d * scale; scale = vunit/boxlen (:937).
    vbc = (C_species, theta_bc_amplitude(|k|) callable, W) adds the relative-velocity term (:918-921).
    tides = (lambda_x, lambda_y, lambda_z) with Dplus0: external tides (only honoured at LPTorder > 1, :190-193) multiply
    the whole mode by the anisotropic velocity growth factor alpha_d^2, alpha_d = 1 - D+ lambda_d (:238-242, :931-934)."""
    p, q = (d + 1) % 3, (d + 2) % 3
    gr, vcorr = _gradients(g, lg)
    tot = vfac * pots["phi"]
    if order > 1:
        tot = tot + (2 * vfac) * pots["phi2"]
    if order > 2:
        tot = tot + (3 * vfac) * pots["phi3"]
    tmp = gr[d] * tot
    if order > 2:
        tmp = tmp + (3 * vfac) * (gr[p] * pots["A3"][q] - gr[q] * pots["A3"][p])
    if vbc is not None:
        C, amp, W = vbc
        k2 = g.k2()
        kn = np.sqrt(k2)
        with np.errstate(divide="ignore", invalid="ignore"):
            term = vfac * C * amp(kn) * W * gr[d] / (kn * kn)
        term[0, 0, 0] = 0.0
        tmp = tmp - term
    if glass:
        tmp = tmp * cic_compensation_kernel(g)      # :923-926
    tmp = tmp * vcorr                               # :929, lg.vfac_corr (1 for the Fourier gradient)
    if tides is not None and order > 1:
        tmp = tmp * (1.0 - Dplus0 * tides[d]) ** 2
    tmp = tmp * scale
    return zero_dc(tmp)


def density_k(g: GridSpec, pots):
    """first-order SPT density delta = k^2 phi  (src/ic_generator.cc:962-973)."""
    return g.k2() * pots["phi"]


def ppt_eulerian(g: GridSpec, Wc, pots, order, Dplus0, vfac, amp_delta_bc, C_species, velocity_unit=1.0, nranks=1):
    """Semiclassical (PPT) Eulerian density and velocity fields, src/ic_generator.cc:685-803.
    `pots` are the growth-scaled potentials of lpt_potentials (phi already contains -D+, :561), `Wc` the conditioned
    white noise in k-space.  Returns ({"density","vx","vy","vz"}, hbar).
    nranks > 1: hbar as the reference's MPI build computes it - the loop of :715-729 runs over the LOCAL planes of every rank
    and stops one short of the last, then MPI_Allreduce(MAX) (:730-733)."""
    n = g.n
    kx, ky, kz = g.kvec()
    kmod = np.sqrt(kx * kx + ky * ky + kz * kz)
    with np.errstate(divide="ignore", invalid="ignore"):
        A = np.asarray(amp_delta_bc(kmod), dtype=np.float64)
    A = np.where(np.isfinite(A), A, 0.0)
    rho = zero_dc(Wc * A)                                             # :697-700
    rho = np.sqrt(1.0 + C_species * g.bwd(rho))                        # :701-705
    phi = g.bwd(pots["phi"])                                           # :710
    # :713-731 - note the x loop stops at size(0)-1: the last x-plane is not visited
    maxd = -1.0
    l0 = n // nranks
    for r in range(nranks):
        blk = phi[r * l0:(r + 1) * l0]
        c = blk[:-1]
        maxd = max(maxd, np.abs(blk[1:] - c).max(), np.abs(np.roll(blk, -1, 1)[:-1] - c).max(),
                   np.abs(np.roll(blk, -1, 2)[:-1] - c).max())
    hbar = maxd / np.pi / Dplus0 * 1.01                                # :733-734
    if order == 1:
        psi = rho * np.exp(1j * ((1.0 / hbar) * (phi / Dplus0)))      # :738-740
        phi2 = None
    else:
        phi2 = g.bwd(pots["phi2"])
        psi = rho * np.exp(1j * (((1.0 / hbar) * (phi + phi2)) / Dplus0))  # :744-747
    # full c2c transforms with the symmetric norm; k from get_k on all three axes (Nyquist = +N/2)
    k1 = g.k1d
    k2 = (k1[:, None, None] ** 2 + k1[None, :, None] ** 2) + k1[None, None, :] ** 2
    norm = g.norm
    cf = lambda x: np.fft.fftn(x) * norm
    cb = lambda X: np.fft.ifftn(X) * (float(n) ** 3 * norm)
    psi = cb(cf(psi) * np.exp(1j * (-((0.5 * hbar) * k2) * Dplus0)))   # :755-761
    if order >= 2:
        psi = psi * np.exp(1j * (((1.0 / hbar) * phi2) / Dplus0))     # :763-767
    dens = psi.real ** 2 + psi.imag ** 2 - 1.0                         # :772-775
    vunit = Dplus0 * vfac / g.L * velocity_unit                        # :779
    out = {"density": dens}
    kk = (k1[:, None, None], k1[None, :, None], k1[None, None, :])
    for d, nm in enumerate(("vx", "vy", "vz")):
        gp = cb(cf(psi) * (1j * kk[d]))                                # :782-793
        out[nm] = vunit * np.real((np.conj(psi) * gp - psi * np.conj(gp)) / (2j / hbar) / (1.0 + dens))  # :796-799
    return out, hbar


def fields(g: GridSpec, pots, order, vfac, lunit=1.0, vunit=1.0):
    """all six real-space component fields + density; defaults = generic output plugin
    (lunit=vunit=1, output_generic.cc:74-76); fields are in box units * unit (:866,:937)."""
    out = {}
    for d, nm in enumerate("xyz"):
        out["d" + nm] = g.bwd(displacement_k(g, pots, order, d, lunit / g.L))
        out["v" + nm] = g.bwd(velocity_k(g, pots, order, d, vfac, vunit / g.L))
    out["delta"] = g.bwd(density_k(g, pots))
    return out


# --------------------------------------------------------------------------------------
# Particle lattice  (include/particle_generator.hh)
# --------------------------------------------------------------------------------------
LATTICE_SHIFTS = {  # :48-56
    0: [(0.0, 0.0, 0.0)],
    1: [(0.0, 0.0, 0.0), (0.5, 0.5, 0.5)],
    2: [(0.0, 0.0, 0.0), (0.0, 0.5, 0.5), (0.5, 0.0, 0.5), (0.5, 0.5, 0.0)],
    3: [(0.0, 0.0, 0.0), (0.0, 0.0, 0.5), (0.0, 0.5, 0.0), (0.0, 0.5, 0.5),
        (0.5, 0.0, 0.0), (0.5, 0.0, 0.5), (0.5, 0.5, 0.0), (0.5, 0.5, 0.5)],
}
SECOND_LATTICE_SHIFT = {0: (0.5, 0.5, 0.5), 1: (0.5, 0.5, 0.0), 2: (0.5, 0.5, 0.5), 3: (0.25, 0.25, 0.25)}  # :58-64
LATTICE_TYPES = {"sc": 0, "bcc": 1, "fcc": 2, "rsc": 3}


def particle_ids(n, lattice_type, id_offset=0, ids64=False):
    """lattice_generator ctor (include/particle_generator.hh:129-170):
    id = IDoffset*ov*N^3 + ov*cell + iload stored at ipcount + iload*N^3."""
    ov = 1 << max(0, lattice_type)
    ncell = n ** 3
    cell = np.arange(ncell, dtype=np.uint64)
    off = np.uint64(id_offset * ov * ncell)
    ids = np.empty(ov * ncell, dtype=np.uint64)
    for iload in range(ov):
        ids[iload * ncell:(iload + 1) * ncell] = off + np.uint64(ov) * cell + np.uint64(iload)
    return ids if ids64 else ids.astype(np.uint32)


def _shift_sequence(lattice_type, lattice_index):
    """absolute shift of each sub-lattice and the *relative* shift applied to the field
    before reading it out (set_positions :433-445)."""
    shifts = LATTICE_SHIFTS[lattice_type]
    sec = SECOND_LATTICE_SHIFT[lattice_type] if lattice_index == 1 else (0.0, 0.0, 0.0)
    seq = []
    for ish, s in enumerate(shifts):
        rel = []
        if ish == 0 and lattice_index == 1:
            rel.append(SECOND_LATTICE_SHIFT[lattice_type])
        if ish > 0:
            rel.append(tuple(a - b for a, b in zip(s, shifts[ish - 1])))
        absolute = tuple(a + b for a, b in zip(s, sec))
        seq.append((absolute, rel))
    return seq


def set_positions(g: GridSpec, field_r, d, lattice_type, lattice_index, lunit, real64=True):
    """lattice_generator::set_positions for Bravais lattices (:422-467).  Returns the d-th
    coordinate of all particles (sub-lattices concatenated) in the storage dtype."""
    n = g.n
    f = field_r
    outs = []
    idx = np.arange(n, dtype=np.float64)
    for absolute, rel in _shift_sequence(lattice_type, lattice_index):
        for s in rel:
            f = shift_field_r(g, f, s)
        shape = [1, 1, 1]
        shape[d] = n
        q = ((idx + absolute[d]) / float(n)).reshape(shape)  # get_unit_r_shifted, grid_fft.hh:243-252
        pos = q * lunit + f
        outs.append(pos.reshape(-1))
    res = np.concatenate(outs)
    return res if real64 else res.astype(np.float32)


def set_velocities(g: GridSpec, field_r, lattice_type, lattice_index, real64=True):
    """lattice_generator::set_velocities (:520-562)."""
    f = field_r
    outs = []
    for _, rel in _shift_sequence(lattice_type, lattice_index):
        for s in rel:
            f = shift_field_r(g, f, s)
        outs.append(f.reshape(-1).copy())
    res = np.concatenate(outs)
    return res if real64 else res.astype(np.float32)


def set_masses(g: GridSpec, field_r, lattice_type, lattice_index, munit, real64=True):
    """lattice_generator::set_masses, Bravais branch (:292-350)."""
    ov = 1 << max(0, lattice_type)
    pmean = munit / float(g.n ** 3 * ov)
    f = field_r
    outs = []
    for _, rel in _shift_sequence(lattice_type, lattice_index):
        for s in rel:
            f = shift_field_r(g, f, s)
        outs.append((pmean * f).reshape(-1))
    res = np.concatenate(outs)
    return res if real64 else res.astype(np.float32)


# --------------------------------------------------------------------------------------
# Masked sc lattice and glass loads (include/particle_generator.hh:171-290, 351-398, 479-515, 557-590)
# --------------------------------------------------------------------------------------
def particle_type_mask(do_baryons, mask_type=3):
    """the 2x2x2 species mask from [setup] DoBaryons / ParticleMaskType (:184-205); index 4a+2b+c"""
    if not do_baryons:
        return [0] * 8
    m = [-1] * 8
    m[0] = 0
    if mask_type & 1:
        m[1] = m[2] = m[4] = 0
    if mask_type & 2:
        m[3] = m[5] = m[6] = 0
    m[7] = 1
    return m


def _mask_select(n, mask, lattice_index):
    """boolean [n][n][n]: get_mask_value(cell) == lattice_index (:84-87)"""
    par = np.arange(n) % 2
    sig = (par[:, None, None] * 2 + par[None, :, None]) * 2 + par[None, None, :]
    return np.asarray(mask)[sig] == lattice_index


def masked_ids(n, mask, lattice_index, id_offset=0, ids64=False):
    """:236-243: IDoffset*N^3 + get_cell_idx_1d of the species' cells in (i,j,k) order"""
    sel = _mask_select(n, mask, lattice_index).reshape(-1)
    ids = np.uint64(id_offset * n ** 3) + np.arange(n ** 3, dtype=np.uint64)[sel]
    return ids if ids64 else ids.astype(np.uint32)


def masked_positions(g: GridSpec, field_r, d, mask, lattice_index, lunit, real64=True):
    """:479-497: get_unit_r(i,j,k)[d] * lunit + field"""
    n = g.n
    shape = [1, 1, 1]
    shape[d] = n
    q = (np.arange(n, dtype=np.float64) / float(n)).reshape(shape)
    res = (q * lunit + field_r)[_mask_select(n, mask, lattice_index)]
    return res if real64 else res.astype(np.float32)


def masked_velocities(g: GridSpec, field_r, mask, lattice_index, real64=True):
    """:557-571"""
    res = field_r[_mask_select(g.n, mask, lattice_index)]
    return res if real64 else res.astype(np.float32)


def masked_masses(g: GridSpec, field_r, mask, lattice_index, munit, real64=True):
    """:351-382 with get_mean_mask_value (:89-112): pmean * (field - (mean over the species' cells of the 2x2x2 block - mean
    over all eight cells)), pmean = munit / number of particles of the species"""
    n = g.n
    sel = _mask_select(n, mask, lattice_index)
    blocks = field_r.reshape(n // 2, 2, n // 2, 2, n // 2, 2)
    sel8 = (np.asarray(mask) == lattice_index).reshape(2, 2, 2)
    full = blocks.sum(axis=(1, 3, 5)) / 8.0
    masked = (blocks * sel8[None, :, None, :, None, :]).sum(axis=(1, 3, 5)) / float(sel8.sum())
    mm = np.repeat(np.repeat(np.repeat(masked - full, 2, 0), 2, 1), 2, 2)
    pmean = munit / float(sel.sum())
    res = (pmean * (field_r - mm))[sel]
    return res if real64 else res.astype(np.float32)


def cic_at(field_r, pos):
    """grid_interpolate<1,.>::get_cic_at (include/grid_interpolate.hh:103-148, non-distributed branch); pos [np][3] in grid
    units"""
    n0, n1, n2 = field_r.shape
    ix, iy, iz = (pos[:, c].astype(np.int64) for c in range(3))
    dx, dy, dz = pos[:, 0] - ix, pos[:, 1] - iy, pos[:, 2] - iz
    tx, ty, tz = 1.0 - dx, 1.0 - dy, 1.0 - dz
    ix1, iy1, iz1 = (ix + 1) % n0, (iy + 1) % n1, (iz + 1) % n2
    v = field_r[ix, iy, iz] * tx * ty * tz
    v = v + field_r[ix, iy, iz1] * tx * ty * dz
    v = v + field_r[ix, iy1, iz] * tx * dy * tz
    v = v + field_r[ix, iy1, iz1] * tx * dy * dz
    v = v + field_r[ix1, iy, iz] * dx * ty * tz
    v = v + field_r[ix1, iy, iz1] * dx * ty * dz
    v = v + field_r[ix1, iy1, iz] * dx * dy * tz
    v = v + field_r[ix1, iy1, iz1] * dx * dy * dz
    return v


def glass_positions(g: GridSpec, field_r, d, posr, lunit, real64=True):
    """:498-515: pos[d] / n * lunit + CIC(field, pos)"""
    res = posr[:, d] / float(g.n) * lunit + cic_at(field_r, posr)
    return res if real64 else res.astype(np.float32)


def glass_velocities(g: GridSpec, field_r, posr, real64=True):
    """:573-590"""
    res = cic_at(field_r, posr)
    return res if real64 else res.astype(np.float32)


# --------------------------------------------------------------------------------------
# Whole pipeline (one species) - used by tests, smoke() and the CPU baseline
# --------------------------------------------------------------------------------------
def run_pipeline(n, boxlen, W, amplitude, order, Dplus0, vfac, lattice="sc", lunit=1.0, vunit=1.0,
                 real64=True, ids64=False, fft=None, particles=True):
    """noise (k-space, already conditioned) -> potentials -> Psi, v (-> particles)."""
    g = GridSpec(n, boxlen, fft)
    phi = phi1_from_noise(g, W, amplitude)
    pots = lpt_potentials(g, phi, order, Dplus0)
    res = {"pots": pots}
    lt = LATTICE_TYPES[lattice]
    pos, vel = [], []
    flds = {}
    for d, nm in enumerate("xyz"):
        dr = g.bwd(displacement_k(g, pots, order, d, lunit / boxlen))
        flds["d" + nm] = dr
        if particles:
            pos.append(set_positions(g, dr, d, lt, 0, lunit, real64))
    for d, nm in enumerate("xyz"):
        vr = g.bwd(velocity_k(g, pots, order, d, vfac, vunit / boxlen))
        flds["v" + nm] = vr
        if particles:
            vel.append(set_velocities(g, vr, lt, 0, real64))
    res["fields"] = flds
    if particles:
        res["pos"] = np.stack(pos, axis=1)
        res["vel"] = np.stack(vel, axis=1)
        res["ids"] = particle_ids(n, lt, 0, ids64)
    return res


# --------------------------------------------------------------------------------------
# Synthetic inputs shared by tests / bench (NGenIC-like Hermitian white noise)
# --------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------
# Device noise fill in the NGenIC layout (src/plugins/random_ngenic.cc:84-154)
# --------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1, _PHILOX_W0, _PHILOX_W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011, Random123): counter words c0..c3 (uint arrays), key (k0, k1) -> four uint32 words.
    Known answers (Random123 kat_vectors): zero counter and key -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0 & 0xffffffff), np.uint64(k1 & 0xffffffff)
    m32, s32 = np.uint64(0xffffffff), np.uint64(32)
    for _ in range(10):
        p0, p1 = np.uint64(_PHILOX_M0) * c0, np.uint64(_PHILOX_M1) * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ k0) & m32, p1 & m32, ((p0 >> s32) ^ c3 ^ k1) & m32, p0 & m32
        k0, k1 = (k0 + np.uint64(_PHILOX_W0)) & m32, (k1 + np.uint64(_PHILOX_W1)) & m32
    return c0, c1, c2, c3


def ngenic_layout_noise(n, seed):
    """White noise in k-space with the layout rules of RNG_ngenic::Fill_Grid (src/plugins/random_ngenic.cc:84-154): one
    independent stream per (k_x, k_y) column, per k_z a phase 2 pi u1 and an amplitude sqrt(-ln u2), unit variance per complex
    mode; DC and every mode with an index N/2 stay zero (:117-118); the k_z = 0 plane is Hermitian: columns with 0 < k_x index
    < N/2, and k_x = 0 with k_y index < N/2, are drawn, their mirrors (N-k_x, N-k_y) get the conjugate (:125-148).
    The generator is NOT the plugin's gsl ranlxd1 (GSL is not available; bit-exact NGenIC phases stay with the host plugin):
    Philox4x32-10 keyed with the 64-bit seed, counter = (column word a, column word b, k_z, 0).  The column words are taken
    from the position of the column in the plugin's ring-ordered seed table (:47-74) - index i < N/2 as it is, i >= N/2 as
    N-1-i with the top bit set - so that, like there, a larger grid draws the same low-k columns."""
    h, nzc = n // 2, n // 2 + 1
    idx = np.arange(n, dtype=np.uint64)
    word = np.where(idx < h, idx, (np.uint64(n - 1) - idx) | np.uint64(0x80000000))

    def draw(a, b, k):
        x0, x1, x2, x3 = philox4x32_10(word[a], word[b], k, np.zeros_like(k), seed & 0xffffffff, (seed >> 32) & 0xffffffff)
        u1 = (((x0 << np.uint64(20)) | (x1 >> np.uint64(12))).astype(np.float64) + 0.5) * 2.0 ** -52
        u2 = (((x2 << np.uint64(20)) | (x3 >> np.uint64(12))).astype(np.float64) + 0.5) * 2.0 ** -52
        amp = np.sqrt(-np.log(u2))
        return amp * (np.cos(2.0 * np.pi * u1) + 1j * np.sin(2.0 * np.pi * u1))
    A, B, K = np.meshgrid(np.arange(n), np.arange(n), np.arange(nzc), indexing="ij")
    W = draw(A, B, K.astype(np.uint64))
    # k_z = 0 plane: mirrors take the conjugate of their primary column
    a, b = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    primary = ((a == 0) & (b < h)) | ((a > 0) & (a < h))
    ma, mb = (n - a) % n, (n - b) % n
    z0 = draw(np.where(primary, a, ma), np.where(primary, b, mb), np.zeros((n, n), dtype=np.uint64))
    W[:, :, 0] = np.where(primary, z0, np.conj(z0))
    W[h, :, :] = 0.0
    W[:, h, :] = 0.0
    W[:, :, h] = 0.0
    W[0, 0, 0] = 0.0
    return W


class Ranlxd1:
    """Luescher's RANLUX in double precision at luxury level 1 (p = 202) - what GSL calls gsl_rng_ranlxd1, the generator of the
    reference's NGenIC plugin (src/plugins/random_ngenic.cc:44-45).  Published algorithm (M. Luescher, Comput. Phys. Commun. 79
    (1994) 100; F. James, ibid. 97 (1996) 357): twelve 48-bit numbers in a ring, x_n = x_{n-5} - x_{n-12} - c (mod 1), 202 updates
    per twelve delivered numbers, seed bits from a 31-bit shift register.  PINNED against the reference: the NGenIC noise it gives
    for the seed of the reference's regression configurations has the phases of the reference's golden density field
    (tests/test_oracle_golden.py::test_ngenic_noise_is_the_reference_s)."""
    ONE_BIT = 2.0 ** -48

    def __init__(self, seed):
        s = int(seed) or 1
        v = s & 0x7FFFFFFF
        bit = [(v >> k) & 1 for k in range(31)]
        ib, jb = 0, 18
        self.x = []
        for _ in range(12):
            acc = 0.0
            for _ in range(48):
                acc += acc + float((bit[ib] + 1) % 2)
                bit[ib] = (bit[ib] + bit[jb]) % 2
                ib, jb = (ib + 1) % 31, (jb + 1) % 31
            self.x.append(acc * self.ONE_BIT)
        self.carry, self.ir, self.jr, self.ir_old = 0.0, 11, 7, 0

    def uniform(self):
        x = self.x
        self.ir = (self.ir + 1) % 12
        if self.ir == self.ir_old:
            ir, jr, c = self.ir, self.jr, self.carry
            for _ in range(202):
                y = x[jr] - x[ir] - c
                if y < 0.0:
                    c = self.ONE_BIT
                    y += 1.0
                else:
                    c = 0.0
                x[ir] = y
                ir, jr = (ir + 1) % 12, (jr + 1) % 12
            self.ir, self.jr, self.carry, self.ir_old = ir, jr, c, ir
        return x[self.ir]


def ngenic_seed_table(n, seed):
    """RNG_ngenic's constructor (src/plugins/random_ngenic.cc:42-75): N^2 column seeds from ONE stream, drawn ring by ring"""
    r = Ranlxd1(seed)
    T = np.zeros((n, n), dtype=np.uint32)
    draw = lambda: np.uint32(int(0x7fffffff * r.uniform()))
    for i in range(n // 2):
        for j in range(i):
            T[i, j] = draw()
        for j in range(i + 1):
            T[j, i] = draw()
        for j in range(i):
            T[n - 1 - i, j] = draw()
        for j in range(i + 1):
            T[n - 1 - j, i] = draw()
        for j in range(i):
            T[i, n - 1 - j] = draw()
        for j in range(i + 1):
            T[j, n - 1 - i] = draw()
        for j in range(i):
            T[n - 1 - i, n - 1 - j] = draw()
        for j in range(i + 1):
            T[n - 1 - j, n - 1 - i] = draw()
    return T


def ngenic_noise(n, seed):
    """RNG_ngenic::Fill_Grid (src/plugins/random_ngenic.cc:84-154, the non-distributed branch; the distributed one fills the same
    logical modes): the reference's white noise for [random] seed, k-space [kx][ky][kz], before any conditioning."""
    T = ngenic_seed_table(n, seed)
    h, nzc = n // 2, n // 2 + 1
    W = np.zeros((n, n, nzc), dtype=np.complex128)
    for i in range(n):
        ii = (n - i) % n
        for j in range(n):
            jj = (n - j) % n
            r = Ranlxd1(int(T[i, j]))
            for k in range(nzc):
                phase = r.uniform() * 2 * np.pi
                while True:
                    ampl = r.uniform()
                    if ampl != 0.0 and ampl != 1.0:
                        break
                if i == h or j == h or k == h or (i == 0 and j == 0 and k == 0):
                    continue
                ampl = np.sqrt(-np.log(ampl))
                z = complex(ampl * np.cos(phase), ampl * np.sin(phase))
                if k > 0:
                    W[i, j, k] = z
                elif i == 0:
                    if j < h:
                        W[i, j, 0] = z
                        W[i, jj, 0] = np.conj(z)
                elif i < h:
                    W[i, j, 0] = z
                    W[ii, jj, 0] = np.conj(z)
    return W


def make_white_noise_k(n, seed, nyquist_zero=True):
    """Unit-variance complex Gaussian white noise on the half spectrum [kx][ky][kz] that is the
    transform of a real field.  With nyquist_zero (NGenIC behaviour, random_ngenic.cc:117) all
    Nyquist planes and DC are zero."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, n, n))
    W = np.fft.rfftn(x) / np.sqrt(float(n) ** 3)
    if nyquist_zero:
        W[n // 2, :, :] = 0.0
        W[:, n // 2, :] = 0.0
        W[:, :, n // 2] = 0.0
        W[0, 0, 0] = 0.0
    return np.ascontiguousarray(W)


def power_law_amplitude(ns=0.9649, kcut=0.5, norm=1.0):
    """A smooth stand-in for k^{ns/2} T(k) sqrt(pnorm) (BBKS-like roll-off)."""

    def amp(k):
        k = np.asarray(k, dtype=np.float64)
        q = k / kcut
        T = np.log1p(2.34 * q) / np.where(q > 0, 2.34 * q, 1.0) * (
            1 + 3.89 * q + (16.1 * q) ** 2 + (5.46 * q) ** 3 + (6.71 * q) ** 4) ** -0.25
        T = np.where(q > 0, T, 1.0)
        return norm * np.power(k, 0.5 * ns) * T

    return amp
