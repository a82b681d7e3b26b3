// grid_ops.cuh - element-wise k-space / real-space operators, reductions and the lattice
// read-out of the LPT pipeline (FP64, bandwidth-bound streaming kernels).
// Each kernel names the reference loop it replaces (paths relative to /root/reference).
#pragma once
#include "fft_core.cuh"
#include "ranlux.cuh"

namespace mfb {

// geometry of a cubic-or-not grid in k-space.  Single GPU: layout [kx][ky][kz] (kz = 0..n2/2).
// Slab-decomposed (tr = 1): the reference's MPI layout [ky_loc][kx][kz] (src/grid_fft.cc:120-128,
// include/grid_fft.hh:282-296), this rank holding ky = off .. off + s0 - 1.
struct KGeom {
    int n0, n1, n2, nzc;
    double kf0, kf1, kf2;  // 2 pi / L_d
    int s0, s1;            // storage extents of the two slow axes
    int tr, off;           // transposed layout, first local ky
};

MFB_HD int signed_idx(int i, int n) { return i - ((i > n / 2) ? n : 0); }
// get_k (include/grid_fft.hh:281-296): Nyquist index carries +n/2
MFB_HD double getk(int i, int n, double kf) { return double(signed_idx(i, n)) * kf; }
// gradient wavenumber (include/grid_fft.hh:376-385, include/operators.hh:50-55): 0 at Nyquist
MFB_HD double gradk(int i, int n, double kf) { return (i == n / 2) ? 0.0 : double(signed_idx(i, n)) * kf; }

constexpr int EW_THREADS = 256;
constexpr int EW_ROWS = 4;  // k-space rows (fixed kx,ky) per block

// iterate over the k-space elements of EW_ROWS rows; opens ONE brace (closed by the caller) and
// defines i, j, k, idx
#define MFB_K_LOOP_BEGIN(G)                                                   \
    for (int t_ = threadIdx.x; t_ < EW_ROWS * (G).nzc; t_ += EW_THREADS) {    \
        const long r_ = long(blockIdx.x) * EW_ROWS + t_ / (G).nzc;            \
        if (r_ >= long((G).s0) * (G).s1) break;                               \
        const int k = t_ % (G).nzc;                                           \
        const int a_ = int(r_ / (G).s1);                                      \
        const int b_ = int(r_ - long(a_) * (G).s1);                           \
        const int i = (G).tr ? b_ : a_;                                       \
        const int j = (G).tr ? a_ + (G).off : b_;                             \
        const long idx = r_ * (G).nzc + k;                                    \
        const bool is_dc = (i | j | k) == 0;                                  \
        (void)i; (void)j; (void)k; (void)is_dc;

inline unsigned kgrid_blocks(const KGeom& g) { return unsigned((long(g.s0) * g.s1 + EW_ROWS - 1) / EW_ROWS); }


// ---------------------------------------------------------------- PPT / Eulerian branch (src/ic_generator.cc:685-803)
// Complex grids (Grid_FFT<ccomplex_t>) are dense [n0][n1][n2] double2; real grids keep the padded rows (pitch n2+2).
struct CGeom {
    int n0, n1, n2;
    double kf0, kf1, kf2;
    ptrdiff_t rpitch;  // doubles per row of the real companions
    int s0, s1;        // storage extents of the two slow axes: (n0 or x_loc, n1) in real space, (ky_loc, n0) in a slab's k-space
    int tr, off;       // transposed k-space layout of a slab, first local ky
};
// i, j, k: indices along x, y, z (wave indices in k-space); t_: element in the complex grid; ridx: same cell in a real companion
#define MFB_C_LOOP_BEGIN(G)                                                                                       \
    const long total_ = long((G).s0) * (G).s1 * (G).n2;                                                           \
    for (long t_ = long(blockIdx.x) * EW_THREADS + threadIdx.x; t_ < total_; t_ += long(gridDim.x) * EW_THREADS) { \
        const int k = int(t_ % (G).n2);                                                                           \
        const long r_ = t_ / (G).n2;                                                                              \
        const int a_ = int(r_ / (G).s1), b_ = int(r_ % (G).s1);                                                   \
        const int i = (G).tr ? b_ : a_;                                                                           \
        const int j = (G).tr ? a_ + (G).off : b_;                                                                 \
        const long ridx = r_ * (G).rpitch + k;                                                                    \
        (void)i; (void)j; (void)ridx;

// rho <- sqrt(1 + C rho) on the logical real elements (:703-705)
__global__ void __launch_bounds__(EW_THREADS) ppt_sqrt1p_kernel(double* __restrict__ rho, CGeom g, double C) {
    MFB_C_LOOP_BEGIN(g)
        rho[ridx] = sqrt(1.0 + C * rho[ridx]);
    }
}
// psi <- rho * exp(i (phi [+ phi2]) / (hbar D))  (:741-751).  The argument is formed exactly as the reference does:
// complex(0,1/hbar) * (phi [+ phi2]) / D  ->  imaginary part ((1/hbar) * sum) / D
__global__ void __launch_bounds__(EW_THREADS) ppt_init_psi_kernel(double2* __restrict__ psi, const double* __restrict__ rho,
                                                                   const double* __restrict__ phi, const double* __restrict__ phi2,
                                                                   CGeom g, double inv_hbar, double D) {
    MFB_C_LOOP_BEGIN(g)
        const double sum = phi2 ? (phi[ridx] + phi2[ridx]) : phi[ridx];
        const double arg = phi2 ? (inv_hbar * sum) / D : inv_hbar * (sum / D);
        double sn, cs;
        sincos(arg, &sn, &cs);
        const double a = rho[ridx];
        psi[t_] = make_double2(a * cs, a * sn);
    }
}
// drift in k-space: psi <- psi * exp(-i/2 hbar k^2 D), k from get_k on all three axes (:757-760)
__global__ void __launch_bounds__(EW_THREADS) ppt_drift_kernel(double2* __restrict__ psi, CGeom g, double hbar, double D) {
    MFB_C_LOOP_BEGIN(g)
        const double kx = getk(i, g.n0, g.kf0), ky = getk(j, g.n1, g.kf1), kz = getk(k, g.n2, g.kf2);
        const double k2 = kx * kx + ky * ky + kz * kz;
        const double arg = -((0.5 * hbar) * k2) * D;
        double sn, cs;
        sincos(arg, &sn, &cs);
        psi[t_] = cmul(psi[t_], make_double2(cs, sn));
    }
}
// kick: psi <- psi * exp(i phi2 / (hbar D))  (:763-767)
__global__ void __launch_bounds__(EW_THREADS) ppt_kick_kernel(double2* __restrict__ psi, const double* __restrict__ phi2, CGeom g,
                                                               double inv_hbar, double D) {
    MFB_C_LOOP_BEGIN(g)
        double sn, cs;
        sincos((inv_hbar * phi2[ridx]) / D, &sn, &cs);
        psi[t_] = cmul(psi[t_], make_double2(cs, sn));
    }
}
// rho <- |psi|^2 - 1  (:772-775)
__global__ void __launch_bounds__(EW_THREADS) ppt_density_kernel(double* __restrict__ rho, const double2* __restrict__ psi, CGeom g) {
    MFB_C_LOOP_BEGIN(g)
        const double2 p = psi[t_];
        rho[ridx] = p.x * p.x + p.y * p.y - 1.0;
    }
}
// g <- g * i k_d, k from get_k (:790-792)
__global__ void __launch_bounds__(EW_THREADS) cgrid_gradient_kernel(double2* __restrict__ d, CGeom g, int idim) {
    MFB_C_LOOP_BEGIN(g)
        const double kd = (idim == 0) ? getk(i, g.n0, g.kf0) : ((idim == 1) ? getk(j, g.n1, g.kf1) : getk(k, g.n2, g.kf2));
        const double2 v = d[t_];
        d[t_] = make_double2(-v.y * kd, v.x * kd);
    }
}
// v <- vunit * Re[(conj(psi) grad - psi conj(grad)) / (2i/hbar)] / (1+rho) = vunit * hbar * Im(conj(psi) grad) / (1+rho)  (:796-799)
__global__ void __launch_bounds__(EW_THREADS) ppt_velocity_kernel(double* __restrict__ out, const double2* __restrict__ psi,
                                                                   const double2* __restrict__ grad, const double* __restrict__ rho,
                                                                   CGeom g, double hbar, double vunit) {
    MFB_C_LOOP_BEGIN(g)
        const double2 p = psi[t_], q = grad[t_];
        const double im2 = 2.0 * (p.x * q.y - p.y * q.x);   // Im(conj(p) q - p conj(q))
        out[ridx] = vunit * ((im2 / (2.0 / hbar)) / (1.0 + rho[ridx]));
    }
}

// ---------------------------------------------------------------- plain streaming ops
__global__ void __launch_bounds__(EW_THREADS) scale_kernel(double* __restrict__ d, size_t n, double a) {
    for (size_t i = size_t(blockIdx.x) * EW_THREADS + threadIdx.x; i < n; i += size_t(gridDim.x) * EW_THREADS) d[i] *= a;
}
__global__ void __launch_bounds__(EW_THREADS) multiply_kernel(double* __restrict__ y, const double* __restrict__ x, size_t n) {
    // y <- y*x over the whole real-indexed buffer (NaiveConvolver product loop, include/convolution.hh:341-345)
    for (size_t i = size_t(blockIdx.x) * EW_THREADS + threadIdx.x; i < n; i += size_t(gridDim.x) * EW_THREADS) y[i] *= x[i];
}
__global__ void __launch_bounds__(EW_THREADS) affine_kernel(double* __restrict__ d, size_t n, double a, double b) {
    // x <- a + b*x  (src/ic_generator.cc:673-675 mass field)
    for (size_t i = size_t(blockIdx.x) * EW_THREADS + threadIdx.x; i < n; i += size_t(gridDim.x) * EW_THREADS)
        d[i] = a + b * d[i];
}
__global__ void __launch_bounds__(EW_THREADS) axpy_kernel(double* __restrict__ y, const double* __restrict__ x, size_t n, double a, double b) {
    // y <- a*y + b*x
    for (size_t i = size_t(blockIdx.x) * EW_THREADS + threadIdx.x; i < n; i += size_t(gridDim.x) * EW_THREADS)
        y[i] = a * y[i] + b * x[i];
}

// ---------------------------------------------------------------- k-space operators
// wnoise conditioning (src/ic_generator.cc:341-358)
__global__ void __launch_bounds__(EW_THREADS) noise_condition_kernel(double2* __restrict__ w, KGeom g, double inv_volfac,
                                                                      int fix, int remove_corners, double kny) {
    MFB_K_LOOP_BEGIN(g)
        double2 v = w[idx];
        if (fix) {
            const double a = sqrt(v.x * v.x + v.y * v.y);
            if (a != 0.0) { v.x /= a; v.y /= a; }
        }
        v.x *= inv_volfac; v.y *= inv_volfac;
        if (remove_corners) {
            const double kx = getk(i, g.n0, g.kf0), ky = getk(j, g.n1, g.kf1), kz = getk(k, g.n2, g.kf2);
            if (sqrt(kx * kx + ky * ky + kz * kz) > kny) v = make_double2(0.0, 0.0);
        }
        w[idx] = v;
    }
}

// out = in * T(|n|^2) * (mode==1 ? -1/|k|^2 : 1), DC = 0.   T = tabulated A(|k|) on integer |n|^2
// mode 1: phi(1) = -W A / k^2 (src/ic_generator.cc:461-468); mode 0: rho = W A (:667-670)
__global__ void __launch_bounds__(EW_THREADS) radial_table_kernel(double2* __restrict__ out, const double2* __restrict__ in,
                                                                   const double* __restrict__ table, KGeom g, int mode) {
    MFB_K_LOOP_BEGIN(g)
        const int a = signed_idx(i, g.n0), b = signed_idx(j, g.n1), c = signed_idx(k, g.n2);
        const long m = long(a) * a + long(b) * b + long(c) * c;
        double f = __ldg(&table[m]);
        if (mode == 1) {
            const double kk = g.kf0 * g.kf0 * double(m);
            f = -f / kk;
        }
        double2 v = in[idx];
        v = (m == 0) ? make_double2(0.0, 0.0) : make_double2(v.x * f, v.y * f);
        out[idx] = v;
    }
}

// x <- x * s * |k|^(2p), p = -1,0,+1, DC = 0
// p=-1,s=-1: apply_InverseLaplacian (include/grid_fft.hh:440-449); p=+1: (negative) Laplacian (:418-438)
__global__ void __launch_bounds__(EW_THREADS) laplacian_pow_kernel(double2* __restrict__ out, const double2* __restrict__ in,
                                                                    KGeom g, int p, double s) {
    MFB_K_LOOP_BEGIN(g)
        const double kx = getk(i, g.n0, g.kf0), ky = getk(j, g.n1, g.kf1), kz = getk(k, g.n2, g.kf2);
        const double k2 = kx * kx + ky * ky + kz * kz;
        double f = s;
        if (p > 0) f *= k2;
        else if (p < 0) f /= k2;
        double2 v = in[idx];
        v = is_dc ? make_double2(0.0, 0.0) : make_double2(v.x * f, v.y * f);
        out[idx] = v;
    }
}

// external-tide correction of phi2 source (src/ic_generator.cc:492-496)
__global__ void __launch_bounds__(EW_THREADS) tides_kernel(double2* __restrict__ phi2, const double2* __restrict__ phi,
                                                            KGeom g, double l0, double l1, double l2) {
    MFB_K_LOOP_BEGIN(g)
        const double kx = getk(i, g.n0, g.kf0), ky = getk(j, g.n1, g.kf1), kz = getk(k, g.n2, g.kf2);
        const double k2 = kx * kx + ky * ky + kz * kz;
        const double fa = kx * kx * l0 + ky * ky * l1 + kz * kz * l2;
        const double f = (l0 + l1 + l2) * k2 + (4.0 / 3.0) * fa;
        const double2 p = phi[idx];
        double2 v = phi2[idx];
        v.x -= f * p.x; v.y -= f * p.y;
        phi2[idx] = v;
    }
}

// sub-cell shift (include/grid_fft.hh:832-842): x <- x * exp(i (s.k) dx)
__global__ void __launch_bounds__(EW_THREADS) shift_phase_kernel(double2* __restrict__ d, KGeom g, double sx, double sy, double sz) {
    MFB_K_LOOP_BEGIN(g)
        // k_d * dx_d = 2 pi n_d / N_d  -> use sincospi for accuracy
        const double ph = 2.0 * (sx * double(signed_idx(i, g.n0)) / g.n0 + sy * double(signed_idx(j, g.n1)) / g.n1 +
                                 sz * double(signed_idx(k, g.n2)) / g.n2);
        double sn, cs;
        sincospi(ph, &sn, &cs);
        const double2 v = d[idx];
        d[idx] = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
    }
}

// Psi / v assembly (src/ic_generator.cc:841-869, 896-940):
//   out = G_d (w1 phi + w2 phi2 + w3 phi3) + wc (G_p A_q - G_q A_p) [- cvbc * T(|k|) W G_d / k^2], * scale, DC=0
// The gradient operator `lg` of src/ic_generator.cc:377-381.  ng == 0: op::fourier_gradient (include/operators.hh:38-61), i k_d
// with 0 at the Nyquist index.  ng > 0: particle::lattice_gradient (include/particle_plt.hh:551-585, build option ENABLE_PLT):
// the eigenvalues of the lattice's dynamical matrix in spherical components (tab[0..2] = Re D_xx_, D_yy_, D_zz_ of that class)
// and the growth-rate correction (tab[3] = Re D_xy_), tabulated by the host on an ng^3 helper mesh in k-space
// ([ng][ng][ng/2+1]) and read out with the k-space CIC rule of Grid_FFT::get_cic_kspace (include/grid_fft.hh:347-373).
struct LatticeGradient {
    const double* tab[4];
    int ng;
    double mapratio, boxlen;
};
MFB_D double plt_cic(const double* __restrict__ T, int ng, double x, double y, double z) {
    const int nzc = ng / 2 + 1;
    const int ix = int(floor(x)), iy = int(floor(y)), iz = int(floor(z));
    const double dx = x - double(ix), tx = 1.0 - dx, dy = y - double(iy), ty = 1.0 - dy, dz = z - double(iz), tz = 1.0 - dz;
    const int ix1 = (ix + 1) % ng, iy1 = (iy + 1) % ng, iz1 = min(iz + 1, nzc - 1);
    auto at = [&](int a, int b, int c) { return __ldg(&T[(long(a) * ng + b) * nzc + c]); };
    double v = 0.0;   // the order of the reference's eight terms
    v += at(ix, iy, iz) * tx * ty * tz;
    v += at(ix, iy, iz1) * tx * ty * dz;
    v += at(ix, iy1, iz) * tx * dy * tz;
    v += at(ix, iy1, iz1) * tx * dy * dz;
    v += at(ix1, iy, iz) * dx * ty * tz;
    v += at(ix1, iy, iz1) * dx * ty * dz;
    v += at(ix1, iy1, iz) * dx * dy * tz;
    v += at(ix1, iy1, iz1) * dx * dy * dz;
    return v;
}
// G[d] = Im lg.gradient(d, ijk); returns lg.vfac_corr(ijk)
MFB_D double plt_gradient(const LatticeGradient& P, int i, int j, int k, double* G) {
    const double x = double(i) * P.mapratio, y = double(j) * P.mapratio, z = double(k) * P.mapratio;
    const double tau = 6.283185307179586476925286766559, h = double(P.ng / 2), n = double(P.ng);
    // D_xx_.get_k<real_t>(x, y, z) on the unit-length helper mesh (include/grid_fft.hh:299-313)
    const double kx = (x - (x > h ? n : 0.0)) * tau, ky = (y - (y > h ? n : 0.0)) * tau, kz = (z - (z > h ? n : 0.0)) * tau;
    const double kr = sqrt(kx * kx + ky * ky + kz * kz);
    const double kmod = kr / P.mapratio / P.boxlen;
    const double Dr = plt_cic(P.tab[0], P.ng, x, y, z), Dt = plt_cic(P.tab[1], P.ng, x, y, z), Dp = plt_cic(P.tab[2], P.ng, x, y, z);
    const double kphi = kr > 0.0 ? atan2(ky, kx) : 0.0, ktheta = kr > 0.0 ? acos(kz / kr) : 0.0;
    const double st = sin(ktheta), ct = cos(ktheta), sp = sin(kphi), cp = cos(kphi);
    G[0] = kmod * (Dr * st * cp + Dt * ct * cp - Dp * sp);
    G[1] = kmod * (Dr * st * sp + Dt * ct * sp + Dp * cp);
    G[2] = kmod * (Dr * ct - Dt * st);
    return plt_cic(P.tab[3], P.ng, x, y, z);
}
struct AssembleArgs {
    const double2 *phi, *phi2, *phi3, *Ap, *Aq, *wnoise;
    const double* vbc_table;
    double w1, w2, w3, wc, cvbc, scale;
    int d, p, q;
    LatticeGradient lg;       // lg.ng == 0: plain spectral gradient
    int vfac_corr;            // velocity fields: multiply the mode by lg.vfac_corr (src/ic_generator.cc:929)
    int cic_comp;             // glass loads: CIC compensation kernel (include/grid_interpolate.hh:195-208, src/ic_generator.cc:861,924)
};
__global__ void __launch_bounds__(EW_THREADS) assemble_kernel(double2* __restrict__ out, AssembleArgs A, KGeom g) {
    MFB_K_LOOP_BEGIN(g)
        double G[3] = {gradk(i, g.n0, g.kf0), gradk(j, g.n1, g.kf1), gradk(k, g.n2, g.kf2)};
        double vcorr = 1.0;
        if (A.lg.ng > 0) vcorr = plt_gradient(A.lg, i, j, k, G);
        const double Gd = (A.d == 0) ? G[0] : ((A.d == 1) ? G[1] : G[2]);
        double2 t = A.phi[idx];
        t.x *= A.w1; t.y *= A.w1;
        if (A.phi2) { const double2 v = A.phi2[idx]; t.x = fma(A.w2, v.x, t.x); t.y = fma(A.w2, v.y, t.y); }
        if (A.phi3) { const double2 v = A.phi3[idx]; t.x = fma(A.w3, v.x, t.x); t.y = fma(A.w3, v.y, t.y); }
        // multiply by i*Gd
        double2 r = make_double2(-Gd * t.y, Gd * t.x);
        if (A.Ap) {
            const double Gp = (A.p == 0) ? G[0] : ((A.p == 1) ? G[1] : G[2]);
            const double Gq = (A.q == 0) ? G[0] : ((A.q == 1) ? G[1] : G[2]);
            const double2 aq = A.Aq[idx], ap = A.Ap[idx];
            // i*(Gp*Aq - Gq*Ap) * wc
            const double cx = Gp * aq.x - Gq * ap.x, cy = Gp * aq.y - Gq * ap.y;
            r.x -= A.wc * cy; r.y += A.wc * cx;
        }
        if (A.vbc_table) {
            const int a = signed_idx(i, g.n0), b = signed_idx(j, g.n1), c = signed_idx(k, g.n2);
            const long m = long(a) * a + long(b) * b + long(c) * c;
            if (m > 0) {
                const double f = A.cvbc * __ldg(&A.vbc_table[m]) * Gd / (g.kf0 * g.kf0 * double(m));
                const double2 w = A.wnoise[idx];
                r.x -= -f * w.y; r.y -= f * w.x;
            }
        }
        if (A.cic_comp) {
            // 1 / (sinc(pi k_x / 2 k_ny) sinc(..) sinc(..))^2 with k = get_k (Nyquist index carries +n/2), k / k_ny = 2 m / n
            auto sinc = [](double x) { return (fabs(x) > 1e-10) ? sin(x) / x : 1.0; };
            const double hp = 1.5707963267948966192313216916398;
            const double kx = getk(i, g.n0, g.kf0), ky = getk(j, g.n1, g.kf1), kz = getk(k, g.n2, g.kf2);
            const double f = sinc(hp * kx / (g.kf0 * double(g.n0 / 2))) * sinc(hp * ky / (g.kf1 * double(g.n1 / 2))) *
                             sinc(hp * kz / (g.kf2 * double(g.n2 / 2)));
            const double c = 1.0 / (f * f);
            r.x *= c; r.y *= c;
        }
        if (A.vfac_corr) { r.x *= vcorr; r.y *= vcorr; }
        r.x *= A.scale; r.y *= A.scale;
        if (is_dc) r = make_double2(0.0, 0.0);
        out[idx] = r;
    }
}

// Fourier-space resample == Grid_FFT::FourierInterpolateCopyTo (src/grid_fft.cc:382-394,432-463):
// target pre-zeroed by the caller; copy modes strictly inside both Nyquist limits.
__global__ void __launch_bounds__(EW_THREADS) resample_kernel(double2* __restrict__ dst, KGeom gd, const double2* __restrict__ src, KGeom gs) {
    MFB_K_LOOP_BEGIN(gd)
        const int h0 = min(gs.n0, gd.n0) / 2, h1 = min(gs.n1, gd.n1) / 2, h2 = min(gs.n2, gd.n2) / 2;
        const int a = signed_idx(i, gd.n0), b = signed_idx(j, gd.n1);
        double2 v = make_double2(0.0, 0.0);
        if (abs(a) < h0 && abs(b) < h1 && k < h2) {
            const int is = (a < 0) ? a + gs.n0 : a, js = (b < 0) ? b + gs.n1 : b;
            v = src[(long(is) * gs.n1 + js) * gs.nzc + k];
        }
        dst[idx] = v;
    }
}

// ---------------------------------------------------------------- reductions (deterministic two-stage)
constexpr int RED_THREADS = 256;
// real view: rows of `pitch` doubles, first n2 valid.  partial[2*b] = sum, partial[2*b+1] = sum of squares
__global__ void __launch_bounds__(RED_THREADS) reduce_real_kernel(const double* __restrict__ d, long nrows, int n2, ptrdiff_t pitch,
                                                                   double* __restrict__ partial) {
    __shared__ double s1[RED_THREADS], s2[RED_THREADS];
    double a = 0.0, b = 0.0;
    const long total = nrows * n2;
    for (long t = long(blockIdx.x) * RED_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * RED_THREADS) {
        const long r = t / n2;
        const double v = d[r * pitch + (t - r * n2)];
        a += v; b += v * v;
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (int(threadIdx.x) < s) { s1[threadIdx.x] += s1[threadIdx.x + s]; s2[threadIdx.x] += s2[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s1[0]; partial[2 * blockIdx.x + 1] = s2[0]; }
}
// complex view (k-space std, include/grid_fft.hh:507-553): sum of real parts, sum of |z|^2
__global__ void __launch_bounds__(RED_THREADS) reduce_cplx_kernel(const double2* __restrict__ d, long n, double* __restrict__ partial) {
    __shared__ double s1[RED_THREADS], s2[RED_THREADS];
    double a = 0.0, b = 0.0;
    for (long t = long(blockIdx.x) * RED_THREADS + threadIdx.x; t < n; t += long(gridDim.x) * RED_THREADS) {
        const double2 v = d[t];
        a += v.x; b += v.x * v.x + v.y * v.y;
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (int(threadIdx.x) < s) { s1[threadIdx.x] += s1[threadIdx.x + s]; s2[threadIdx.x] += s2[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s1[0]; partial[2 * blockIdx.x + 1] = s2[0]; }
}
// max |phi(x+1)-phi(x)| over the three axes (src/ic_generator.cc:713-731; note i < size(0)-1)
__global__ void __launch_bounds__(RED_THREADS) max_neighbour_diff_kernel(const double* __restrict__ d, int n0, int n1, int n2,
                                                                          ptrdiff_t pitch, double* __restrict__ partial) {
    __shared__ double s1[RED_THREADS];
    double m = -1.0;
    const long total = long(n0 - 1) * n1 * n2;
    for (long t = long(blockIdx.x) * RED_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * RED_THREADS) {
        const int k = int(t % n2);
        const long r = t / n2;
        const int j = int(r % n1), i = int(r / n1);
        const int ir = (i + 1) % n0, jr = (j + 1) % n1, kr = (k + 1) % n2;
        const double c = d[(long(i) * n1 + j) * pitch + k];
        m = fmax(m, fabs(d[(long(ir) * n1 + j) * pitch + k] - c));
        m = fmax(m, fabs(d[(long(i) * n1 + jr) * pitch + k] - c));
        m = fmax(m, fabs(d[(long(i) * n1 + j) * pitch + kr] - c));
    }
    s1[threadIdx.x] = m;
    __syncthreads();
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (int(threadIdx.x) < s) s1[threadIdx.x] = fmax(s1[threadIdx.x], s1[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s1[0];
}
// final stage: one block folds `nb` partial pairs. mode 0: sum pairs; mode 1: max singles
__global__ void __launch_bounds__(RED_THREADS) reduce_final_kernel(const double* __restrict__ partial, int nb, int mode, double* __restrict__ out) {
    __shared__ double s1[RED_THREADS], s2[RED_THREADS];
    double a = (mode == 1) ? -1.0 : 0.0, b = 0.0;
    for (int t = threadIdx.x; t < nb; t += RED_THREADS) {
        if (mode == 1) a = fmax(a, partial[t]);
        else { a += partial[2 * t]; b += partial[2 * t + 1]; }
    }
    s1[threadIdx.x] = a; s2[threadIdx.x] = b;
    __syncthreads();
    for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
        if (int(threadIdx.x) < s) {
            if (mode == 1) s1[threadIdx.x] = fmax(s1[threadIdx.x], s1[threadIdx.x + s]);
            else { s1[threadIdx.x] += s1[threadIdx.x + s]; s2[threadIdx.x] += s2[threadIdx.x + s]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s1[0]; out[1] = s2[0]; }
}
// real-space DC removal (include/grid_fft.hh:851-886)
__global__ void __launch_bounds__(EW_THREADS) subtract_real_kernel(double* __restrict__ d, long nrows, int n2, ptrdiff_t pitch, double m) {
    const long total = nrows * n2;
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * EW_THREADS) {
        const long r = t / n2;
        d[r * pitch + (t - r * n2)] -= m;
    }
}

// ---------------------------------------------------------------- particles (include/particle_generator.hh)
// IDs (:150-169): id = idoffset + ov*cell + iload at slot iload*ncell + cell
template <typename id_t>
__global__ void __launch_bounds__(EW_THREADS) particle_ids_kernel(id_t* __restrict__ ids, long ncell, int ov, unsigned long long idoffset,
                                                                   long cell0) {
    const long total = ncell * ov;
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * EW_THREADS) {
        const long iload = t / ncell, c = t - iload * ncell;
        ids[t] = id_t(idoffset + (unsigned long long)ov * (unsigned long long)(cell0 + c) + (unsigned long long)iload);
    }
}
// lattice read-out (:448-465 positions, :544-560 velocities, :327-344 masses):
//   dst[stride*(slot0 + cell) + comp] = real_t( add_lattice ? ((idx_d + shift)/n_d)*lunit + f : mul*f )
template <typename real_out>
__global__ void __launch_bounds__(EW_THREADS) lattice_readout_kernel(real_out* __restrict__ dst, const double* __restrict__ f,
                                                                      int n0loc, int n0, int n1, int n2, ptrdiff_t pitch, int i0,
                                                                      long slot0, int stride, int comp, int add_lattice, int d,
                                                                      double shift, double lunit, double mul) {
    const long total = long(n0loc) * n1 * n2;  // this rank's slab of lattice planes (i0 = local_0_start_)
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * EW_THREADS) {
        const int k = int(t % n2);
        const long r = t / n2;
        const int j = int(r % n1), i = int(r / n1);
        const double v = f[r * pitch + k];
        double o;
        if (add_lattice) {
            const int id = (d == 0) ? (i + i0) : ((d == 1) ? j : k);
            const int nd = (d == 0) ? n0 : ((d == 1) ? n1 : n2);
            const double qpos = (double(id) + shift) / double(nd);
#ifdef MFB_EMU
            o = qpos * lunit + v;
#else
            o = __dadd_rn(__dmul_rn(qpos, lunit), v);  // no FMA contraction: bit-identical to the CPU expression
#endif
        } else {
            o = mul * v;
        }
        dst[size_t(stride) * size_t(slot0 + t) + comp] = real_out(o);
    }
}

// ---------------------------------------------------------------- device noise fill, NGenIC layout
// RNG_ngenic::Fill_Grid (src/plugins/random_ngenic.cc:84-154) with a counter-based generator: every mode is computed on its
// own, so the fill is one streaming write of the grid instead of N^2 sequential column streams + an H2D copy.
MFB_D void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = 0xD2511F53ull * c0, p1 = 0xCD9E8D57ull * c2;
        const unsigned n0 = unsigned(p1 >> 32) ^ c1 ^ k0, n1 = unsigned(p1), n2 = unsigned(p0 >> 32) ^ c3 ^ k1, n3 = unsigned(p0);
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// column word: position in the plugin's ring-ordered seed table (:47-74): i < N/2 as it is, else N-1-i with the top bit set
MFB_D unsigned ngenic_word(int i, int n) { return i < n / 2 ? unsigned(i) : (unsigned(n - 1 - i) | 0x80000000u); }
MFB_D double2 ngenic_draw(int a, int b, int k, int n, unsigned long long seed) {
    unsigned x[4];
    philox4x32_10(ngenic_word(a, n), ngenic_word(b, n), unsigned(k), 0u, unsigned(seed), unsigned(seed >> 32), x);
    const double u1 = (double((((unsigned long long)x[0]) << 20) | (x[1] >> 12)) + 0.5) * 2.220446049250313e-16;   // 2^-52
    const double u2 = (double((((unsigned long long)x[2]) << 20) | (x[3] >> 12)) + 0.5) * 2.220446049250313e-16;
    const double amp = sqrt(-log(u2));
    double sn, cs;
    sincos(6.283185307179586476925286766559 * u1, &sn, &cs);   // the reference takes cos / sin of 2 pi u (:120-121)
    return make_double2(amp * cs, amp * sn);
}
__global__ void __launch_bounds__(EW_THREADS) ngenic_fill_kernel(double2* __restrict__ w, KGeom g, unsigned long long seed) {
    MFB_K_LOOP_BEGIN(g)
        const int n = g.n0, h = n / 2;
        double2 z = make_double2(0.0, 0.0);
        if (!(i == h || j == h || k == h || is_dc)) {       // :117-118
            if (k > 0) {
                z = ngenic_draw(i, j, k, n, seed);
            } else {                                         // Hermitian k_z = 0 plane (:125-148)
                const bool primary = (i == 0 && j < h) || (i > 0 && i < h);
                z = primary ? ngenic_draw(i, j, 0, n, seed) : ngenic_draw((n - i) % n, (n - j) % n, 0, n, seed);
                if (!primary) z.y = -z.y;
            }
        }
        w[idx] = z;
    }
}

// The same fill with the reference's OWN generator: one RANLUX stream per (k_x, k_y) column, seeded from the plugin's seed table,
// two draws per k_z (phase, amplitude) in the plugin's order (src/plugins/random_ngenic.cc:96-121) - the reference's NGenIC white
// noise itself, to the last bit of log / cos / sin.  One thread per stored column walks k_z; its k_z = 0 element is the first draw
// of its own stream where the column is the primary one of the Hermitian pair (:125-148), else the conjugate of the partner
// column's first draw (that stream is seeded here a second time instead of writing across columns: no race, slab-local).
MFB_D double2 ngenic_ranlux_mode(Ranlxd1& r) {   // consumes the two draws of one mode (:110-115)
    const double phase = r.uniform() * 2 * M_PI;
    double ampl;
    do { ampl = r.uniform(); } while (ampl == 0.0 || ampl == 1.0);
    ampl = sqrt(-log(ampl));
    return make_double2(ampl * cos(phase), ampl * sin(phase));
}
__global__ void __launch_bounds__(128) ngenic_ranlux_fill_kernel(double2* __restrict__ w, KGeom g, const unsigned* __restrict__ seed_table) {
    const long col = long(blockIdx.x) * blockDim.x + threadIdx.x;
    if (col >= long(g.s0) * g.s1) return;
    const int a = int(col / g.s1), b = int(col - long(a) * g.s1);
    const int i = g.tr ? b : a, j = g.tr ? a + g.off : b;       // k_x, k_y of this stored column
    const int n = g.n0, h = n / 2;
    double2* row = w + col * g.nzc;
    const bool dead = (i == h || j == h);                        // :117 - Nyquist columns stay zero
    Ranlxd1 r;
    r.seed(seed_table[size_t(i) * n + j]);
    for (int k = 0; k < g.nzc; ++k) {
        const double2 z = ngenic_ranlux_mode(r);
        const bool zero = dead || k == h || (i | j | k) == 0;
        if (k > 0) row[k] = zero ? make_double2(0.0, 0.0) : z;
        else {
            const bool primary = (i == 0 && j < h) || (i > 0 && i < h);
            double2 z0 = z;
            if (!primary && !zero) {
                Ranlxd1 p;
                p.seed(seed_table[size_t((n - i) % n) * n + (n - j) % n]);
                z0 = ngenic_ranlux_mode(p);
                z0.y = -z0.y;
            }
            row[0] = zero ? make_double2(0.0, 0.0) : z0;
        }
    }
}

// ---------------------------------------------------------------- masked sc lattice (include/particle_generator.hh:171-260)
// A 2x2x2 mask assigns every cell of the N^3 lattice to a species (or to nobody); a species' particles are the cells of its
// type in (i,j,k) order.  The mask is periodic, so the running particle count of the reference's loops has a closed form.
struct LatticeMask {
    int sel[8];      // 1: cell parity (a,b,c) -> index 4a+2b+c belongs to this species
    long plane[2];   // selected cells in a plane with i % 2 == a
    long row[4];     // ... in a row with (i % 2, j % 2) == (a, b): index 2a+b
    int pair[4];     // ... in a pair of cells (k, k+1), k even, of such a row
};
// particles of this species before global cell (gi, j, k) in this rank's slab starting at plane i0 (even)
MFB_D long mask_rank(const LatticeMask& M, int gi, int i0, int j, int k) {
    const int il = gi - i0, a = gi & 1, b = j & 1;
    long c = long(il / 2) * (M.plane[0] + M.plane[1]) + ((il & 1) ? M.plane[i0 & 1] : 0);
    c += long(j / 2) * (M.row[2 * a] + M.row[2 * a + 1]) + (b ? M.row[2 * a] : 0);
    c += long(k / 2) * M.pair[2 * a + b] + ((k & 1) ? M.sel[4 * a + 2 * b] : 0);
    return c;
}
template <typename id_t>
__global__ void __launch_bounds__(EW_THREADS) masked_ids_kernel(id_t* __restrict__ ids, LatticeMask M, int n0loc, int n1, int n2, int i0,
                                                                 unsigned long long idoffset) {
    const long total = long(n0loc) * n1 * n2;
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * EW_THREADS) {
        const int k = int(t % n2);
        const long r = t / n2;
        const int j = int(r % n1), gi = int(r / n1) + i0;
        if (!M.sel[4 * (gi & 1) + 2 * (j & 1) + (k & 1)]) continue;
        // :236-243: IDoffset * global_size + get_cell_idx_1d
        ids[mask_rank(M, gi, i0, j, k)] = id_t(idoffset + (unsigned long long)((long(gi) * n1 + j) * n2 + k));
    }
}
// kind 0: positions (:479-497, get_unit_r * lunit + f), 1: velocities (:557-571), 2: masses (:351-382, pmean * (f - mean_mask))
template <typename real_out>
__global__ void __launch_bounds__(EW_THREADS) masked_readout_kernel(real_out* __restrict__ dst, const double* __restrict__ f, LatticeMask M,
                                                                     int n0loc, int n0, int n1, int n2, ptrdiff_t pitch, int i0, int kind,
                                                                     int d, double unit) {
    const long total = long(n0loc) * n1 * n2;
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < total; t += long(gridDim.x) * EW_THREADS) {
        const int k = int(t % n2);
        const long r = t / n2;
        const int j = int(r % n1), il = int(r / n1), gi = il + i0;
        if (!M.sel[4 * (gi & 1) + 2 * (j & 1) + (k & 1)]) continue;
        const long ip = mask_rank(M, gi, i0, j, k);
        const double v = f[r * pitch + k];
        if (kind == 2) {
            // get_mean_mask_value (:89-112): mean over the species' cells of the 2x2x2 block minus the mean over all eight
            const int oi = il - (gi & 1), oj = j - (j & 1), ok = k - (k & 1);
            double full = 0.0, masked = 0.0;
            int cnt = 0;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b)
                    for (int c = 0; c < 2; ++c) {
                        const double w = f[(long(oi + a) * n1 + (oj + b)) * pitch + ok + c];
                        full += w;
                        if (M.sel[4 * a + 2 * b + c]) { masked += w; ++cnt; }
                    }
            full /= 8.0;
            masked /= double(cnt);
            dst[ip] = real_out(unit * (v - (masked - full)));
        } else if (kind == 0) {
            const int id = (d == 0) ? gi : ((d == 1) ? j : k);
            const int nd = (d == 0) ? n0 : ((d == 1) ? n1 : n2);
            const double qpos = double(id) / double(nd);
#ifdef MFB_EMU
            dst[3 * ip + d] = real_out(qpos * unit + v);
#else
            dst[3 * ip + d] = real_out(__dadd_rn(__dmul_rn(qpos, unit), v));
#endif
        } else {
            dst[3 * ip + d] = real_out(v);
        }
    }
}

// ---------------------------------------------------------------- glass load (include/particle_generator.hh:498-515, 573-590)
// particle positions in grid units (glass_loader::glass_posr); the field is read out with the CIC rule of
// grid_interpolate::get_cic_at (include/grid_interpolate.hh:103-148, single slab).
template <typename id_t>
__global__ void __launch_bounds__(EW_THREADS) glass_ids_kernel(id_t* __restrict__ ids, long np, unsigned long long idoffset) {
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < np; t += long(gridDim.x) * EW_THREADS)
        ids[t] = id_t(idoffset + (unsigned long long)t);
}
template <typename real_out>
__global__ void __launch_bounds__(EW_THREADS) glass_readout_kernel(real_out* __restrict__ dst, const double* __restrict__ f,
                                                                    const double* __restrict__ posr, long np, int n0, int n1, int n2,
                                                                    ptrdiff_t pitch, int add_lattice, int d, double lunit) {
    for (long t = long(blockIdx.x) * EW_THREADS + threadIdx.x; t < np; t += long(gridDim.x) * EW_THREADS) {
        const double px = posr[3 * t], py = posr[3 * t + 1], pz = posr[3 * t + 2];
        const int ix = int(px), iy = int(py), iz = int(pz);
        const double dx = px - double(ix), tx = 1.0 - dx, dy = py - double(iy), ty = 1.0 - dy, dz = pz - double(iz), tz = 1.0 - dz;
        const int ix1 = (ix + 1) % n0, iy1 = (iy + 1) % n1, iz1 = (iz + 1) % n2;
        auto at = [&](int a, int b, int c) { return f[(long(a) * n1 + b) * pitch + c]; };
        double v = 0.0;
        v += at(ix, iy, iz) * tx * ty * tz;
        v += at(ix, iy, iz1) * tx * ty * dz;
        v += at(ix, iy1, iz) * tx * dy * tz;
        v += at(ix, iy1, iz1) * tx * dy * dz;
        v += at(ix1, iy, iz) * dx * ty * tz;
        v += at(ix1, iy, iz1) * dx * ty * dz;
        v += at(ix1, iy1, iz) * dx * dy * tz;
        v += at(ix1, iy1, iz1) * dx * dy * dz;
        if (add_lattice) {
            const double pd = (d == 0) ? px : ((d == 1) ? py : pz);
            const int nd = (d == 0) ? n0 : ((d == 1) ? n1 : n2);
#ifdef MFB_EMU
            v = pd / double(nd) * lunit + v;
#else
            v = __dadd_rn(__dmul_rn(__ddiv_rn(pd, double(nd)), lunit), v);
#endif
        }
        dst[3 * t + d] = real_out(v);
    }
}

}  // namespace mfb
