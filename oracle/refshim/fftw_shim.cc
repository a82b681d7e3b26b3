// oracle/refshim/fftw_shim.cc - TEST INFRASTRUCTURE (see fftw3.h in this directory): exact DFTs behind the FFTW calls of
// the reference's Grid_FFT.  Every 1-D transform is the plain O(n^2) sum with a table of exp(-+2 pi i m/n) evaluated in long
// double, accumulated in long double - slower than an FFT, but free of any algorithmic choice, which is the point of a pin.
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

struct fftw_plan_s {
    int kind;  // 0 r2c, 1 c2r, 2 c2c
    int n0, n1, n2, sign;
    void* data;
    int transposed;  // FFTW-MPI TRANSPOSED_OUT (forward) / TRANSPOSED_IN (backward): k-space stored [k1][k0][k2]
};
typedef std::complex<long double> lc;

static std::vector<lc> table(int n, int sign) {
    std::vector<lc> w(n);
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    for (int m = 0; m < n; ++m) w[m] = lc(cosl(tau * m / n), sign * sinl(tau * m / n));
    return w;
}
// in-place DFT of `n` elements spaced `stride` complex values apart
static void dft_line(std::complex<double>* x, int n, long stride, const std::vector<lc>& w) {
    std::vector<lc> in(n), out(n);
    for (int j = 0; j < n; ++j) in[j] = lc(x[j * stride].real(), x[j * stride].imag());
    for (int k = 0; k < n; ++k) {
        lc s = 0;
        for (int j = 0; j < n; ++j) s += in[j] * w[(long(j) * k) % n];
        out[k] = s;
    }
    for (int k = 0; k < n; ++k) x[k * stride] = std::complex<double>((double)out[k].real(), (double)out[k].imag());
}

extern "C" {
void* fftw_malloc(size_t n) { return aligned_alloc(64, (n + 63) / 64 * 64); }
void fftw_free(void* p) { free(p); }
static fftw_plan mk(int kind, int n0, int n1, int n2, int sign, void* d, int transposed = 0) {
    fftw_plan p = new fftw_plan_s{kind, n0, n1, n2, sign, d, transposed};
    return p;
}
// [a][b][m] <-> [b][a][m] for the first two axes of a complex array
static void swap01(std::complex<double>* c, int a, int b, int m) {
    std::vector<std::complex<double>> t(size_t(a) * b * m);
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j)
            for (int k = 0; k < m; ++k) t[(size_t(j) * a + i) * m + k] = c[(size_t(i) * b + j) * m + k];
    for (size_t i = 0; i < t.size(); ++i) c[i] = t[i];
}
static void execute_plain(const fftw_plan p);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double* in, fftw_complex*, unsigned) { return mk(0, n0, n1, n2, -1, in); }
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex* in, double*, unsigned) { return mk(1, n0, n1, n2, +1, in); }
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex* in, fftw_complex*, int sign, unsigned) { return mk(2, n0, n1, n2, sign, in); }
void fftw_destroy_plan(fftw_plan p) { delete p; }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}

void fftw_execute(const fftw_plan p) {
    if (!p->transposed) { execute_plain(p); return; }
    const int m = (p->kind == 2) ? p->n2 : p->n2 / 2 + 1;
    auto* c = reinterpret_cast<std::complex<double>*>(p->data);
    const bool to_k = (p->kind == 0) || (p->kind == 2 && p->sign == FFTW_FORWARD);
    if (to_k) { execute_plain(p); swap01(c, p->n0, p->n1, m); }      // natural in, [k1][k0][k2] out
    else { swap01(c, p->n1, p->n0, m); execute_plain(p); }           // [k1][k0][k2] in, natural out
}
void fftw_mpi_init(void) {}
ptrdiff_t fftw_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int, ptrdiff_t* l0, ptrdiff_t* s0, ptrdiff_t* l1,
                                            ptrdiff_t* s1) {
    *l0 = n0; *s0 = 0; *l1 = n1; *s1 = 0;
    return n0 * n1 * n2;   // complex elements this rank must be able to hold (callers pass n2/2+1 or n2 as they need)
}
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double* in, fftw_complex*, int, unsigned flags) {
    return mk(0, int(n0), int(n1), int(n2), -1, in, (flags & (1u << 30)) ? 1 : 0);
}
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, double*, int, unsigned flags) {
    return mk(1, int(n0), int(n1), int(n2), +1, in, (flags & (1u << 29)) ? 1 : 0);
}
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, fftw_complex*, int, int sign, unsigned flags) {
    return mk(2, int(n0), int(n1), int(n2), sign, in, (flags & ((1u << 29) | (1u << 30))) ? 1 : 0);
}
}   // extern "C"

static void execute_plain(const fftw_plan p) {
    const int n0 = p->n0, n1 = p->n1, n2 = p->n2, nzc = n2 / 2 + 1;
    auto* c = reinterpret_cast<std::complex<double>*>(p->data);
    auto* r = reinterpret_cast<double*>(p->data);
    if (p->kind == 2) {  // complex 3-D, dense [n0][n1][n2]
        const auto w0 = table(n0, p->sign), w1 = table(n1, p->sign), w2 = table(n2, p->sign);
        for (long i = 0; i < long(n0) * n1; ++i) dft_line(c + i * n2, n2, 1, w2);
        for (int i = 0; i < n0; ++i) for (int k = 0; k < n2; ++k) dft_line(c + (long(i) * n1) * n2 + k, n1, n2, w1);
        for (int j = 0; j < n1; ++j) for (int k = 0; k < n2; ++k) dft_line(c + long(j) * n2 + k, n0, long(n1) * n2, w0);
        return;
    }
    const auto w0 = table(n0, p->sign), w1 = table(n1, p->sign);
    if (p->kind == 0) {  // real [n0][n1][n2 (+2 pad)] -> half-complex [n0][n1][nzc], in place
        const auto w2 = table(n2, -1);
        std::vector<lc> row(nzc);
        for (long i = 0; i < long(n0) * n1; ++i) {
            const double* x = r + i * (n2 + 2);
            for (int k = 0; k < nzc; ++k) {
                lc s = 0;
                for (int j = 0; j < n2; ++j) s += (long double)x[j] * w2[(long(j) * k) % n2];
                row[k] = s;
            }
            for (int k = 0; k < nzc; ++k) c[i * nzc + k] = std::complex<double>((double)row[k].real(), (double)row[k].imag());
        }
        for (int i = 0; i < n0; ++i) for (int k = 0; k < nzc; ++k) dft_line(c + (long(i) * n1) * nzc + k, n1, nzc, w1);
        for (int j = 0; j < n1; ++j) for (int k = 0; k < nzc; ++k) dft_line(c + long(j) * nzc + k, n0, long(n1) * nzc, w0);
    } else {             // half-complex -> real, in place (x, y complex backward; z Hermitian-completed, Im of k=0 / Nyquist dropped)
        for (int j = 0; j < n1; ++j) for (int k = 0; k < nzc; ++k) dft_line(c + long(j) * nzc + k, n0, long(n1) * nzc, w0);
        for (int i = 0; i < n0; ++i) for (int k = 0; k < nzc; ++k) dft_line(c + (long(i) * n1) * nzc + k, n1, nzc, w1);
        const auto w2 = table(n2, +1);
        std::vector<lc> X(nzc);
        std::vector<long double> out(n2);
        for (long i = 0; i < long(n0) * n1; ++i) {
            for (int k = 0; k < nzc; ++k) X[k] = lc(c[i * nzc + k].real(), c[i * nzc + k].imag());
            for (int j = 0; j < n2; ++j) {
                long double s = X[0].real();
                if (n2 % 2 == 0) s += X[n2 / 2].real() * ((j & 1) ? -1.0L : 1.0L);
                for (int k = 1; k < (n2 + 1) / 2; ++k) s += 2.0L * (X[k] * w2[(long(j) * k) % n2]).real();
                out[j] = s;
            }
            double* x = r + i * (n2 + 2);
            for (int j = 0; j < n2; ++j) x[j] = (double)out[j];
        }
    }
}
