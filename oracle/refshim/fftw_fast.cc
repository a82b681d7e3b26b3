// oracle/refshim/fftw_fast.cc - TEST / BASELINE INFRASTRUCTURE (see fftw3.h in this directory): the FFTW3 calls of the
// reference's Grid_FFT served by a plain mixed-radix (4, 2, 3) FFT with OpenMP over the lines, so that the reference's OWN
// Grid_FFT / OrszagConvolver sources can be TIMED on the host cores (oracle/_ref/ref_bench, bench.py's cpu_baseline of kind
// "reference").  fftw_shim.cc is the exact O(n^2) DFT used to pin results; this file trades that for speed and is itself
// checked against the oracle (tests/test_oracle_vs_reference.py::test_fast_fft_backend).  Same conventions as FFTW:
// unnormalised, forward sign -1, in-place r2c / c2r with the padded last dimension, c2r ignores the imaginary parts of the
// self-conjugate modes.  It is NOT FFTW: no SIMD codelets, no cache-oblivious planning - expect FFTW or MKL to be faster.
#include "fftw3.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

typedef std::complex<double> cd;

struct fftw_plan_s {
    int kind;  // 0 r2c, 1 c2r, 2 c2c
    int n0, n1, n2, sign;
    void* data;
    int transposed = 0;   // one-rank FFTW-MPI plans with TRANSPOSED_OUT / _IN: k-space stored [k1][k0][k2]
};

namespace {
struct Tw {
    int n;
    std::vector<cd> w;   // w[m] = exp(-2 pi i m / n)
};
const Tw& table(int n) {
    static std::map<int, Tw> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    Tw t;
    t.n = n;
    t.w.resize(n);
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    for (int m = 0; m < n; ++m) t.w[m] = cd((double)cosl(tau * m / n), (double)-sinl(tau * m / n));
    return cache.emplace(n, std::move(t)).first->second;
}
inline cd twid(const Tw& t, long m, int sign) {
    const cd w = t.w[m % t.n];
    return sign < 0 ? w : std::conj(w);
}
inline cd mul_i(cd a, int sign) { return sign < 0 ? cd(a.imag(), -a.real()) : cd(-a.imag(), a.real()); }   // * exp(sign i pi/2)

// decimation in time, out of place: out[0..n) = DFT_n(in[0], in[stride], ...); N = length of the whole transform (twiddle table)
void fft_rec(const cd* in, long stride, cd* out, int n, const Tw& T, int sign) {
    if (n == 1) { out[0] = in[0]; return; }
    if (n == 2) { const cd a = in[0], b = in[stride]; out[0] = a + b; out[1] = a - b; return; }
    if (n == 3) {
        const cd a = in[0], b = in[stride], c = in[2 * stride];
        const double s = sign < 0 ? -0.86602540378443864676 : 0.86602540378443864676;
        const cd t = b + c, m = a - 0.5 * t, d = cd(0.0, s) * (b - c);
        out[0] = a + t; out[1] = m + d; out[2] = m - d;
        return;
    }
    if (n == 4) {
        const cd a = in[0], b = in[stride], c = in[2 * stride], d = in[3 * stride];
        const cd t0 = a + c, t1 = a - c, t2 = b + d, t3 = mul_i(b - d, sign);
        out[0] = t0 + t2; out[2] = t0 - t2; out[1] = t1 + t3; out[3] = t1 - t3;
        return;
    }
    const int r = (n % 4 == 0) ? 4 : ((n % 2 == 0) ? 2 : 3);
    const int m = n / r;
    for (int q = 0; q < r; ++q) fft_rec(in + q * stride, stride * r, out + q * m, m, T, sign);
    const long step = T.n / n;
    if (r == 4) {
        for (int k = 0; k < m; ++k) {
            const cd a = out[k], b = out[k + m] * twid(T, step * k, sign), c = out[k + 2 * m] * twid(T, 2 * step * k, sign),
                     d = out[k + 3 * m] * twid(T, 3 * step * k, sign);
            const cd t0 = a + c, t1 = a - c, t2 = b + d, t3 = mul_i(b - d, sign);
            out[k] = t0 + t2; out[k + 2 * m] = t0 - t2; out[k + m] = t1 + t3; out[k + 3 * m] = t1 - t3;
        }
    } else if (r == 2) {
        for (int k = 0; k < m; ++k) {
            const cd a = out[k], b = out[k + m] * twid(T, step * k, sign);
            out[k] = a + b; out[k + m] = a - b;
        }
    } else {
        const double s = sign < 0 ? -0.86602540378443864676 : 0.86602540378443864676;
        for (int k = 0; k < m; ++k) {
            const cd a = out[k], b = out[k + m] * twid(T, step * k, sign), c = out[k + 2 * m] * twid(T, 2 * step * k, sign);
            const cd t = b + c, mm = a - 0.5 * t, d = cd(0.0, s) * (b - c);
            out[k] = a + t; out[k + m] = mm + d; out[k + 2 * m] = mm - d;
        }
    }
}
bool supported(int n) { while (n % 2 == 0) n /= 2; while (n % 3 == 0) n /= 3; return n == 1; }

// `count` lines of length n: line l starts at base + l * dist, elements `stride` apart; blocks of 8 adjacent lines are
// gathered together so that strided passes touch whole cache lines
void fft_lines(cd* base, int n, long stride, long count, long dist, int sign) {
    if (!supported(n)) { std::abort(); }
    const Tw& T = table(n);
    constexpr int B = 8;
#pragma omp parallel
    {
        std::vector<cd> in(size_t(n) * B), out(size_t(n) * B);
#pragma omp for schedule(static)
        for (long l0 = 0; l0 < count; l0 += B) {
            const int nb = int(std::min<long>(B, count - l0));
            if (stride == 1) {
                for (int b = 0; b < nb; ++b) {
                    cd* x = base + (l0 + b) * dist;
                    fft_rec(x, 1, out.data(), n, T, sign);
                    std::memcpy(x, out.data(), sizeof(cd) * n);
                }
            } else {
                for (int j = 0; j < n; ++j)
                    for (int b = 0; b < nb; ++b) in[size_t(b) * n + j] = base[(l0 + b) * dist + j * stride];
                for (int b = 0; b < nb; ++b) fft_rec(in.data() + size_t(b) * n, 1, out.data() + size_t(b) * n, n, T, sign);
                for (int j = 0; j < n; ++j)
                    for (int b = 0; b < nb; ++b) base[(l0 + b) * dist + j * stride] = out[size_t(b) * n + j];
            }
        }
    }
}

void execute(const fftw_plan p) {
    const int n0 = p->n0, n1 = p->n1, n2 = p->n2, nzc = n2 / 2 + 1;
    cd* c = reinterpret_cast<cd*>(p->data);
    double* r = reinterpret_cast<double*>(p->data);
    if (p->kind == 2) {   // complex 3-D, dense [n0][n1][n2]
        fft_lines(c, n2, 1, long(n0) * n1, n2, p->sign);
        for (int i = 0; i < n0; ++i) fft_lines(c + long(i) * n1 * n2, n1, n2, n2, 1, p->sign);
        fft_lines(c, n0, long(n1) * n2, long(n1) * n2, 1, p->sign);
        return;
    }
    const long nrows = long(n0) * n1;
    const Tw& T2 = table(n2);
    if (p->kind == 0) {   // real [n0][n1][n2 (+2 pad)] -> half-complex [n0][n1][nzc], in place; two rows per complex FFT
#pragma omp parallel
        {
            std::vector<cd> in(n2), out(n2);
#pragma omp for schedule(static)
            for (long i = 0; i < nrows; i += 2) {
                const bool two = i + 1 < nrows;
                const double* x1 = r + i * (n2 + 2);
                const double* x2 = r + (i + 1) * (n2 + 2);
                for (int j = 0; j < n2; ++j) in[j] = cd(x1[j], two ? x2[j] : 0.0);
                fft_rec(in.data(), 1, out.data(), n2, T2, -1);
                for (int k = 0; k < nzc; ++k) {
                    const cd zk = out[k], zm = std::conj(out[(n2 - k) % n2]);
                    c[i * nzc + k] = 0.5 * (zk + zm);
                    if (two) c[(i + 1) * nzc + k] = cd(0.0, -0.5) * (zk - zm);
                }
            }
        }
        for (int i = 0; i < n0; ++i) fft_lines(c + long(i) * n1 * nzc, n1, nzc, nzc, 1, -1);
        fft_lines(c, n0, long(n1) * nzc, long(n1) * nzc, 1, -1);
    } else {              // half-complex -> real, in place
        fft_lines(c, n0, long(n1) * nzc, long(n1) * nzc, 1, +1);
        for (int i = 0; i < n0; ++i) fft_lines(c + long(i) * n1 * nzc, n1, nzc, nzc, 1, +1);
#pragma omp parallel
        {
            std::vector<cd> in(n2), out(n2);
#pragma omp for schedule(static)
            for (long i = 0; i < nrows; i += 2) {
                const bool two = i + 1 < nrows;
                // Z = X1 + i X2, Hermitian-extended; Im of the self-conjugate modes is ignored (FFTW c2r)
                for (int k = 0; k < nzc; ++k) {
                    cd a = c[i * nzc + k], b = two ? c[(i + 1) * nzc + k] : cd(0.0, 0.0);
                    if (k == 0 || (n2 % 2 == 0 && k == n2 / 2)) { a = cd(a.real(), 0.0); b = cd(b.real(), 0.0); }
                    in[k] = a + cd(0.0, 1.0) * b;
                    if (k > 0 && k < n2 - k) in[n2 - k] = std::conj(a) + cd(0.0, 1.0) * std::conj(b);
                }
                fft_rec(in.data(), 1, out.data(), n2, T2, +1);
                double* x1 = r + i * (n2 + 2);
                double* x2 = r + (i + 1) * (n2 + 2);
                for (int j = 0; j < n2; ++j) { x1[j] = out[j].real(); if (two) x2[j] = out[j].imag(); }
            }
        }
    }
}
}  // namespace

extern "C" {
void* fftw_malloc(size_t n) { return aligned_alloc(64, (n + 63) / 64 * 64); }
void fftw_free(void* p) { free(p); }
static fftw_plan mk(int kind, int n0, int n1, int n2, int sign, void* d) { return new fftw_plan_s{kind, n0, n1, n2, sign, d}; }
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double* in, fftw_complex*, unsigned) { return mk(0, n0, n1, n2, -1, in); }
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex* in, double*, unsigned) { return mk(1, n0, n1, n2, +1, in); }
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex* in, fftw_complex*, int sign, unsigned) { return mk(2, n0, n1, n2, sign, in); }
void fftw_destroy_plan(fftw_plan p) { delete p; }
int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int) {}
void fftw_cleanup_threads(void) {}
// [a][b][m] <-> [b][a][m] for the first two axes of a complex array
static void swap01(cd* c, int a, int b, int m) {
    std::vector<cd> t(size_t(a) * b * m);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < a; ++i)
        for (int j = 0; j < b; ++j) std::memcpy(&t[(size_t(j) * a + i) * m], &c[(size_t(i) * b + j) * m], sizeof(cd) * m);
    std::memcpy(c, t.data(), sizeof(cd) * t.size());
}
void fftw_execute(const fftw_plan p) {
    if (!p->transposed) { execute(p); return; }
    const int m = (p->kind == 2) ? p->n2 : p->n2 / 2 + 1;
    cd* c = reinterpret_cast<cd*>(p->data);
    const bool to_k = (p->kind == 0) || (p->kind == 2 && p->sign == -1);
    if (to_k) { execute(p); swap01(c, p->n0, p->n1, m); }      // natural in, [k1][k0][k2] out
    else { swap01(c, p->n1, p->n0, m); execute(p); }           // [k1][k0][k2] in, natural out
}
// FFTW-MPI on ONE rank (the reference's USE_MPI build, src/grid_fft.cc:120-140): the whole grid is the local slab; the integer
// communicator argument is ignored.  Declared in fftw3-mpi.h of this directory.
void fftw_mpi_init(void) {}
ptrdiff_t fftw_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int, ptrdiff_t* l0, ptrdiff_t* s0, ptrdiff_t* l1, ptrdiff_t* s1) {
    *l0 = n0; *s0 = 0; *l1 = n1; *s1 = 0;
    return n0 * n1 * n2;
}
static fftw_plan mkt(int kind, ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, int sign, void* d, bool transposed) {
    fftw_plan p = mk(kind, int(n0), int(n1), int(n2), sign, d);
    p->transposed = transposed ? 1 : 0;
    return p;
}
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double* in, fftw_complex*, int, unsigned flags) {
    return mkt(0, n0, n1, n2, -1, in, flags & (1u << 30));
}
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, double*, int, unsigned flags) {
    return mkt(1, n0, n1, n2, +1, in, flags & (1u << 29));
}
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, fftw_complex*, int, int sign, unsigned flags) {
    return mkt(2, n0, n1, n2, sign, in, flags & ((1u << 29) | (1u << 30)));
}
}   // extern "C"
