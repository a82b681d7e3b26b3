// tests/cpp/gslmini.cc - TEST INFRASTRUCTURE: small functional stand-ins for the GSL entry points the reference's HOST code
// calls (cosmology calculator: natural cubic splines and adaptive quadrature; NGenIC plugin: a uniform generator), so that the
// reference's unmodified executable can be LINKED and RUN against the drop-in headers in a container without GSL
// (tests/test_reference_driver.py).  Not GSL: the quadrature is a plain adaptive Gauss-Kronrod(7,15) bisection driven to
// convergence (GSL's QAG with the 61-point rule stops at the requested 1e-10; the two agree to 3.5e-12 on the sigma_8 integral), the
// splines are natural cubic splines, and "ranlxd1" is Luescher's RANLUX written from the published algorithm - that it delivers
// GSL's sequence is checked against the reference's golden files, whose NGenIC white noise it reproduces.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <vector>

#include <gsl/gsl_errno.h>
#include <gsl/gsl_integration.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_spline.h>

extern "C" {
// ------------------------------------------------------------------ splines
struct gsl_interp_type_s { int kind; };
struct gsl_interp_accel_s { size_t hint; };
struct gsl_spline_s { const gsl_interp_type* type; std::vector<double> x, y, y2; };
static const gsl_interp_type k_cspline{0}, k_cspline_p{1}, k_linear{2};
const gsl_interp_type* gsl_interp_cspline = &k_cspline;
const gsl_interp_type* gsl_interp_cspline_periodic = &k_cspline_p;
const gsl_interp_type* gsl_interp_linear = &k_linear;

gsl_interp_accel* gsl_interp_accel_alloc(void) { return new gsl_interp_accel_s{0}; }
void gsl_interp_accel_free(gsl_interp_accel* a) { delete a; }
gsl_spline* gsl_spline_alloc(const gsl_interp_type* T, size_t size) {
    auto* s = new gsl_spline_s();
    s->type = T; s->x.reserve(size);
    return s;
}
int gsl_spline_init(gsl_spline* s, const double xa[], const double ya[], size_t n) {
    s->x.assign(xa, xa + n); s->y.assign(ya, ya + n); s->y2.assign(n, 0.0);
    if (s->type->kind == 2 || n < 3) return GSL_SUCCESS;
    // natural cubic spline: tridiagonal system for the second derivatives
    std::vector<double> u(n, 0.0);
    for (size_t i = 1; i + 1 < n; ++i) {
        const double sig = (s->x[i] - s->x[i - 1]) / (s->x[i + 1] - s->x[i - 1]);
        const double p = sig * s->y2[i - 1] + 2.0;
        s->y2[i] = (sig - 1.0) / p;
        const double d = (s->y[i + 1] - s->y[i]) / (s->x[i + 1] - s->x[i]) - (s->y[i] - s->y[i - 1]) / (s->x[i] - s->x[i - 1]);
        u[i] = (6.0 * d / (s->x[i + 1] - s->x[i - 1]) - sig * u[i - 1]) / p;
    }
    for (size_t k = n - 1; k-- > 0;) s->y2[k] = s->y2[k] * s->y2[k + 1] + u[k];
    return GSL_SUCCESS;
}
void gsl_spline_free(gsl_spline* s) { delete s; }
static size_t interval(const gsl_spline* s, double x) {
    const size_t n = s->x.size();
    size_t hi = size_t(std::upper_bound(s->x.begin(), s->x.end(), x) - s->x.begin());
    if (hi < 1) hi = 1;
    if (hi > n - 1) hi = n - 1;
    return hi - 1;
}
double gsl_spline_eval(const gsl_spline* s, double x, gsl_interp_accel*) {
    const size_t i = interval(s, x);
    const double h = s->x[i + 1] - s->x[i], a = (s->x[i + 1] - x) / h, b = (x - s->x[i]) / h;
    return a * s->y[i] + b * s->y[i + 1] + ((a * a * a - a) * s->y2[i] + (b * b * b - b) * s->y2[i + 1]) * h * h / 6.0;
}
double gsl_spline_eval_deriv(const gsl_spline* s, double x, gsl_interp_accel*) {
    const size_t i = interval(s, x);
    const double h = s->x[i + 1] - s->x[i], a = (s->x[i + 1] - x) / h, b = (x - s->x[i]) / h;
    return (s->y[i + 1] - s->y[i]) / h - (3.0 * a * a - 1.0) / 6.0 * h * s->y2[i] + (3.0 * b * b - 1.0) / 6.0 * h * s->y2[i + 1];
}

// ------------------------------------------------------------------ adaptive quadrature
struct gsl_integration_workspace_s { size_t limit; };
gsl_integration_workspace* gsl_integration_workspace_alloc(const size_t n) { return new gsl_integration_workspace_s{n}; }
void gsl_integration_workspace_free(gsl_integration_workspace* w) { delete w; }
namespace {
struct Seg { double a, b, val, err; bool operator<(const Seg& o) const { return err < o.err; } };
Seg gk15(const gsl_function* f, double a, double b) {
    static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851, 0.864864423359769072789712788640926,
                                  0.741531185599394439863864773280788, 0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                  0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
    static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204, 0.104790010322250183839876322541518,
                                  0.140653259715525918745189590510238, 0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                  0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
    static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780, 0.381830050505118944950369775488975,
                                 0.417959183673469387755102040816327};
    const double c = 0.5 * (a + b), h = 0.5 * (b - a);
    const double fc = f->function(c, f->params);
    double rk = wgk[7] * fc, rg = wg[3] * fc;
    for (int j = 0; j < 7; ++j) {
        const double dx = h * xgk[j];
        const double s = f->function(c - dx, f->params) + f->function(c + dx, f->params);
        rk += wgk[j] * s;
        if (j % 2 == 1) rg += wg[j / 2] * s;
    }
    return Seg{a, b, rk * h, std::fabs((rk - rg) * h)};
}
}  // namespace
int gsl_integration_qag(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit, int, gsl_integration_workspace*,
                        double* result, double* abserr) {
    std::priority_queue<Seg> q;
    Seg s0 = gk15(f, a, b);
    double val = s0.val, err = s0.err;
    q.push(s0);
    epsrel = std::min(epsrel, 1e-13);   // converge well past the request: the reference's numbers were made by GSL's 61-point QAG, whose
                                        // true error is far below its own estimate - only a converged value can agree with it to 1e-12
    for (size_t it = 1; it < std::max<size_t>(limit, 1) && err > std::max(epsabs, epsrel * std::fabs(val)); ++it) {
        const Seg s = q.top(); q.pop();
        const double m = 0.5 * (s.a + s.b);
        const Seg l = gk15(f, s.a, m), r = gk15(f, m, s.b);
        val += l.val + r.val - s.val; err += l.err + r.err - s.err;
        q.push(l); q.push(r);
    }
    *result = val; *abserr = err;
    return GSL_SUCCESS;
}
int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit, gsl_integration_workspace* w,
                         double* result, double* abserr) {
    return gsl_integration_qag(f, a, b, epsabs, epsrel, limit, 1, w, result, abserr);
}

// ------------------------------------------------------------------ uniform generators
// "ranlxd1": Luescher's RANLUX in double precision, luxury level 1 (p = 202) - the published algorithm (M. Luescher, Comput. Phys.
// Commun. 79 (1994) 100; F. James, ibid. 97 (1996) 357) in the state layout GSL documents for gsl_rng_ranlxd1: twelve 48-bit
// numbers, subtract-with-borrow x_n = x_{n-7} - x_{n-12} - c, seed bits from a 31-bit shift register.  Whether it reproduces GSL bit
// for bit is CHECKED, not assumed: tests/test_reference_driver.py compares the phases of the NGenIC white noise it produces
// through the reference's plugin with the reference's own golden file.  Any other type: xoshiro256**.
struct gsl_rng_type_s { int id; };
struct gsl_rng_s {
    int type;
    mutable uint64_t s[4];
    mutable double x[12], carry;
    mutable unsigned ir, jr, ir_old, pr;
};
static const gsl_rng_type k_ranlxd1{0}, k_mt{1};
const gsl_rng_type* gsl_rng_ranlxd1 = &k_ranlxd1;
const gsl_rng_type* gsl_rng_mt19937 = &k_mt;
namespace {
const double kOneBit = 1.0 / 281474976710656.0;   // 2^-48
inline unsigned nxt(unsigned i) { return i == 11 ? 0 : i + 1; }
void lux_step(const gsl_rng* r) {   // one subtract-with-borrow update at (ir, jr)
    double y = r->x[r->jr] - r->x[r->ir] - r->carry;
    if (y < 0) { r->carry = kOneBit; y += 1.0; } else r->carry = 0.0;
    r->x[r->ir] = y;
    r->ir = nxt(r->ir); r->jr = nxt(r->jr);
}
void lux_increment(const gsl_rng* r) {   // p updates between blocks of twelve delivered numbers
    for (unsigned k = 0; k < r->pr; ++k) lux_step(r);
    r->ir_old = r->ir;
}
void lux_seed(const gsl_rng* r, unsigned long s) {
    if (s == 0) s = 1;
    long i = long(s & 0x7FFFFFFFUL);
    int xbit[31];
    for (int k = 0; k < 31; ++k) { xbit[k] = int(i % 2); i /= 2; }
    int ibit = 0, jbit = 18;
    for (int k = 0; k < 12; ++k) {
        double x = 0;
        for (int l = 1; l <= 48; ++l) {
            const double y = double((xbit[ibit] + 1) % 2);
            x += x + y;
            xbit[ibit] = (xbit[ibit] + xbit[jbit]) % 2;
            ibit = (ibit + 1) % 31; jbit = (jbit + 1) % 31;
        }
        r->x[k] = kOneBit * x;
    }
    r->carry = 0; r->ir = 11; r->jr = 7; r->ir_old = 0; r->pr = 202;
}
double lux_get(const gsl_rng* r) {
    r->ir = nxt(r->ir);
    if (r->ir == r->ir_old) lux_increment(r);
    return r->x[r->ir];
}
}  // namespace
void gsl_rng_set(const gsl_rng* r, unsigned long int seed) {
    if (r->type == 0) { lux_seed(r, seed); return; }
    uint64_t z = seed + 0x9e3779b97f4a7c15ull;
    for (int i = 0; i < 4; ++i) {   // splitmix64 seeding
        z += 0x9e3779b97f4a7c15ull;
        uint64_t x = z;
        x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
        r->s[i] = x ^ (x >> 31);
    }
}
gsl_rng* gsl_rng_alloc(const gsl_rng_type* T) { auto* r = new gsl_rng_s(); r->type = T->id; gsl_rng_set(r, 0); return r; }
void gsl_rng_free(gsl_rng* r) { delete r; }
unsigned long int gsl_rng_get(const gsl_rng* r) {
    if (r->type == 0) return (unsigned long)(lux_get(r) * 4294967296.0);   // ranlxd delivers doubles; integers are their top 32 bits
    auto rotl = [](uint64_t x, int k) { return (x << k) | (x >> (64 - k)); };
    uint64_t* s = r->s;
    const uint64_t res = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return res;
}
double gsl_rng_uniform(const gsl_rng* r) {
    if (r->type == 0) return lux_get(r);
    return double(gsl_rng_get(r) >> 11) * (1.0 / 9007199254740992.0);
}
double gsl_rng_uniform_pos(const gsl_rng* r) { double x; do x = gsl_rng_uniform(r); while (x == 0.0); return x; }

// Gaussian deviates (MUSIC1 white-noise generator): Leva's ratio-of-uniforms method
double gsl_ran_ugaussian_ratio_method(const gsl_rng* r) {
    for (;;) {
        const double u = 1.0 - gsl_rng_uniform(r), v = 1.7156 * (gsl_rng_uniform(r) - 0.5);
        const double x = u - 0.449871, y = std::fabs(v) + 0.386595, q = x * x + y * (0.19600 * y - 0.25472 * x);
        if (q < 0.27597 || (q <= 0.27846 && v * v <= -4.0 * u * u * std::log(u))) return v / u;
    }
}
double gsl_ran_ugaussian(const gsl_rng* r) { return gsl_ran_ugaussian_ratio_method(r); }
double gsl_ran_gaussian(const gsl_rng* r, const double sigma) { return sigma * gsl_ran_ugaussian_ratio_method(r); }
double gsl_ran_flat(const gsl_rng* r, const double a, const double b) { return a + (b - a) * gsl_rng_uniform(r); }

// ------------------------------------------------------------------ error handling
gsl_error_handler_t* gsl_set_error_handler_off(void) { return nullptr; }
gsl_error_handler_t* gsl_set_error_handler(gsl_error_handler_t*) { return nullptr; }
const char* gsl_strerror(const int) { return "gslmini"; }

// ------------------------------------------------------------------ strings src/main.cc prints
extern const char fftw_version[];
const char fftw_version[] = "none (libmfb200 drop-in)";
}
// build-system generated in the reference (CMake configure_file of version.cc)
const char* GIT_TAG = "dropin";
const char* GIT_REV = "dropin";
const char* GIT_BRANCH = "dropin";
