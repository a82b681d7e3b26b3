#pragma once
#include <cstddef>
extern "C" {
typedef struct gsl_interp_type_s gsl_interp_type;
typedef struct gsl_interp_accel_s gsl_interp_accel;
typedef struct gsl_spline_s gsl_spline;
extern const gsl_interp_type* gsl_interp_cspline;
extern const gsl_interp_type* gsl_interp_cspline_periodic;
extern const gsl_interp_type* gsl_interp_linear;
gsl_interp_accel* gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel* a);
gsl_spline* gsl_spline_alloc(const gsl_interp_type* T, size_t size);
int gsl_spline_init(gsl_spline* spline, const double xa[], const double ya[], size_t size);
void gsl_spline_free(gsl_spline* spline);
double gsl_spline_eval(const gsl_spline* spline, double x, gsl_interp_accel* a);
double gsl_spline_eval_deriv(const gsl_spline* spline, double x, gsl_interp_accel* a);
}
