#pragma once
#include <cstddef>
extern "C" {
typedef struct { double (*function)(double x, void* params); void* params; } gsl_function;
typedef struct gsl_integration_workspace_s gsl_integration_workspace;
gsl_integration_workspace* gsl_integration_workspace_alloc(const size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace* w);
enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3, GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };
int gsl_integration_qag(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit, int key,
                        gsl_integration_workspace* workspace, double* result, double* abserr);
int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace* workspace, double* result, double* abserr);
}
