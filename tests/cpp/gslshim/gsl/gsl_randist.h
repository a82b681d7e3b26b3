#pragma once
#include "gsl_rng.h"
extern "C" {
double gsl_ran_flat(const gsl_rng* r, const double a, const double b);
double gsl_ran_gaussian(const gsl_rng* r, const double sigma);
double gsl_ran_ugaussian(const gsl_rng* r);
double gsl_ran_ugaussian_ratio_method(const gsl_rng* r);
}
