#pragma once
extern "C" {
typedef struct gsl_rng_type_s gsl_rng_type;
typedef struct gsl_rng_s gsl_rng;
extern const gsl_rng_type* gsl_rng_ranlxd1;
extern const gsl_rng_type* gsl_rng_mt19937;
gsl_rng* gsl_rng_alloc(const gsl_rng_type* T);
void gsl_rng_free(gsl_rng* r);
void gsl_rng_set(const gsl_rng* r, unsigned long int seed);
double gsl_rng_uniform(const gsl_rng* r);
double gsl_rng_uniform_pos(const gsl_rng* r);
unsigned long int gsl_rng_get(const gsl_rng* r);
}
