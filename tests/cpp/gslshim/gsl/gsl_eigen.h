#pragma once
#include <cstddef>
extern "C" {
typedef struct { size_t size; size_t stride; double* data; void* block; int owner; } gsl_vector;
typedef struct { size_t size1; size_t size2; size_t tda; double* data; void* block; int owner; } gsl_matrix;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct gsl_eigen_symmv_workspace_s gsl_eigen_symmv_workspace;
gsl_vector* gsl_vector_alloc(size_t n);
void gsl_vector_free(gsl_vector* v);
double gsl_vector_get(const gsl_vector* v, size_t i);
gsl_matrix* gsl_matrix_alloc(size_t n1, size_t n2);
void gsl_matrix_free(gsl_matrix* m);
double gsl_matrix_get(const gsl_matrix* m, size_t i, size_t j);
void gsl_matrix_set(gsl_matrix* m, size_t i, size_t j, double x);
gsl_matrix_view gsl_matrix_view_array(double* base, size_t n1, size_t n2);
gsl_vector_view gsl_matrix_column(gsl_matrix* m, size_t j);
gsl_eigen_symmv_workspace* gsl_eigen_symmv_alloc(size_t n);
void gsl_eigen_symmv_free(gsl_eigen_symmv_workspace* w);
int gsl_eigen_symmv(gsl_matrix* A, gsl_vector* eval, gsl_matrix* evec, gsl_eigen_symmv_workspace* w);
typedef enum { GSL_EIGEN_SORT_VAL_ASC, GSL_EIGEN_SORT_VAL_DESC, GSL_EIGEN_SORT_ABS_ASC, GSL_EIGEN_SORT_ABS_DESC } gsl_eigen_sort_t;
int gsl_eigen_symmv_sort(gsl_vector* eval, gsl_matrix* evec, gsl_eigen_sort_t sort_type);
}
