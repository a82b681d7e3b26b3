// oracle/refshim/fftw3.h - TEST INFRASTRUCTURE.  The subset of the FFTW3 API that the reference's Grid_FFT uses
// (src/grid_fft.cc:50-62,222,256; include/grid_fft.hh:97-99), declared here and implemented in fftw_shim.cc as an exact
// O(n^2)-per-line DFT, so that the reference's OWN grid / convolver sources can be compiled and run in a container that has
// no FFTW.  Same conventions as FFTW: unnormalised, forward sign -1, in-place r2c/c2r with the padded last dimension,
// c2r ignores the imaginary parts it cannot represent.  Only meant for the 32^3 / 48^3 grids of the parity fixtures.
#pragma once
#include <cstddef>
extern "C" {
typedef double fftw_complex[2];
struct fftw_plan_s;
typedef fftw_plan_s* fftw_plan;
enum { FFTW_FORWARD = -1, FFTW_BACKWARD = +1 };
enum { FFTW_MEASURE = 0, FFTW_PATIENT = 1u << 5, FFTW_ESTIMATE = 1u << 6 };
void* fftw_malloc(size_t n);
void fftw_free(void* p);
fftw_plan fftw_plan_dft_r2c_3d(int n0, int n1, int n2, double* in, fftw_complex* out, unsigned flags);
fftw_plan fftw_plan_dft_c2r_3d(int n0, int n1, int n2, fftw_complex* in, double* out, unsigned flags);
fftw_plan fftw_plan_dft_3d(int n0, int n1, int n2, fftw_complex* in, fftw_complex* out, int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
extern const char fftw_version[];   /* printed by src/main.cc:209 */
int fftw_init_threads(void);
void fftw_plan_with_nthreads(int);
void fftw_cleanup_threads(void);
}
