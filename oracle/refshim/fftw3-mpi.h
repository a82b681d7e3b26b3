// oracle/refshim/fftw3-mpi.h - TEST INFRASTRUCTURE.  One-rank stand-in for the FFTW-MPI calls of the reference's Grid_FFT
// (src/grid_fft.cc:120-140): the whole grid is the local slab, FFTW_MPI_TRANSPOSED_OUT / _IN keep their meaning (k-space
// is stored [ky][kx][kz]).  Implemented in fftw_shim.cc.
#pragma once
#include <mpi.h>

#include "fftw3.h"
extern "C" {
enum { FFTW_MPI_TRANSPOSED_IN = 1u << 29, FFTW_MPI_TRANSPOSED_OUT = 1u << 30 };
void fftw_mpi_init(void);
ptrdiff_t fftw_mpi_local_size_3d_transposed(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, MPI_Comm comm, ptrdiff_t* local_n0,
                                            ptrdiff_t* local_0_start, ptrdiff_t* local_n1, ptrdiff_t* local_1_start);
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double* in, fftw_complex* out, MPI_Comm comm, unsigned flags);
fftw_plan fftw_mpi_plan_dft_c2r_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, double* out, MPI_Comm comm, unsigned flags);
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex* in, fftw_complex* out, MPI_Comm comm, int sign,
                               unsigned flags);
}
