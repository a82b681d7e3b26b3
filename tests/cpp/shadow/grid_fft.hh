// tests/cpp/shadow - the reference header name resolved to the GPU-backed drop-in (tests/test_plugin_dropin.py)
#pragma once
#include <mfb/grid_fft.hh>
#include <mfb/grid_fft_complex.hh>   // Grid_FFT<ccomplex_t> of the PPT branch
