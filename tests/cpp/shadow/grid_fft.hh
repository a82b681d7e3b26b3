// tests/cpp/shadow - the reference header name resolved to the GPU-backed drop-in (tests/test_plugin_dropin.py)
#include <mfb/grid_fft.hh>
