// tests/cpp/shadow - the reference header name resolved to the GPU-backed drop-in (tests/test_plugin_dropin.py).
// grid_interpolate.hh (+ grid_ghosts.hh) stay the reference's own: host-side helpers of the glass branch of ic_generator::run
// (compensation_kernel, src/ic_generator.cc:817,862,925) that only need the grid's public host interface.
#pragma once
#include <mfb/particle_generator.hh>
#include <grid_interpolate.hh>
