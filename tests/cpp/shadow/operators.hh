// tests/cpp/shadow - the reference header name resolved to the GPU-backed drop-in: the output operators are the tagged
// descriptors of <mfb/operators.hh>; op::fourier_gradient keeps the reference's interface (include/operators.hh:37-61)
#pragma once
#include <mfb/operators.hh>
