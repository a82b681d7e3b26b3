// include/mfb/operators.hh - output operators of the reference (include/operators.hh:25-35) as tagged
// descriptors: the accumulate happens inside the last FFT pass on the device instead of a host lambda.
#pragma once
#include "grid_fft.hh"

namespace op {
template <typename field> struct out_op { field* g; int code; double alpha; };
template <typename field> inline out_op<field> assign_to(field& g) { return {&g, MFB_OP_ASSIGN, 0.0}; }
template <typename field, typename val> inline out_op<field> multiply_add_to(field& g, val x) { return {&g, MFB_OP_MULADD, double(x)}; }
template <typename field> inline out_op<field> add_to(field& g) { return {&g, MFB_OP_ADD, 0.0}; }
template <typename field> inline out_op<field> subtract_from(field& g) { return {&g, MFB_OP_SUB, 0.0}; }
}  // namespace op
