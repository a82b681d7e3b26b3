// ranlux.cuh - Luescher's RANLUX in double precision, luxury level 1 ("ranlxd1": p = 202), host and device.
// M. Luescher, Comput. Phys. Commun. 79 (1994) 100; F. James, ibid. 97 (1996) 357: twelve 48-bit numbers in a ring,
// x_n = x_{n-5} - x_{n-12} - c (mod 1), p updates for every twelve delivered numbers, the seed spread over the state by a 31-bit
// shift register.  This is the generator behind the reference's NGenIC noise plugin (src/plugins/random_ngenic.cc:44-45, GSL's
// gsl_rng_ranlxd1): with it the device fill below draws the reference's own white noise, which is checked against the
// reference's golden files (tests/test_oracle_golden.py: the phases of the golden density field).
#pragma once
#include "mfb_platform.h"

namespace mfb {

struct Ranlxd1 {
    double x[12];
    double carry;
    int ir, jr, ir_old;

    MFB_HD void seed(unsigned long s) {
        if (s == 0) s = 1;
        long v = long(s & 0x7FFFFFFFUL);
        int bit[31];
        for (int k = 0; k < 31; ++k) { bit[k] = int(v % 2); v /= 2; }
        int ib = 0, jb = 18;
        for (int k = 0; k < 12; ++k) {
            double acc = 0.0;
            for (int l = 0; l < 48; ++l) {
                acc += acc + double((bit[ib] + 1) % 2);
                bit[ib] = (bit[ib] + bit[jb]) % 2;
                ib = (ib + 1) % 31; jb = (jb + 1) % 31;
            }
            x[k] = acc * (1.0 / 281474976710656.0);   // 2^-48
        }
        carry = 0.0; ir = 11; jr = 7; ir_old = 0;
    }
    MFB_HD double uniform() {   // in [0, 1), a multiple of 2^-48
        ir = (ir == 11) ? 0 : ir + 1;
        if (ir == ir_old) {
            for (int k = 0; k < 202; ++k) {
                double y = x[jr] - x[ir] - carry;
                if (y < 0.0) { carry = 1.0 / 281474976710656.0; y += 1.0; } else carry = 0.0;
                x[ir] = y;
                ir = (ir == 11) ? 0 : ir + 1;
                jr = (jr == 11) ? 0 : jr + 1;
            }
            ir_old = ir;
        }
        return x[ir];
    }
};

}  // namespace mfb
