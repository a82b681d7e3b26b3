// fft_kernels.cuh - the FFT pass kernels of the LPT pipeline (FP64, sm_100a).
//
//   strided_pass_kernel : complex FFT along a strided axis (x or y) for a tile of T adjacent-k_z
//                         lines; optional spectral zero-padding of the input (N -> 3N/2), spectral
//                         truncation of the output (3N/2 -> N), k-space operand factor at load
//                         (Hessian / gradient of the potential), scale + output operator
//                         (=, +=, -=, += alpha*) at store.          [replaces fftw_execute x/y sweeps,
//                         OrszagConvolver::pad_insert/unpad, op::assign_to & friends]
//   z_c2r_kernel/z_r2c_kernel : the contiguous-axis real<->half-complex pass, two real rows per
//                         complex FFT, in place in the reference's padded-row layout
//                         (src/grid_fft.cc:45-54,73-74).
//   (the fused z passes of the dealiased products live in fft_kernels_z2.cuh / fft_kernels_zm.cuh)
#pragma once
#include "fft_core.cuh"

namespace mfb {

constexpr int MFB_MAX_RANKS = 8;

enum OutOp { OP_ASSIGN = 0, OP_ADD = 1, OP_SUB = 2, OP_MULADD = 3 };

// k-space factor applied to an operand element at integer wave index (ix,iy,iz):
//   f = c1*K[a1]*K[b1] + c2*K[a2]*K[b2],   K[d] = kf[d] * (i_d - n[d]*[i_d > n[d]/2]), 0 at i_d == n[d]/2
// axis index -1 means "1".  imag: multiply by i on top (single gradient).
struct OperandFactor {
    int enabled;
    int a1, b1, a2, b2;
    double c1, c2;
    int imag;
    int n[3];      // grid size per axis (wave-number folding)
    double kf[3];  // 2 pi / L per axis
    int keep_nyq;  // 1: K at the Nyquist index is +n/2*kf (get_k, include/grid_fft.hh:282-296) instead of 0 (gradient, :376-385);
                   //    only the un-padded NaiveConvolver ever feeds a Nyquist mode through a factor
};

struct StridedArgs {
    const double2* in;
    double2* out;
    ptrdiff_t in_ps, in_as;    // element strides of the position axis / the "a" (other) axis
    ptrdiff_t out_ps, out_as;
    int na;                    // number of a-indices
    int nz;                    // number of contiguous lines per a (tiles cover [0,nz))
    int nz_fft;                // lines >= nz_fft are not transformed (only zero-filled on assign)
    int n_in, n_out;           // logical input / output lengths (<= n): spectral pad / truncate
    int inverse;               // backward transform (conj at load and store)
    double scale;              // applied at store
    OperandFactor f;
    int pos_axis, a_axis;      // which grid axis (0=x,1=y) the position / a index is (for f)
    int op;                    // OutOp
    double alpha;
    int a_skip;                // (global) a-index whose lines are identically zero (Nyquist), -1: none
    int zero_fill;             // on OP_ASSIGN also write zeros to skipped lines / Nyquist outputs
    // ---- slab-decomposed (multi-GPU) passes
    int a_off;                 // global index of local a == 0 (wave numbers, a_skip)
    // z-blocked layouts of exchange buffers: line iz of a-index a sits at a*as + (iz / zb)*zbs + iz % zb instead of a*as + iz
    // (zb == 0: plain).  With zb = the producer's tile height and position stride ps = zb, the segments a tile writes for
    // consecutive positions are adjacent in memory: peer stores of a routed pass become long contiguous runs.
    int in_zb, out_zb;
    ptrdiff_t in_zbs, out_zbs;
    int out_blk;               // > 0: output position k goes to rank k / out_blk, at local position k % out_blk:
    double2* out_peer[MFB_MAX_RANKS];  //   out_peer[rank] + a*out_as + (k % out_blk)*out_ps + iz  (peer memory over NVLink,
                               //   or a staging chunk when the exchange is done by a host collective)
};

MFB_HD ptrdiff_t zline_off(int iz, int zb, ptrdiff_t zbs) { return zb > 0 ? ptrdiff_t(iz / zb) * zbs + (iz % zb) : ptrdiff_t(iz); }

MFB_HD double wave_k(int i, int n, double kfac, int keep_nyq = 0) {
    return (i == n / 2 && !keep_nyq) ? 0.0 : kfac * double(i - ((i > n / 2) ? n : 0));
}

template <class P>
__global__ void __launch_bounds__(P::NTHREADS) strided_pass_kernel(StridedArgs A, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N;
    constexpr int T = P::T;
    const int tid = threadIdx.x;
    const int line = tid % T;
    const int q = tid / T;
    const int tiles_per_a = (A.nz + T - 1) / T;
    const int a = blockIdx.x / tiles_per_a;
    const int z0 = (blockIdx.x - a * tiles_per_a) * T;
    const int iz = z0 + line;
    const bool line_valid = iz < A.nz;
    const ptrdiff_t ooff = ptrdiff_t(a) * A.out_as + zline_off(iz, A.out_zb, A.out_zbs);
    // address of output position k (routed to the owning rank's buffer for slab-decomposed passes)
    auto optr = [&](int k) -> double2* {
        if (A.out_blk > 0) {
            const int b = k / A.out_blk;
            return A.out_peer[b] + ooff + ptrdiff_t(k - b * A.out_blk) * A.out_ps;
        }
        return A.out + ooff + ptrdiff_t(k) * A.out_ps;
    };

    // ---- lines that carry no data: optional zero fill, no transform
    const bool skip_tile = (a + A.a_off == A.a_skip) || (z0 >= A.nz_fft);
    if (skip_tile) {
        if (A.zero_fill && A.op == OP_ASSIGN && line_valid) {
            for (int k = q; k < A.n_out; k += P::TPL) mfb_st_stream(optr(k), make_double2(0.0, 0.0));
        }
        return;
    }
    const bool fft_line = line_valid && iz < A.nz_fft;

    // ---- per-thread operand factor polynomial in K_pos:  f = q0 + q1*Kp + q2*Kp^2
    double q0 = 1.0, q1 = 0.0, q2 = 0.0;
    if (A.f.enabled) {
        const double Ka = wave_k(a + A.a_off, A.f.n[A.a_axis], A.f.kf[A.a_axis], A.f.keep_nyq);
        const double Kz = wave_k(iz, A.f.n[2], A.f.kf[2], A.f.keep_nyq);
        q0 = 0.0;
        auto term = [&](int d1, int d2, double c) {
            if (c == 0.0) return;
            int np = 0;
            double rest = c;
            int dd[2] = {d1, d2};
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int d = dd[t];
                if (d < 0) continue;
                if (d == A.pos_axis) ++np;
                else if (d == A.a_axis) rest *= Ka;
                else rest *= Kz;
            }
            if (np == 0) q0 += rest;
            else if (np == 1) q1 += rest;
            else q2 += rest;
        };
        term(A.f.a1, A.f.b1, A.f.c1);
        term(A.f.a2, A.f.b2, A.f.c2);
    }

    const double2* ibase = A.in + ptrdiff_t(a) * A.in_as + zline_off(iz, A.in_zb, A.in_zbs);
    const int hin = A.n_in / 2;
    const int hout = A.n_out / 2;

    auto loader = [&](int pos, int /*line*/) -> double2 {
        if (!fft_line) return make_double2(0.0, 0.0);
        int src = pos;
        if (A.n_in != N) {
            if (pos < hin) src = pos;
            else if (pos > N - hin) src = pos - (N - A.n_in);
            else return make_double2(0.0, 0.0);
        }
        double2 v = mfb_ld_stream(ibase + ptrdiff_t(src) * A.in_ps);
        if (A.f.enabled) {
            const double Kp = wave_k(src, A.f.n[A.pos_axis], A.f.kf[A.pos_axis], A.f.keep_nyq);
            const double fac = fma(fma(q2, Kp, q1), Kp, q0);
            v = A.f.imag ? make_double2(-v.y * fac, v.x * fac) : make_double2(v.x * fac, v.y * fac);
        }
        if (A.inverse) v.y = -v.y;
        return v;
    };
    auto storer = [&](int k, int /*line*/, double2 v) {
        if (!fft_line) return;
        int dst = k;
        if (A.n_out != N) {
            if (k < hout) dst = k;
            else if (k > N - hout) dst = k - (N - A.n_out);
            else return;
        }
        v = make_double2(v.x * A.scale, (A.inverse ? -v.y : v.y) * A.scale);
        double2* p = optr(dst);
        if (A.op != OP_ASSIGN) {
            const double2 o = mfb_ld_stream(p);
            const double al = (A.op == OP_ADD) ? 1.0 : ((A.op == OP_SUB) ? -1.0 : A.alpha);
            v = make_double2(fma(al, v.x, o.x), fma(al, v.y, o.y));
        }
        mfb_st_stream(p, v);
    };

    fft_tile<P, false, true, true>(sm, tw, line, q, true, loader, storer);

    // Nyquist output of a truncated transform is defined to be zero (include/convolution.hh:641-643)
    if (A.zero_fill && A.op == OP_ASSIGN && A.n_out != N && q == 0 && line_valid) {
        mfb_st_stream(optr(hout), make_double2(0.0, 0.0));
    }
    // lines of a partially transformed tile beyond nz_fft
    if (A.zero_fill && A.op == OP_ASSIGN && line_valid && !fft_line) {
        for (int k = q; k < A.n_out; k += P::TPL) mfb_st_stream(optr(k), make_double2(0.0, 0.0));
    }
}

// ------------------------------------------------------------------------------------------
// Optional epilogue of the c2r pass: instead of storing the real field, read it out into a particle attribute right away -
// the lattice read-out of particle::lattice_generator for a single-shift (sc) load (include/particle_generator.hh:448-465
// positions, :544-560 velocities, :327-344 masses; "invalidates field" there too).  Saves one write and one read of the grid
// per component.  dst[stride*(slot0 + cell) + comp] = real_t( add_lattice ? ((idx_d + shift)/n_d)*lunit + f : mul*f )
struct ZReadout {
    void* dst;
    int real64;
    int stride, comp;
    long slot0;
    int add_lattice, d;
    double shift, lunit, mul;
    int n0, n1, n2, i0;   // grid extents, local_0_start_
};
MFB_D void zreadout_store(const ZReadout& R, long row, int k, double v) {
    double o;
    if (R.add_lattice) {
        const int j = int(row % R.n1), i = int(row / R.n1);
        const int id = (R.d == 0) ? (i + R.i0) : ((R.d == 1) ? j : k);
        const int nd = (R.d == 0) ? R.n0 : ((R.d == 1) ? R.n1 : R.n2);
        const double qpos = (double(id) + R.shift) / double(nd);
#ifdef MFB_EMU
        o = qpos * R.lunit + v;
#else
        o = __dadd_rn(__dmul_rn(qpos, R.lunit), v);  // no FMA contraction: bit-identical to the CPU expression
#endif
    } else {
        o = R.mul * v;
    }
    const size_t at = size_t(R.stride) * size_t(R.slot0 + row * R.n2 + k) + R.comp;
    if (R.real64) reinterpret_cast<double*>(R.dst)[at] = o;
    else reinterpret_cast<float*>(R.dst)[at] = float(o);
}

// z pass, half-complex -> real, in place.  rows: nrows rows of `pitch` doubles, each holding
// n/2+1 complex on input and n reals (+2 pad) on output.  One complex FFT serves two rows.
template <class P, bool RO = false>
__global__ void __launch_bounds__(P::NTHREADS) z_c2r_kernel(double* __restrict__ data, ptrdiff_t pitch, long nrows,
                                                            double scale, const double2* __restrict__ tw, ZReadout R = ZReadout()) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, T = P::T, NC = N / 2 + 1;
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * (2 * T);
    // pack: Z[k] = X1[k] + i X2[k], Z[n-k] = conj(X1[k]) + i conj(X2[k]); store conj(Z) (inverse trick)
    for (int idx = tid; idx < T * NC; idx += P::NTHREADS) {
        const int l = idx / NC, k = idx - l * NC;
        const long r1 = row0 + 2 * l;
        double2 x1 = make_double2(0.0, 0.0), x2 = x1;
        if (r1 < nrows) x1 = mfb_ld_stream(reinterpret_cast<const double2*>(data + r1 * pitch) + k);
        if (r1 + 1 < nrows) x2 = mfb_ld_stream(reinterpret_cast<const double2*>(data + (r1 + 1) * pitch) + k);
        if (k == 0 || k == N / 2) { x1.y = 0.0; x2.y = 0.0; }  // c2r ignores these imaginary parts
        sm[P::smem_addr(k, l)] = make_double2(x1.x - x2.y, -(x1.y + x2.x));
        if (k > 0 && k < N / 2) sm[P::smem_addr(N - k, l)] = make_double2(x1.x + x2.y, -(x2.x - x1.y));
    }
    __syncthreads();
    NoLoader nl; NoStorer ns;
    fft_tile<P, false, false, false>(sm, tw, tid % T, tid / T, true, nl, ns);
    __syncthreads();
    for (int idx = tid; idx < T * N; idx += P::NTHREADS) {
        const int l = idx / N, j = idx - l * N;
        const long r1 = row0 + 2 * l;
        const double2 z = sm[P::smem_addr(P::pos_of_freq(j), l)];
        if constexpr (RO) {
            if (r1 < nrows) zreadout_store(R, r1, j, z.x * scale);
            if (r1 + 1 < nrows) zreadout_store(R, r1 + 1, j, -z.y * scale);
        } else {
            if (r1 < nrows) data[r1 * pitch + j] = z.x * scale;
            if (r1 + 1 < nrows) data[(r1 + 1) * pitch + j] = -z.y * scale;
        }
    }
}

// z pass, real -> half-complex, in place.
template <class P>
__global__ void __launch_bounds__(P::NTHREADS) z_r2c_kernel(double* __restrict__ data, ptrdiff_t pitch, long nrows,
                                                            double scale, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, T = P::T, NC = N / 2 + 1;
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * (2 * T);
    for (int idx = tid; idx < T * N; idx += P::NTHREADS) {
        const int l = idx / N, j = idx - l * N;
        const long r1 = row0 + 2 * l;
        double2 z = make_double2(0.0, 0.0);
        if (r1 < nrows) z.x = mfb_ld_stream1(data + r1 * pitch + j);
        if (r1 + 1 < nrows) z.y = mfb_ld_stream1(data + (r1 + 1) * pitch + j);
        sm[P::smem_addr(j, l)] = z;
    }
    __syncthreads();
    NoLoader nl; NoStorer ns;
    fft_tile<P, false, false, false>(sm, tw, tid % T, tid / T, true, nl, ns);
    __syncthreads();
    for (int idx = tid; idx < T * NC; idx += P::NTHREADS) {
        const int l = idx / NC, k = idx - l * NC;
        const long r1 = row0 + 2 * l;
        const double2 zk = sm[P::smem_addr(P::pos_of_freq(k), l)];
        const double2 zm = sm[P::smem_addr(P::pos_of_freq((N - k) % N), l)];
        const double h = 0.5 * scale;
        const double2 x1 = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
        const double2 x2 = make_double2(h * (zk.y + zm.y), -h * (zk.x - zm.x));
        if (r1 < nrows) mfb_st_stream(reinterpret_cast<double2*>(data + r1 * pitch) + k, x1);
        if (r1 + 1 < nrows) mfb_st_stream(reinterpret_cast<double2*>(data + (r1 + 1) * pitch) + k, x2);
    }
}

// z pass of a complex-to-complex transform (Grid_FFT<ccomplex_t>, PPT branch src/ic_generator.cc:685-803): rows of n
// contiguous complex values, in place.  Row-major coalesced loads/stores transpose through the swizzled tile.
template <class P>
__global__ void __launch_bounds__(P::NTHREADS) z_c2c_kernel(double2* __restrict__ data, long nrows, double scale, int inverse,
                                                            const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, T = P::T;
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * T;
    for (int idx = tid; idx < T * N; idx += P::NTHREADS) {
        const int l = idx / N, j = idx - l * N;
        double2 z = make_double2(0.0, 0.0);
        if (row0 + l < nrows) z = mfb_ld_stream(data + (row0 + l) * N + j);
        if (inverse) z.y = -z.y;
        sm[P::smem_addr(j, l)] = z;
    }
    __syncthreads();
    NoLoader nl; NoStorer ns;
    fft_tile<P, false, false, false>(sm, tw, tid % T, tid / T, true, nl, ns);
    __syncthreads();
    const double sy = inverse ? -scale : scale;
    for (int idx = tid; idx < T * N; idx += P::NTHREADS) {
        const int l = idx / N, k = idx - l * N;
        const double2 z = sm[P::smem_addr(P::pos_of_freq(k), l)];
        if (row0 + l < nrows) mfb_st_stream(data + (row0 + l) * N + k, make_double2(z.x * scale, z.y * sy));
    }
}

}  // namespace mfb
