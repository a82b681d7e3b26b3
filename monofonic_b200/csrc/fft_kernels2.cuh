// fft_kernels2.cuh - second-generation strided FFT pass (FP64, sm_100a).
//
// Same contract as strided_pass_kernel (fft_kernels.cuh) - complex FFT along a strided axis for a tile of T
// adjacent-k_z lines with spectral pad / truncate, operand factor, scale and output operator fused - but with
// the per-element instruction stream cut down (ncu on the first version: 47 instructions per element visit,
// 11 of them FP64, long-scoreboard stalls on address-dependent loads; profiles/r01_ncu_full_strided_768.txt):
//   * the pad (N -> 3N/2) and truncate (3N/2 -> N) index maps are resolved at COMPILE time: with the first
//     (pad) or last (truncate) radix a multiple of 3, "which butterfly leg is in the zero band" depends only on
//     the leg number, so there are no per-element compares/branches and no loads for zero legs;
//   * no shared-memory swizzle (a quarter warp = 8 lines of one position = 128 contiguous bytes): every
//     shared-memory address is one per-butterfly base plus compile-time immediates;
//   * global addresses advance by one precomputed 64-bit stride per leg;
//   * the operand factor is a polynomial in the signed integer wave index, obtained from one int->double
//     conversion per butterfly plus compile-time offsets.
// The arithmetic (butterflies, twiddles, order of operations) is identical to the first version.
#pragma once
#include "fft_kernels.cuh"

namespace mfb {

enum StridedMode { SM_PLAIN = 0, SM_PAD = 1, SM_TRUNC = 2 };

template <class P, int MODE, int s>
MFB_D void strided2_mid_stage(double2* __restrict__ sm, const double2* __restrict__ tw, int line, int q) {
    constexpr int T = P::T;
    constexpr int r = P::radix(s), L = P::len(s), M = L / r, Bn = P::N / r;
#pragma unroll 1
    for (int b = q; b < Bn; b += P::TPL) {
        const int blk = b / M, j = b - blk * M;
        double2* sp = sm + (blk * L + j) * T + line;
        double2 v[r];
#pragma unroll
        for (int t = 0; t < r; ++t) v[t] = sp[t * M * T];
        Dft<r>::run(v);
        const double2* twp = tw + P::tw2_off(s) + j;
#pragma unroll
        for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(twp + (u - 1) * M));
#pragma unroll
        for (int u = 0; u < r; ++u) sp[u * M * T] = v[u];
    }
}
struct CtaSync {   // whole-CTA barrier (one tile per CTA)
    MFB_D void operator()() const { __syncthreads(); }
};
template <class P, int MODE, int s, class Sync>
MFB_D void strided2_mid_stages(double2* __restrict__ sm, const double2* __restrict__ tw, int line, int q, const Sync& sync) {
    if constexpr (s < P::S - 1) {
        sync();
        strided2_mid_stage<P, MODE, s>(sm, tw, line, q);
        strided2_mid_stages<P, MODE, s + 1>(sm, tw, line, q, sync);
    }
}

// One tile (fixed a, T adjacent k_z starting at z0) of a strided pass, executed by P::NTHREADS threads that share `sm` and
// synchronise with `sync`.  `after_load` is called once by every thread when stage 0 has consumed its inputs.
// SMEM_IN: the tile's input already sits in `sm` at its stage-0 leg positions (staged there asynchronously, e.g. by the
// double-buffered strided3_kernel of scripts/kbench/kbench_legacy_kernels.cuh); stage 0 then runs in place like the middle stages.
template <class P, int MODE, bool BATCH0, bool SMEM_IN = false, class Sync, class AfterLoad>
MFB_D void strided2_tile(const StridedArgs& A, const double2* __restrict__ tw, double2* __restrict__ sm, int tid, int a, int z0,
                         const Sync& sync, const AfterLoad& after_load) {
    constexpr int N = P::N, T = P::T, S = P::S;
    static_assert(S >= 2, "at least two stages");
    const int line = tid % T;
    const int q = tid / T;
    const int iz = z0 + line;
    const bool line_valid = iz < A.nz;
    const ptrdiff_t ooff = ptrdiff_t(a) * A.out_as + zline_off(iz, A.out_zb, A.out_zbs);
    const bool routed = A.out_blk > 0;
    auto optr = [&](int k) -> double2* {
        if (routed) {
            const int b = k / A.out_blk;
            return A.out_peer[b] + ooff + ptrdiff_t(k - b * A.out_blk) * A.out_ps;
        }
        return A.out + ooff + ptrdiff_t(k) * A.out_ps;
    };
    const bool zf = A.zero_fill && A.op == OP_ASSIGN;

    // ---- lines that carry no data: optional zero fill, no transform
    if ((a + A.a_off == A.a_skip) || (z0 >= A.nz_fft)) {
        if (zf && line_valid)
            for (int k = q; k < A.n_out; k += P::TPL) mfb_st_stream(optr(k), make_double2(0.0, 0.0));
        after_load();
        return;
    }
    const bool fft_line = line_valid && iz < A.nz_fft;

    // ================= stage 0: global -> butterfly -> shared
    {
        constexpr int r = P::radix(0), M = N / r;
        constexpr int K3 = (MODE == SM_PAD) ? r / 3 : 0;
        static_assert(MODE != SM_PAD || r % 3 == 0, "pad mode needs a first radix that is a multiple of 3");
        // operand factor  f = q0 + q1*m + q2*m^2  in the signed integer wave index m of the transformed axis
        double q0 = 1.0, q1 = 0.0, q2 = 0.0;
        if (A.f.enabled) {
            const double Ka = wave_k(a + A.a_off, A.f.n[A.a_axis], A.f.kf[A.a_axis]);
            const double Kz = wave_k(iz, A.f.n[2], A.f.kf[2]);
            const double kp = A.f.kf[A.pos_axis];
            q0 = 0.0;
            auto term = [&](int d1, int d2, double c) {
                if (c == 0.0) return;
                int np = 0;
                double rest = c;
                const int dd[2] = {d1, d2};
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int d = dd[t];
                    if (d < 0) continue;
                    if (d == A.pos_axis) { ++np; rest *= kp; }
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
        const double2* gin = A.in + ptrdiff_t(a) * A.in_as + zline_off(iz, A.in_zb, A.in_zbs);
        const ptrdiff_t gstep = ptrdiff_t(M) * A.in_ps;
        const bool use_f = A.f.enabled != 0, f_imag = A.f.imag != 0, inv = A.inverse != 0;
        // global loads of one butterfly (zero legs are not loaded)
        auto load = [&](int j, double2* v) {
            if constexpr (SMEM_IN) {
                const double2* sp = sm + j * T + line;
#pragma unroll
                for (int t = 0; t < r; ++t) {
                    if constexpr (MODE == SM_PAD) { if (t >= K3 && t < 2 * K3) continue; }   // zero band: never staged
                    v[t] = fft_line ? sp[t * M * T] : make_double2(0.0, 0.0);
                }
                return;
            }
            const double2* p = gin + ptrdiff_t(j) * A.in_ps;
            if constexpr (MODE == SM_PAD) {
                // source line has 2r/3 * M samples; leg k of the source maps to leg t = k (k < r/3) or k + r/3
#pragma unroll
                for (int k = 0; k < 2 * K3; ++k) {
                    const int t = (k < K3) ? k : k + K3;
                    v[t] = fft_line ? mfb_ld_stream(p + ptrdiff_t(k) * gstep) : make_double2(0.0, 0.0);
                }
            } else {
#pragma unroll
                for (int t = 0; t < r; ++t) v[t] = fft_line ? mfb_ld_stream(p + ptrdiff_t(t) * gstep) : make_double2(0.0, 0.0);
            }
        };
        auto process = [&](int j, double2* v) {
            if constexpr (MODE == SM_PAD) {
#pragma unroll
                for (int t = K3; t < 2 * K3; ++t) v[t] = make_double2(0.0, 0.0);
                if (j == 0) v[2 * K3] = make_double2(0.0, 0.0);  // the source's Nyquist sample is dropped (MPI pad semantics)
                if (use_f) {
                    const double dj = double(j);
#pragma unroll
                    for (int k = 0; k < 2 * K3; ++k) {
                        const int t = (k < K3) ? k : k + K3;
                        const double m = dj + double((k < K3) ? k * M : k * M - 2 * K3 * M);
                        const double fac = fma(fma(q2, m, q1), m, q0);
                        v[t] = f_imag ? make_double2(-v[t].y * fac, v[t].x * fac) : make_double2(v[t].x * fac, v[t].y * fac);
                    }
                }
            }
            if (inv) {
#pragma unroll
                for (int t = 0; t < r; ++t) v[t].y = -v[t].y;
            }
            Dft<r>::run(v);
#pragma unroll
            for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(tw + (u - 1) * M + j));
            double2* sp = sm + j * T + line;
#pragma unroll
            for (int u = 0; u < r; ++u) sp[u * M * T] = v[u];
        };
        constexpr int PER0 = (M + P::TPL - 1) / P::TPL;
        if constexpr (BATCH0 && PER0 == 2) {
            // both butterflies' loads are issued before any arithmetic: twice the bytes in flight per thread
            double2 va[r], vb[r];
            const int j1 = q + P::TPL;
            load(q, va);
            if (j1 < M) load(j1, vb);
            process(q, va);
            if (j1 < M) process(j1, vb);
        } else {
#pragma unroll 1
            for (int j = q; j < M; j += P::TPL) {
                double2 v[r];
                load(j, v);
                process(j, v);
            }
        }
    }
    after_load();
    // ================= middle stages: shared -> shared
    strided2_mid_stages<P, MODE, 1>(sm, tw, line, q, sync);
    sync();
    // ================= last stage: shared -> butterfly -> global
    {
        constexpr int r = P::radix(S - 1), Bn = N / r, fs = N / r;
        constexpr int K3 = (MODE == SM_TRUNC) ? r / 3 : 0;
        static_assert(MODE != SM_TRUNC || r % 3 == 0, "truncate mode needs a last radix that is a multiple of 3");
        const double sc = A.scale, scy = A.inverse ? -A.scale : A.scale;
        const int op = A.op;
        const double al = (op == OP_ADD) ? 1.0 : ((op == OP_SUB) ? -1.0 : A.alpha);
        const ptrdiff_t ostep = ptrdiff_t(fs) * A.out_ps;
        auto emit = [&](double2* p, double2 v, double2 o) {
            v = make_double2(v.x * sc, v.y * scy);
            if (op != OP_ASSIGN) v = make_double2(fma(al, v.x, o.x), fma(al, v.y, o.y));
            mfb_st_stream(p, v);
        };
#pragma unroll 1
        for (int b = q; b < Bn; b += P::TPL) {
            const double2* sp = sm + b * r * T + line;
            double2 v[r];
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = sp[t * T];
            if (!fft_line) continue;
            const int k0 = P::freq_of_pos(b * r);  // < fs; output u has frequency k0 + u*fs
            if constexpr (MODE == SM_TRUNC) {
                // keep u < r/3 (dst = k) and u >= 2r/3 (dst = k - N/3); dst == n_out/2 is defined to be zero
                double2* p[2 * K3];
                double2 o[2 * K3];
#pragma unroll
                for (int m = 0; m < 2 * K3; ++m)
                    p[m] = routed ? optr(k0 + m * fs) : A.out + ooff + ptrdiff_t(k0) * A.out_ps + ptrdiff_t(m) * ostep;
                if (op != OP_ASSIGN) {   // read-modify-write: all old values are requested before the butterfly
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) o[m] = mfb_ld_stream(p[m]);
                }
                Dft<r>::run(v);
#pragma unroll
                for (int m = 0; m < 2 * K3; ++m) {
                    const int u = (m < K3) ? m : m + K3;
                    if (m == K3 && k0 == 0) {
                        if (zf) mfb_st_stream(p[m], make_double2(0.0, 0.0));
                    } else {
                        emit(p[m], v[u], o[m]);
                    }
                }
            } else {
                double2* p[r];
                double2 o[r];
#pragma unroll
                for (int u = 0; u < r; ++u)
                    p[u] = routed ? optr(k0 + u * fs) : A.out + ooff + ptrdiff_t(k0) * A.out_ps + ptrdiff_t(u) * ostep;
                if (op != OP_ASSIGN) {
#pragma unroll
                    for (int u = 0; u < r; ++u) o[u] = mfb_ld_stream(p[u]);
                }
                Dft<r>::run(v);
#pragma unroll
                for (int u = 0; u < r; ++u) emit(p[u], v[u], o[u]);
            }
        }
    }
    // lines of a partially transformed tile beyond nz_fft
    if (zf && line_valid && !fft_line)
        for (int k = q; k < A.n_out; k += P::TPL) mfb_st_stream(optr(k), make_double2(0.0, 0.0));
}

template <class P, int MODE, int MINB, bool BATCH0 = false>
__global__ void __launch_bounds__(P::NTHREADS, MINB) strided2_kernel(StridedArgs A, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    const int tiles_per_a = (A.nz + P::T - 1) / P::T;
    const int a = blockIdx.x / tiles_per_a;
    const int z0 = (blockIdx.x - a * tiles_per_a) * P::T;
    strided2_tile<P, MODE, BATCH0>(A, tw, sm, threadIdx.x, a, z0, CtaSync(), [] {});
}

}  // namespace mfb
