// fft_core.cuh - FP64 mixed-radix FFT building blocks for sm_100a.
//
// Design:
//   * A CTA transforms a TILE of T independent lines of length n (n = 2^a * 3^b) that sits in
//     shared memory as sm[pos][line] (line index XOR-swizzled, see smem_addr) so that the T
//     lines that are adjacent in global memory (consecutive k_z) are adjacent lanes: every
//     global access is a T*16-byte contiguous segment and every shared-memory access of a
//     quarter warp hits 8 distinct 16-byte bank groups.
//   * In-place decimation-in-frequency with a radix sequence (r_0 .. r_{S-1}), radices from
//     {2,3,4,6,8,12,16}: one barrier per stage, registers hold one butterfly at a time.
//     The first stage can pull its inputs straight from global memory through a Loader
//     functor (pad / prune / k-space factor fused in), the last stage can push its outputs
//     through a Storer functor (prune / scale / output-op fused in).  Outputs of the last
//     stage come in digit-reversed position order; freq_of_pos/pos_of_freq map between them.
//   * Only the forward transform (e^{-2 pi i jk/n}) is implemented; the backward transform is
//     conj(FFT(conj(x))) with the conjugations folded into the loader / storer.
//   * Inter-stage twiddles come from a global table tw[m] = exp(-2 pi i m / n) (L1-resident,
//     broadcast across the T lanes of a tile); butterfly-internal twiddles are compile-time
//     48th roots of unity.
#pragma once
#include "mfb_platform.h"

namespace mfb {

// ------------------------------------------------------------------ complex helpers
MFB_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
MFB_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
MFB_HD double2 cmul(double2 a, double2 b) {
    return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y));
}
MFB_HD double2 cmulc(double2 a, double cr, double ci) {  // a * (cr + i ci)
    return make_double2(fma(-a.y, ci, a.x * cr), fma(a.y, cr, a.x * ci));
}
MFB_HD double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
MFB_HD double2 cscale(double2 a, double s) { return make_double2(a.x * s, a.y * s); }
MFB_HD double2 cmul_i(double2 a) { return make_double2(-a.y, a.x); }    // * (+i)
MFB_HD double2 cmul_mi(double2 a) { return make_double2(a.y, -a.x); }   // * (-i)

// ------------------------------------------------------------------ 48th roots of unity
__host__ __device__ constexpr double cos48_q(int j) {  // cos(2*pi*j/48), j = 0..12
    switch (j) {
        case 0: return 1.0;
        case 1: return 0.99144486137381041114;
        case 2: return 0.96592582628906828675;
        case 3: return 0.92387953251128675613;
        case 4: return 0.86602540378443864676;
        case 5: return 0.79335334029123516458;
        case 6: return 0.70710678118654752440;
        case 7: return 0.60876142900872063942;
        case 8: return 0.5;
        case 9: return 0.38268343236508977173;
        case 10: return 0.25881904510252076235;
        case 11: return 0.13052619222005159155;
        default: return 0.0;
    }
}
__host__ __device__ constexpr double cos48(int j) {
    j = ((j % 48) + 48) % 48;
    if (j > 24) j = 48 - j;
    return (j > 12) ? -cos48_q(24 - j) : cos48_q(j);
}
__host__ __device__ constexpr double sin48(int j) { return cos48(j - 12); }

// ------------------------------------------------------------------ small forward DFTs (in registers)
template <int R> struct Dft;

template <> struct Dft<1> { static MFB_HD void run(double2*) {} };

template <> struct Dft<2> {
    static MFB_HD void run(double2* v) {
        double2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <> struct Dft<3> {
    static MFB_HD void run(double2* v) {
        constexpr double s = 0.86602540378443864676;
        double2 t = cadd(v[1], v[2]);
        double2 d = csub(v[1], v[2]);
        double2 m = make_double2(fma(-0.5, t.x, v[0].x), fma(-0.5, t.y, v[0].y));
        v[0] = cadd(v[0], t);
        // X1 = m - i s d ; X2 = m + i s d
        v[1] = make_double2(fma(s, d.y, m.x), fma(-s, d.x, m.y));
        v[2] = make_double2(fma(-s, d.y, m.x), fma(s, d.x, m.y));
    }
};

template <> struct Dft<4> {
    static MFB_HD void run(double2* v) {
        double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        double2 t2 = cadd(v[1], v[3]), t3 = csub(v[1], v[3]);
        v[0] = cadd(t0, t2);
        v[2] = csub(t0, t2);
        v[1] = cadd(t1, cmul_mi(t3));  // t1 - i t3
        v[3] = cadd(t1, cmul_i(t3));   // t1 + i t3
    }
};

// multiply by w_R^m = exp(-2 pi i m / R), compile-time m
template <int R, int M> MFB_HD double2 tw_const(double2 a) {
    static_assert(48 % R == 0, "radix must divide 48");
    constexpr int j = ((M * (48 / R)) % 48 + 48) % 48;
    if constexpr (j == 0) return a;
    else if constexpr (j == 12) return cmul_mi(a);
    else if constexpr (j == 24) return make_double2(-a.x, -a.y);
    else if constexpr (j == 36) return cmul_i(a);
    else {
        constexpr double c = cos48(j);
        constexpr double s = -sin48(j);
        return cmulc(a, c, s);
    }
}

// Cooley-Tukey composition R = A*B: input index i = B*a + b, output k = k1 + A*k2
template <int A, int B> struct DftCT {
    static constexpr int R = A * B;
    template <int b, int k1> static MFB_HD void twid(double2* t) {
        if constexpr (k1 < A) {
            t[k1] = tw_const<R, b * k1>(t[k1]);
            twid<b, k1 + 1>(t);
        }
    }
    template <int b> static MFB_HD void cols(double2* v) {
        if constexpr (b < B) {
            double2 t[A];
#pragma unroll
            for (int a = 0; a < A; ++a) t[a] = v[b + B * a];
            Dft<A>::run(t);
            twid<b, 1>(t);
#pragma unroll
            for (int a = 0; a < A; ++a) v[b + B * a] = t[a];
            cols<b + 1>(v);
        }
    }
    static MFB_HD void run(double2* v) {
        cols<0>(v);
        double2 o[R];
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            double2 t[B];
#pragma unroll
            for (int b = 0; b < B; ++b) t[b] = v[B * k1 + b];
            Dft<B>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) o[k1 + A * k2] = t[k2];
        }
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = o[i];
    }
};

template <> struct Dft<6> { static MFB_HD void run(double2* v) { DftCT<2, 3>::run(v); } };
template <> struct Dft<8> { static MFB_HD void run(double2* v) { DftCT<2, 4>::run(v); } };
template <> struct Dft<12> { static MFB_HD void run(double2* v) { DftCT<3, 4>::run(v); } };
template <> struct Dft<16> { static MFB_HD void run(double2* v) { DftCT<4, 4>::run(v); } };

// ------------------------------------------------------------------ twiddles of one butterfly from two table entries
// w[u] = W^u for u = 1..R-1, W = exp(-2 pi i j / L), from the table entries W^1 and W^4 (stage-major table: entry (u, j) sits at
// twp[(u - 1) * M]): products of depth <= 3 instead of R-1 loads.  The loads are what costs here - the L1/shared data pipe
// is the busiest unit of the FFT passes (ncu: 60-66 % of peak with the FP64 pipe at 32-37 %), and with the shared-memory
// carve-out at its maximum the table falls out of L1 (14 % hit rate) - while the FP64 pipe has room for the extra
// multiplications.  The result differs from the table by a few 1e-16 (relative).
MFB_HD double2 csqr(double2 a) { return make_double2(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y); }
template <int R> MFB_D void tw_powers(const double2* __restrict__ twp, int M, double2* w) {
    w[1] = __ldg(twp);
    if constexpr (R > 2) w[2] = csqr(w[1]);
    if constexpr (R > 3) w[3] = cmul(w[2], w[1]);
    if constexpr (R > 4) w[4] = __ldg(twp + 3 * M);
    if constexpr (R > 5) w[5] = cmul(w[4], w[1]);
    if constexpr (R > 6) w[6] = cmul(w[4], w[2]);
    if constexpr (R > 7) w[7] = cmul(w[4], w[3]);
    if constexpr (R > 8) w[8] = csqr(w[4]);
    if constexpr (R > 9) w[9] = cmul(w[8], w[1]);
    if constexpr (R > 10) w[10] = cmul(w[8], w[2]);
    if constexpr (R > 11) w[11] = cmul(w[8], w[3]);
    if constexpr (R > 12) w[12] = cmul(w[8], w[4]);
    if constexpr (R > 13) w[13] = cmul(w[12], w[1]);
    if constexpr (R > 14) w[14] = cmul(w[12], w[2]);
    if constexpr (R > 15) w[15] = cmul(w[12], w[3]);
    static_assert(R <= 16, "radix");
}

#ifndef MFB_TW_COMPUTE
#define MFB_TW_COMPUTE 1   // 0: every twiddle from the table (kbench A/B)
#endif
// v[u] *= W^u (u = 1..R-1) for the butterfly whose table entries start at twp (stride M between u)
template <int R> MFB_D void apply_tw(double2* v, const double2* __restrict__ twp, int M) {
#if MFB_TW_COMPUTE
    if constexpr (R > 4) {
        double2 w[R];
        tw_powers<R>(twp, M, w);
#pragma unroll
        for (int u = 1; u < R; ++u) v[u] = cmul(v[u], w[u]);
        return;
    }
#endif
#pragma unroll
    for (int u = 1; u < R; ++u) v[u] = cmul(v[u], __ldg(twp + (u - 1) * M));
}

// ------------------------------------------------------------------ plan traits
template <int... Rs> __host__ __device__ constexpr int pack_get(int s) {
    constexpr int r[] = {Rs...};
    return r[s];
}
template <int... Rs> __host__ __device__ constexpr int pack_prod(int upto) {
    int p = 1;
    for (int i = 0; i < upto; ++i) p *= pack_get<Rs...>(i);
    return p;
}

// element offset of stage s in the stage-major twiddle table (see Plan::tw2_off)
template <int... Rs> __host__ __device__ constexpr int pack_tw2_off(int n, int s) {
    int o = 0;
    for (int i = 0; i < s; ++i) {
        const int len = n / pack_prod<Rs...>(i), r = pack_get<Rs...>(i);
        o += (r - 1) * (len / r);
    }
    return o;
}

template <int N_, int TPL_, int T_, int... Rs> struct Plan {
    static constexpr int N = N_;      // line length
    static constexpr int TPL = TPL_;  // threads per line
    static constexpr int T = T_;      // lines per tile (power of two)
    static constexpr int S = sizeof...(Rs);
    static constexpr int NTHREADS = TPL * T;
    static constexpr size_t SMEM_BYTES = size_t(N) * T * sizeof(double2);
    static_assert(pack_prod<Rs...>(S) == N_, "radices must multiply to N");
    static_assert((T_ & (T_ - 1)) == 0, "T must be a power of two");
    static constexpr int M0 = N_ / pack_get<Rs...>(0);
    __host__ __device__ static constexpr int radix(int s) { return pack_get<Rs...>(s); }
    // sub-transform length at stage s
    __host__ __device__ static constexpr int len(int s) { return N_ / pack_prod<Rs...>(s); }
    // frequency stride of digit s
    __host__ __device__ static constexpr int fstride(int s) { return pack_prod<Rs...>(s); }

    // digit-reversed position <-> frequency
    MFB_HD static int freq_of_pos(int p) {
        int k = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int Ms = len(s) / radix(s);
            k += ((p / Ms) % radix(s)) * fstride(s);
        }
        return k;
    }
    MFB_HD static int pos_of_freq(int k) {
        int p = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int Ms = len(s) / radix(s);
            p += ((k / fstride(s)) % radix(s)) * Ms;
        }
        return p;
    }
    // Stage-major twiddle table ("tw2"): for every twiddled stage s the block [u-1][j] = exp(-2 pi i u j / len(s)),
    // u = 1..radix(s)-1, j = 0..len(s)/radix(s)-1.  Lanes that work on consecutive butterflies j read consecutive
    // 16-byte entries (one 128-byte wavefront per 8 lanes) instead of entries u*step apart in the plain table
    // (one wavefront PER lane: the L1 data pipe was 89 % busy in the fused z pass, profiles/r01_ncu_z3_768.txt).
    __host__ __device__ static constexpr int tw2_off(int s) { return pack_tw2_off<Rs...>(N_, s); }
    static constexpr int TW2_ELEMS = pack_tw2_off<Rs...>(N_, int(sizeof...(Rs)) - 1);
    // swizzled shared-memory slot of (position, line)
    MFB_HD static int smem_addr(int pos, int line) {
        return pos * T + (line ^ ((pos + pos / M0) & (T - 1)));
    }
};

// ------------------------------------------------------------------ one stage
// DIF (DIT == false): natural-order input, butterfly then twiddle, stages 0..S-1, output in
//                     digit-reversed position order.
// DIT (DIT == true) : the transposed flow graph: input in digit-reversed position order
//                     (x[j] sits at pos_of_freq(j)), twiddle then butterfly, stages S-1..0,
//                     natural-order output.  DIF followed by DIT needs no permutation at all,
//                     which is what the fused convolution z-pass exploits.
// FROM_LOADER: inputs of the first executed stage come from ld(index, line) - index is the
//              natural sample index - instead of shared memory.
// TO_STORER  : outputs of the last executed stage go to st(freq, line, value).
// `active` == false: the thread only takes part in the barriers.
template <class P, int s, bool DIT, bool FROM_LOADER, bool TO_STORER, class Loader, class Storer>
MFB_D void fft_stage(double2* __restrict__ sm, const double2* __restrict__ tw, int line, int q, bool active,
                     Loader& ld, Storer& st) {
    constexpr int r = P::radix(s);
    constexpr int L = P::len(s);
    constexpr int M = L / r;
    constexpr int B = P::N / r;
    static_assert(B % P::TPL == 0, "threads per line must divide the butterfly count");
    constexpr int PER = B / P::TPL;
    constexpr bool innermost = (s == P::S - 1);  // M == 1: no twiddles
    constexpr int twstep = P::N / L;
    constexpr int fs = P::fstride(s);
    if (!active) return;

#pragma unroll 1
    for (int i = 0; i < PER; ++i) {
        const int b = q + i * P::TPL;
        const int blk = b / M;
        const int j = b - blk * M;
        const int base = blk * L + j;
        double2 v[r];
        if constexpr (FROM_LOADER && DIT) {
            static_assert(innermost, "DIT loader only at stage S-1");
            const int k0 = P::freq_of_pos(base);
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = ld(k0 + t * fs, line);
        } else {
#pragma unroll
            for (int t = 0; t < r; ++t) {
                if constexpr (FROM_LOADER) v[t] = ld(base + t * M, line);
                else v[t] = sm[P::smem_addr(base + t * M, line)];
            }
        }
        if constexpr (DIT && !innermost) {
#pragma unroll
            for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(&tw[(j * u) * twstep]));
        }
        Dft<r>::run(v);
        if constexpr (!DIT && !innermost) {
#pragma unroll
            for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(&tw[(j * u) * twstep]));
        }
        if constexpr (TO_STORER && !DIT) {
            static_assert(innermost, "DIF storer only at stage S-1");
            const int k0 = P::freq_of_pos(base);
#pragma unroll
            for (int u = 0; u < r; ++u) st(k0 + u * fs, line, v[u]);
        } else if constexpr (TO_STORER && DIT) {
            static_assert(s == 0, "DIT storer only at stage 0");
#pragma unroll
            for (int u = 0; u < r; ++u) st(base + u * M, line, v[u]);
        } else {
#pragma unroll
            for (int u = 0; u < r; ++u) sm[P::smem_addr(base + u * M, line)] = v[u];
        }
    }
}

template <class P, int i, bool DIT, bool FROM_LOADER, bool TO_STORER, class Loader, class Storer>
MFB_D void fft_stages_from(double2* __restrict__ sm, const double2* __restrict__ tw, int line, int q,
                           bool active, Loader& ld, Storer& st) {
    if constexpr (i < P::S) {
        constexpr int s = DIT ? (P::S - 1 - i) : i;
        constexpr bool first = (i == 0), last = (i == P::S - 1);
        if constexpr (!first) __syncthreads();
        fft_stage<P, s, DIT, first && FROM_LOADER, last && TO_STORER>(sm, tw, line, q, active, ld, st);
        fft_stages_from<P, i + 1, DIT, FROM_LOADER, TO_STORER>(sm, tw, line, q, active, ld, st);
    }
}

// Full forward FFT of the tile.  If !FROM_LOADER the tile must already be in shared memory
// (caller has synchronised); if !TO_STORER the result stays in shared memory (caller
// synchronises before reading it).  All threads of the CTA must call this (barriers inside).
template <class P, bool DIT, bool FROM_LOADER, bool TO_STORER, class Loader, class Storer>
MFB_D void fft_tile(double2* __restrict__ sm, const double2* __restrict__ tw, int line, int q, bool active,
                    Loader& ld, Storer& st) {
    static_assert(P::S >= 2, "at least two stages (a barrier must separate loads and stores)");
    fft_stages_from<P, 0, DIT, FROM_LOADER, TO_STORER>(sm, tw, line, q, active, ld, st);
}

struct NoLoader {
    MFB_D double2 operator()(int, int) const { return make_double2(0.0, 0.0); }
};
struct NoStorer {
    MFB_D void operator()(int, int, double2) const {}
};

}  // namespace mfb
