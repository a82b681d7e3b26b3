// kbench_legacy_kernels.cuh - kernel generations that left the product build (commit 1017958), kept here so that the
// micro-benchmark can still time them against their successors: the gen-1 fused z pass (z_conv_kernel), the
// double-buffered persistent strided pass (strided3_kernel), the gen-2 fused z pass (z_conv2_kernel) and the
// pair-parallel multi-product pass (z_multi_kernel).  Development tool, not product; logs of these under profiles/r01_kbench_*.
// (Restored from the repository history after an ignore rule had kept this header out of it; the TMA bulk variant of
// the strided pass, strided4_kernel, was lost with it - its measurements stay in profiles/r02_kbench_strided_tma_bulk_negative.txt.)
#pragma once
#include "fft_core.cuh"
#include "fft_kernels.cuh"
#include "fft_kernels2.cuh"
#include "fft_kernels_z2.cuh"
#include "fft_kernels_zm.cuh"

namespace mfb {

// ---- from fft_kernels.cuh (as of the commit before 1017958) ----
// ------------------------------------------------------------------------------------------
// Fused z pass of a dealiased product on the padded grid (n = 3N/2):
//   rows of A and B hold the nzin = N/2 retained k_z modes; per row
//     a + i b = IFFT_n( A^ + i B^ ) (Hermitian extension, zero padded), p = a*b,
//   two rows' products are packed into one complex line, FFT_n, unpacked, and the first nzin
//   modes are written to `out`.  DIF backward + DIT forward => no permutation in between.
template <class P>
__global__ void __launch_bounds__(P::NTHREADS) z_conv_kernel(const double2* __restrict__ Ain, const double2* __restrict__ Bin,
                                                             double2* __restrict__ out, ptrdiff_t in_pitch, ptrdiff_t out_pitch,
                                                             long nrows, int nzin, double scale, const double2* __restrict__ tw,
                                                             int accumulate = 0) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, T = P::T, H = T / 2;
    static_assert(T >= 2, "need at least two lines per tile");
    const int tid = threadIdx.x;
    const int line = tid % T, q = tid / T;
    const long row0 = long(blockIdx.x) * T;
    // zero the padded band [nzin, N - nzin]
    const int nzero = N - 2 * nzin + 1;
    for (int idx = tid; idx < T * nzero; idx += P::NTHREADS) {
        const int l = idx / nzero, k = nzin + (idx - l * nzero);
        sm[P::smem_addr(k, l)] = make_double2(0.0, 0.0);
    }
    for (int idx = tid; idx < T * nzin; idx += P::NTHREADS) {
        const int l = idx / nzin, k = idx - l * nzin;
        const long r = row0 + l;
        double2 x1 = make_double2(0.0, 0.0), x2 = x1;
        if (r < nrows) {
            x1 = mfb_ld_stream(Ain + r * in_pitch + k);
            x2 = mfb_ld_stream(Bin + r * in_pitch + k);
        }
        if (k == 0) { x1.y = 0.0; x2.y = 0.0; }
        sm[P::smem_addr(k, l)] = make_double2(x1.x - x2.y, -(x1.y + x2.x));
        if (k > 0) sm[P::smem_addr(N - k, l)] = make_double2(x1.x + x2.y, -(x2.x - x1.y));
    }
    __syncthreads();
    NoLoader nl; NoStorer ns;
    fft_tile<P, false, false, false>(sm, tw, line, q, true, nl, ns);   // DIF: result digit-reversed
    __syncthreads();
    // product, in digit-reversed order: a = Re, b = -Im (conj trick); pack lines l and l+H
    for (int idx = tid; idx < H * N; idx += P::NTHREADS) {
        const int p = idx / H, l = idx - p * H;
        const double2 u = sm[P::smem_addr(p, l)];
        const double2 w = sm[P::smem_addr(p, l + H)];
        sm[P::smem_addr(p, l)] = make_double2(-u.x * u.y, -w.x * w.y);
    }
    __syncthreads();
    fft_tile<P, true, false, false>(sm, tw, line, q, line < H, nl, ns);  // DIT: natural order out
    __syncthreads();
    for (int idx = tid; idx < H * nzin; idx += P::NTHREADS) {
        const int l = idx / nzin, k = idx - l * nzin;
        const long r1 = row0 + l, r2 = row0 + l + H;
        const double2 zk = sm[P::smem_addr(k, l)];
        const double2 zm = sm[P::smem_addr((N - k) % N, l)];
        const double h = 0.5 * scale;
        double2 x1 = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
        double2 x2 = make_double2(h * (zk.y + zm.y), -h * (zk.x - zm.x));
        if (accumulate) {
            if (r1 < nrows) { const double2 o = mfb_ld_stream(out + r1 * out_pitch + k); x1.x += o.x; x1.y += o.y; }
            if (r2 < nrows) { const double2 o = mfb_ld_stream(out + r2 * out_pitch + k); x2.x += o.x; x2.y += o.y; }
        }
        if (r1 < nrows) mfb_st_stream(out + r1 * out_pitch + k, x1);
        if (r2 < nrows) mfb_st_stream(out + r2 * out_pitch + k, x2);
    }
}


// ---- from fft_kernels2.cuh (as of the commit before 1017958) ----
// ------------------------------------------------------------------------------------------------------------------
// strided3_kernel - persistent, double-buffered version of the strided pass.
//
// What bounds strided2_kernel is not the access pattern - a bare copy with the same 64..128-byte segments at an 8 KB
// stride runs at 6.0-6.5 TB/s (profiles/r01_kbench_segment_copy.txt) - but bytes in flight: one tile per CTA issues its
// global loads in stage 0, waits for them, and has nothing outstanding during the shared-memory stages, so on average
// ~20 KB per SM are in flight where ~45 KB are needed (6.5 TB/s x ~1 us / 148 SMs), and the pass sustains 3.1-4.0 TB/s.
// (Splitting the CTA into two warp groups that take turns loading made it worse - half the loads in flight again:
// profiles/r01_kbench_pingpong_negative.txt.)
// Here a CTA owns TWO tile buffers.  While the butterflies run on one, the input of the CTA's next tile streams into
// the other with 16-byte cp.async copies (LDGSTS, no register staging), landing directly at the stage-0 leg positions,
// so stage 0 is an in-place shared-memory stage too and a full tile of loads is in flight for the whole compute phase.
template <class P, int MODE>
__global__ void __launch_bounds__(P::NTHREADS, 1) strided3_kernel(StridedArgs A, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm_all);
    constexpr int N = P::N, T = P::T, NT = P::NTHREADS;
    constexpr int TILE = N * T;                      // complex elements per buffer
    const int tid = threadIdx.x;
    const int tiles_per_a = (A.nz + T - 1) / T;
    const long ntiles = long(A.na) * tiles_per_a;
    const int hin = A.n_in / 2, shift = N - A.n_in;  // pad map: source sample s sits at position s (+ shift in the upper half)

    // asynchronous staging of one tile's input; lines / tiles that carry no data are not touched (stage 0 reads zeros)
    auto stage_in = [&](long tile, double2* buf) {
        const int a = int(tile / tiles_per_a);
        const int z0 = int(tile - long(a) * tiles_per_a) * T;
        if ((a + A.a_off == A.a_skip) || (z0 >= A.nz_fft)) return;
        const int l = tid % T, p0 = tid / T;
        const int iz = z0 + l;
        if (iz >= A.nz || iz >= A.nz_fft) return;
        const double2* g = A.in + ptrdiff_t(a) * A.in_as + iz;
        for (int s_ = p0; s_ < A.n_in; s_ += NT / T) {
            const int pos = (MODE == SM_PAD && s_ >= hin) ? s_ + shift : s_;
            mfb_cp_async16(buf + pos * T + l, g + ptrdiff_t(s_) * A.in_ps);
        }
    };

    long tile = blockIdx.x;
    if (tile >= ntiles) return;
    int cur = 0;
    stage_in(tile, sm_all);
    mfb_cp_async_commit();
    for (; tile < ntiles; tile += gridDim.x) {
        const long next = tile + gridDim.x;
        if (next < ntiles) stage_in(next, sm_all + (cur ^ 1) * TILE);
        mfb_cp_async_commit();                       // (possibly empty) group of the next tile
        mfb_cp_async_wait_but_one();                 // this tile's input has landed (this thread's copies) ...
        __syncthreads();                             // ... and everybody else's
        const int a = int(tile / tiles_per_a);
        const int z0 = int(tile - long(a) * tiles_per_a) * T;
        strided2_tile<P, MODE, false, true>(A, tw, sm_all + cur * TILE, tid, a, z0, CtaSync(), [] {});
        __syncthreads();                             // the buffer is refilled two iterations from now at the earliest,
        cur ^= 1;                                    // but the next iteration's staging targets it right away
    }
}


// ---- from fft_kernels_z2.cuh (as of the commit before 1017958) ----
// P::TPL threads per row, RPC rows per CTA (even).  P::T is not used.
template <class P, int RPC, int MINB>
__global__ void __launch_bounds__(P::TPL* RPC, MINB)
    z_conv2_kernel(const double2* __restrict__ Ain, const double2* __restrict__ Bin, double2* __restrict__ out, ptrdiff_t in_pitch,
                   ptrdiff_t out_pitch, long nrows, double scale, const double2* __restrict__ tw, int accumulate = 0) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, S = P::S, TPL = P::TPL, NT = TPL * RPC;
    constexpr int NZ = N / 3;             // retained k_z modes of each operand row
    constexpr int LW = zpad(N) + 1;       // padded line length
    static_assert(P::radix(0) % 3 == 0, "the z pass needs a first radix that is a multiple of 3");
    static_assert(RPC % 2 == 0 && S >= 2, "rows per CTA must be even");
    double2* work = sm;                   // RPC lines of LW
    double2* raw = sm + RPC * LW;         // RPC x {A,B} rows of NZ
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * RPC;

    // ---- raw operand rows: global -> shared, asynchronous 16-byte copies
    for (int idx = tid; idx < RPC * 2 * NZ; idx += NT) {
        const int o = idx / NZ, k = idx - o * NZ;
        const long r = row0 + (o >> 1);
        double2* d = raw + idx;
        if (r < nrows) mfb_cp_async16(d, ((o & 1) ? Bin : Ain) + r * in_pitch + k);
        else *d = make_double2(0.0, 0.0);
    }
    mfb_cp_async_wait_all();
    __syncthreads();

    // ---- backward transform of conj(A^ + i B^) (Hermitian-extended, zero padded), DIF
    {
        const int l = tid / TPL, q = tid - l * TPL;
        double2* ln = work + l * LW;
        const double2* ra = raw + (2 * l) * NZ;
        const double2* rb = ra + NZ;
        constexpr int r = P::radix(0), M = N / r, K3 = r / 3;
#pragma unroll 1
        for (int j = q; j < M; j += TPL) {
            double2 v[r];
#pragma unroll
            for (int t = 0; t < K3; ++t) {   // positions p = j + t*M < NZ: conj(X1[p] + i X2[p])
                double2 x1 = ra[j + t * M], x2 = rb[j + t * M];
                if (t == 0 && j == 0) { x1.y = 0.0; x2.y = 0.0; }  // c2r ignores Im of the self-conjugate mode
                v[t] = make_double2(x1.x - x2.y, -(x1.y + x2.x));
            }
#pragma unroll
            for (int t = K3; t < 2 * K3; ++t) v[t] = make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 2 * K3; t < r; ++t) {  // positions p = N - k, k = (r-t)*M - j in (0, NZ]
                const int k = (r - t) * M - j;
                if (t == 2 * K3 && j == 0) { v[t] = make_double2(0.0, 0.0); continue; }  // k == NZ: dropped Nyquist
                const double2 x1 = ra[k], x2 = rb[k];
                v[t] = make_double2(x1.x + x2.y, -(x2.x - x1.y));
            }
            Dft<r>::run(v);
#pragma unroll
            for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(tw + (u - 1) * M + j));
            double2* sp = ln + zpad(j);
#pragma unroll
            for (int u = 0; u < r; ++u) sp[zleg<M, r>(j, u)] = v[u];
        }
        z2_dif_rest<P, 1, TPL>(ln, tw, q);
    }
    __syncthreads();

    // ---- forward transform of the product pair (p_even + i p_odd), DIT; the product is formed on the fly
    {
        constexpr int PT = 2 * TPL;           // threads per pair
        const int m = tid / PT, q = tid - m * PT;
        double2* l0 = work + (2 * m) * LW;
        const double2* l1 = l0 + LW;
        {
            constexpr int r = P::radix(S - 1), Bn = N / r;  // innermost stage: M == 1, no twiddles
#pragma unroll 1
            for (int b = q; b < Bn; b += PT) {
                const int base = b * r;
                double2* sp = l0 + zpad(base);
                const double2* sq = l1 + zpad(base);
                double2 v[r];
#pragma unroll
                for (int t = 0; t < r; ++t) {
                    const double2 u = sp[zleg<1, r>(base, t)], w = sq[zleg<1, r>(base, t)];
                    v[t] = make_double2(-u.x * u.y, -w.x * w.y);  // a = Re, b = -Im (conj trick)
                }
                Dft<r>::run(v);
#pragma unroll
                for (int u = 0; u < r; ++u) sp[zleg<1, r>(base, u)] = v[u];
            }
        }
        z2_dit_mid<P, S - 2, PT>(l0, tw, q);
        __syncthreads();
        {
            constexpr int r = P::radix(0), M = N / r;  // outermost stage: natural-order output
#pragma unroll 1
            for (int j = q; j < M; j += PT) {
                double2* sp = l0 + zpad(j);
                double2 v[r];
#pragma unroll
                for (int t = 0; t < r; ++t) v[t] = sp[zleg<M, r>(j, t)];
#pragma unroll
                for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(tw + (u - 1) * M + j));
                Dft<r>::run(v);
#pragma unroll
                for (int u = 0; u < r; ++u) sp[zleg<M, r>(j, u)] = v[u];
            }
        }
    }
    __syncthreads();

    // ---- unpack the two half-spectra and store the retained modes
    {
        const double h = 0.5 * scale;
        for (int idx = tid; idx < (RPC / 2) * NZ; idx += NT) {
            const int m = idx / NZ, k = idx - m * NZ;
            const double2* l0 = work + (2 * m) * LW;
            const double2 zk = l0[zpad(k)];
            const double2 zm = l0[zpad((N - k) % N)];
            const long r1 = row0 + 2 * m, r2 = r1 + 1;
            double2 v1 = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
            double2 v2 = make_double2(h * (zk.y + zm.y), -h * (zk.x - zm.x));
            if (accumulate) {   // sum of products of one target, formed before the (single) forward y/x passes
                if (r1 < nrows) { const double2 o = mfb_ld_stream(out + r1 * out_pitch + k); v1.x += o.x; v1.y += o.y; }
                if (r2 < nrows) { const double2 o = mfb_ld_stream(out + r2 * out_pitch + k); v2.x += o.x; v2.y += o.y; }
            }
            if (r1 < nrows) mfb_st_stream(out + r1 * out_pitch + k, v1);
            if (r2 < nrows) mfb_st_stream(out + r2 * out_pitch + k, v2);
        }
    }
}



// ---- from fft_kernels_zm.cuh (as of the commit before 1017958) ----
template <class P, int MINB>
__global__ void __launch_bounds__(2 * P::TPL, MINB) z_multi_kernel(ZmArgs A, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, S = P::S, TPL = P::TPL, PT = 2 * TPL, NT = PT;
    constexpr int NZ = N / 3;
    constexpr int LW = zpad(N) + 1;
    static_assert(P::radix(0) % 3 == 0, "the z pass needs a first radix that is a multiple of 3");
    static_assert(S >= 2, "at least two stages");
    double2* work = sm;              // two padded lines: the rows of the pair
    double2* acc = sm + 2 * LW;      // (p_even, p_odd) summed over the products, digit-reversed order (only if nprod > 1)
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * 2;
    const int l = tid / TPL, ql = tid - l * TPL;     // stage 0: one row per half of the CTA
    const int q = tid;                               // pair stages: all threads on the pair
    double2* ln = work + l * LW;
    const bool row_ok = row0 + l < A.nrows;
    const ptrdiff_t roff = (row0 + l) * A.in_pitch;
    const int nprod = A.nprod;

#pragma unroll 1
    for (int ip = 0; ip < nprod; ++ip) {
        const ZmProduct& pr = A.prod[ip];
        if (ip > 0) __syncthreads();   // the innermost step of the previous product has read both work lines
        // ---- DIF stage 0 of this thread's row: conj(A^ + i B^), Hermitian-extended and zero padded, built in the legs
        {
            constexpr int r = P::radix(0), M = N / r, K3 = r / 3;
            const bool square = pr.square != 0;
            const int na = pr.a.n, nb = square ? 0 : pr.b.n;
            const double2* pa0 = pr.a.t[0].p + roff;
            const double2* pa1 = pr.a.t[1].p + roff;
            const double2* pb0 = pr.b.t[0].p + roff;
            const double2* pb1 = pr.b.t[1].p + roff;
            const double ca0 = pr.a.t[0].c, ca1 = pr.a.t[1].c, cb0 = pr.b.t[0].c, cb1 = pr.b.t[1].c;
            const int ea0 = pr.a.t[0].e, ea1 = pr.a.t[1].e, eb0 = pr.b.t[0].e, eb1 = pr.b.t[1].e;
            auto ld = [&](const double2* p) -> double2 { return row_ok ? mfb_ld_stream(p) : make_double2(0.0, 0.0); };
            auto kpow = [&](double kz, int e, double c) -> double { return e == 0 ? c : (e == 1 ? c * kz : c * kz * kz); };
#pragma unroll 1
            for (int j = ql; j < M; j += TPL) {
                double2 xa[2 * K3], xb[2 * K3];
                int kk[2 * K3];
#pragma unroll
                for (int t = 0; t < K3; ++t) kk[t] = j + t * M;                 // lower half: mode k = position
#pragma unroll
                for (int t = 2 * K3; t < r; ++t) kk[t - K3] = (r - t) * M - j;  // upper half: position N - k
                // all first-term loads of the butterfly are in flight before the first use
#pragma unroll
                for (int m = 0; m < 2 * K3; ++m) {
                    const bool dropped = (m == K3 && j == 0);                   // k == NZ: the dropped Nyquist sample (not stored)
                    xa[m] = dropped ? make_double2(0.0, 0.0) : ld(pa0 + kk[m]);
                    if (nb > 0) xb[m] = dropped ? make_double2(0.0, 0.0) : ld(pb0 + kk[m]);
                }
                if (na > 1 || nb > 1) {   // sums / differences of Hessians (convolution.hh:214-274): second term
                    double2 ya[2 * K3], yb[2 * K3];
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) {
                        const bool dropped = (m == K3 && j == 0);
                        ya[m] = (na > 1 && !dropped) ? ld(pa1 + kk[m]) : make_double2(0.0, 0.0);
                        yb[m] = (nb > 1 && !dropped) ? ld(pb1 + kk[m]) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) {
                        const double kz = A.kfz * double(kk[m]);
                        const double fa0 = kpow(kz, ea0, ca0), fa1 = (na > 1) ? kpow(kz, ea1, ca1) : 0.0;
                        xa[m] = make_double2(fma(fa1, ya[m].x, fa0 * xa[m].x), fma(fa1, ya[m].y, fa0 * xa[m].y));
                        if (nb > 0) {
                            const double fb0 = kpow(kz, eb0, cb0), fb1 = (nb > 1) ? kpow(kz, eb1, cb1) : 0.0;
                            xb[m] = make_double2(fma(fb1, yb[m].x, fb0 * xb[m].x), fma(fb1, yb[m].y, fb0 * xb[m].y));
                        }
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) {
                        const double kz = A.kfz * double(kk[m]);
                        const double fa0 = kpow(kz, ea0, ca0);
                        xa[m] = make_double2(fa0 * xa[m].x, fa0 * xa[m].y);
                        if (nb > 0) {
                            const double fb0 = kpow(kz, eb0, cb0);
                            xb[m] = make_double2(fb0 * xb[m].x, fb0 * xb[m].y);
                        }
                    }
                }
                if (nb == 0) {
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) xb[m] = xa[m];
                }
                double2 v[r];
#pragma unroll
                for (int t = 0; t < K3; ++t) {
                    double2 x1 = xa[t], x2 = xb[t];
                    if (t == 0 && j == 0) { x1.y = 0.0; x2.y = 0.0; }   // c2r ignores Im of the self-conjugate mode
                    v[t] = make_double2(x1.x - x2.y, -(x1.y + x2.x));
                }
#pragma unroll
                for (int t = K3; t < 2 * K3; ++t) v[t] = make_double2(0.0, 0.0);
#pragma unroll
                for (int t = 2 * K3; t < r; ++t) {
                    const double2 x1 = xa[t - K3], x2 = xb[t - K3];
                    v[t] = make_double2(x1.x + x2.y, -(x2.x - x1.y));
                }
                Dft<r>::run(v);
                apply_tw<r>(v, tw + j, M);
                double2* sp = ln + zpad(j);
#pragma unroll
                for (int u = 0; u < r; ++u) sp[zleg<M, r>(j, u)] = v[u];
            }
        }
        // ---- middle DIF stages, both rows of the pair in one thread (twiddles fetched once)
        z3_dif_mid_pair<P, 1, PT>(work, tw, q);
        __syncthreads();
        // ---- innermost DIF stage of both rows -> product -> accumulate; after the last product: innermost DIT stage
        {
            constexpr int r = P::radix(S - 1), Bn = N / r;
            const double cf = pr.coeff;
            const bool first = ip == 0, last = ip == nprod - 1;
#pragma unroll 1
            for (int b = q; b < Bn; b += PT) {
                const int base = b * r;
                double2* sp = work + zpad(base);
                double2* ap = acc + zpad(base);
                double2 v[r];
                double p0[r];
#pragma unroll
                for (int t = 0; t < r; ++t) v[t] = sp[zleg<1, r>(base, t)];
                Dft<r>::run(v);
#pragma unroll
                for (int t = 0; t < r; ++t) p0[t] = -v[t].x * v[t].y;   // a = Re, b = -Im (conj trick)
#pragma unroll
                for (int t = 0; t < r; ++t) v[t] = sp[LW + zleg<1, r>(base, t)];
                Dft<r>::run(v);
#pragma unroll
                for (int t = 0; t < r; ++t) v[t] = make_double2(cf * p0[t], cf * (-v[t].x * v[t].y));
                if (!first) {
#pragma unroll
                    for (int t = 0; t < r; ++t) {
                        const double2 o = ap[zleg<1, r>(base, t)];
                        v[t].x += o.x; v[t].y += o.y;
                    }
                }
                if (!last) {
#pragma unroll
                    for (int t = 0; t < r; ++t) ap[zleg<1, r>(base, t)] = v[t];
                } else {
                    Dft<r>::run(v);
#pragma unroll
                    for (int u = 0; u < r; ++u) sp[zleg<1, r>(base, u)] = v[u];
                }
            }
        }
    }
    // ---- forward transform of the packed pair (DIT), once per target
    double2* l0 = work;
    z2_dit_mid<P, S - 2, PT>(l0, tw, q);
    __syncthreads();
    {
        constexpr int r = P::radix(0), M = N / r;
#pragma unroll 1
        for (int j = q; j < M; j += PT) {
            double2* sp = l0 + zpad(j);
            double2 v[r];
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = sp[zleg<M, r>(j, t)];
            apply_tw<r>(v, tw + j, M);
            Dft<r>::run(v);
#pragma unroll
            for (int u = 0; u < r; ++u) sp[zleg<M, r>(j, u)] = v[u];
        }
    }
    __syncthreads();
    {   // unpack the two half-spectra, store the retained modes
        const long r1 = row0, r2 = row0 + 1;
        for (int k = tid; k < NZ; k += NT) {
            const double2 zk = l0[zpad(k)];
            const double2 zm = l0[zpad((N - k) % N)];
            const double2 v1 = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double2 v2 = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
            if (r1 < A.nrows) mfb_st_stream(A.out + r1 * A.out_pitch + k, v1);
            if (r2 < A.nrows) mfb_st_stream(A.out + r2 * A.out_pitch + k, v2);
        }
    }
}

template <class P> constexpr size_t z_multi_smem(bool with_acc) { return size_t(with_acc ? 3 : 2) * (zpad(P::N) + 1) * sizeof(double2); }

template <class P, int RPC> constexpr size_t z_conv2_smem() { return size_t(RPC) * (zpad(P::N) + 1 + 2 * (P::N / 3)) * sizeof(double2); }

}  // namespace mfb
