// fft_kernels_zm.cuh - multi-product fused z pass of the LPT source stage (FP64, sm_100a).
//
// One launch forms, per (x,y) row of a chunk of the padded grid,
//     out^(k_z) = r2c_z [ sum_p coeff_p * a_p(z) * b_p(z) ],      a_p = c2r_z[ sum_t c_t k_z^{e_t} F_t^(k_z) ]
// i.e. ALL the dealiased products of one target (src/ic_generator.cc:482-486, 522-538, 551-554) with the sum taken in
// REAL space, so the forward transform runs once per target instead of once per product
// (include/convolution.hh:524-534 does c2r, c2r, multiply, r2c per product).  The operands F_t are "base fields": the
// x- and y-transformed spectra of k_x^mx k_y^my f^ of a source potential; the k_z part of a Hessian factor
// (include/convolution.hh:137-160, -k_a k_b) and sums / differences of Hessians (:214-274) are applied while the row is
// loaded, so six base fields per potential serve every product of the stage.
//
// Per product the data flow of z_conv3_kernel is kept: Hermitian extension of conj(A^ + i B^) in the legs of DIF stage 0
// straight from global memory, both rows of a pair in one thread in the middle stages, innermost DIF stage of both rows
// -> pointwise product in registers.  The product then goes to a per-pair accumulator line in shared memory (digit-
// reversed order, thread-private positions: no barrier); after the last product the accumulator feeds the innermost DIT
// stage and the forward transform / unpack / store run once.
#pragma once
#include "fft_kernels_z2.cuh"

namespace mfb {

constexpr int ZM_MAXP = 12;   // products per launch (phi3 3b: 6, A3_d: 4, phi2: 5, 3a outer: 4)

struct ZmTerm {
    const double2* p;   // base field chunk: rows of `in_pitch` complex, retained k_z modes only
    double c;           // coefficient (includes the sign of (i k_a)(i k_b))
    int e;              // power of k_z: 0, 1, 2
    int pad_;
};
struct ZmOperand {
    ZmTerm t[2];
    int n;              // 1 or 2 terms
    int pad_;
};
struct ZmProduct {
    ZmOperand a, b;
    double coeff;
    int square;         // b is the same operand as a: its row is loaded once
    int pad_;
};
struct ZmArgs {
    ZmProduct prod[ZM_MAXP];
    int nprod;
    int pad_;
    double2* out;       // target chunk: rows of `out_pitch` complex
    ptrdiff_t in_pitch, out_pitch;
    long nrows;
    double kfz;         // 2 pi / L_z
};

// ------------------------------------------------------------------------------------------------------------------
// z_multi2_kernel - the same contract with ONE work line per CTA: the two rows of a pair go through the backward transform
// one after the other (NT threads on one row), their products land in two padded REAL accumulator lines (sum over the
// products of p_even and p_odd), and the forward transform of the packed pair starts from those.  Work line + accumulators
// = 2 complex lines (55 KB at n = 1536), so FOUR CTAs share an SM where the pair-parallel version above fits two:
// the first B200 run of z_multi_kernel (246 registers, 8 warps per SM) sustained 21.6 ms per complex-FFT sweep of the padded
// grid against 15.7 ms for the per-product kernel with 12 warps per SM - the pass is bound by latency, not by a pipe.
template <class P, int s, int NT> MFB_D void zm_dif_mid(double2* ln, const double2* tw, int q) {
    if constexpr (s < P::S - 1) {
        __syncthreads();
        z2_dif_stage<P, s, NT>(ln, tw, q);
        zm_dif_mid<P, s + 1, NT>(ln, tw, q);
    }
}

//
// PF (prefetch): the first-term rows of the NEXT (product, row) item are fetched by the TMA unit - two cp.async.bulk copies of a
// whole row each, issued by one thread right after the barrier that ends stage 0, completion on an mbarrier - into a 2-row
// staging area while the middle and innermost stages of the current item run; stage 0 then reads them from shared memory.
// The wait for HBM (long-scoreboard stalls: 17 % of all stall cycles of the direct version, profiles/r02_ncu_zmulti2_*.txt)
// leaves the critical path of the CTA and no registers hold loads in flight.  Second terms of two-term operands stay direct
// loads (a third and fourth staged row would cost the third resident CTA).
// the same with the barrier AFTER each stage (the caller has synchronised before the first one)
template <class P, int s, int NT> MFB_D void zm_dif_mid_nosync(double2* ln, const double2* tw, int q) {
    if constexpr (s < P::S - 1) {
        z2_dif_stage<P, s, NT>(ln, tw, q);
        __syncthreads();
        zm_dif_mid_nosync<P, s + 1, NT>(ln, tw, q);
    }
}

template <class P, int NT, int MINB, bool PF = false>
__global__ void __launch_bounds__(NT, MINB) z_multi2_kernel(ZmArgs A, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, S = P::S;
    constexpr int NZ = N / 3;
    constexpr int LW = zpad(N) + 1;
    static_assert(P::radix(0) % 3 == 0, "the z pass needs a first radix that is a multiple of 3");
    static_assert(S >= 2, "at least two stages");
    double2* work = sm;                                         // one padded complex line
    double* accx = reinterpret_cast<double*>(sm + LW);          // sum of the products of the even row (padded like a line)
    double* accy = accx + LW;                                   // ... of the odd row
    double2* raw = sm + 2 * LW;                                 // PF: staged first-term rows of A and B (NZ complex each)
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 2 * NZ);
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * 2;
    const int nprod = A.nprod;
    // PF: thread 0 starts the bulk copies of item (ip, h); rows beyond nrows are not copied (stage 0 substitutes zeros)
    auto prefetch = [&](int ip, int h) {
        const ZmProduct& pr = A.prod[ip];
        const bool sq = pr.square != 0;
        const long ra = row0 + h, rb = sq ? row0 + 1 : row0 + h;
        const bool oka = ra < A.nrows, okb = rb < A.nrows;
        constexpr unsigned ROWB = unsigned(NZ) * unsigned(sizeof(double2));
        mfb_mbar_expect_tx(bar, (oka ? ROWB : 0u) + (okb ? ROWB : 0u));
        if (oka) mfb_bulk_g2s(raw, pr.a.t[0].p + ra * A.in_pitch, ROWB, bar);
        if (okb) mfb_bulk_g2s(raw + NZ, (sq ? pr.a.t[0].p : pr.b.t[0].p) + rb * A.in_pitch, ROWB, bar);
    };
    unsigned phase = 0;
    if constexpr (PF) {
        if (tid == 0) {
            mfb_mbar_init(bar, 1);
            if (nprod > 0) prefetch(0, 0);
        }
        __syncthreads();
    }

#pragma unroll 1
    for (int ip = 0; ip < nprod; ++ip) {
        const ZmProduct& pr = A.prod[ip];
        const bool square = pr.square != 0;
        const int na = pr.a.n, nb = square ? pr.a.n : pr.b.n;
        const double ca0 = pr.a.t[0].c, ca1 = pr.a.t[1].c, cb0 = square ? ca0 : pr.b.t[0].c, cb1 = square ? ca1 : pr.b.t[1].c;
        const int ea0 = pr.a.t[0].e, ea1 = pr.a.t[1].e, eb0 = square ? ea0 : pr.b.t[0].e, eb1 = square ? ea1 : pr.b.t[1].e;
        const double cf = pr.coeff;
        // A SQUARE (xy^2, yz^2, ... : 6 of the 32 products of a 3LPT step) needs one real field per row: the two ROWS of the pair
        // share one complex transform (a_even + i a_odd) and the pair is done in one pass instead of two.
        const int npass = square ? 1 : 2;
#pragma unroll 1
        for (int h = 0; h < npass; ++h) {
            const bool row_ok = row0 + h < A.nrows;
            const bool row1_ok = row0 + 1 < A.nrows;
            const ptrdiff_t roff = (row0 + h) * A.in_pitch;
            if (ip > 0 || h > 0) __syncthreads();   // the innermost step of the previous row has read the work line
            if constexpr (PF) { mfb_mbar_wait(bar, phase); phase ^= 1u; }
            // ---- DIF stage 0: conj(A^ + i B^), Hermitian-extended and zero padded, built in the legs
            {
                constexpr int r = P::radix(0), M = N / r, K3 = r / 3;
                const double2* pa0 = pr.a.t[0].p + roff;
                const double2* pa1 = pr.a.t[1].p + roff;
                // square: "B" is the same operand on the odd row of the pair
                const double2* pb0 = square ? pr.a.t[0].p + roff + A.in_pitch : pr.b.t[0].p + roff;
                const double2* pb1 = square ? pr.a.t[1].p + roff + A.in_pitch : pr.b.t[1].p + roff;
                auto ld = [&](const double2* p) -> double2 { return row_ok ? mfb_ld_stream(p) : make_double2(0.0, 0.0); };
                auto ldb = [&](const double2* p) -> double2 { return (square ? row1_ok : row_ok) ? mfb_ld_stream(p) : make_double2(0.0, 0.0); };
                // first terms: staged rows (PF) or direct loads
                auto ld0 = [&](int k) -> double2 {
                    if constexpr (PF) return row_ok ? raw[k] : make_double2(0.0, 0.0);
                    else return ld(pa0 + k);
                };
                auto ldb0 = [&](int k) -> double2 {
                    if constexpr (PF) return (square ? row1_ok : row_ok) ? raw[NZ + k] : make_double2(0.0, 0.0);
                    else return ldb(pb0 + k);
                };
                auto kpow = [&](double kz, int e, double c) -> double { return e == 0 ? c : (e == 1 ? c * kz : c * kz * kz); };
#pragma unroll 1
                for (int j = tid; j < M; j += NT) {
                    double2 xa[2 * K3], xb[2 * K3];
                    auto mode_k = [&](int m) -> int { return m < K3 ? j + m * M : (2 * K3 - m) * M - j; };   // leg m <-> mode k
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) {
                        const bool dropped = (m == K3 && j == 0);   // k == NZ: the dropped Nyquist sample (not stored)
                        xa[m] = dropped ? make_double2(0.0, 0.0) : ld0(mode_k(m));
                        xb[m] = dropped ? make_double2(0.0, 0.0) : ldb0(mode_k(m));
                    }
                    if (na > 1) {   // sums / differences of Hessians (convolution.hh:214-274): second term of A
                        double2 ya[2 * K3];
#pragma unroll
                        for (int m = 0; m < 2 * K3; ++m) ya[m] = (m == K3 && j == 0) ? make_double2(0.0, 0.0) : ld(pa1 + mode_k(m));
#pragma unroll
                        for (int m = 0; m < 2 * K3; ++m) {
                            const double kz = A.kfz * double(mode_k(m));
                            const double f0 = kpow(kz, ea0, ca0), f1 = kpow(kz, ea1, ca1);
                            xa[m] = make_double2(fma(f1, ya[m].x, f0 * xa[m].x), fma(f1, ya[m].y, f0 * xa[m].y));
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 2 * K3; ++m) {
                            const double f0 = kpow(A.kfz * double(mode_k(m)), ea0, ca0);
                            xa[m] = make_double2(f0 * xa[m].x, f0 * xa[m].y);
                        }
                    }
                    if (nb > 1) {
                        double2 yb[2 * K3];
#pragma unroll
                        for (int m = 0; m < 2 * K3; ++m) yb[m] = (m == K3 && j == 0) ? make_double2(0.0, 0.0) : ldb(pb1 + mode_k(m));
#pragma unroll
                        for (int m = 0; m < 2 * K3; ++m) {
                            const double kz = A.kfz * double(mode_k(m));
                            const double f0 = kpow(kz, eb0, cb0), f1 = kpow(kz, eb1, cb1);
                            xb[m] = make_double2(fma(f1, yb[m].x, f0 * xb[m].x), fma(f1, yb[m].y, f0 * xb[m].y));
                        }
                    } else {
#pragma unroll
                        for (int m = 0; m < 2 * K3; ++m) {
                            const double f0 = kpow(A.kfz * double(mode_k(m)), eb0, cb0);
                            xb[m] = make_double2(f0 * xb[m].x, f0 * xb[m].y);
                        }
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
                    double2* sp = work + zpad(j);
#pragma unroll
                    for (int u = 0; u < r; ++u) sp[zleg<M, r>(j, u)] = v[u];
                }
            }
            if constexpr (PF) {
                __syncthreads();   // stage 0 has consumed the staged rows: the TMA unit may overwrite them with the next item's
                if (tid == 0) {
                    if (h + 1 < npass) prefetch(ip, h + 1);
                    else if (ip + 1 < nprod) prefetch(ip + 1, 0);
                }
                zm_dif_mid_nosync<P, 1, NT>(work, tw, tid);
            } else {
                zm_dif_mid<P, 1, NT>(work, tw, tid);
                __syncthreads();
            }
            // ---- innermost DIF stage -> a * b -> real accumulator of this row; last product, odd row: innermost DIT stage
            {
                constexpr int r = P::radix(S - 1), Bn = N / r;
                double* acc = h ? accy : accx;
                const bool first = ip == 0, finish = (ip == nprod - 1) && (h == 1 || square);
#pragma unroll 1
                for (int b = tid; b < Bn; b += NT) {
                    const int base = b * r;
                    double2* sp = work + zpad(base);
                    double* ap = acc + zpad(base);
                    double2 v[r];
#pragma unroll
                    for (int t = 0; t < r; ++t) v[t] = sp[zleg<1, r>(base, t)];
                    Dft<r>::run(v);
                    double p[r];
                    if (square) {   // Re = a_even, -Im = a_odd: both rows' squares in one go
                        double* ay = accy + zpad(base);
                        double pe[r];
#pragma unroll
                        for (int t = 0; t < r; ++t) { pe[t] = cf * v[t].x * v[t].x; p[t] = cf * v[t].y * v[t].y; }
                        if (!first) {
#pragma unroll
                            for (int t = 0; t < r; ++t) { pe[t] += ap[zleg<1, r>(base, t)]; p[t] += ay[zleg<1, r>(base, t)]; }
                        }
                        if (!finish) {
#pragma unroll
                            for (int t = 0; t < r; ++t) { ap[zleg<1, r>(base, t)] = pe[t]; ay[zleg<1, r>(base, t)] = p[t]; }
                        } else {
#pragma unroll
                            for (int t = 0; t < r; ++t) v[t] = make_double2(pe[t], p[t]);
                            Dft<r>::run(v);
#pragma unroll
                            for (int u = 0; u < r; ++u) sp[zleg<1, r>(base, u)] = v[u];
                        }
                        continue;
                    }
#pragma unroll
                    for (int t = 0; t < r; ++t) p[t] = cf * (-v[t].x * v[t].y);   // a = Re, b = -Im (conj trick)
                    if (!first) {
#pragma unroll
                        for (int t = 0; t < r; ++t) p[t] += ap[zleg<1, r>(base, t)];
                    }
                    if (!finish) {
#pragma unroll
                        for (int t = 0; t < r; ++t) ap[zleg<1, r>(base, t)] = p[t];
                    } else {   // forward transform of (p_even + i p_odd): same thread, same positions - no barrier
                        const double* ax = accx + zpad(base);
#pragma unroll
                        for (int t = 0; t < r; ++t) v[t] = make_double2(ax[zleg<1, r>(base, t)], p[t]);
                        Dft<r>::run(v);
#pragma unroll
                        for (int u = 0; u < r; ++u) sp[zleg<1, r>(base, u)] = v[u];
                    }
                }
            }
        }
    }
    // ---- forward transform of the packed pair (DIT), once per target
    z2_dit_mid<P, S - 2, NT>(work, tw, tid);
    __syncthreads();
    {
        constexpr int r = P::radix(0), M = N / r;
#pragma unroll 1
        for (int j = tid; j < M; j += NT) {
            double2* sp = work + zpad(j);
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
            const double2 zk = work[zpad(k)];
            const double2 zm = work[zpad((N - k) % N)];
            const double2 v1 = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double2 v2 = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
            if (r1 < A.nrows) mfb_st_stream(A.out + r1 * A.out_pitch + k, v1);
            if (r2 < A.nrows) mfb_st_stream(A.out + r2 * A.out_pitch + k, v2);
        }
    }
}
template <class P, bool PF = false> constexpr size_t z_multi2_smem() {
    return size_t(2) * (zpad(P::N) + 1) * sizeof(double2) + (PF ? size_t(2) * (P::N / 3) * sizeof(double2) + 16 : 0);
}

}  // namespace mfb
