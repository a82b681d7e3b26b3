// fft_kernels_zf.cuh - shared-field z pass of the LPT source stage (FP64, sm_100a).
//
// The multi-product pass (fft_kernels_zm.cuh) transforms the operands of every PRODUCT: the twelve Hessian components of phi
// and phi2 that the last third-order stage multiplies (src/ic_generator.cc:533-556: 3b and the three components of the
// transverse term) go through 36 real c2r transforms per row for 18 products.  Here every distinct real FIELD of a stage is
// transformed exactly once per row, all of them stay in shared memory, the targets are formed pointwise from the resident
// real lines and transformed forward two at a time:
//
//     stage A (:482-486, :522-526)  6 fields  phi,ab              -> phi2 source, I1..I4     3 + 3  complex FFTs per row (9.5 before)
//     stage C (:533-556)           12 fields  phi,ab and phi2,ab  -> 3b, A_x, A_y, A_z       6 + 2  complex FFTs per row (20 before)
//
// and every base row is read from HBM once (12 + 4 rows moved per row of stage C instead of 33 + 4).
//
// One CTA per SM holds the NL = NF/2 complex work lines of ONE row (two real fields per complex transform, the conj trick of
// fft_kernels_z2.cuh).  The CTA is G groups of NTG threads; a group is what a CTA of z_multi2_kernel was: it owns NL/G lines,
// runs their backward transforms one after the other with barriers of its own (bar.sync id, NTG), and keeps one TMA bulk copy
// (cp.async.bulk + mbarrier) of its next pair of base rows in flight in its 2-row staging area.  Two CTA-wide barriers per row
// frame the pointwise step.  The forward transforms run in the groups' own first lines, so a group that has no target pair to
// transform (stage C: 2 pairs, 3 groups) is already in the next row's backward transforms.  Persistent CTAs walk the rows with
// stride gridDim.x, so the staging pipeline of a group never drains between rows.
#pragma once
#include "fft_kernels_zm.cuh"

namespace mfb {

constexpr int ZF_MAXF = 12, ZF_MAXT = 6;
enum { ZF_PROG_A = 0, ZF_PROG_C = 1 };

struct ZfField {
    const double2* p;   // base field chunk: rows of `in_pitch` complex, retained k_z modes only
    double c;           // coefficient
    int e;              // power of k_z: 0, 1, 2
    int pad_;
};
struct ZfArgs {
    ZfField f[ZF_MAXF];      // real field i = c2r_z[ c k_z^e F ]; fields 2l and 2l+1 share complex line l
    double2* out[ZF_MAXT];   // target chunks: rows of `out_pitch` complex
    ptrdiff_t in_pitch, out_pitch;
    long nrows;
    double kfz;              // 2 pi / L_z
    double coef[4];          // program constants
};

// field order of both programs: xx yy zz xy xz yz (of phi; stage C: then the same of phi2)
template <int PROG> struct ZfProg;
template <> struct ZfProg<ZF_PROG_A> {
    static constexpr int NF = 6, NT = 5;
    // phi2 source (src/ic_generator.cc:482-486) and the inner products of the determinant (:522-526, convolve3 truncates
    // between inner and outer product, include/convolution.hh:542-549): I1 = xx yy - xy^2, I2 = xy xz, I3 = yz^2, I4 = xz^2
    static MFB_HD void run(const double* f, const double*, double* t) {
        const double xx = f[0], yy = f[1], zz = f[2], xy = f[3], xz = f[4], yz = f[5];
        const double sxy = xy * xy, sxz = xz * xz, syz = yz * yz;
        const double xxyy = xx * yy;
        t[0] = fma(xx + yy, zz, xxyy) - sxy - sxz - syz;
        t[1] = xxyy - sxy;
        t[2] = xy * xz;
        t[3] = syz;
        t[4] = sxz;
    }
};
template <> struct ZfProg<ZF_PROG_C> {
    static constexpr int NF = 12, NT = 4;
    // 3b (:533-538; coef[0] = -5/7 on the diagonal terms, coef[1] = 10/7 on the mixed ones) and the transverse term
    // A_d = phi2,dp phi,dq - phi2,dq phi,dp + phi,pq (phi2,pp - phi2,qq) - phi2,pq (phi,pp - phi,qq),  (d, p, q) cyclic (:545-556)
    static MFB_HD void run(const double* f, const double* cf, double* t) {
        const double xx = f[0], yy = f[1], zz = f[2], xy = f[3], xz = f[4], yz = f[5];
        const double XX = f[6], YY = f[7], ZZ = f[8], XY = f[9], XZ = f[10], YZ = f[11];
        const double dg = fma(xx, YY + ZZ, fma(yy, ZZ + XX, zz * (XX + YY)));
        const double mx = fma(xy, XY, fma(xz, XZ, yz * YZ));
        t[0] = fma(cf[0], dg, cf[1] * mx);
        t[1] = fma(XY, xz, fma(-XZ, xy, fma(yz, YY - ZZ, -YZ * (yy - zz))));
        t[2] = fma(YZ, xy, fma(-XY, yz, fma(xz, ZZ - XX, -XZ * (zz - xx))));
        t[3] = fma(XZ, yz, fma(-YZ, xz, fma(xy, XX - YY, -XY * (xx - yy))));
    }
};

// barrier of one group (PTX named barriers 1..15; 0 is __syncthreads): ids 1..G for the groups, 8.. for the forward teams
#if !defined(MFB_EMU)
MFB_D void zf_group_sync(int g, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nthreads) : "memory"); }
#else
inline void zf_group_sync(int g, int nthreads) { mfb_emu::bar_sync(g + 1, nthreads); }
#endif
constexpr int ZF_TEAM_BAR = 7;   // zf_group_sync(ZF_TEAM_BAR + team, ..) -> barrier ids 8, 9, 10

// the remaining DIF stages of one line (s >= 1): a group barrier after each stage but the innermost
template <class P, int s, int NTG> MFB_D void zf_dif_rest(double2* ln, const double2* tw, int q, int g) {
    if constexpr (s < P::S) {
        z2_dif_stage<P, s, NTG>(ln, tw, q);
        if constexpr (s < P::S - 1) zf_group_sync(g, NTG);
        zf_dif_rest<P, s + 1, NTG>(ln, tw, q, g);
    }
}
// the middle DIT stages S-2 .. 1, a group barrier before each
template <class P, int s, int NTG> MFB_D void zf_dit_mid(double2* ln, const double2* tw, int q, int g) {
    if constexpr (s >= 1) {
        zf_group_sync(g, NTG);
        z2_dit_stage<P, s, NTG>(ln, tw, q);
        zf_dit_mid<P, s - 1, NTG>(ln, tw, q, g);
    }
}

// every forward team contains a whole group (whose first line is the team's work line)
constexpr bool zf_teams_ok(int ft, int ntg, int nfw) {
    for (int j = 0; j < nfw; ++j)
        if (((j * ft + ntg - 1) / ntg + 1) * ntg > (j + 1) * ft) return false;
    return true;
}

// COOP: the forward transforms are shared out over ALL threads - NFW teams of NTH / NFW threads, one target pair each - instead
// of one group per pair (which leaves G - NFW groups without forward work: they run ahead into the next row's backward
// transforms and then wait).  A team's work line is the first line of the first group that lies entirely inside the team, so
// that only that group's own next stage 0 overwrites it, after the team's last barrier.
template <class P, int NTG, int G, int PROG, bool PF, bool COOP = false>
__global__ void __launch_bounds__(NTG* G, 1) z_fields_kernel(ZfArgs A, const double2* __restrict__ tw) {
    MFB_DYN_SMEM(double2, sm);
    using PR = ZfProg<PROG>;
    constexpr int N = P::N, S = P::S, NZ = N / 3, LW = zpad(N) + 1;
    constexpr int NF = PR::NF, NL = NF / 2, LPG = NL / G, NTGT = PR::NT, NFW = (NTGT + 1) / 2, NTH = NTG * G;
    static_assert(NL % G == 0 && NFW <= G, "lines must divide over the groups, one target pair per group at most");
    constexpr int FT = COOP ? NTH / NFW : NTG;   // threads per forward team
    static_assert(FT % 32 == 0 && FT * NFW <= NTH && zf_teams_ok(FT, NTG, NFW), "forward teams are whole warps and hold a whole group");
    auto fwd_line = [](int j) { return ((j * FT + NTG - 1) / NTG) * LPG; };   // first line of the first group inside team j
    static_assert(P::radix(0) % 3 == 0 && S >= 2, "the z pass needs a first radix that is a multiple of 3");
    const int tid = threadIdx.x, g = tid / NTG, q = tid - g * NTG;
    double2* lines = sm;                                        // NL padded complex lines
    double2* raw = sm + NL * LW + size_t(g) * 2 * NZ;           // PF: this group's staged pair of base rows
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + NL * LW + size_t(G) * 2 * NZ) + g;
    const long rstep = gridDim.x;

    auto prefetch = [&](long row, int li) {   // one thread of the group: bulk copies of the two base rows of line g*LPG + li
        const int l = g * LPG + li;
        constexpr unsigned ROWB = unsigned(NZ) * unsigned(sizeof(double2));
        mfb_mbar_expect_tx(bar, 2u * ROWB);
        mfb_bulk_g2s(raw, A.f[2 * l].p + row * A.in_pitch, ROWB, bar);
        mfb_bulk_g2s(raw + NZ, A.f[2 * l + 1].p + row * A.in_pitch, ROWB, bar);
    };
    unsigned phase = 0;
    if constexpr (PF) {
        if (q == 0) {
            mfb_mbar_init(bar, 1);
            if (long(blockIdx.x) < A.nrows) prefetch(blockIdx.x, 0);
        }
        __syncthreads();
    }

#pragma unroll 1
    for (long row = blockIdx.x; row < A.nrows; row += rstep) {
        // ---- backward transforms of this group's lines
#pragma unroll 1
        for (int li = 0; li < LPG; ++li) {
            const int l = g * LPG + li;
            double2* work = lines + l * LW;
            const ZfField fa = A.f[2 * l], fb = A.f[2 * l + 1];
            if constexpr (PF) { mfb_mbar_wait(bar, phase); phase ^= 1u; }
            {   // DIF stage 0: conj(A^ + i B^), Hermitian-extended and zero padded, built in the legs (as in z_multi2_kernel)
                constexpr int r = P::radix(0), M = N / r, K3 = r / 3;
                const double2* pa = fa.p + row * A.in_pitch;
                const double2* pb = fb.p + row * A.in_pitch;
                auto kpow = [&](double kz, int e, double c) -> double { return e == 0 ? c : (e == 1 ? c * kz : c * kz * kz); };
#pragma unroll 1
                for (int j = q; j < M; j += NTG) {
                    double2 xa[2 * K3], xb[2 * K3];
                    auto mode_k = [&](int m) -> int { return m < K3 ? j + m * M : (2 * K3 - m) * M - j; };   // leg m <-> mode k
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) {
                        const bool dropped = (m == K3 && j == 0);   // k == NZ: the dropped Nyquist sample (not stored)
                        const int k = mode_k(m);
                        if constexpr (PF) {
                            xa[m] = dropped ? make_double2(0.0, 0.0) : raw[k];
                            xb[m] = dropped ? make_double2(0.0, 0.0) : raw[NZ + k];
                        } else {
                            xa[m] = dropped ? make_double2(0.0, 0.0) : mfb_ld_stream(pa + k);
                            xb[m] = dropped ? make_double2(0.0, 0.0) : mfb_ld_stream(pb + k);
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 2 * K3; ++m) {
                        const double kz = A.kfz * double(mode_k(m));
                        const double f0 = kpow(kz, fa.e, fa.c), f1 = kpow(kz, fb.e, fb.c);
                        xa[m] = make_double2(f0 * xa[m].x, f0 * xa[m].y);
                        xb[m] = make_double2(f1 * xb[m].x, f1 * xb[m].y);
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
            zf_group_sync(g, NTG);   // stage 0 done; the staged rows are consumed: the TMA unit may overwrite them
            if constexpr (PF) {
                if (q == 0) {
                    if (li + 1 < LPG) prefetch(row, li + 1);
                    else if (row + rstep < A.nrows) prefetch(row + rstep, 0);
                }
#ifdef MFB_EMU
                zf_group_sync(g, NTG);   // emulation: the "asynchronous" copy is a memcpy of the issuing fiber
#endif
            }
            zf_dif_rest<P, 1, NTG>(work, tw, q, g);   // middle stages and the innermost one, a group barrier after each but the last
        }
        __syncthreads();   // every real line of the row is in place: Re = field 2l, -Im = field 2l+1, digit-reversed order
        // ---- targets, pointwise; pair j goes to the first line of group j as (t_2j, t_2j+1)
        {
            double cf[4] = {A.coef[0], A.coef[1], A.coef[2], A.coef[3]};
#pragma unroll 1
            for (int p = tid; p < N; p += NTH) {
                const int o = zpad(p);
                double f[NF], t[2 * NFW];
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const double2 v = lines[l * LW + o];
                    f[2 * l] = v.x;
                    f[2 * l + 1] = -v.y;
                }
                if constexpr (NTGT < 2 * NFW) t[2 * NFW - 1] = 0.0;
                PR::run(f, cf, t);
#pragma unroll
                for (int j = 0; j < NFW; ++j) lines[fwd_line(j) * LW + o] = make_double2(t[2 * j], t[2 * j + 1]);
            }
        }
        __syncthreads();
        // ---- forward transform of this group's target pair (DIT), unpack, store
        const int team = COOP ? tid / FT : g, qf = COOP ? tid - team * FT : q, tbar = COOP ? ZF_TEAM_BAR + team : g;
        if (team < NFW) {
            double2* work = lines + fwd_line(team) * LW;
            z2_dif_stage<P, S - 1, FT>(work, tw, qf);   // innermost stage: no twiddles, DIT == DIF
            zf_dit_mid<P, S - 2, FT>(work, tw, qf, tbar);
            zf_group_sync(tbar, FT);
            {
                constexpr int r = P::radix(0), M = N / r;
#pragma unroll 1
                for (int j = qf; j < M; j += FT) {
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
            zf_group_sync(tbar, FT);
            double2* o1 = A.out[2 * team] + row * A.out_pitch;
            double2* o2 = (2 * team + 1 < NTGT) ? A.out[2 * team + 1] + row * A.out_pitch : nullptr;
            for (int k = qf; k < NZ; k += FT) {
                const double2 zk = work[zpad(k)];
                const double2 zm = work[zpad((N - k) % N)];
                mfb_st_stream(o1 + k, make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y)));
                if (o2) mfb_st_stream(o2 + k, make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x)));
            }
            zf_group_sync(tbar, FT);   // the line is free for the next row's stage 0
        }
    }
}
template <class P, int G, int PROG, bool PF> constexpr size_t z_fields_smem() {
    return size_t(ZfProg<PROG>::NF / 2) * (zpad(P::N) + 1) * sizeof(double2) + (PF ? size_t(G) * 2 * (P::N / 3) * sizeof(double2) : 0) + size_t(G) * 8 + 8;
}

}  // namespace mfb
