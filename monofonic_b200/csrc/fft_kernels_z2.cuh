// fft_kernels_z2.cuh - fused z pass of ONE dealiased product (FP64, sm_100a): z_conv3_kernel, the per-product path of the
// convolver (named products of BaseConvolver, and the source stage when the slot workspace cannot be had).
//
// Per (x,y) row of the padded grid: c2r of both operands, pointwise product, r2c, all in shared memory
// (include/convolution.hh:524-532 collapsed into one pass; the padded real-space fields never touch HBM).
//   * line-major shared memory: each row is ONE contiguous line (1 pad slot per 8 elements), a warp works on
//     32 consecutive butterflies of one row; a CTA holds one product pair;
//   * the Hermitian extension  conj(A^ + iB^)  is built inside the first butterfly stage straight from global memory, the
//     product inside the innermost stages: no separate pack / product sweeps over shared memory;
//   * two-for-one packing, DIF backward -> product in digit-reversed order -> DIT forward, unpacking: no permutation.
// Earlier generations (tile-major gen 1, cp.async-staged gen 2) are in scripts/kbench/kbench_legacy_kernels.cuh.
#pragma once
#include "fft_kernels.cuh"

namespace mfb {

MFB_HD constexpr int zpad(int p) { return p + (p >> 3); }

// offset of leg t (position base + t*M) relative to zpad(base), when that is independent of base
template <int M, int r> MFB_HD constexpr bool zsep() { return (M % 8 == 0) || (M == 1 && r % 8 == 0) || (M == 4 && (r * M) % 8 == 0); }
template <int M, int r> MFB_D int zleg(int base, int t) {
    if constexpr (zsep<M, r>()) return zpad(t * M);  // M%8==0: exact; M==1,r%8==0: base%8==0; M==4: j<4
    else return zpad(base + t * M) - zpad(base);
}

// DIF stage s (s >= 1) of one line, NT threads
template <class P, int s, int NT> MFB_D void z2_dif_stage(double2* __restrict__ ln, const double2* __restrict__ tw, int q) {
    constexpr int r = P::radix(s), L = P::len(s), M = L / r, Bn = P::N / r;
    constexpr bool innermost = (s == P::S - 1);
#pragma unroll 1
    for (int b = q; b < Bn; b += NT) {
        const int blk = b / M, j = b - blk * M, base = blk * L + j;
        double2* sp = ln + zpad(base);
        double2 v[r];
#pragma unroll
        for (int t = 0; t < r; ++t) v[t] = sp[zleg<M, r>(base, t)];
        Dft<r>::run(v);
        if constexpr (!innermost) apply_tw<r>(v, tw + P::tw2_off(s) + j, M);
#pragma unroll
        for (int u = 0; u < r; ++u) sp[zleg<M, r>(base, u)] = v[u];
    }
}
// DIT stage s (S-1 > s >= 1): twiddle, butterfly
template <class P, int s, int NT> MFB_D void z2_dit_stage(double2* __restrict__ ln, const double2* __restrict__ tw, int q) {
    constexpr int r = P::radix(s), L = P::len(s), M = L / r, Bn = P::N / r;
#pragma unroll 1
    for (int b = q; b < Bn; b += NT) {
        const int blk = b / M, j = b - blk * M, base = blk * L + j;
        double2* sp = ln + zpad(base);
        double2 v[r];
#pragma unroll
        for (int t = 0; t < r; ++t) v[t] = sp[zleg<M, r>(base, t)];
        apply_tw<r>(v, tw + P::tw2_off(s) + j, M);
        Dft<r>::run(v);
#pragma unroll
        for (int u = 0; u < r; ++u) sp[zleg<M, r>(base, u)] = v[u];
    }
}
template <class P, int s, int NT> MFB_D void z2_dif_rest(double2* ln, const double2* tw, int q) {
    if constexpr (s < P::S) {
        __syncthreads();
        z2_dif_stage<P, s, NT>(ln, tw, q);
        z2_dif_rest<P, s + 1, NT>(ln, tw, q);
    }
}
template <class P, int s, int NT> MFB_D void z2_dit_mid(double2* ln, const double2* tw, int q) {
    if constexpr (s >= 1) {
        __syncthreads();
        z2_dit_stage<P, s, NT>(ln, tw, q);
        z2_dit_mid<P, s - 1, NT>(ln, tw, q);
    }
}

// ------------------------------------------------------------------------------------------
// Third step (z_conv3_kernel): the fused z pass is bound by shared-memory / L1 bandwidth (about 8600 16-byte
// accesses per row at n = 768 against 512 + 256 HBM accesses), so
//   * the middle DIF stages process BOTH rows of a product pair in one thread: each twiddle is fetched once;
//   * the innermost DIF stage of both rows, the product and the innermost DIT stage of the packed pair are one
//     register-resident step: the backward result is never written back and the product never stored.
template <class P, int s, int NT> MFB_D void z3_dif_stage_pair(double2* __restrict__ l0, const double2* __restrict__ tw, int q) {
    constexpr int r = P::radix(s), L = P::len(s), M = L / r, Bn = P::N / r;
    constexpr int LW = zpad(P::N) + 1;
#pragma unroll 1
    for (int b = q; b < Bn; b += NT) {
        const int blk = b / M, j = b - blk * M, base = blk * L + j;
        double2 w[r];
#if MFB_TW_COMPUTE
        if constexpr (r > 4) tw_powers<r>(tw + P::tw2_off(s) + j, M, w);
        else
#endif
        {
#pragma unroll
            for (int u = 1; u < r; ++u) w[u] = __ldg(tw + P::tw2_off(s) + (u - 1) * M + j);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double2* sp = l0 + h * LW + zpad(base);
            double2 v[r];
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = sp[zleg<M, r>(base, t)];
            Dft<r>::run(v);
#pragma unroll
            for (int u = 1; u < r; ++u) v[u] = cmul(v[u], w[u]);
#pragma unroll
            for (int u = 0; u < r; ++u) sp[zleg<M, r>(base, u)] = v[u];
        }
    }
}
template <class P, int s, int NT> MFB_D void z3_dif_mid_pair(double2* l0, const double2* tw, int q) {
    if constexpr (s < P::S - 1) {
        __syncthreads();
        z3_dif_stage_pair<P, s, NT>(l0, tw, q);
        z3_dif_mid_pair<P, s + 1, NT>(l0, tw, q);
    }
}

// DIRECT: the operand rows are read straight from global memory in the legs of stage 0 instead of being staged in shared
// memory by cp.async: 55 KB instead of 88 KB per CTA at n = 1536 (4 CTAs per SM instead of 2 - the in-pipeline capture
// profiles/r01_ncu_full_z3_1536_in_pipeline.txt shows barrier stalls of 3.3 per issue with two 256-thread CTAs).
// (Prefetching the old output rows of an accumulate launch into shared memory with cp.async was measured and dropped:
// 7.05 vs 6.34 ms at n = 1536, profiles/r01_kbench_z4_3072_and_accbuf.txt - the extra 16 KB per CTA cost more than the tail
// latency.  So was a persistent CTA that stages the NEXT pair's rows under the current pair's stages: 6.42 vs 5.70 ms,
// profiles/r01_kbench_z_persistent_prefetch_negative.txt.  Resident warps matter more here than hidden HBM latency.)
// EARLY_ACC: how an accumulate launch reads the old sum rows.  0: inside the final loop, one element at a time - every
// iteration then waits a full HBM round trip (the SASS view of the in-pipeline capture attributes 32 % of all stall samples
// to the two DADDs behind those loads, profiles/r01_ncu_z3_1536_sass_mix.txt); 1: all of a thread's loads are issued back to
// back after the last barrier; 2: they are issued before the outermost forward stage and consumed after it.
template <class P, int RPC, int MINB, bool DIRECT = false, int EARLY_ACC = 2>
__global__ void __launch_bounds__(P::TPL* RPC, MINB)
    z_conv3_kernel(const double2* __restrict__ Ain, const double2* __restrict__ Bin, double2* __restrict__ out, ptrdiff_t in_pitch,
                   ptrdiff_t out_pitch, long nrows, double scale, const double2* __restrict__ tw, int accumulate = 0) {
    MFB_DYN_SMEM(double2, sm);
    constexpr int N = P::N, S = P::S, TPL = P::TPL, NT = TPL * RPC, PT = 2 * TPL;
    constexpr int NZ = N / 3;
    constexpr int LW = zpad(N) + 1;
    static_assert(P::radix(0) % 3 == 0, "the z pass needs a first radix that is a multiple of 3");
    static_assert(RPC % 2 == 0 && S >= 2, "rows per CTA must be even");
    double2* work = sm;
    double2* raw = sm + RPC * LW;
    const int tid = threadIdx.x;
    const long row0 = long(blockIdx.x) * RPC;

    if constexpr (!DIRECT) {
        for (int idx = tid; idx < RPC * 2 * NZ; idx += NT) {
            const int o = idx / NZ, k = idx - o * NZ;
            const long r = row0 + (o >> 1);
            double2* d = raw + idx;
            if (r < nrows) mfb_cp_async16(d, ((o & 1) ? Bin : Ain) + r * in_pitch + k);
            else *d = make_double2(0.0, 0.0);
        }
        mfb_cp_async_wait_all();
        __syncthreads();
    }

    // ---- DIF stage 0 per row: Hermitian extension of conj(A^ + i B^) built in the butterfly legs
    {
        const int l = tid / TPL, q = tid - l * TPL;
        double2* ln = work + l * LW;
        const bool row_ok = row0 + l < nrows;
        const double2* ra = DIRECT ? Ain + (row0 + l) * in_pitch : raw + (2 * l) * NZ;
        const double2* rb = DIRECT ? Bin + (row0 + l) * in_pitch : ra + NZ;
        auto ld = [&](const double2* p) -> double2 {
            if constexpr (DIRECT) return row_ok ? mfb_ld_stream(p) : make_double2(0.0, 0.0);
            else return *p;
        };
        constexpr int r = P::radix(0), M = N / r, K3 = r / 3;
#pragma unroll 1
        for (int j = q; j < M; j += TPL) {
            double2 v[r];
            if constexpr (DIRECT) {   // all 4*K3 loads of the butterfly are in flight before the first use
                double2 xa[2 * K3], xb[2 * K3];
#pragma unroll
                for (int t = 0; t < K3; ++t) { xa[t] = ld(ra + j + t * M); xb[t] = ld(rb + j + t * M); }
#pragma unroll
                for (int t = 2 * K3; t < r; ++t) {
                    const int k = (r - t) * M - j;
                    const bool dropped = (t == 2 * K3 && j == 0);      // k == NZ: the dropped Nyquist sample (not stored)
                    xa[t - K3] = dropped ? make_double2(0.0, 0.0) : ld(ra + k);
                    xb[t - K3] = dropped ? make_double2(0.0, 0.0) : ld(rb + k);
                }
#pragma unroll
                for (int t = 0; t < K3; ++t) {
                    double2 x1 = xa[t], x2 = xb[t];
                    if (t == 0 && j == 0) { x1.y = 0.0; x2.y = 0.0; }
                    v[t] = make_double2(x1.x - x2.y, -(x1.y + x2.x));
                }
#pragma unroll
                for (int t = K3; t < 2 * K3; ++t) v[t] = make_double2(0.0, 0.0);
#pragma unroll
                for (int t = 2 * K3; t < r; ++t) {
                    const double2 x1 = xa[t - K3], x2 = xb[t - K3];
                    v[t] = make_double2(x1.x + x2.y, -(x2.x - x1.y));
                }
            } else {
#pragma unroll
            for (int t = 0; t < K3; ++t) {
                double2 x1 = ra[j + t * M], x2 = rb[j + t * M];
                if (t == 0 && j == 0) { x1.y = 0.0; x2.y = 0.0; }
                v[t] = make_double2(x1.x - x2.y, -(x1.y + x2.x));
            }
#pragma unroll
            for (int t = K3; t < 2 * K3; ++t) v[t] = make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 2 * K3; t < r; ++t) {
                const int k = (r - t) * M - j;
                if (t == 2 * K3 && j == 0) { v[t] = make_double2(0.0, 0.0); continue; }
                const double2 x1 = ra[k], x2 = rb[k];
                v[t] = make_double2(x1.x + x2.y, -(x2.x - x1.y));
            }
            }
            Dft<r>::run(v);
#pragma unroll
            for (int u = 1; u < r; ++u) v[u] = cmul(v[u], __ldg(tw + (u - 1) * M + j));
            double2* sp = ln + zpad(j);
#pragma unroll
            for (int u = 0; u < r; ++u) sp[zleg<M, r>(j, u)] = v[u];
        }
    }
    // ---- per product pair from here on
    const int m = tid / PT, q = tid - m * PT;
    double2* l0 = work + (2 * m) * LW;
    z3_dif_mid_pair<P, 1, PT>(l0, tw, q);
    __syncthreads();
    {   // innermost DIF stage of both rows -> product -> innermost DIT stage of the packed pair, in registers
        constexpr int r = P::radix(S - 1), Bn = N / r;
#pragma unroll 1
        for (int b = q; b < Bn; b += PT) {
            const int base = b * r;
            double2* sp = l0 + zpad(base);
            double2 v[r];
            double p0[r];
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = sp[zleg<1, r>(base, t)];
            Dft<r>::run(v);
#pragma unroll
            for (int t = 0; t < r; ++t) p0[t] = -v[t].x * v[t].y;  // a = Re, b = -Im (conj trick)
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = sp[LW + zleg<1, r>(base, t)];
            Dft<r>::run(v);
#pragma unroll
            for (int t = 0; t < r; ++t) v[t] = make_double2(p0[t], -v[t].x * v[t].y);
            Dft<r>::run(v);
#pragma unroll
            for (int u = 0; u < r; ++u) sp[zleg<1, r>(base, u)] = v[u];
        }
    }
    z2_dit_mid<P, S - 2, PT>(l0, tw, q);
    constexpr int TOTAL = (RPC / 2) * NZ, PERT = (TOTAL + NT - 1) / NT;   // output pairs per CTA / per thread
    double2 o1[PERT], o2[PERT];
    auto fetch_old = [&]() {
#pragma unroll
        for (int u = 0; u < PERT; ++u) {
            const int idx = tid + u * NT;
            o1[u] = o2[u] = make_double2(0.0, 0.0);
            if (idx < TOTAL) {
                const int mm = idx / NZ, k = idx - mm * NZ;
                const long r1 = row0 + 2 * mm, r2 = r1 + 1;
                if (r1 < nrows) o1[u] = mfb_ld_stream(out + r1 * out_pitch + k);
                if (r2 < nrows) o2[u] = mfb_ld_stream(out + r2 * out_pitch + k);
            }
        }
    };
    if (EARLY_ACC == 2 && accumulate) fetch_old();
    __syncthreads();
    {
        constexpr int r = P::radix(0), M = N / r;
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
    __syncthreads();
    if (EARLY_ACC == 1 && accumulate) fetch_old();
    {
        const double h = 0.5 * scale;
#pragma unroll
        for (int u = 0; u < PERT; ++u) {
            const int idx = tid + u * NT;
            if (idx >= TOTAL) break;
            const int mm = idx / NZ, k = idx - mm * NZ;
            const double2* lz = work + (2 * mm) * LW;
            const double2 zk = lz[zpad(k)];
            const double2 zm = lz[zpad((N - k) % N)];
            const long r1 = row0 + 2 * mm, r2 = r1 + 1;
            double2 v1 = make_double2(h * (zk.x + zm.x), h * (zk.y - zm.y));
            double2 v2 = make_double2(h * (zk.y + zm.y), -h * (zk.x - zm.x));
            if (accumulate) {   // sum of products of one target, formed before the (single) forward y/x passes
                if constexpr (EARLY_ACC != 0) {
                    v1.x += o1[u].x; v1.y += o1[u].y; v2.x += o2[u].x; v2.y += o2[u].y;
                } else {
                    if (r1 < nrows) { const double2 o = mfb_ld_stream(out + r1 * out_pitch + k); v1.x += o.x; v1.y += o.y; }
                    if (r2 < nrows) { const double2 o = mfb_ld_stream(out + r2 * out_pitch + k); v2.x += o.x; v2.y += o.y; }
                }
            }
            if (r1 < nrows) mfb_st_stream(out + r1 * out_pitch + k, v1);
            if (r2 < nrows) mfb_st_stream(out + r2 * out_pitch + k, v2);
        }
    }
}

template <class P, int RPC> constexpr size_t z_conv3_direct_smem() { return size_t(RPC) * (zpad(P::N) + 1) * sizeof(double2); }

}  // namespace mfb
