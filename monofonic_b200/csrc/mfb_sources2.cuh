// mfb_sources2.cuh - second-generation LPT source stage ("v2"): phi2, phi3, A3 of src/ic_generator.cc:475-574 with the
// transforms shared between the products (exact by linearity):
//
//   * X buffers.  Every operand of the stage is a second derivative (or a sum / difference of two) of phi or phi2, or a
//     plain inner product I.  Its padded backward x pass depends only on the power of k_x it carries, so ONE x pass per
//     (potential, mx in {0,1,2}) serves all of them: 10 x passes (and, on slabs, 10 exchanges) per 3LPT step instead of 49.
//     The result ([P0loc][N1][N2/2] complex, "U" bytes) is kept in a slot of the workspace.
//   * x-plane chunks.  The y and z passes are independent per x plane.  For a chunk of planes the y pass turns the X
//     buffers into the base fields F_{mx,my} = bwd_xy[k_x^mx k_y^my f^] (6 per potential), which all fit in a small chunk
//     workspace at the same time; the multi-product z pass (fft_kernels_zm.cuh) applies the k_z part of each factor while
//     loading, sums ALL products of a target in real space and transforms forward once; the forward y pass brings the
//     chunk of the target back to N1 modes.  Several targets share one loop (and its y passes) when slots allow.
//   * the forward x pass (truncation, unpad scale, Nyquist = 0, output operator) runs once per target after the loop.
//
// 3LPT: 10 x passes, 6 + 8 + 24 backward y passes (49 + 49 before), 36.5 complex z FFTs per row (48 before), 10 forward
// y/x passes; the intermediate truncation of the three-field products (include/convolution.hh:542-549) is kept: the inner
// products I1..I4 are ordinary targets of the first loop.  Needs >= 7 U-sized slots (own HBM + the idle particle arrays);
// with less memory mfb_lpt_sources falls back to the per-product path (ConvGroup in mfb_api.cu).
#pragma once

namespace mfb {

// multi-product z pass (z_multi2_kernel, fft_kernels_zm.cuh): N, threads per CTA, min CTAs/SM, radices (first radix a multiple
// of 3, last radix 8 or 16: conflict-free padded lines).  One work line + the accumulators = 2 complex lines per CTA.
// n = 1536 on B200 (profiles/r02_kbench_zm_variants.txt, ms per complex-FFT sweep of 1536^2 rows): pair-parallel version
// (246 regs, 8 warps, now scripts/kbench/kbench_legacy_kernels.cuh) 22.2; sequential rows 64 threads x 4 CTAs 21.1; 128 x 4
// (128 regs, 16 warps) 20.9; 128 x 3 (164 regs, 12 warps) 19.1; 256 x 2 24.8; 96 x 4 21.4.  n = 3072: 192 x 2 and 256 x 2 equal.
// PF_: rows staged by TMA bulk copies (z_multi2_kernel<.., PF = true>) where the 2-row staging area does not cost a resident
// CTA (n = 3072: two lines + staging = 143 KB would leave one CTA per SM); mfb_set_option("zm_prefetch", 0) selects the
// direct-load kernel everywhere (A/B runs, check_kernel_generations).
#define MFB_FOR_EACH_PLANZM2(X)      \
    X(48, 32, 1, 1, 6, 8)            \
    X(96, 32, 1, 1, 12, 8)           \
    X(192, 32, 2, 1, 6, 4, 8)        \
    X(384, 64, 4, 1, 6, 8, 8)        \
    X(768, 64, 4, 1, 12, 8, 8)       \
    X(1536, 128, 3, 1, 12, 16, 8)    \
    X(3072, 192, 2, 0, 12, 16, 16)

template <class P, int NT, int MINB, bool PF> void launch_zmulti2_t(const ZmArgs& a, double bytes) {
    static bool done = false;
    constexpr size_t smem = z_multi2_smem<P, PF>();
    optin_smem(z_multi2_kernel<P, NT, MINB, PF>, smem, done);
    const unsigned grid = unsigned((a.nrows + 1) / 2);
    if (grid == 0) return;
    const double2* tw = twiddles2<P>();
    RoleScope rs(ROLE_CONV_Z, bytes);
    MFB_LAUNCH((z_multi2_kernel<P, NT, MINB, PF>), grid, NT, smem, g_ctx.stream, a, tw);
    check_launch("z_multi2_kernel");
}
static void launch_zmulti(int n, const ZmArgs& a) {
    // algorithmic bytes: every operand-term row read once per product + the output row
    double rows_read = 0;
    for (int i = 0; i < a.nprod; ++i) rows_read += a.prod[i].a.n + (a.prod[i].square ? 0 : a.prod[i].b.n);
    const double bytes = double(a.nrows) * double(n / 3) * 16.0 * (rows_read + 1.0);
    switch (n) {
#define X(N_, NT_, MINB_, PF_, ...)                                                                              \
    case N_:                                                                                                     \
        if (PF_ && g_ctx.zm_prefetch) launch_zmulti2_t<Plan<N_, NT_, 2, __VA_ARGS__>, NT_, MINB_, PF_ != 0>(a, bytes); \
        else launch_zmulti2_t<Plan<N_, NT_, 2, __VA_ARGS__>, NT_, MINB_, false>(a, bytes);                        \
        return;
        MFB_FOR_EACH_PLANZM2(X)
#undef X
        default: throw std::runtime_error("unsupported padded length " + std::to_string(n) + " for the multi-product z pass");
    }
}

// shared-field z pass (z_fields_kernel, fft_kernels_zf.cuh): N, threads per group (three groups per CTA, one CTA per SM),
// radices as above.  n = 3072 has no plan: six work lines of 55 KB do not fit one SM (3LPT 2048^3 keeps the multi-product pass).
#define MFB_FOR_EACH_PLANZF(X)   \
    X(48, 32, 6, 8)              \
    X(96, 32, 12, 8)             \
    X(192, 32, 6, 4, 8)          \
    X(384, 64, 6, 8, 8)          \
    X(768, 64, 12, 8, 8)         \
    X(1536, 128, 12, 16, 8)
constexpr int ZF_GROUPS = 3;
// stage A: three groups of NTG threads, one line and one target pair each, rows staged by TMA.  Stage C: SIX groups of NTG / 2
// threads, one line each, direct loads (six staging areas do not fit beside six lines) - the four groups without a target pair
// run ahead into the next row (profiles/r02_kbench_zf_variants.txt, n = 1536, ms per launch of 147456 rows: 6 x 64 direct 8.94,
// 3 x 128 TMA 9.50, 3 x 128 direct 9.79, 2 x 128 TMA 10.18; forward transforms shared out over all threads - COOP - 9.59 / 9.73)
constexpr int zf_ntg_c(int ntg) { return ntg >= 64 ? ntg / 2 : 32; }

static bool zfields_has_plan(int n) {
    switch (n) {
#define X(N_, ...) case N_:
        MFB_FOR_EACH_PLANZF(X)
#undef X
        return true;
        default: return false;
    }
}
template <class P, int NTG, int G, int PROG, bool PF, bool COOP> void launch_zfields_t(const ZfArgs& a, double bytes) {
    static bool done = false;
    constexpr size_t smem = z_fields_smem<P, G, PROG, PF>();
    optin_smem(z_fields_kernel<P, NTG, G, PROG, PF, COOP>, smem, done);
    if (a.nrows <= 0) return;
    const unsigned grid = unsigned(std::min<long>(a.nrows, g_ctx.num_sms));   // persistent: one CTA per SM walks the rows
    const double2* tw = twiddles2<P>();
    RoleScope rs(ROLE_CONV_Z, bytes);
    MFB_LAUNCH((z_fields_kernel<P, NTG, G, PROG, PF, COOP>), grid, NTG * G, smem, g_ctx.stream, a, tw);
    check_launch("z_fields_kernel");
}
static void launch_zfields(int n, int prog, const ZfArgs& a) {
    // algorithmic bytes: every base row read once + one row written per target
    const int nf = prog == ZF_PROG_A ? ZfProg<ZF_PROG_A>::NF : ZfProg<ZF_PROG_C>::NF;
    const int nt = prog == ZF_PROG_A ? ZfProg<ZF_PROG_A>::NT : ZfProg<ZF_PROG_C>::NT;
    const double bytes = double(a.nrows) * double(n / 3) * 16.0 * double(nf + nt);
    switch (n) {
#define X(N_, NTG_, ...)                                                                                   \
    case N_:                                                                                               \
        if (prog == ZF_PROG_A) launch_zfields_t<Plan<N_, NTG_, 2, __VA_ARGS__>, NTG_, ZF_GROUPS, ZF_PROG_A, true, false>(a, bytes);  \
        else launch_zfields_t<Plan<N_, NTG_, 2, __VA_ARGS__>, zf_ntg_c(NTG_), 6, ZF_PROG_C, false, false>(a, bytes);                 \
        return;
        MFB_FOR_EACH_PLANZF(X)
#undef X
        default: throw std::runtime_error("unsupported padded length " + std::to_string(n) + " for the shared-field z pass");
    }
}

// ------------------------------------------------------------------------------------------ workspace modes
static size_t conv_slot_elems(const mfb_conv* c) { return size_t(c->p0loc) * c->n[1] * (c->n[2] / 2); }

static void conv_free_slots(mfb_conv* c) {
    for (int i = 0; i < c->nxs; ++i)
        if (c->xs[i].owned && c->xs[i].p) cudaFree(c->xs[i].p);
    c->nxs = 0;
    if (c->chunk && c->chunk_owned) cudaFree(c->chunk);
    c->chunk = nullptr; c->chunk_elems = 0; c->chunk_owned = 0;
}
static void conv_free_legacy(mfb_conv* c, bool keep_w1) {
    if (c->W1 && !keep_w1) { cudaFree(c->W1); c->W1 = nullptr; }
    if (c->W2a) { cudaFree(c->W2a); c->W2a = nullptr; }
    if (c->W2b) { cudaFree(c->W2b); c->W2b = nullptr; }
    if (c->Wacc && !c->acc_borrowed) cudaFree(c->Wacc);
    c->Wacc = nullptr; c->acc_borrowed = false;
    if (c->fbuf) { mfb_grid_destroy(c->fbuf); c->fbuf = nullptr; }
    c->legacy_ready = 0;
    conv_cache_clear(c);
}
// the per-product workspace (what OrszagConvolver's constructor owns, include/convolution.hh:463-506, pruned): W1, W2a, W2b,
// fbuf, and the accumulator of product groups when HBM is plentiful - or borrowed from the lender's particle arrays
static void conv_legacy_ensure(mfb_conv* c) {
    if (c->naive) return;
    if (c->mode == 1) {
        cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
        conv_free_slots(c);   // W1 survives: it is slot 0 in single-slab mode
        c->mode = 0;
    }
    if (c->legacy_ready && (c->Wacc || !c->lender)) return;
    const size_t nzin = c->n[2] / 2;
    c->w1_elems = c->dist ? 0 : size_t(c->np[0]) * c->n[1] * nzin;   // slab mode: the exchange buffers play W1's role
    c->w2_elems = size_t(c->p0loc) * c->np[1] * nzin;
    if (!c->dist && !c->W1) c->W1 = (double2*)dev_alloc(c->w1_elems * sizeof(double2));
    if (!c->W2a) c->W2a = (double2*)dev_alloc(c->w2_elems * sizeof(double2));
    if (!c->W2b) c->W2b = (double2*)dev_alloc(c->w2_elems * sizeof(double2));
    if (!c->fbuf && mfb_grid_create(&c->fbuf, c->n, c->L, MFB_KSPACE)) throw std::runtime_error(g_err);
    if (!c->Wacc) {
        // the accumulator of product groups is an optimisation, not a requirement: take it only if HBM is plentiful after it
        // (room for the particle arrays and the remaining grids of a 3LPT run: 3.75 R(N) of this slab)
        size_t free_b = 0, total_b = 0;
        cuda_check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        const size_t w2_bytes = c->w2_elems * sizeof(double2);
        const size_t reserve = size_t(3.75 * double(c->fbuf->ntot) * sizeof(double));
        if (g_ctx.conv_own_accumulator && free_b > w2_bytes + reserve) {
            try { c->Wacc = (double2*)dev_alloc(w2_bytes); }
            catch (...) { c->Wacc = nullptr; }   // free memory was not contiguous: stay ungrouped
        } else if (c->lender && c->lender->posvel_bytes >= w2_bytes) {
            c->Wacc = (double2*)c->lender->pos;
            c->acc_borrowed = true;
        }
    }
    c->legacy_ready = 1;
}

// Switch to the slot workspace.  Returns the number of U-sized slots (0: not enough memory for `min_slots`, the per-product
// workspace is back in place).
static int conv_enter_slots(mfb_conv* c, int want_slots, int min_slots, size_t want_chunk_elems, size_t min_chunk_elems) {
    if (c->mode == 1 && c->nxs >= min_slots && c->chunk_elems >= min_chunk_elems) return c->nxs;
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    if (c->mode == 1) conv_free_slots(c);
    conv_free_legacy(c, true);
    c->mode = 1;
    c->nxs = 0;
    const size_t U = conv_slot_elems(c) * sizeof(double2);
    auto add = [&](double2* p, int owned, int xslot, int lent = 0) {
        mfb_conv_slot& s = c->xs[c->nxs++];
        s.p = p; s.owned = owned; s.xslot = xslot; s.g = nullptr; s.mx = 0; s.valid = 0; s.pinned = 0; s.stamp = 0; s.lent = lent; s.zb = 0;
    };
    const size_t margin = size_t(3) << 29;   // 1.5 GiB stay free
    size_t lent_left = 0; char* lent_p = nullptr;
    if (c->lender) { lent_left = c->lender->posvel_bytes; lent_p = (char*)c->lender->pos; }
    if (c->dist) {
        DistState& d = D();
        // every exchange buffer of the block, the ping-pong pair of the 3-D FFTs included: no other exchange runs inside the
        // source stage, and barriers at its start and end separate it from the neighbouring transforms
        for (int k = 0; k < d.nxbuf && c->nxs < MFB_CONV_MAX_SLOTS; ++k) add(d.xbuf[k], 0, k);
    } else {
        if (!c->W1) {
            c->w1_elems = conv_slot_elems(c);
            try { c->W1 = (double2*)dev_alloc(U); } catch (...) { c->W1 = nullptr; }
        }
        if (c->W1) add(c->W1, 0, -1);
        while (lent_left >= U && c->nxs < want_slots) { add((double2*)lent_p, 0, -1, 1); lent_p += U; lent_left -= U; }
    }
    // chunk workspace: what is left of the lent particle memory if it is enough, else an allocation of its own
    const size_t min_chunk = min_chunk_elems * sizeof(double2), want_chunk = want_chunk_elems * sizeof(double2);
    c->chunk_lent = 0;
    // where HBM is plentiful (1024^3 on 4-8 GPUs) the chunk workspace is an allocation of its own even if lent memory would
    // do: a lent chunk makes the whole stage wait for the previous run's particle hand-off
    bool own_chunk_fits = false;
    {
        size_t free_b = 0, total_b = 0;
        cuda_check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        own_chunk_fits = g_ctx.conv_own_accumulator && free_b > want_chunk + margin + total_b / 8;   // option 0: as if HBM were tight
    }
    if (!own_chunk_fits && lent_left >= min_chunk && (c->dist || lent_left >= want_chunk)) {
        c->chunk = (double2*)lent_p; c->chunk_elems = lent_left / sizeof(double2); c->chunk_owned = 0; c->chunk_lent = 1;
    } else {
        size_t free_b = 0, total_b = 0;
        cuda_check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
        size_t take = want_chunk;
        if (free_b < take + margin) take = free_b > margin + min_chunk ? std::max(min_chunk, (free_b - margin) / 2) : 0;
        if (take >= min_chunk) {
            try { c->chunk = (double2*)dev_alloc(take); c->chunk_elems = take / sizeof(double2); c->chunk_owned = 1; }
            catch (...) { c->chunk = nullptr; c->chunk_elems = 0; }
        }
    }
    if (!c->dist) {
        while (c->nxs < want_slots && c->nxs < MFB_CONV_MAX_SLOTS) {
            size_t free_b = 0, total_b = 0;
            cuda_check(cudaMemGetInfo(&free_b, &total_b), "cudaMemGetInfo");
            if (free_b < U + margin) break;
            double2* p = nullptr;
            try { p = (double2*)dev_alloc(U); } catch (...) { break; }
            add(p, 1, -1);
        }
    }
    if (g_ctx.sources_max_slots > 0 && c->nxs > g_ctx.sources_max_slots) {   // tests: behave as under memory pressure
        for (int i = g_ctx.sources_max_slots; i < c->nxs; ++i)
            if (c->xs[i].owned) cudaFree(c->xs[i].p);
        c->nxs = g_ctx.sources_max_slots;
    }
    if (c->nxs < min_slots || c->chunk == nullptr) {
        conv_free_slots(c);
        c->mode = 0;
        conv_legacy_ensure(c);
        return 0;
    }
    return c->nxs;
}

// ------------------------------------------------------------------------------------------ the stage as data
struct S2Base { mfb_grid* g; int mx, my; };
struct S2Term { mfb_grid* g; int mx, my, e; double c; };
struct S2Operand { S2Term t[2]; int n; };
struct S2Product { S2Operand a, b; double coeff; bool square; };
struct S2Target {
    std::vector<S2Product> prods;
    mfb_grid* out; int op; double alpha;
};

// (i k_a)(i k_b) f = -k_a k_b f   (include/convolution.hh:137-160; gradient() zeroes its own Nyquist, and the pad drops every
// Nyquist plane anyway)
static S2Term s2_hess_term(mfb_grid* g, int a, int b, double c) {
    S2Term t{g, 0, 0, 0, c};
    for (int d : {a, b}) { if (d == 0) ++t.mx; else if (d == 1) ++t.my; else if (d == 2) ++t.e; }
    return t;
}
static S2Operand s2_hess(mfb_grid* g, int a, int b) { S2Operand o; o.n = 1; o.t[0] = s2_hess_term(g, a, b, -1.0); o.t[1] = o.t[0]; return o; }
// f,ab + sgn * f,cd   (convolve_SumOfHessians / convolve_DifferenceOfHessians, include/convolution.hh:214-274)
static S2Operand s2_hess2(mfb_grid* g, int a, int b, int cc, int d, double sgn) {
    S2Operand o; o.n = 2; o.t[0] = s2_hess_term(g, a, b, -1.0); o.t[1] = s2_hess_term(g, cc, d, -sgn); return o;
}
// the six Hessian components of a potential in the field order of the shared-field programs: xx yy zz xy xz yz
static std::vector<S2Term> s2_hess_fields(mfb_grid* g) {
    return {s2_hess_term(g, 0, 0, -1.0), s2_hess_term(g, 1, 1, -1.0), s2_hess_term(g, 2, 2, -1.0),
            s2_hess_term(g, 0, 1, -1.0), s2_hess_term(g, 0, 2, -1.0), s2_hess_term(g, 1, 2, -1.0)};
}
static S2Operand s2_plain(mfb_grid* g) { S2Operand o; o.n = 1; o.t[0] = S2Term{g, 0, 0, 0, 1.0}; o.t[1] = o.t[0]; return o; }
static bool s2_same_term(const S2Term& x, const S2Term& y) {
    return x.g == y.g && x.mx == y.mx && x.my == y.my && x.e == y.e && x.c == y.c;
}
static S2Product s2_prod(const S2Operand& a, const S2Operand& b, double coeff) {
    S2Product p{a, b, coeff, false};
    p.square = a.n == b.n && s2_same_term(a.t[0], b.t[0]) && (a.n == 1 || s2_same_term(a.t[1], b.t[1]));
    return p;
}

struct S2Runner {
    mfb_conv* c;
    int N0, N1, N2, P0, P1, P2, nzc, nzin, p0l;
    double s1, unpad_scale;
    long n_x = 0, n_y = 0, n_z = 0, n_loops = 0;
    // Slots carved from the lender's particle arrays may still be read by streamed hand-off copies of the previous run
    // (mfb_particles_stream_out).  `limit` < nxs: such slots are left alone (restricted stage - the hand-off drains under it);
    // otherwise the compute stream waits for the copy when the slot is first taken, not at the start of the stage.
    int limit = 0;
    explicit S2Runner(mfb_conv* c_) : c(c_) {
        N0 = c->n[0]; N1 = c->n[1]; N2 = c->n[2]; P0 = c->np[0]; P1 = c->np[1]; P2 = c->np[2];
        nzc = N2 / 2 + 1; nzin = N2 / 2; p0l = c->p0loc;
        const double ptot = double(P0) * P1 * P2, ntot = double(N0) * N1 * N2;
        s1 = std::pow(1.5, 1.5) / std::sqrt(ptot);                                       // pad factor (convolution.hh:572) * bwd norm
        unpad_scale = (1.0 / std::sqrt(ptot)) / (std::sqrt(ptot) / std::sqrt(ntot));     // fwd norm / rfac (convolution.hh:620,664)
    }
    size_t plane_elems() const { return size_t(P1) * nzin; }

    // ---- slots
    int find_x(const mfb_grid* g, int mx) const {
        for (int i = 0; i < c->nxs; ++i) if (c->xs[i].valid && c->xs[i].g == g && c->xs[i].mx == mx) return i;
        return -1;
    }
    // bytes [lo, hi) of the lender's pos/vel block behind slot i still have a hand-off copy pending
    bool slot_busy(int i) const {
        const mfb_conv_slot& s = c->xs[i];
        if (!s.lent || !c->lender) return false;
        const mfb_particles* l = c->lender;
        const size_t half = l->posvel_bytes / 2;
        const size_t lo = size_t((const char*)s.p - (const char*)l->pos), hi = lo + conv_slot_elems(c) * sizeof(double2);
        return (l->copy_pending[MFB_ATTR_POS] && lo < half) || (l->copy_pending[MFB_ATTR_VEL] && hi > half);
    }
    int free_of_handoff() const {
        int k = 0;
        for (int i = 0; i < c->nxs; ++i) k += slot_busy(i) ? 0 : 1;
        return k;
    }
    void wait_slot(int i) {   // compute stream waits for the hand-off copies that still read slot i
        if (!slot_busy(i)) return;
        mfb_particles* l = c->lender;
        const size_t half = l->posvel_bytes / 2;
        const size_t lo = size_t((const char*)c->xs[i].p - (const char*)l->pos), hi = lo + conv_slot_elems(c) * sizeof(double2);
        for (int at : {MFB_ATTR_POS, MFB_ATTR_VEL}) {
            const bool overlaps = at == MFB_ATTR_POS ? lo < half : hi > half;
            if (overlaps && l->copy_pending[at]) {
                cuda_check(cudaStreamWaitEvent(g_ctx.stream, l->copied[at], 0), "cudaStreamWaitEvent");
                l->copy_pending[at] = false;
            }
        }
    }
    int take_slot() {   // a free slot, else the least recently used unpinned X buffer; slots under a pending hand-off last
        const bool restricted = limit > 0 && limit < c->nxs;
        int best = -1, best_busy = -1;
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = 0; i < c->nxs; ++i) {
                if (c->xs[i].pinned) continue;
                const bool busy = slot_busy(i);
                if (busy != (pass == 1)) continue;
                if (busy && restricted) continue;
                if (!c->xs[i].valid) { wait_slot(i); return i; }
                int& b = busy ? best_busy : best;
                if (b < 0 || c->xs[i].stamp < c->xs[b].stamp) b = i;
            }
            if (pass == 0 && best >= 0) break;
        }
        if (best < 0) best = best_busy;
        require(best >= 0, "internal: no free slot in the source-stage workspace");
        wait_slot(best);
        c->xs[best].valid = 0;
        return best;
    }
    void invalidate(const mfb_grid* g) {
        for (int i = 0; i < c->nxs; ++i) if (c->xs[i].g == g) c->xs[i].valid = 0;
    }
    void barrier() {   // slab mode: separates the last reads of a slot from the next peer writes into it
        if (!c->dist || D().staged) return;
#ifndef MFB_EMU
        DistState& d = D();
        ++d.epoch;
        PeerFlags pf;
        for (int s = 0; s < MFB_MAX_RANKS; ++s) pf.p[s] = d.peer_flags[s];
        RoleScope rs(ROLE_BARRIER, 0.0);
        dist_barrier_kernel<<<1, 32, 0, g_ctx.stream>>>(pf, d.flags, d.rank, d.nranks, d.epoch, d.err);
        check_launch("dist_barrier_kernel");
#endif
    }

    // ---- x pass: X(g, mx) = bwd_x[ pad( s1 * k_x^mx g^ ) ]  -> slot, layout [x][ky][kz] (slabs: [ky][x_loc][kz] after the exchange)
    int ensure_x(mfb_grid* g, int mx) {
        int i = find_x(g, mx);
        if (i >= 0) { c->xs[i].stamp = ++c->stamp; c->xs[i].pinned = 1; return i; }
        i = take_slot();
        to_kspace(g);
        const Operand o{g, mx >= 1 ? 0 : -1, mx >= 2 ? 0 : -1, -1, -1, 1.0, 0.0, 0};
        if (c->dist) {
            const int l1 = g->nloc1;
            Exchange x = slot_exchange(c->xs[i].xslot, size_t(l1) * p0l * nzin);
            StridedArgs a = plain_strided(reinterpret_cast<const double2*>(g->d), nullptr, nzc, ptrdiff_t(N0) * nzc, l1, nzin, P0, 1);
            a.n_in = N0;
            a.a_off = g->off1;
            a.a_skip = N1 / 2;
            set_operand_factor(a, c, o, s1);
            a.pos_axis = 0; a.a_axis = 1;
            route(a, x, p0l, nzin, ptrdiff_t(p0l) * nzin);                      // chunk layout [ky_loc][x_loc][kz]
            c->xs[i].zb = route_blocked(a, P0, nzin, p0l);                       // ... or z-blocked [ky_loc][kz/T][x_loc][T]
            launch_strided(P0, a, ROLE_CONV_X_BWD);
            end_exchange(x);                                                     // slot: [ky][x_loc][kz]
        } else {
            StridedArgs a = plain_strided(reinterpret_cast<const double2*>(g->d), c->xs[i].p, 0, 0, N1, nzin, P0, 1);
            a.in_ps = ptrdiff_t(N1) * nzc; a.in_as = nzc;
            a.out_ps = ptrdiff_t(N1) * nzin; a.out_as = nzin;
            a.n_in = N0;
            a.a_skip = N1 / 2;
            set_operand_factor(a, c, o, s1);
            a.pos_axis = 0; a.a_axis = 1;
            launch_strided(P0, a, ROLE_CONV_X_BWD);
        }
        ++n_x;
        c->xs[i].g = g; c->xs[i].mx = mx; c->xs[i].valid = 1; c->xs[i].pinned = 1; c->xs[i].stamp = ++c->stamp;
        return i;
    }
    // ---- y pass on the planes [x0, x0 + cx): F = bwd_y[ pad( k_y^my X ) ]  -> dst [cx][P1][kz]
    void ypass(int slot, int my, int x0, int cx, double2* dst) {
        const double2* X = c->xs[slot].p;
        StridedArgs b;
        if (c->dist) {
            const int zb = c->xs[slot].zb;
            if (zb > 0) {   // z-blocked exchange layout [ky][kz/zb][x_loc][zb]
                b = plain_strided(X + ptrdiff_t(x0) * zb, dst, ptrdiff_t(p0l) * nzin, zb, cx, nzin, P1, 1);
                b.in_zb = zb; b.in_zbs = ptrdiff_t(p0l) * zb;
            } else {
                b = plain_strided(X + ptrdiff_t(x0) * nzin, dst, ptrdiff_t(p0l) * nzin, nzin, cx, nzin, P1, 1);
            }
            b.out_ps = nzin; b.out_as = ptrdiff_t(P1) * nzin;
        } else {
            b = plain_strided(X + ptrdiff_t(x0) * N1 * nzin, dst, nzin, 0, cx, nzin, P1, 1);
            b.in_as = ptrdiff_t(N1) * nzin;
            b.out_as = ptrdiff_t(P1) * nzin;
        }
        b.n_in = N1;
        if (my > 0) {
            const Operand o{nullptr, 1, my >= 2 ? 1 : -1, -1, -1, 1.0, 0.0, 0};
            set_operand_factor(b, c, o, 1.0);
            b.pos_axis = 1; b.a_axis = 0;
        }
        launch_strided(P1, b, ROLE_CONV_Y_BWD);
        ++n_y;
    }

    // ---- one x-plane loop over a set of targets that share their base fields
    // prog >= 0 (shared-field loop): the z pass transforms every base field once and forms ALL targets of the loop
    // (z_fields_kernel); `fields` lists the program's real fields in its order.  alias_out: the targets' forward-y output goes
    // into the x-planes of X buffers that the chunk's backward y passes have already consumed (single slab, last stage).
    void run_loop(S2Target* const* T, int nt, int prog = -1, const std::vector<S2Term>* fields = nullptr, bool alias_out = false) {
        // base fields and X buffers of the loop
        std::vector<S2Base> bases;
        auto base_index = [&](const S2Term& t) -> int {
            for (size_t i = 0; i < bases.size(); ++i)
                if (bases[i].g == t.g && bases[i].mx == t.mx && bases[i].my == t.my) return int(i);
            bases.push_back(S2Base{t.g, t.mx, t.my});
            return int(bases.size()) - 1;
        };
        if (prog >= 0)
            for (const S2Term& t : *fields) base_index(t);
        for (int it = 0; it < nt && prog < 0; ++it)
            for (const S2Product& p : T[it]->prods) {
                require(int(T[it]->prods.size()) <= ZM_MAXP, "internal: too many products for one target");
                for (int k = 0; k < p.a.n; ++k) base_index(p.a.t[k]);
                if (!p.square) for (int k = 0; k < p.b.n; ++k) base_index(p.b.t[k]);
            }
        for (int i = 0; i < c->nxs; ++i) c->xs[i].pinned = 0;
        barrier();   // every rank has finished reading the slots of the previous loop
        std::vector<int> xslot(bases.size());
        for (size_t i = 0; i < bases.size(); ++i) xslot[i] = ensure_x(bases[i].g, bases[i].mx);
        std::vector<int> yout(nt);
        if (alias_out) {
            std::vector<int> distinct;
            for (int sl : xslot) if (std::find(distinct.begin(), distinct.end(), sl) == distinct.end()) distinct.push_back(sl);
            require(!c->dist && int(distinct.size()) >= nt, "internal: not enough X buffers to take the targets of the loop");
            for (int it = 0; it < nt; ++it) yout[it] = distinct[it];
        } else {
            for (int it = 0; it < nt; ++it) { yout[it] = take_slot(); c->xs[yout[it]].pinned = 1; c->xs[yout[it]].valid = 0; }
        }
        // chunk geometry
        const size_t pe = plane_elems();
        const size_t nbuf = bases.size() + (prog >= 0 ? size_t(nt) : 1);
        long cx = long(c->chunk_elems / (nbuf * pe));
        if (g_ctx.sources_chunk > 0 && cx > g_ctx.sources_chunk) cx = g_ctx.sources_chunk;
        if (cx > p0l) cx = p0l;
        require(cx >= 1, "internal: chunk workspace of the source stage is too small");
        auto wout_of = [&](int it) -> double2* { return c->chunk + (bases.size() + (prog >= 0 ? size_t(it) : 0)) * size_t(cx) * pe; };
        std::vector<Exchange> xch(nt);
        std::vector<int> yzb(nt, 0);   // z-block height of each target's exchange layout
        if (c->dist)
            for (int it = 0; it < nt; ++it) xch[it] = slot_exchange(c->xs[yout[it]].xslot, size_t(p0l) * T[it]->out->nloc1 * nzin);
        for (int x0 = 0; x0 < p0l; x0 += int(cx)) {
            const int cxx = int(std::min<long>(cx, p0l - x0));
            for (size_t i = 0; i < bases.size(); ++i) ypass(xslot[i], bases[i].my, x0, cxx, c->chunk + i * size_t(cx) * pe);
            if (prog >= 0) {
                ZfArgs zf;
                memset(&zf, 0, sizeof(zf));
                for (size_t i = 0; i < fields->size(); ++i) {
                    const S2Term& t = (*fields)[i];
                    zf.f[i].p = c->chunk + size_t(base_index(t)) * size_t(cx) * pe;
                    zf.f[i].c = t.c; zf.f[i].e = t.e;
                }
                for (int it = 0; it < nt; ++it) zf.out[it] = wout_of(it);
                zf.in_pitch = zf.out_pitch = nzin;
                zf.nrows = long(cxx) * P1;
                zf.kfz = 2.0 * M_PI / c->L[2];
                zf.coef[0] = -5.0 / 7.0; zf.coef[1] = 10.0 / 7.0;   // 3b (src/ic_generator.cc:533-538)
                launch_zfields(P2, prog, zf);
                ++n_z;
                for (int it = 0; it < nt; ++it) c->nconv += int(T[it]->prods.size());
            }
            for (int it = 0; it < nt; ++it) {
                double2* wout = wout_of(it);
                if (prog < 0) {
                ZmArgs za;
                memset(&za, 0, sizeof(za));
                za.nprod = int(T[it]->prods.size());
                auto fill = [&](ZmOperand& zo, const S2Operand& so) {
                    zo.n = so.n;
                    for (int k = 0; k < 2; ++k) {
                        const S2Term& t = so.t[k < so.n ? k : 0];
                        zo.t[k].p = c->chunk + size_t(base_index(t)) * size_t(cx) * pe;
                        zo.t[k].c = t.c; zo.t[k].e = t.e;
                    }
                };
                for (int ip = 0; ip < za.nprod; ++ip) {
                    const S2Product& p = T[it]->prods[ip];
                    fill(za.prod[ip].a, p.a);
                    fill(za.prod[ip].b, p.b);
                    za.prod[ip].coeff = p.coeff;
                    za.prod[ip].square = p.square ? 1 : 0;
                }
                za.out = wout;
                za.in_pitch = za.out_pitch = nzin;
                za.nrows = long(cxx) * P1;
                za.kfz = 2.0 * M_PI / c->L[2];
                launch_zmulti(P2, za);
                ++n_z;
                c->nconv += za.nprod;
                }
                // forward y pass of the chunk, P1 -> N1
                if (c->dist) {
                    const int l1 = T[it]->out->nloc1;
                    StridedArgs d = plain_strided(wout, nullptr, nzin, ptrdiff_t(P1) * nzin, cxx, nzin, P1, 0);
                    d.n_out = N1;
                    Exchange x = xch[it];
                    for (int s = 0; s < MFB_MAX_RANKS; ++s) if (x.base[s]) x.base[s] += ptrdiff_t(x0) * l1 * nzin;
                    route(d, x, l1, nzin, ptrdiff_t(l1) * nzin);                    // chunk layout [x_loc][ky_loc][kz]
                    const int zb = route_blocked(d, P1, nzin, l1);                  // ... or z-blocked [x_loc][kz/T][ky_loc][T]
                    require(x0 == 0 || zb == yzb[it], "internal: exchange layout changed inside a loop");
                    yzb[it] = zb;
                    launch_strided(P1, d, ROLE_CONV_Y_FWD);
                } else {
                    StridedArgs d = plain_strided(wout, c->xs[yout[it]].p + ptrdiff_t(x0) * N1 * nzin, nzin, 0, cxx, nzin, P1, 0);
                    d.in_as = ptrdiff_t(P1) * nzin;
                    d.out_as = ptrdiff_t(N1) * nzin;
                    d.n_out = N1;
                    launch_strided(P1, d, ROLE_CONV_Y_FWD);
                }
            }
        }
        // forward x pass per target: P0 -> N0, unpad scale, Nyquist = 0, output operator
        for (int it = 0; it < nt; ++it) {
            mfb_grid* out = T[it]->out;
            require(out->space == MFB_KSPACE, "convolution output grid must be in k-space");
            const int op = T[it]->op;
            if (c->dist) {
                const int l1 = out->nloc1;
                end_exchange(xch[it]);                                               // slot: [x (P0)][ky_loc][kz]
                StridedArgs e = plain_strided(xch[it].recv, reinterpret_cast<double2*>(out->d), ptrdiff_t(l1) * nzin, nzin, l1,
                                              (op == OP_ASSIGN) ? nzc : nzin, P0, 0);
                e.out_ps = nzc; e.out_as = ptrdiff_t(N0) * nzc;
                if (yzb[it] > 0) { e.in_as = yzb[it]; e.in_zb = yzb[it]; e.in_zbs = ptrdiff_t(l1) * yzb[it]; }
                e.nz_fft = nzin;
                e.n_out = N0;
                e.scale = unpad_scale;
                e.op = op; e.alpha = T[it]->alpha;
                e.a_off = out->off1;
                e.a_skip = N1 / 2;
                e.zero_fill = 1;
                launch_strided(P0, e, ROLE_CONV_X_FWD);
            } else {
                StridedArgs e = plain_strided(c->xs[yout[it]].p, reinterpret_cast<double2*>(out->d), 0, 0, N1, (op == OP_ASSIGN) ? nzc : nzin, P0, 0);
                e.in_ps = ptrdiff_t(N1) * nzin; e.in_as = nzin;
                e.out_ps = ptrdiff_t(N1) * nzc; e.out_as = nzc;
                e.nz_fft = nzin;
                e.n_out = N0;
                e.scale = unpad_scale;
                e.op = op; e.alpha = T[it]->alpha;
                e.a_skip = N1 / 2;
                e.zero_fill = 1;
                launch_strided(P0, e, ROLE_CONV_X_FWD);
            }
            invalidate(out);   // the target changed: X buffers made from its old contents are stale
            c->xs[yout[it]].pinned = 0;
            if (alias_out) c->xs[yout[it]].valid = 0;   // the X buffer was overwritten by the target
        }
        for (int i = 0; i < c->nxs; ++i) c->xs[i].pinned = 0;
        ++n_loops;
    }

    // A whole stage as ONE shared-field loop, if the program has a plan for this grid and the slots are there: the X buffers
    // of its base fields + one slot per target (none when the targets may overwrite consumed X planes).  Returns false if not
    // (the caller runs the stage product by product: run_stage).
    bool run_shared(std::vector<S2Target>& st, int prog, const std::vector<S2Term>& fields, bool last_stage, bool may_restrict) {
        if (!g_ctx.sources_shared || !zfields_has_plan(P2)) return false;
        if (may_restrict && free_of_handoff() < c->nxs) return false;   // hand-off copies pending: leave the lent slots alone
        XSet xs;
        for (const S2Term& t : fields) {
            const std::pair<const mfb_grid*, int> key(t.g, t.mx);
            if (std::find(xs.begin(), xs.end(), key) == xs.end()) xs.push_back(key);
        }
        const int nt = int(st.size());
        const bool alias_out = last_stage && !c->dist && int(xs.size()) >= nt && g_ctx.sources_alias;
        if (int(xs.size()) + (alias_out ? 0 : nt) > c->nxs) return false;
        std::vector<S2Base> bases;
        for (const S2Term& t : fields) {
            bool seen = false;
            for (const S2Base& b : bases) seen = seen || (b.g == t.g && b.mx == t.mx && b.my == t.my);
            if (!seen) bases.push_back(S2Base{t.g, t.mx, t.my});
        }
        if (c->chunk_elems / ((bases.size() + size_t(nt)) * plane_elems()) < 1) return false;
        limit = c->nxs;
        std::vector<S2Target*> ptr(nt);
        for (int k = 0; k < nt; ++k) ptr[k] = &st[k];
        run_loop(ptr.data(), nt, prog, &fields, alias_out);
        return true;
    }

    // X buffers (grid, mx) a set of products needs, added to `set`
    typedef std::vector<std::pair<const mfb_grid*, int>> XSet;
    static XSet x_of(const S2Product& p, XSet set) {
        for (const S2Operand* o : {&p.a, &p.b})
            for (int k = 0; k < o->n; ++k) {
                const std::pair<const mfb_grid*, int> key(o->t[k].g, o->t[k].mx);
                if (std::find(set.begin(), set.end(), key) == set.end()) set.push_back(key);
            }
        return set;
    }
    static XSet x_of(const S2Target& t, XSet set) {
        for (const S2Product& p : t.prods) set = x_of(p, set);
        return set;
    }
    // targets in order, as many per loop as the slots allow (each needs one; the X buffers of the loop must all be resident).
    // A target whose products need more X buffers than there are slots is split: its products go through several loops and
    // the later ones ADD to the target (linear; costs extra forward passes - the tight-memory variant, 3LPT 2048^3 on 8 GPUs).
    void run_stage(std::vector<S2Target>& in, bool may_restrict = false) {
        // restricted stage: only the slots no hand-off copy is reading (if they are enough for the stage's widest product)
        limit = c->nxs;
        if (may_restrict) {
            const int k = free_of_handoff();
            size_t widest = 0;
            for (const S2Target& t : in)
                for (const S2Product& p : t.prods) widest = std::max(widest, x_of(p, XSet()).size());
            if (k < c->nxs && k >= int(widest) + 1 && k >= 4) limit = k;
        }
        const int nslots = limit;
        std::vector<S2Target> targets;
        for (S2Target& t : in) {
            if (int(x_of(t, XSet()).size()) + 1 <= nslots) { targets.push_back(t); continue; }
            S2Target part = t;
            part.prods.clear();
            bool first = true;
            for (const S2Product& p : t.prods) {
                S2Target trial = part;
                trial.prods.push_back(p);
                if (!part.prods.empty() && int(x_of(trial, XSet()).size()) + 1 > nslots) {
                    targets.push_back(part);
                    first = false;
                    part.prods.clear();
                    // whatever the target's operator, the remaining products accumulate with the same sign / factor
                    if (t.op == OP_ASSIGN) part.op = OP_ADD;
                }
                require(int(x_of(p, XSet()).size()) + 1 <= nslots, "internal: the source-stage workspace has too few slots for one product");
                // a target that is also an operand (the first inner product lives in phi3) must be consumed before it is written
                for (const S2Operand* o : {&p.a, &p.b})
                    for (int k = 0; k < o->n; ++k) require(first || o->t[k].g != t.out, "internal: split target would overwrite its own operand");
                part.prods.push_back(p);
            }
            targets.push_back(part);
        }
        size_t i = 0;
        while (i < targets.size()) {
            int nt = 0;
            XSet xs;
            while (i + nt < targets.size()) {
                // two parts of one target cannot share a loop (the second adds to what the first assigns)
                bool same_out = false;
                for (int k = 0; k < nt; ++k) same_out = same_out || targets[i + k].out == targets[i + nt].out;
                if (same_out) break;
                XSet cand = x_of(targets[i + nt], xs);
                if (int(cand.size()) + nt + 1 > nslots) break;
                xs = cand;
                ++nt;
            }
            require(nt >= 1, "internal: the source-stage workspace has too few slots for this target");
            std::vector<S2Target*> ptr(nt);
            for (int k = 0; k < nt; ++k) ptr[k] = &targets[i + k];
            run_loop(ptr.data(), nt);
            i += nt;
        }
        limit = c->nxs;
    }
};

// src/ic_generator.cc:475-574 on the slot workspace.  Returns false (nothing done) if the workspace cannot be had.
static bool lpt_sources_v2(mfb_conv* c, int order, double Dplus, mfb_grid* phi, mfb_grid* phi2, mfb_grid* phi3, mfb_grid* const A3[3],
                           const double* lam) {
    if (c->naive || !g_ctx.sources_v2 || order < 2) return false;
    const size_t pe = size_t(c->np[1]) * (c->n[2] / 2);
    const int p0l = c->p0loc;
    // base fields of the widest loop + the target chunk(s): 12 + 1 (multi-product pass), 12 + 4 (shared-field last stage)
    const int kmax = order > 2 ? (g_ctx.sources_shared ? 16 : 13) : 7;
    const long cx_want = std::max<long>(1, std::min<long>(p0l, std::max<long>(32, p0l / 32)));   // >= 32 planes: few, large launches
    // 8 slots: everything in 4 loops (3LPT).  7: one target per loop.  5-6: targets are split (more forward passes).
    // (10: the shared last stage with a slot per target - what slabs need; a single slab lets the targets overwrite consumed X planes)
    const int min_slots = order > 2 ? 5 : 4, want_slots = order > 2 ? ((g_ctx.sources_shared && !g_ctx.sources_alias) ? 10 : 8) : 4;
    if (c->dist && D().nxbuf < min_slots) return false;
    if (!conv_enter_slots(c, want_slots, min_slots, size_t(kmax) * cx_want * pe, size_t(kmax) * pe)) return false;
    // lent particle arrays: streamed hand-off copies of the previous run must have drained before their memory is written.
    // A lent CHUNK workspace is needed from the first launch on; lent SLOTS wait when they are first taken (S2Runner::take_slot),
    // and the first two stages leave them alone while a copy is pending, so that the hand-off drains under them.
    if (c->chunk_lent) conv_wait_lender(c);

    const double g1 = -Dplus;
    const double g2 = -3.0 / 7.0 * Dplus * Dplus;
    const double g3 = (order > 2) ? 1.0 / 3.0 * Dplus * Dplus * Dplus : 0.0;
    const double g3c = (order > 2) ? 1.0 / 7.0 * Dplus * Dplus * Dplus : 0.0;
    S2Runner R(c);
    for (int i = 0; i < c->nxs; ++i) { c->xs[i].valid = 0; c->xs[i].pinned = 0; }
    auto H = [](mfb_grid* g, int a, int b) { return s2_hess(g, a, b); };
    require(phi2 != nullptr, "phi2 grid required for LPT order > 1");
    phi2->space = MFB_KSPACE;
    mfb_grid* I[4] = {nullptr, nullptr, nullptr, nullptr};
    {
        // ---- loop 1: phi2 source (:482-486) and, at third order, the inner products of the determinant (:522-526).  The
        // truncation between inner and outer product of convolve3 (include/convolution.hh:542-549) is kept - the inner
        // products are targets of their own; it is linear, so inner products that meet the same outer field are summed
        // first:  det = T(xx yy - xy xy) zz + 2 T(xy xz) yz - T(yz yz) xx - T(xz xz) yy.
        std::vector<S2Target> st(1);
        st[0].out = phi2; st[0].op = OP_ASSIGN; st[0].alpha = 0.0;
        st[0].prods = {s2_prod(H(phi, 0, 0), s2_hess2(phi, 1, 1, 2, 2, +1.0), 1.0), s2_prod(H(phi, 1, 1), H(phi, 2, 2), 1.0),
                       s2_prod(H(phi, 0, 1), H(phi, 0, 1), -1.0), s2_prod(H(phi, 0, 2), H(phi, 0, 2), -1.0),
                       s2_prod(H(phi, 1, 2), H(phi, 1, 2), -1.0)};
        if (order > 2) {
            require(phi3 && A3 && A3[0] && A3[1] && A3[2], "phi3 and A3 grids required for LPT order > 2");
            // the inner products live in the grids of later targets until the outer products have consumed them
            I[0] = phi3; I[1] = A3[0]; I[2] = A3[1]; I[3] = A3[2];
            for (int k = 0; k < 4; ++k) I[k]->space = MFB_KSPACE;
            st.resize(5);
            for (int k = 1; k < 5; ++k) { st[k].out = I[k - 1]; st[k].op = OP_ASSIGN; st[k].alpha = 0.0; }
            st[1].prods = {s2_prod(H(phi, 0, 0), H(phi, 1, 1), 1.0), s2_prod(H(phi, 0, 1), H(phi, 0, 1), -1.0)};
            st[2].prods = {s2_prod(H(phi, 0, 1), H(phi, 0, 2), 1.0)};
            st[3].prods = {s2_prod(H(phi, 1, 2), H(phi, 1, 2), 1.0)};
            st[4].prods = {s2_prod(H(phi, 0, 2), H(phi, 0, 2), 1.0)};
        }
        // third order: the six Hessian components of phi transformed once for all five targets (z_fields_kernel, program A)
        if (!(order > 2 && R.run_shared(st, ZF_PROG_A, s2_hess_fields(phi), false, true))) R.run_stage(st, true);
    }
    if (lam) {
        KGeom k = kgeom(phi2);
        MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, tides_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(phi2->d),
                   reinterpret_cast<const double2*>(phi->d), k, lam[0], lam[1], lam[2]);
        check_launch("tides");
    }
    grid_laplacian_pow(phi2, -1, -1.0);  // :499
    if (order > 2) {
        {
            // ---- loop 2: the outer products of the determinant -> phi3.  Its x-forward pass reads the slot, not the grid, so
            // I1 may sit in phi3 itself: every X buffer of the loop is made before the target is written.
            std::vector<S2Target> st(1);
            st[0].out = phi3; st[0].op = OP_ASSIGN; st[0].alpha = 0.0;
            st[0].prods = {s2_prod(s2_plain(I[0]), H(phi, 2, 2), 1.0), s2_prod(s2_plain(I[1]), H(phi, 1, 2), 2.0),
                           s2_prod(s2_plain(I[2]), H(phi, 0, 0), -1.0), s2_prod(s2_plain(I[3]), H(phi, 1, 1), -1.0)};
            R.run_stage(st, true);
        }
        {
            // ---- loops 3..: 3b (:533-538, added to phi3) and the transverse term A_d = P1 - P2 + P3 - P4 (:545-556)
            std::vector<S2Target> st(4);
            st[0].out = phi3; st[0].op = OP_ADD; st[0].alpha = 0.0;
            st[0].prods = {s2_prod(H(phi, 0, 0), s2_hess2(phi2, 1, 1, 2, 2, +1.0), -5.0 / 7.0),
                           s2_prod(H(phi, 1, 1), s2_hess2(phi2, 2, 2, 0, 0, +1.0), -5.0 / 7.0),
                           s2_prod(H(phi, 2, 2), s2_hess2(phi2, 0, 0, 1, 1, +1.0), -5.0 / 7.0),
                           s2_prod(H(phi, 0, 1), H(phi2, 0, 1), 10.0 / 7.0), s2_prod(H(phi, 0, 2), H(phi2, 0, 2), 10.0 / 7.0),
                           s2_prod(H(phi, 1, 2), H(phi2, 1, 2), 10.0 / 7.0)};
            for (int d = 0; d < 3; ++d) {
                const int p = (d + 1) % 3, q = (d + 2) % 3;
                S2Target& t = st[1 + d];
                t.out = A3[d]; t.op = OP_ASSIGN; t.alpha = 0.0;
                A3[d]->space = MFB_KSPACE;
                t.prods = {s2_prod(H(phi2, d, p), H(phi, d, q), 1.0), s2_prod(H(phi2, d, q), H(phi, d, p), -1.0),
                           s2_prod(H(phi, p, q), s2_hess2(phi2, p, p, q, q, -1.0), 1.0),
                           s2_prod(H(phi2, p, q), s2_hess2(phi, p, p, q, q, -1.0), -1.0)};
            }
            // the twelve Hessian components of phi and phi2 transformed once for all four targets (program C)
            std::vector<S2Term> fields = s2_hess_fields(phi), f2 = s2_hess_fields(phi2);
            fields.insert(fields.end(), f2.begin(), f2.end());
            if (!R.run_shared(st, ZF_PROG_C, fields, true, false)) R.run_stage(st);
        }
        grid_laplacian_pow(phi3, -1, -g3);   // :539 and the growth factor of :566 in one sweep
        for (int d = 0; d < 3; ++d) grid_laplacian_pow(A3[d], -1, -g3c);   // :555 and :568-570
    }
    for (int i = 0; i < c->nxs; ++i) { c->xs[i].valid = 0; c->xs[i].pinned = 0; }
    R.barrier();   // slab mode: the next 3-D FFT may write a peer's ping-pong buffer that served as a slot here
    // :561-574
    grid_scale(phi, g1);
    grid_scale(phi2, g2);
    c->last_path = 2; c->last_slots = c->nxs; c->last_loops = int(R.n_loops);
    c->last_xpass = R.n_x; c->last_ypass = R.n_y; c->last_zlaunch = R.n_z;
    return true;
}

}  // namespace mfb
