// mfb_api.cu - libmfb200: context, grids, FFT orchestration, Orszag convolver, LPT source stage,
// lattice read-out and the extern "C" boundary declared in include/mfb200.h.
//
// Single-GPU part.  (The multi-GPU slab layer lives in mfb_dist.cu.)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "fft_kernels.cuh"
#include "fft_kernels2.cuh"
#include "fft_kernels_z2.cuh"
#include "fft_kernels_zm.cuh"
#include "fft_kernels_zf.cuh"
#include "grid_ops.cuh"
#include "mfb_internal.h"
#include "../../include/mfb200.h"

#ifdef MFB_EMU
using std::abs;
using std::max;
using std::min;
#endif

namespace mfb {

// ------------------------------------------------------------------------------------------
// plan table: n, threads-per-line, lines-per-tile, radix sequence
// ------------------------------------------------------------------------------------------
#define MFB_FOR_EACH_PLAN(X)      \
    X(32, 4, 16, 8, 4)            \
    X(48, 2, 16, 8, 6)            \
    X(64, 8, 16, 8, 8)            \
    X(96, 4, 16, 12, 8)           \
    X(128, 8, 16, 16, 8)          \
    X(192, 4, 16, 16, 12)         \
    X(256, 16, 8, 16, 16)         \
    X(384, 16, 8, 8, 8, 6)        \
    X(512, 32, 8, 8, 8, 8)        \
    X(768, 32, 8, 12, 8, 8)       \
    X(1024, 32, 8, 16, 8, 8)      \
    X(1536, 32, 8, 16, 12, 8)     \
    X(2048, 64, 4, 16, 16, 8)     \
    X(3072, 64, 4, 16, 16, 12)

// contiguous-axis (z) passes of N-grids: rows are contiguous in HBM whatever the tile height, so the long lines use narrow
// tiles (T = 4 / 2 lines, half / quarter of the shared memory of the strided plans) and 2-3 CTAs share an SM instead of one
#define MFB_FOR_EACH_PLAN_ZROW(X) \
    X(32, 4, 16, 8, 4)            \
    X(48, 2, 16, 8, 6)            \
    X(64, 8, 16, 8, 8)            \
    X(96, 4, 16, 12, 8)           \
    X(128, 8, 16, 16, 8)          \
    X(192, 4, 16, 16, 12)         \
    X(256, 16, 8, 16, 16)         \
    X(384, 16, 8, 8, 8, 6)        \
    X(512, 32, 8, 8, 8, 8)        \
    X(768, 32, 4, 12, 8, 8)       \
    X(1024, 64, 4, 16, 8, 8)      \
    X(1536, 32, 4, 16, 12, 8)     \
    X(2048, 128, 2, 16, 16, 8)    \
    X(3072, 64, 2, 16, 16, 12)

Context g_ctx;
thread_local std::string g_err;
}  // namespace mfb
#include "mfb_dist.cuh"
namespace mfb {

void set_error(const std::string& s) { g_err = s; }

#define MFB_TRY try {
#define MFB_CATCH                              \
    }                                          \
    catch (const std::exception& e) {          \
        mfb::set_error(e.what());              \
        return 1;                              \
    }                                          \
    catch (...) {                              \
        mfb::set_error("unknown error");       \
        return 1;                              \
    }                                          \
    return 0;

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
void require(bool c, const char* msg) {
    if (!c) throw std::runtime_error(msg);
}
#ifndef MFB_EMU
static cudaEvent_t prof_event() {
    if (!g_ctx.event_pool.empty()) { cudaEvent_t e = g_ctx.event_pool.back(); g_ctx.event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cuda_check(cudaEventCreate(&e), "cudaEventCreate");
    return e;
}
#endif
// every launch goes through RoleScope (begin event) and check_launch (end event)
RoleScope::RoleScope(int role, double bytes) : prev_role(g_ctx.cur_role), prev_bytes(g_ctx.cur_bytes), armed(false) {
    g_ctx.cur_role = role;
    g_ctx.cur_bytes = bytes;
#ifndef MFB_EMU
    if (g_ctx.profiling) {
        ProfEvent pe;
        pe.a = prof_event(); pe.b = nullptr; pe.role = role; pe.bytes = bytes;
        cudaEventRecord(pe.a, g_ctx.stream);
        g_ctx.pending.push_back(pe);
        armed = true;
    }
#endif
}
RoleScope::~RoleScope() {
#ifndef MFB_EMU
    if (armed) {
        ProfEvent& pe = g_ctx.pending.back();
        pe.b = prof_event();
        cudaEventRecord(pe.b, g_ctx.stream);
    }
#endif
    g_ctx.cur_role = prev_role;
    g_ctx.cur_bytes = prev_bytes;
}
void prof_collect() {
#ifndef MFB_EMU
    if (g_ctx.pending.empty()) return;
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    for (auto& pe : g_ctx.pending) {
        float ms = 0.f;
        if (pe.b && cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) {
            g_ctx.role_ms[pe.role] += ms;
            g_ctx.role_bytes[pe.role] += pe.bytes;
            g_ctx.role_count[pe.role] += 1;
        }
        g_ctx.event_pool.push_back(pe.a);
        if (pe.b) g_ctx.event_pool.push_back(pe.b);
    }
    g_ctx.pending.clear();
#endif
}
void check_launch(const char* what) {
    ++g_ctx.launches;
    cuda_check(cudaGetLastError(), what);
}

#define MFB_LAUNCH_ROLE(role, kernel, grid, block, ...)                           \
    do {                                                                          \
        RoleScope rs_(role, 0.0);                                                 \
        MFB_LAUNCH(kernel, grid, block, 0, g_ctx.stream, __VA_ARGS__);            \
    } while (0)

void* dev_alloc(size_t bytes) {
    void* p = nullptr;
    const cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();   // clear the sticky allocation error: the library stays usable after an out-of-memory report
        throw std::runtime_error("cudaMalloc of " + std::to_string(bytes >> 20) + " MiB failed: " + cudaGetErrorString(e));
    }
    return p;
}

const double2* twiddles(int n) {
    auto it = g_ctx.tw.find(n);
    if (it != g_ctx.tw.end()) return it->second;
    std::vector<double2> h(n);
    for (int m = 0; m < n; ++m) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
        h[m].x = (double)cosl(a);
        h[m].y = (double)sinl(a);
    }
    double2* d = (double2*)dev_alloc(sizeof(double2) * n);
    cuda_check(cudaMemcpy(d, h.data(), sizeof(double2) * n, cudaMemcpyHostToDevice), "twiddle upload");
    g_ctx.tw[n] = d;
    return d;
}

// stage-major twiddle table of a plan (Plan::tw2_off): [stage][u-1][j] = exp(-2 pi i u j / len(stage))
template <class P> const double2* twiddles2() {
    std::string key = std::to_string(P::N);
    for (int s = 0; s < P::S; ++s) key += "." + std::to_string(P::radix(s));
    auto it = g_ctx.tw2.find(key);
    if (it != g_ctx.tw2.end()) return it->second;
    std::vector<double2> h(P::TW2_ELEMS > 0 ? P::TW2_ELEMS : 1);
    const long double tau = 2.0L * 3.14159265358979323846264338327950288L;
    for (int s = 0; s + 1 < P::S; ++s) {
        const int r = P::radix(s), L = P::len(s), M = L / r;
        for (int u = 1; u < r; ++u)
            for (int j = 0; j < M; ++j) {
                const long double a = -tau * (long double)((long)u * j % L) / (long double)L;
                h[P::tw2_off(s) + (u - 1) * M + j] = make_double2((double)cosl(a), (double)sinl(a));
            }
    }
    double2* d = (double2*)dev_alloc(sizeof(double2) * h.size());
    cuda_check(cudaMemcpy(d, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice), "twiddle upload");
    g_ctx.tw2[key] = d;
    return d;
}

bool supported_size(int n) {
    switch (n) {
#define X(N_, TPL_, T_, ...) case N_:
        MFB_FOR_EACH_PLAN(X)
#undef X
        return true;
        default: return false;
    }
}

template <class K> void optin_smem(K kernel, size_t bytes, bool& done) {
    if (!done) {
        if (bytes > 48 * 1024) cuda_check(MFB_KERNEL_SMEM_OPTIN(kernel, bytes), "cudaFuncSetAttribute(smem)");
        done = true;
    }
}

// ------------------------------------------------------------------------------------------
// kernel launchers
// ------------------------------------------------------------------------------------------
static int* g_query_T = nullptr;   // non-null: launch_strided only reports the tile height (lines) of the plan it would use
template <class P> void launch_strided_t(const StridedArgs& a, int role) {
    if (g_query_T) { *g_query_T = P::T; return; }
    static bool done = false;
    optin_smem(strided_pass_kernel<P>, P::SMEM_BYTES, done);
    const int tiles = (a.nz + P::T - 1) / P::T;
    const unsigned grid = unsigned(a.na) * unsigned(tiles);
    if (grid == 0) return;
    const double2* tw = twiddles(P::N);
    // algorithmic bytes: every transformed line reads n_in and writes n_out complex (+ RMW read)
    const double lines = double(a.na - (a.a_skip >= 0 ? 1 : 0)) * double(a.nz_fft);
    const double bytes = lines * 16.0 * (double(a.n_in) + double(a.n_out) * (a.op == OP_ASSIGN ? 1.0 : 2.0));
    RoleScope rs(role, bytes);
    MFB_LAUNCH(strided_pass_kernel<P>, grid, P::NTHREADS, P::SMEM_BYTES, g_ctx.stream, a, tw);
    check_launch("strided_pass_kernel");
}
// second-generation plans (fft_kernels2.cuh).  "B": first radix carries the factor 3 (pad N -> 3N/2 resolved at
// compile time; also used for plain passes); "F": last radix carries it (truncate 3N/2 -> N).
#define MFB_FOR_EACH_PLAN2_B(X)   \
    X(32, 4, 16, 1, 8, 4)            \
    X(48, 4, 16, 1, 6, 8)            \
    X(64, 8, 16, 1, 8, 8)            \
    X(96, 4, 16, 1, 12, 8)           \
    X(128, 8, 16, 1, 16, 8)          \
    X(192, 4, 16, 1, 12, 16)         \
    X(256, 16, 8, 2, 16, 16)         \
    X(384, 16, 8, 2, 6, 8, 8)        \
    X(512, 32, 8, 2, 8, 8, 8)        \
    X(768, 32, 8, 2, 12, 8, 8)       \
    X(1024, 64, 8, 1, 16, 8, 8)      \
    X(1536, 64, 8, 1, 12, 16, 8)     \
    X(2048, 64, 4, 1, 16, 16, 8)     \
    X(3072, 64, 4, 1, 12, 16, 16)
#define MFB_FOR_EACH_PLAN2_F(X)   \
    X(48, 4, 16, 1, 8, 6)            \
    X(96, 4, 16, 1, 8, 12)           \
    X(192, 4, 16, 1, 16, 12)         \
    X(384, 16, 8, 2, 8, 8, 6)        \
    X(768, 32, 8, 2, 8, 8, 12)       \
    X(1536, 64, 4, 2, 8, 16, 12)     \
    X(3072, 64, 4, 1, 16, 16, 12)
// fused z pass of the dealiased product (fft_kernels_z2.cuh): N, threads per row, rows per CTA, min CTAs/SM, radices
// (first radix a multiple of 3, last radix 8: conflict-free padded lines)
// Operand rows are loaded straight from global memory in stage 0 (z_conv3_kernel<..., DIRECT = true>): kbench on B200,
// 589824 rows: n = 1536  6.25 -> 5.54 ms (TPL 64, 3 CTAs/SM), n = 768  2.58 -> 2.48 ms (profiles/r01_kbench_z4_direct.txt).
#define MFB_FOR_EACH_PLANZ(X)     \
    X(48, 32, 2, 1, 6, 8)         \
    X(96, 32, 2, 1, 12, 8)        \
    X(192, 32, 2, 1, 6, 4, 8)     \
    X(384, 32, 2, 4, 6, 8, 8)     \
    X(768, 64, 2, 5, 12, 8, 8)    \
    X(1536, 64, 2, 3, 12, 16, 8)  \
    X(3072, 128, 2, 2, 12, 16, 16)

template <class P, int MODE, int MINB> void launch_strided2_t(const StridedArgs& a, int role) {
    if (g_query_T) { *g_query_T = P::T; return; }
    static bool done = false;
    optin_smem(strided2_kernel<P, MODE, MINB>, P::SMEM_BYTES, done);
    const int tiles = (a.nz + P::T - 1) / P::T;
    const unsigned grid = unsigned(a.na) * unsigned(tiles);
    if (grid == 0) return;
    const double2* tw = twiddles2<P>();
    const double lines = double(a.na - (a.a_skip >= 0 ? 1 : 0)) * double(a.nz_fft);
    const double bytes = lines * 16.0 * (double(a.n_in) + double(a.n_out) * (a.op == OP_ASSIGN ? 1.0 : 2.0));
    RoleScope rs(role, bytes);
    MFB_LAUNCH((strided2_kernel<P, MODE, MINB>), grid, P::NTHREADS, P::SMEM_BYTES, g_ctx.stream, a, tw);
    check_launch("strided2_kernel");
}
// returns false if the second-generation kernel does not cover this pass (caller falls back to the first)
static bool launch_strided2(int n, const StridedArgs& a, int role) {
    int mode;
    if (a.n_in == n && a.n_out == n) mode = SM_PLAIN;
    else if (a.n_out == n && 3 * a.n_in == 2 * n) mode = SM_PAD;
    else if (a.n_in == n && 3 * a.n_out == 2 * n) mode = SM_TRUNC;
    else return false;
    if (a.f.enabled && mode != SM_PAD) return false;
    if (mode == SM_TRUNC) {
        // n = 1536: one 8-line tile per SM instead of two 4-line tiles (a) for a read-modify-write output (kbench: 3518 vs
        // 3239 GB/s) and (b) for x passes: with a position stride of megabytes 64-byte segments copy at 2.9 TB/s, 128-byte
        // ones at 5.6 TB/s (profiles/r01_kbench_segment_copy.txt, "x-like")
        const bool far_stride = a.in_ps * ptrdiff_t(sizeof(double2)) >= (1 << 20) || a.out_ps * ptrdiff_t(sizeof(double2)) >= (1 << 20);
        // (c) for routed passes (slabs: the forward y pass stores straight into the peers' receive buffers): 128-byte instead of
        // 64-byte peer stores - the padded x pass with 8-line tiles reaches 76 % of the NVLink rate, 4-line tiles 48 %
        const bool routed = a.out_blk > 0;
        if (n == 1536 && (a.op != OP_ASSIGN || far_stride || routed)) { launch_strided2_t<Plan<1536, 64, 8, 8, 16, 12>, SM_TRUNC, 1>(a, role); return true; }
        switch (n) {
#define X(N_, TPL_, T_, MINB_, ...)                                      \
    case N_: launch_strided2_t<Plan<N_, TPL_, T_, __VA_ARGS__>, SM_TRUNC, MINB_>(a, role); return true;
            MFB_FOR_EACH_PLAN2_F(X)
#undef X
            default: return false;
        }
    }
    if (mode == SM_PAD && n == 1536 && g_ctx.y_pad_tile4) {
        // y-like padded passes (position stride of kilobytes, not routed to peers): two 4-line tiles per SM instead of one 8-line
        // tile - the two CTAs overlap each other's load, butterfly and store phases (kbench 3.36 -> 3.75 TB/s,
        // profiles/r02_kbench_strided_tma_bulk_negative.txt); x passes and routed passes keep 128-byte segments
        const bool far_stride = a.in_ps * ptrdiff_t(sizeof(double2)) >= (1 << 20) || a.out_ps * ptrdiff_t(sizeof(double2)) >= (1 << 20);
        if (!far_stride && a.out_blk == 0) { launch_strided2_t<Plan<1536, 64, 4, 12, 16, 8>, SM_PAD, 2>(a, role); return true; }
    }
    switch (n) {
#define X(N_, TPL_, T_, MINB_, ...)                                                  \
    case N_:                                                                         \
        if (mode == SM_PAD) {                                                        \
            if constexpr (N_ % 3 == 0) launch_strided2_t<Plan<N_, TPL_, T_, __VA_ARGS__>, SM_PAD, MINB_>(a, role); \
            else return false;                                                       \
        } else launch_strided2_t<Plan<N_, TPL_, T_, __VA_ARGS__>, SM_PLAIN, MINB_>(a, role); \
        return true;
        MFB_FOR_EACH_PLAN2_B(X)
#undef X
        default: return false;
    }
}

void launch_strided(int n, const StridedArgs& a, int role);
// tile height (adjacent k_z lines per CTA) of the plan launch_strided picks for this pass
static int strided_tile_lines(int n, const StridedArgs& a) {
    int T = 0;
    g_query_T = &T;
    try { launch_strided(n, a, ROLE_ELEMENTWISE); } catch (...) { g_query_T = nullptr; throw; }
    g_query_T = nullptr;
    return T;
}
// z-blocked chunk layout for a routed pass and its consumer: block height = the producer's tile height, if the line count
// allows it (returns 0 otherwise: plain layout).  Sets the producer's output fields; the consumer uses consumer_blocked().
static int route_blocked(StridedArgs& prod, int n, int nlines, int blk) {
    if (!g_ctx.exchange_blocked) return 0;
    const int T = strided_tile_lines(n, prod);
    if (T <= 1 || nlines % T != 0 || prod.nz % T != 0) return 0;
    prod.out_zb = T;
    prod.out_ps = T;                              // consecutive positions of a tile: adjacent T-line segments
    prod.out_zbs = ptrdiff_t(blk) * T;            // next z block of the same a-index
    return T;
}
void launch_strided(int n, const StridedArgs& a, int role) {
    if (g_ctx.strided_v2 && launch_strided2(n, a, role)) return;
    switch (n) {
#define X(N_, TPL_, T_, ...)                              \
    case N_: {                                            \
        using P = Plan<N_, TPL_, T_, __VA_ARGS__>;        \
        launch_strided_t<P>(a, role);                     \
        break;                                            \
    }
        MFB_FOR_EACH_PLAN(X)
#undef X
        default: throw std::runtime_error("unsupported FFT length " + std::to_string(n));
    }
}

template <class P> void launch_z_t(int kind, double* data, ptrdiff_t pitch, long nrows, double scale, const ZReadout* ro) {
    static bool done_c2r = false, done_r2c = false, done_ro = false;
    const unsigned grid = unsigned((nrows + 2 * P::T - 1) / (2 * P::T));
    const double2* tw = twiddles(P::N);
    RoleScope rs(kind == 0 ? ROLE_FFT_Z_C2R : ROLE_FFT_Z_R2C, double(nrows) * (16.0 * P::N + 16.0));
    if (kind == 0 && ro) {   // c2r with the lattice read-out as its epilogue: the real field is never stored
        optin_smem((z_c2r_kernel<P, true>), P::SMEM_BYTES, done_ro);
        MFB_LAUNCH((z_c2r_kernel<P, true>), grid, P::NTHREADS, P::SMEM_BYTES, g_ctx.stream, data, pitch, nrows, scale, tw, *ro);
        check_launch("z_c2r_kernel(readout)");
    } else if (kind == 0) {
        optin_smem((z_c2r_kernel<P, false>), P::SMEM_BYTES, done_c2r);
        MFB_LAUNCH((z_c2r_kernel<P, false>), grid, P::NTHREADS, P::SMEM_BYTES, g_ctx.stream, data, pitch, nrows, scale, tw, ZReadout());
        check_launch("z_c2r_kernel");
    } else {
        optin_smem(z_r2c_kernel<P>, P::SMEM_BYTES, done_r2c);
        MFB_LAUNCH(z_r2c_kernel<P>, grid, P::NTHREADS, P::SMEM_BYTES, g_ctx.stream, data, pitch, nrows, scale, tw);
        check_launch("z_r2c_kernel");
    }
}
void launch_z(int n, int kind, double* data, ptrdiff_t pitch, long nrows, double scale, const ZReadout* ro = nullptr) {
    switch (n) {
#define X(N_, TPL_, T_, ...)                                  \
    case N_: {                                                \
        using P = Plan<N_, TPL_, T_, __VA_ARGS__>;            \
        launch_z_t<P>(kind, data, pitch, nrows, scale, ro);   \
        break;                                                \
    }
        MFB_FOR_EACH_PLAN_ZROW(X)
#undef X
        default: throw std::runtime_error("unsupported FFT length " + std::to_string(n));
    }
}

template <class P> void launch_zc2c_t(double2* data, long nrows, double scale, int inverse) {
    static bool done = false;
    optin_smem(z_c2c_kernel<P>, P::SMEM_BYTES, done);
    const unsigned grid = unsigned((nrows + P::T - 1) / P::T);
    RoleScope rs(ROLE_FFT_Z_C2C, double(nrows) * 32.0 * P::N);
    MFB_LAUNCH(z_c2c_kernel<P>, grid, P::NTHREADS, P::SMEM_BYTES, g_ctx.stream, data, nrows, scale, inverse, twiddles(P::N));
    check_launch("z_c2c_kernel");
}
void launch_zc2c(int n, double2* data, long nrows, double scale, int inverse) {
    switch (n) {
#define X(N_, TPL_, T_, ...)                                                     \
    case N_: launch_zc2c_t<Plan<N_, TPL_, T_, __VA_ARGS__>>(data, nrows, scale, inverse); break;
        MFB_FOR_EACH_PLAN_ZROW(X)
#undef X
        default: throw std::runtime_error("unsupported FFT length " + std::to_string(n));
    }
}

template <class P, int RPC, int MINB>
void launch_zconv2_t(const double2* A, const double2* B, double2* out, ptrdiff_t in_pitch, ptrdiff_t out_pitch, long nrows, double scale,
                     int acc) {
    static bool done = false;
    constexpr size_t smem = z_conv3_direct_smem<P, RPC>();
    optin_smem(z_conv3_kernel<P, RPC, MINB, true>, smem, done);
    const unsigned grid = unsigned((nrows + RPC - 1) / RPC);
    const double2* tw = twiddles2<P>();
    // algorithmic bytes: both operand rows (one if A == B) + the output row (+ its old value when accumulating)
    RoleScope rs(ROLE_CONV_Z, double(nrows) * double(P::N / 3) * 16.0 * ((A == B ? 1.0 : 2.0) + (acc ? 2.0 : 1.0)));
    MFB_LAUNCH((z_conv3_kernel<P, RPC, MINB, true>), grid, P::TPL * RPC, smem, g_ctx.stream, A, B, out, in_pitch, out_pitch, nrows, scale, tw,
               acc);
    check_launch("z_conv3_kernel");
}
static bool launch_zconv2(int n, const double2* A, const double2* B, double2* out, ptrdiff_t in_pitch, ptrdiff_t out_pitch,
                          long nrows, int nzin, double scale, int acc) {
    if (3 * nzin != n) return false;
    switch (n) {
#define X(N_, TPL_, RPC_, MINB_, ...)                                                                          \
    case N_: launch_zconv2_t<Plan<N_, TPL_, 2, __VA_ARGS__>, RPC_, MINB_>(A, B, out, in_pitch, out_pitch, nrows, scale, acc); return true;
        MFB_FOR_EACH_PLANZ(X)
#undef X
        default: return false;
    }
}
void launch_zconv(int n, const double2* A, const double2* B, double2* out, ptrdiff_t in_pitch, ptrdiff_t out_pitch, long nrows,
                  int nzin, double scale, int acc = 0) {
    if (!launch_zconv2(n, A, B, out, in_pitch, out_pitch, nrows, nzin, scale, acc))
        throw std::runtime_error("unsupported padded length " + std::to_string(n) + " for the fused z pass");
}

static StridedArgs plain_strided(const double2* in, double2* out, ptrdiff_t ps, ptrdiff_t as, int na, int nz, int n,
                                 int inverse) {
    StridedArgs a;
    memset(&a, 0, sizeof(a));
    a.in = in; a.out = out;
    a.in_ps = a.out_ps = ps;
    a.in_as = a.out_as = as;
    a.na = na; a.nz = nz; a.nz_fft = nz;
    a.n_in = a.n_out = n;
    a.inverse = inverse;
    a.scale = 1.0;
    a.op = OP_ASSIGN;
    a.a_skip = -1;
    return a;
}

// ------------------------------------------------------------------------------------------
// Grid
// ------------------------------------------------------------------------------------------
KGeom kgeom(const mfb_grid* g) {
    KGeom k;
    k.n0 = g->n[0]; k.n1 = g->n[1]; k.n2 = g->n[2]; k.nzc = g->nzc;
    k.kf0 = 2.0 * M_PI / g->L[0]; k.kf1 = 2.0 * M_PI / g->L[1]; k.kf2 = 2.0 * M_PI / g->L[2];
    if (g->dist) { k.s0 = g->nloc1; k.s1 = g->n[0]; k.tr = 1; k.off = g->off1; }
    else { k.s0 = g->n[0]; k.s1 = g->n[1]; k.tr = 0; k.off = 0; }
    return k;
}
static unsigned stream_blocks(size_t n) {
    size_t b = (n + EW_THREADS - 1) / EW_THREADS;
    const size_t cap = size_t(148) * 16;
    return unsigned(b < cap ? (b ? b : 1) : cap);
}

// routed output of a slab-decomposed pass: position k of the transformed axis belongs to rank k / blk
static void route(StridedArgs& a, const Exchange& x, int blk, ptrdiff_t ps, ptrdiff_t as) {
    a.out = nullptr;
    a.out_blk = blk;
    a.out_ps = ps; a.out_as = as;
    for (int s = 0; s < MFB_MAX_RANKS; ++s) a.out_peer[s] = x.base[s];
}

// z pitch of the exchange chunks of the 3-D FFTs: N/2+1 lines rounded up to a multiple of 8, so that the 128-byte segments
// an 8-line tile stores into a peer's buffer start on 128-byte boundaries (with the grid's own pitch of N/2+1 = 513 elements
// every segment straddled two lines: the routed x pass sustained 47 % of the NVLink rate, profiles/r02_bench_3lpt1024_n8_v2.json)
static int dist_zpitch(int nzc) { return (nzc + 7) & ~7; }

// slab-decomposed r2c: z (local rows), y (local planes) -> exchange -> x.   Replaces fftw_mpi_plan_dft_r2c_3d with
// FFTW_MPI_TRANSPOSED_OUT (src/grid_fft.cc:125-128): the result is [ky_loc][kx][kz].
static void dist_fft_forward(mfb_grid* g, double norm) {
    double2* c = reinterpret_cast<double2*>(g->d);
    const int n0 = g->n[0], n1 = g->n[1], nzc = g->nzc, l0 = g->nloc0, l1 = g->nloc1;
    const int zp = dist_zpitch(nzc);
    launch_z(g->n[2], 1, g->d, g->pitch, long(l0) * n1, norm);
    Exchange x = begin_exchange(size_t(l0) * l1 * zp);
    StridedArgs a = plain_strided(c, nullptr, nzc, ptrdiff_t(n1) * nzc, l0, nzc, n1, 0);
    route(a, x, l1, zp, ptrdiff_t(l1) * zp);                           // chunk layout [x_loc][ky_loc][kz (pitch zp)]
    launch_strided(n1, a, ROLE_FFT_Y);
    end_exchange(x);                                                    // recv: [x][ky_loc][kz]
    StridedArgs b = plain_strided(x.recv, c, ptrdiff_t(l1) * zp, zp, l1, nzc, n0, 0);
    b.out_ps = nzc; b.out_as = ptrdiff_t(n0) * nzc;                     // -> [ky_loc][kx][kz]
    launch_strided(n0, b, ROLE_FFT_X);
}
static void dist_fft_backward(mfb_grid* g, double norm, const ZReadout* ro = nullptr) {
    double2* c = reinterpret_cast<double2*>(g->d);
    const int n0 = g->n[0], n1 = g->n[1], nzc = g->nzc, l0 = g->nloc0, l1 = g->nloc1;
    const int zp = dist_zpitch(nzc);
    Exchange x = begin_exchange(size_t(l0) * l1 * zp);
    StridedArgs a = plain_strided(c, nullptr, nzc, ptrdiff_t(n0) * nzc, l1, nzc, n0, 1);
    route(a, x, l0, zp, ptrdiff_t(l0) * zp);                           // chunk layout [ky_loc][x_loc][kz (pitch zp)]
    launch_strided(n0, a, ROLE_FFT_X);
    end_exchange(x);                                                    // recv: [ky][x_loc][kz]
    StridedArgs b = plain_strided(x.recv, c, ptrdiff_t(l0) * zp, zp, l0, nzc, n1, 1);
    b.out_ps = nzc; b.out_as = ptrdiff_t(n1) * nzc;                     // -> [x_loc][y][kz]
    launch_strided(n1, b, ROLE_FFT_Y);
    launch_z(g->n[2], 0, g->d, g->pitch, long(l0) * n1, norm, ro);
}

void grid_fft_forward(mfb_grid* g) {
    const double norm = 1.0 / std::sqrt(double(g->n[0]) * double(g->n[1]) * double(g->n[2]));
    if (g->dist) { dist_fft_forward(g, norm); return; }
    double2* c = reinterpret_cast<double2*>(g->d);
    const int n0 = g->n[0], n1 = g->n[1], nzc = g->nzc;
    launch_z(g->n[2], 1, g->d, g->pitch, long(n0) * n1, norm);
    launch_strided(n1, plain_strided(c, c, nzc, ptrdiff_t(n1) * nzc, n0, nzc, n1, 0), ROLE_FFT_Y);
    launch_strided(n0, plain_strided(c, c, ptrdiff_t(n1) * nzc, nzc, n1, nzc, n0, 0), ROLE_FFT_X);
}
void grid_fft_backward(mfb_grid* g, const ZReadout* ro = nullptr) {
    const double norm = 1.0 / std::sqrt(double(g->n[0]) * double(g->n[1]) * double(g->n[2]));
    if (g->dist) { dist_fft_backward(g, norm, ro); return; }
    double2* c = reinterpret_cast<double2*>(g->d);
    const int n0 = g->n[0], n1 = g->n[1], nzc = g->nzc;
    launch_strided(n0, plain_strided(c, c, ptrdiff_t(n1) * nzc, nzc, n1, nzc, n0, 1), ROLE_FFT_X);
    launch_strided(n1, plain_strided(c, c, nzc, ptrdiff_t(n1) * nzc, n0, nzc, n1, 1), ROLE_FFT_Y);
    launch_z(g->n[2], 0, g->d, g->pitch, long(n0) * n1, norm, ro);
}
void to_kspace(mfb_grid* g) {
    if (g->space != MFB_KSPACE) { grid_fft_forward(g); g->space = MFB_KSPACE; }
}
void to_rspace(mfb_grid* g) {
    if (g->space != MFB_RSPACE) { grid_fft_backward(g); g->space = MFB_RSPACE; }
}

double* partial_buffer(size_t ndoubles) {
    if (g_ctx.partial_cap < ndoubles) {
        if (g_ctx.partial) cudaFree(g_ctx.partial);
        g_ctx.partial = (double*)dev_alloc(ndoubles * sizeof(double));
        g_ctx.partial_cap = ndoubles;
    }
    return g_ctx.partial;
}

// returns (sum, sumsq) over the logical elements of the grid in its current space
static void grid_sums(mfb_grid* g, double out[2]) {
    const int nb = 148 * 4;
    double* part = partial_buffer(2 * nb + 2);
    const int x0 = g->dist ? g->nloc0 : g->n[0], k0 = g->dist ? g->nloc1 : g->n[0];
    if (g->space == MFB_RSPACE) {
        MFB_LAUNCH_ROLE(ROLE_REDUCE, reduce_real_kernel, nb, RED_THREADS, g->d, long(x0) * g->n[1], g->n[2], g->pitch, part);
    } else {
        // slab k-layout is [ky_loc][kx][kz]: nloc1 * n0 * nzc elements (n0 * n1 * nzc on a single slab)
        MFB_LAUNCH_ROLE(ROLE_REDUCE, reduce_cplx_kernel, nb, RED_THREADS, reinterpret_cast<const double2*>(g->d),
                   long(k0) * (g->dist ? g->n[0] : g->n[1]) * g->nzc, part);
    }
    check_launch("reduce");
    MFB_LAUNCH_ROLE(ROLE_REDUCE, reduce_final_kernel, 1, RED_THREADS, part, nb, 0, part + 2 * nb);
    check_launch("reduce_final");
    if (g->dist) dist_allreduce_sum2(part + 2 * nb);  // include/grid_fft.hh:532-542: MPI_Allreduce(SUM)
    cuda_check(cudaMemcpyAsync(out, part + 2 * nb, 2 * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream), "reduce D2H");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
}

void grid_zero_dc(mfb_grid* g) {
    if (g->space == MFB_KSPACE) {
        if (!g->dist || g->off1 == 0) cuda_check(cudaMemsetAsync(g->d, 0, sizeof(double2), g_ctx.stream), "zero DC");
    } else {
        double s[2];
        grid_sums(g, s);
        const double mean = s[0] / (double(g->n[0]) * g->n[1] * g->n[2]);
        const long nrows = long(g->dist ? g->nloc0 : g->n[0]) * g->n[1];
        MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, subtract_real_kernel, stream_blocks(size_t(nrows) * g->n[2]), EW_THREADS, g->d, nrows, g->n[2],
                   g->pitch, mean);
        check_launch("subtract_real");
    }
}

void grid_laplacian_pow(mfb_grid* g, int p, double s) {
    to_kspace(g);
    KGeom k = kgeom(g);
    double2* c = reinterpret_cast<double2*>(g->d);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, laplacian_pow_kernel, kgrid_blocks(k), EW_THREADS, c, c, k, p, s);
    check_launch("laplacian_pow");
}

void grid_scale(mfb_grid* g, double a) {
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, scale_kernel, stream_blocks(g->ntot), EW_THREADS, g->d, g->ntot, a);
    check_launch("scale");
}

void grid_shift_field(mfb_grid* g, const double s[3], bool back) {
    to_kspace(g);
    KGeom k = kgeom(g);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, shift_phase_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(g->d), k, s[0], s[1], s[2]);
    check_launch("shift_phase");
    if (back) to_rspace(g);
}

// ------------------------------------------------------------------------------------------
// complex grids (Grid_FFT<ccomplex_t>): c2c transforms with the symmetric norm (src/grid_fft.cc:199-270)
// ------------------------------------------------------------------------------------------
static void cgrid_fft(mfb_cgrid* g, int inverse) {
    const int n0 = g->n[0], n1 = g->n[1], n2 = g->n[2];
    const double norm = 1.0 / std::sqrt(double(n0) * n1 * n2);
    double2* c = g->d;
    if (g->dist) {
        // slabs (FFTW-MPI c2c with transposed output, src/grid_fft.cc:130-133): z on the local rows, y on the local planes with
        // the exchange fused into its stores, x on the received [x][ky_loc] block.  The exchange buffers are sized for the
        // half-spectra of the real grids, so the complex grid goes through in two halves of the z range; a half's x pass
        // rewrites only that half of every row, which the other half's y pass never reads.
        const int l0 = g->nloc0, l1 = g->nloc1, hz = n2 / 2;
        if (!inverse) {
            launch_zc2c(n2, c, long(l0) * n1, norm, 0);
            for (int z0 = 0; z0 < n2; z0 += hz) {
                Exchange x = begin_exchange(size_t(l0) * l1 * hz);
                StridedArgs a = plain_strided(c + z0, nullptr, n2, ptrdiff_t(n1) * n2, l0, hz, n1, 0);
                route(a, x, l1, hz, ptrdiff_t(l1) * hz);                          // chunk layout [x_loc][ky_loc][z half]
                launch_strided(n1, a, ROLE_FFT_Y);
                end_exchange(x);                                                   // recv: [x][ky_loc][z half]
                StridedArgs b = plain_strided(x.recv, c + z0, ptrdiff_t(l1) * hz, hz, l1, hz, n0, 0);
                b.out_ps = n2; b.out_as = ptrdiff_t(n0) * n2;                      // -> [ky_loc][kx][kz]
                launch_strided(n0, b, ROLE_FFT_X);
            }
        } else {
            for (int z0 = 0; z0 < n2; z0 += hz) {
                Exchange x = begin_exchange(size_t(l0) * l1 * hz);
                StridedArgs a = plain_strided(c + z0, nullptr, n2, ptrdiff_t(n0) * n2, l1, hz, n0, 1);
                route(a, x, l0, hz, ptrdiff_t(l0) * hz);                          // chunk layout [ky_loc][x_loc][z half]
                launch_strided(n0, a, ROLE_FFT_X);
                end_exchange(x);                                                   // recv: [ky][x_loc][z half]
                StridedArgs b = plain_strided(x.recv, c + z0, ptrdiff_t(l0) * hz, hz, l0, hz, n1, 1);
                b.out_ps = n2; b.out_as = ptrdiff_t(n1) * n2;                      // -> [x_loc][y][z]
                launch_strided(n1, b, ROLE_FFT_Y);
            }
            launch_zc2c(n2, c, long(l0) * n1, norm, 1);
        }
        return;
    }
    if (!inverse) {
        launch_zc2c(n2, c, long(n0) * n1, norm, 0);
        launch_strided(n1, plain_strided(c, c, n2, ptrdiff_t(n1) * n2, n0, n2, n1, 0), ROLE_FFT_Y);
        launch_strided(n0, plain_strided(c, c, ptrdiff_t(n1) * n2, n2, n1, n2, n0, 0), ROLE_FFT_X);
    } else {
        launch_strided(n0, plain_strided(c, c, ptrdiff_t(n1) * n2, n2, n1, n2, n0, 1), ROLE_FFT_X);
        launch_strided(n1, plain_strided(c, c, n2, ptrdiff_t(n1) * n2, n0, n2, n1, 1), ROLE_FFT_Y);
        launch_zc2c(n2, c, long(n0) * n1, norm, 1);
    }
}
static CGeom cgeom(const mfb_cgrid* g) {
    CGeom c;
    c.n0 = g->n[0]; c.n1 = g->n[1]; c.n2 = g->n[2];
    c.kf0 = 2.0 * M_PI / g->L[0]; c.kf1 = 2.0 * M_PI / g->L[1]; c.kf2 = 2.0 * M_PI / g->L[2];
    c.rpitch = g->n[2] + 2;
    c.s0 = g->n[0]; c.s1 = g->n[1]; c.tr = 0; c.off = 0;
    if (g->dist) {
        if (g->space == MFB_KSPACE) { c.s0 = g->nloc1; c.s1 = g->n[0]; c.tr = 1; c.off = g->off1; }
        else c.s0 = g->nloc0;
    }
    return c;
}
static size_t cgrid_local_elems(const mfb_cgrid* g) {   // elements of the current layout
    if (!g->dist) return g->nelem;
    return (g->space == MFB_KSPACE ? size_t(g->nloc1) * g->n[0] : size_t(g->nloc0) * g->n[1]) * g->n[2];
}
static void check_ppt_real(const mfb_cgrid* c, const mfb_grid* r, const char* name) {
    if (!r || r->dist != c->dist || r->n[0] != c->n[0] || r->n[1] != c->n[1] || r->n[2] != c->n[2])
        throw std::runtime_error(std::string("PPT: real grid '") + name + "' must be a grid of the same size and decomposition");
    if (r->space != MFB_RSPACE) throw std::runtime_error(std::string("PPT: real grid '") + name + "' must be in real space");
}

// ------------------------------------------------------------------------------------------
// Orszag convolver
// ------------------------------------------------------------------------------------------
struct Operand {
    mfb_grid* g;
    int a1, b1, a2, b2;
    double c1, c2;
    int imag;
    int keep_nyq = 0;  // factor built from get_k (Nyquist = +N/2) instead of gradient (Nyquist = 0); NaiveConvolver only
};

static void set_operand_factor(StridedArgs& a, mfb_conv* c, const Operand& o, double s1) {
    a.f.enabled = 1;
    a.f.a1 = o.a1; a.f.b1 = o.b1; a.f.a2 = o.a2; a.f.b2 = o.b2;
    a.f.c1 = o.c1 * s1; a.f.c2 = o.c2 * s1;
    a.f.imag = o.imag;
    a.f.keep_nyq = o.keep_nyq;
    for (int d = 0; d < 3; ++d) { a.f.n[d] = c->n[d]; a.f.kf[d] = 2.0 * M_PI / c->L[d]; }
}

// slab-decomposed operand transform: x pass (N0 -> P0, on the local ky rows) -> exchange of the pruned
// 1.5 R(N)/P intermediate -> y pass (N1 -> P1) on the local x planes of the padded grid.
static void dist_conv_backward_operand(mfb_conv* c, const Operand& o, double2* W2) {
    const int N0 = c->n[0], N1 = c->n[1], N2 = c->n[2];
    const int P0 = c->np[0], P1 = c->np[1];
    const int nzc = N2 / 2 + 1, nzin = N2 / 2;
    const int l1 = o.g->nloc1, p0l = c->p0loc;
    const double ptot = double(c->np[0]) * c->np[1] * c->np[2];
    const double s1 = std::pow(1.5, 1.5) / std::sqrt(ptot);
    Exchange x = begin_exchange(size_t(l1) * p0l * nzin);
    StridedArgs a = plain_strided(reinterpret_cast<const double2*>(o.g->d), nullptr, nzc, ptrdiff_t(N0) * nzc, l1, nzin, P0, 1);
    a.n_in = N0;
    a.a_off = o.g->off1;
    a.a_skip = N1 / 2;
    set_operand_factor(a, c, o, s1);
    a.pos_axis = 0; a.a_axis = 1;
    route(a, x, p0l, nzin, ptrdiff_t(p0l) * nzin);                      // chunk layout [ky_loc][x_loc][kz]
    launch_strided(P0, a, ROLE_CONV_X_BWD);
    end_exchange(x);                                                     // recv: [ky][x_loc][kz]
    StridedArgs b = plain_strided(x.recv, W2, ptrdiff_t(p0l) * nzin, nzin, p0l, nzin, P1, 1);
    b.out_ps = nzin; b.out_as = ptrdiff_t(P1) * nzin;                    // -> [x_loc][P1][kz]
    b.n_in = N1;
    launch_strided(P1, b, ROLE_CONV_Y_BWD);
}

static void conv_backward_operand(mfb_conv* c, const Operand& o, double2* W2) {
    if (c->dist) { dist_conv_backward_operand(c, o, W2); return; }
    const int N0 = c->n[0], N1 = c->n[1], N2 = c->n[2];
    const int P0 = c->np[0], P1 = c->np[1];
    const int nzc = N2 / 2 + 1, nzin = N2 / 2;
    const double ptot = double(c->np[0]) * c->np[1] * c->np[2];
    const double s1 = std::pow(1.5, 1.5) / std::sqrt(ptot);  // pad factor (convolution.hh:572) * bwd norm
    // pass 1: x axis, N0 -> P0, operand factor fused into the load
    StridedArgs a = plain_strided(reinterpret_cast<const double2*>(o.g->d), c->W1, 0, 0, N1, nzin, P0, 1);
    a.in_ps = ptrdiff_t(N1) * nzc; a.in_as = nzc;
    a.out_ps = ptrdiff_t(N1) * nzin; a.out_as = nzin;
    a.n_in = N0;
    a.a_skip = N1 / 2;
    set_operand_factor(a, c, o, s1);
    a.pos_axis = 0; a.a_axis = 1;
    launch_strided(P0, a, ROLE_CONV_X_BWD);
    // pass 2: y axis, N1 -> P1
    StridedArgs b = plain_strided(c->W1, W2, nzin, 0, P0, nzin, P1, 1);
    b.in_as = ptrdiff_t(N1) * nzin;
    b.out_as = ptrdiff_t(P1) * nzin;
    b.n_in = N1;
    launch_strided(P1, b, ROLE_CONV_Y_BWD);
}

// ---- NaiveConvolver (include/convolution.hh:296-426, compile-time CONVOLVER_TYPE=NAIVE): same named products,
// no padding - copy_in(kfunc) over EVERY stored mode incl. the Nyquist planes, backward FFT, product over the
// real-indexed buffer, forward FFT, output operator over every stored mode (no Nyquist zeroing, no rescale).
// copy_in is fused into the x sweep of the backward transform, the output operator into the x sweep of the forward one.
static void naive_backward_operand(mfb_conv* c, const Operand& o, mfb_grid* dst) {
    const int N0 = c->n[0], N1 = c->n[1], N2 = c->n[2], nzc = N2 / 2 + 1;
    to_kspace(o.g);
    double2* d = reinterpret_cast<double2*>(dst->d);
    if (o.g->dist) {
        // slabs: the operand factor rides on the x pass of the transposed layout [ky_loc][kx][kz], whose output is routed to the
        // ranks owning the x planes; then y and z on the local planes (dist_fft_backward, out of place)
        const int l0 = dst->nloc0, l1 = o.g->nloc1;
        const int zp = dist_zpitch(nzc);
        Exchange x = begin_exchange(size_t(l0) * l1 * zp);
        StridedArgs a = plain_strided(reinterpret_cast<const double2*>(o.g->d), nullptr, nzc, ptrdiff_t(N0) * nzc, l1, nzc, N0, 1);
        set_operand_factor(a, c, o, 1.0);
        a.pos_axis = 0; a.a_axis = 1;
        a.a_off = o.g->off1;
        route(a, x, l0, zp, ptrdiff_t(l0) * zp);                           // chunk layout [ky_loc][x_loc][kz (pitch zp)]
        launch_strided(N0, a, ROLE_FFT_X);
        end_exchange(x);                                                    // recv: [ky][x_loc][kz]
        StridedArgs b = plain_strided(x.recv, d, ptrdiff_t(l0) * zp, zp, l0, nzc, N1, 1);
        b.out_ps = nzc; b.out_as = ptrdiff_t(N1) * nzc;                     // -> [x_loc][y][kz]
        launch_strided(N1, b, ROLE_FFT_Y);
        launch_z(N2, 0, dst->d, dst->pitch, long(l0) * N1, 1.0 / std::sqrt(double(N0) * N1 * N2));
        dst->space = MFB_RSPACE;
        return;
    }
    StridedArgs a = plain_strided(reinterpret_cast<const double2*>(o.g->d), d, ptrdiff_t(N1) * nzc, nzc, N1, nzc, N0, 1);
    set_operand_factor(a, c, o, 1.0);
    a.pos_axis = 0; a.a_axis = 1;
    launch_strided(N0, a, ROLE_FFT_X);
    launch_strided(N1, plain_strided(d, d, nzc, ptrdiff_t(N1) * nzc, N0, nzc, N1, 1), ROLE_FFT_Y);
    launch_z(N2, 0, dst->d, dst->pitch, long(N0) * N1, 1.0 / std::sqrt(double(N0) * N1 * N2));
    dst->space = MFB_RSPACE;
}
static void naive_multiply(mfb_grid* y, const mfb_grid* x) {
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, multiply_kernel, stream_blocks(y->ntot), EW_THREADS, y->d, x->d, y->ntot);
    check_launch("multiply");
}
static void naive_forward_output(mfb_conv* c, mfb_grid* src, mfb_grid* out, int op, double alpha) {
    const int N0 = c->n[0], N1 = c->n[1], N2 = c->n[2], nzc = N2 / 2 + 1;
    require(out->space == MFB_KSPACE, "convolution output grid must be in k-space");
    double2* d = reinterpret_cast<double2*>(src->d);
    if (src->dist) {   // dist_fft_forward with the output operator on the last (x) pass
        const int l0 = src->nloc0, l1 = out->nloc1;
        launch_z(N2, 1, src->d, src->pitch, long(l0) * N1, 1.0 / std::sqrt(double(N0) * N1 * N2));
        const int zp = dist_zpitch(nzc);
        Exchange x = begin_exchange(size_t(l0) * l1 * zp);
        StridedArgs a = plain_strided(d, nullptr, nzc, ptrdiff_t(N1) * nzc, l0, nzc, N1, 0);
        route(a, x, l1, zp, ptrdiff_t(l1) * zp);                           // chunk layout [x_loc][ky_loc][kz (pitch zp)]
        launch_strided(N1, a, ROLE_FFT_Y);
        end_exchange(x);                                                    // recv: [x][ky_loc][kz]
        StridedArgs e = plain_strided(x.recv, reinterpret_cast<double2*>(out->d), ptrdiff_t(l1) * zp, zp, l1, nzc, N0, 0);
        e.out_ps = nzc; e.out_as = ptrdiff_t(N0) * nzc;                     // -> [ky_loc][kx][kz]
        e.op = op; e.alpha = alpha;
        launch_strided(N0, e, ROLE_FFT_X);
        src->space = MFB_KSPACE;
        ++c->nconv;
        return;
    }
    launch_z(N2, 1, src->d, src->pitch, long(N0) * N1, 1.0 / std::sqrt(double(N0) * N1 * N2));
    launch_strided(N1, plain_strided(d, d, nzc, ptrdiff_t(N1) * nzc, N0, nzc, N1, 0), ROLE_FFT_Y);
    StridedArgs e = plain_strided(d, reinterpret_cast<double2*>(out->d), ptrdiff_t(N1) * nzc, nzc, N1, nzc, N0, 0);
    e.op = op; e.alpha = alpha;
    launch_strided(N0, e, ROLE_FFT_X);
    src->space = MFB_KSPACE;
    ++c->nconv;
}
static void naive_conv(mfb_conv* c, const Operand& A, const Operand& B, const Operand* C3, mfb_grid* out, int op, double alpha) {
    naive_backward_operand(c, A, c->nb1);
    naive_backward_operand(c, B, c->nb2);
    naive_multiply(c->nb2, c->nb1);
    if (C3) {   // convolve3: all three fields multiplied in real space (:361-401), no intermediate truncation
        naive_backward_operand(c, *C3, c->nb1);
        naive_multiply(c->nb2, c->nb1);
    }
    naive_forward_output(c, c->nb2, out, op, alpha);
}

// a borrowed accumulator lives in the particle arrays: streamed hand-off copies of the previous run must have drained
static void conv_acquire_accumulator(mfb_conv* c) {
    if (!c->acc_borrowed || !c->lender) return;
    for (int at : {MFB_ATTR_POS, MFB_ATTR_VEL}) {
        if (c->lender->copy_pending[at]) {
            cuda_check(cudaStreamWaitEvent(g_ctx.stream, c->lender->copied[at], 0), "cudaStreamWaitEvent");
            c->lender->copy_pending[at] = false;
        }
    }
}
// lent particle memory in any role (accumulator, slots, chunk workspace)
static void conv_wait_lender(mfb_conv* c) {
    if (!c->lender) return;
    for (int at : {MFB_ATTR_POS, MFB_ATTR_VEL}) {
        if (c->lender->copy_pending[at]) {
            cuda_check(cudaStreamWaitEvent(g_ctx.stream, c->lender->copied[at], 0), "cudaStreamWaitEvent");
            c->lender->copy_pending[at] = false;
        }
    }
}
static mfb_operand_key operand_key(const Operand& o) {
    // (a,b) and (b,a) name the same Hessian: store the axis pairs sorted
    mfb_operand_key k;
    k.g = o.g;
    k.a1 = o.a1 < o.b1 ? o.a1 : o.b1; k.b1 = o.a1 < o.b1 ? o.b1 : o.a1;
    k.a2 = o.a2 < o.b2 ? o.a2 : o.b2; k.b2 = o.a2 < o.b2 ? o.b2 : o.a2;
    k.c1 = o.c1; k.c2 = o.c2; k.imag = o.imag; k.keep_nyq = o.keep_nyq; k.valid = true;
    return k;
}
static bool same_key(const mfb_operand_key& a, const mfb_operand_key& b) {
    return a.valid && b.valid && a.g == b.g && a.a1 == b.a1 && a.b1 == b.b1 && a.a2 == b.a2 && a.b2 == b.b2 && a.c1 == b.c1 &&
           a.c2 == b.c2 && a.imag == b.imag && a.keep_nyq == b.keep_nyq;
}
static bool same_operand(const Operand& a, const Operand& b) { return same_key(operand_key(a), operand_key(b)); }
static void conv_cache_clear(mfb_conv* c) { c->slot[0].valid = c->slot[1].valid = false; }
static void conv_cache_invalidate(mfb_conv* c, const mfb_grid* g) {
    for (int i = 0; i < 2; ++i) if (c->slot[i].g == g) c->slot[i].valid = false;
}

// ---- a GROUP of dealiased products with one target: out (op)= unpad( F[ sum_i coeff_i * a_i * b_i ] ).
// The per-product part is the two operand transforms and the fused z pass, which adds coeff_i * a_i b_i to the running
// sum in z-Fourier space (the padded accumulator Wacc, [P0][P1][N2/2]); the forward y and x passes, the truncation and the
// output operator run ONCE per group.  Exact by linearity of the transforms and of the truncation; the
// reference does one forward transform per product (include/convolution.hh:524-534).  Without Wacc (not enough HBM)
// every product is its own group and the z pass works in place on an operand buffer.
// A product of an operand with itself transforms it once, and inside mfb_lpt_sources (cache_on) an operand that is still
// sitting in W2a / W2b from the previous product is not transformed again.
struct ConvGroup {
    mfb_conv* c;
    bool inplace;   // single product whose z pass overwrites its first operand buffer (leaves Wacc to an enclosing group)
    int count = 0;
    double2* sum = nullptr;
    explicit ConvGroup(mfb_conv* c_, bool inplace_ = false) : c(c_), inplace(inplace_ || c_->Wacc == nullptr || !g_ctx.conv_grouping) {}
    int cached(const mfb_operand_key& k) const {
        if (!c->cache_on) return -1;
        for (int i = 0; i < 2; ++i) if (same_key(c->slot[i], k)) return i;
        return -1;
    }
    void add(const Operand& A, const Operand& B, double coeff) {
        require(count == 0 || !inplace, "internal: product groups need the accumulator workspace");
        to_kspace(A.g);
        to_kspace(B.g);
        const int P1 = c->np[1], P2 = c->np[2], nzin = c->n[2] / 2;
        double2* const buf[2] = {c->W2a, c->W2b};
        const mfb_operand_key ka = operand_key(A), kb = operand_key(B);
        const bool square = same_key(ka, kb);
        int ia = cached(ka), ib = square ? ia : cached(kb);
        auto transform = [&](const Operand& o, const mfb_operand_key& k, int i) {
            conv_backward_operand(c, o, buf[i]);
            c->slot[i] = k;
        };
        if (square) {
            if (ia < 0) { ia = c->slot[0].valid && !c->slot[1].valid ? 1 : 0; transform(A, ka, ia); }
            ib = ia;
        } else if (ia >= 0 && ib >= 0) {
            // both operands are still there
        } else if (ia >= 0) { ib = 1 - ia; transform(B, kb, ib); }
        else if (ib >= 0) { ia = 1 - ib; transform(A, ka, ia); }
        else { ia = 0; ib = 1; transform(A, ka, 0); transform(B, kb, 1); }
        sum = inplace ? buf[ia] : c->Wacc;
        if (!inplace && count == 0) conv_acquire_accumulator(c);
        launch_zconv(P2, buf[ia], buf[ib], sum, nzin, nzin, long(c->p0loc) * P1, nzin, coeff, count > 0);
        if (inplace) c->slot[ia].valid = false;   // overwritten by the product
        ++count;
        ++c->nconv;
    }
    void finish(mfb_grid* out, int op, double alpha) {
        require(count > 0, "internal: empty product group");
        require(out->space == MFB_KSPACE, "convolution output grid must be in k-space");
        conv_cache_invalidate(c, out);   // the target changes: an operand transformed from its old contents is stale
        const int N0 = c->n[0], N1 = c->n[1], N2 = c->n[2];
        const int P0 = c->np[0], P1 = c->np[1], P2 = c->np[2];
        const int nzc = N2 / 2 + 1, nzin = N2 / 2;
        const double ptot = double(P0) * P1 * P2, ntot = double(N0) * N1 * N2;
        const double unpad_scale = (1.0 / std::sqrt(ptot)) / (std::sqrt(ptot) / std::sqrt(ntot));  // fwd norm / rfac (convolution.hh:620,664)
        if (c->dist) {
            const int p0l = c->p0loc, l1 = out->nloc1;
            // y forward P1 -> N1 on the local x planes, routed to the owners of the ky rows
            Exchange x = begin_exchange(size_t(p0l) * l1 * nzin);
            StridedArgs d = plain_strided(sum, nullptr, nzin, ptrdiff_t(P1) * nzin, p0l, nzin, P1, 0);
            d.n_out = N1;
            route(d, x, l1, nzin, ptrdiff_t(l1) * nzin);                    // chunk layout [x_loc][ky_loc][kz]
            launch_strided(P1, d, ROLE_CONV_Y_FWD);
            end_exchange(x);                                                 // recv: [x (P0)][ky_loc][kz]
            // x forward P0 -> N0 + unpad scale + Nyquist zero + output operator, into [ky_loc][kx][kz]
            StridedArgs e = plain_strided(x.recv, reinterpret_cast<double2*>(out->d), ptrdiff_t(l1) * nzin, nzin, l1,
                                          (op == OP_ASSIGN) ? nzc : nzin, P0, 0);
            e.out_ps = nzc; e.out_as = ptrdiff_t(N0) * nzc;
            e.nz_fft = nzin;
            e.n_out = N0;
            e.scale = unpad_scale;
            e.op = op; e.alpha = alpha;
            e.a_off = out->off1;
            e.a_skip = N1 / 2;
            e.zero_fill = 1;
            launch_strided(P0, e, ROLE_CONV_X_FWD);
            return;
        }
        // y axis forward, P1 -> N1
        StridedArgs d = plain_strided(sum, c->W1, nzin, 0, P0, nzin, P1, 0);
        d.in_as = ptrdiff_t(P1) * nzin;
        d.out_as = ptrdiff_t(N1) * nzin;
        d.n_out = N1;
        launch_strided(P1, d, ROLE_CONV_Y_FWD);
        // x axis forward, P0 -> N0, unpad scale + Nyquist zero + output operator
        StridedArgs e = plain_strided(c->W1, reinterpret_cast<double2*>(out->d), 0, 0, N1, (op == OP_ASSIGN) ? nzc : nzin, P0, 0);
        e.in_ps = ptrdiff_t(N1) * nzin; e.in_as = nzin;
        e.out_ps = ptrdiff_t(N1) * nzc; e.out_as = nzc;
        e.nz_fft = nzin;
        e.n_out = N0;
        e.scale = unpad_scale;
        e.op = op; e.alpha = alpha;
        e.a_skip = N1 / 2;
        e.zero_fill = 1;
        launch_strided(P0, e, ROLE_CONV_X_FWD);
    }
};
static bool can_group(const mfb_conv* c) { return !c->naive && c->Wacc != nullptr && g_ctx.conv_grouping; }
static void conv_legacy_ensure(mfb_conv* c);   // mfb_sources2.cuh: the per-product workspace is in place (not the slot workspace)

void conv2(mfb_conv* c, const Operand& A, const Operand& B, mfb_grid* out, int op, double alpha) {
    require(out->space == MFB_KSPACE, "convolution output grid must be in k-space");
    if (c->naive) { naive_conv(c, A, B, nullptr, out, op, alpha); return; }
    conv_legacy_ensure(c);
    c->cache_on = false;
    ConvGroup g(c);
    g.add(A, B, 1.0);
    g.finish(out, op, alpha);
}

static Operand hessian_operand(mfb_grid* g, const int d[2]) { return Operand{g, d[0], d[1], -1, -1, -1.0, 0.0, 0}; }

// three-field product (include/convolution.hh:542-549 Orszag: two dealiased products with the intermediate truncated
// to the N-grid; :361-401 Naive: one real-space triple product)
void conv3(mfb_conv* c, const Operand& A, const Operand& B, const Operand& C3, mfb_grid* out, int op, double alpha) {
    require(out->space == MFB_KSPACE, "convolution output grid must be in k-space");
    if (c->naive) { naive_conv(c, A, B, &C3, out, op, alpha); return; }
    conv_legacy_ensure(c);
    to_kspace(C3.g);
    c->fbuf->space = MFB_KSPACE;
    conv2(c, A, B, c->fbuf, OP_ASSIGN, 0.0);
    conv2(c, Operand{c->fbuf, -1, -1, -1, -1, 1.0, 0.0, 0}, C3, out, op, alpha);
}

}  // namespace mfb
#include "mfb_sources2.cuh"
namespace mfb {

void lpt_sources(mfb_conv* c, int order, double D, mfb_grid* phi, mfb_grid* phi2, mfb_grid* phi3, mfb_grid* const A3[3],
                 const double* lam) {
    require(phi->space == MFB_KSPACE, "phi must be in k-space");
    if (lpt_sources_v2(c, order, D, phi, phi2, phi3, A3, lam)) return;
    conv_legacy_ensure(c);
    c->last_path = 1; c->last_slots = 0; c->last_loops = 0; c->last_xpass = c->last_ypass = c->last_zlaunch = 0;
    const double g1 = -D;
    const double g2 = (order > 1) ? -3.0 / 7.0 * D * D : 0.0;
    const double g3 = (order > 2) ? 1.0 / 3.0 * D * D * D : 0.0;
    const double g3c = (order > 2) ? 1.0 / 7.0 * D * D * D : 0.0;
    auto H = [](mfb_grid* g, int a, int b) { int d[2] = {a, b}; return hessian_operand(g, d); };
    auto HS = [](mfb_grid* g, int a, int b, int cc, int d, double sgn) { return Operand{g, a, b, cc, d, -1.0, -sgn, 0}; };
    const bool grouped = can_group(c);
    // The product ORDER inside a group (and of the groups) is free; it is chosen so that consecutive products share an
    // operand, which then stays in its buffer (cache_on): 49 operand transforms instead of 64 at 3LPT.
    conv_cache_clear(c);
    c->cache_on = grouped;
    if (order > 1 && grouped) {
        require(phi2 != nullptr, "phi2 grid required for LPT order > 1");
        phi2->space = MFB_KSPACE;
        // src/ic_generator.cc:482-486, the five products summed before ONE forward transform
        ConvGroup g(c);
        g.add(H(phi, 0, 0), HS(phi, 1, 1, 2, 2, +1.0), 1.0);
        g.add(H(phi, 1, 1), H(phi, 2, 2), 1.0);
        g.add(H(phi, 0, 2), H(phi, 0, 2), -1.0);
        g.add(H(phi, 1, 2), H(phi, 1, 2), -1.0);
        g.add(H(phi, 0, 1), H(phi, 0, 1), -1.0);
        g.finish(phi2, OP_ASSIGN, 0.0);
    } else if (order > 1) {
        require(phi2 != nullptr, "phi2 grid required for LPT order > 1");
        phi2->space = MFB_KSPACE;
        // src/ic_generator.cc:482-486
        conv2(c, H(phi, 0, 0), HS(phi, 1, 1, 2, 2, +1.0), phi2, OP_ASSIGN, 0.0);
        conv2(c, H(phi, 1, 1), H(phi, 2, 2), phi2, OP_ADD, 0.0);
        conv2(c, H(phi, 0, 1), H(phi, 0, 1), phi2, OP_SUB, 0.0);
        conv2(c, H(phi, 0, 2), H(phi, 0, 2), phi2, OP_SUB, 0.0);
        conv2(c, H(phi, 1, 2), H(phi, 1, 2), phi2, OP_SUB, 0.0);
    }
    if (order > 1) {
        if (lam) {
            KGeom k = kgeom(phi2);
            MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, tides_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(phi2->d),
                       reinterpret_cast<const double2*>(phi->d), k, lam[0], lam[1], lam[2]);
            check_launch("tides");
        }
        grid_laplacian_pow(phi2, -1, -1.0);  // :499
    }
    if (order > 2) require(phi3 && A3 && A3[0] && A3[1] && A3[2], "phi3 and A3 grids required for LPT order > 2");
    if (order > 2 && grouped) {
        phi3->space = MFB_KSPACE;
        mfb_grid* fb = c->fbuf;
        fb->space = MFB_KSPACE;
        const Operand FB{fb, -1, -1, -1, -1, 1.0, 0.0, 0};
        // 3a :522-526.  The inner products of convolve3 are truncated to the N-grid before the outer product
        // (include/convolution.hh:542-549) - that stays; but truncation is linear, so inner products that meet the same
        // outer field are summed first:  det = T(xx yy - xy xy) zz + 2 T(xy xz) yz - T(yz yz) xx - T(xz xz) yy.
        {
            ConvGroup in(c);
            in.add(H(phi, 0, 1), H(phi, 0, 1), -1.0);   // still in its buffer when phi2 was the previous target
            in.add(H(phi, 0, 0), H(phi, 1, 1), 1.0);
            in.finish(fb, OP_ASSIGN, 0.0);
        }
        ConvGroup gp3(c);   // every term of phi3 (3a outer products and 3b :533-538) goes into one sum
        auto inner = [&](const Operand& a, const Operand& b) {
            ConvGroup in(c, true);   // the accumulator belongs to gp3: single product, in place
            in.add(a, b, 1.0);
            in.finish(fb, OP_ASSIGN, 0.0);
        };
        gp3.add(FB, H(phi, 2, 2), 1.0);
        gp3.add(H(phi, 2, 2), HS(phi2, 0, 0, 1, 1, +1.0), -5.0 / 7.0);
        inner(H(phi, 0, 1), H(phi, 0, 2));
        gp3.add(FB, H(phi, 1, 2), 2.0);
        gp3.add(H(phi, 1, 2), H(phi2, 1, 2), 10.0 / 7.0);
        inner(H(phi, 1, 2), H(phi, 1, 2));
        gp3.add(FB, H(phi, 0, 0), -1.0);
        gp3.add(H(phi, 0, 0), HS(phi2, 1, 1, 2, 2, +1.0), -5.0 / 7.0);
        inner(H(phi, 0, 2), H(phi, 0, 2));
        gp3.add(FB, H(phi, 1, 1), -1.0);
        gp3.add(H(phi, 1, 1), HS(phi2, 2, 2, 0, 0, +1.0), -5.0 / 7.0);
        gp3.add(H(phi, 0, 1), H(phi2, 0, 1), 10.0 / 7.0);
        gp3.add(H(phi, 0, 2), H(phi2, 0, 2), 10.0 / 7.0);
        gp3.finish(phi3, OP_ASSIGN, 0.0);
        grid_laplacian_pow(phi3, -1, -g3);   // :539 and the growth factor of :566 (phi3 is not read again before it) in one sweep
        // transverse term  :545-556:  A_d = P1 - P2 + P3 - P4
        for (int d = 0; d < 3; ++d) {
            const int p = (d + 1) % 3, q = (d + 2) % 3;
            mfb_grid* A = A3[d];
            A->space = MFB_KSPACE;
            ConvGroup ga(c);
            if (d == 0) {   // start with phi2,xz (left over from the last phi3 product), end with phi2,yz (first of d = 1)
                ga.add(H(phi2, d, q), H(phi, d, p), -1.0);
                ga.add(H(phi2, d, p), H(phi, d, q), 1.0);
            } else {        // d = 1 starts with phi2,yz and ends with phi2,zx, which d = 2 starts with
                ga.add(H(phi2, d, p), H(phi, d, q), 1.0);
                ga.add(H(phi2, d, q), H(phi, d, p), -1.0);
            }
            ga.add(H(phi, p, q), HS(phi2, p, p, q, q, -1.0), 1.0);
            ga.add(H(phi2, p, q), HS(phi, p, p, q, q, -1.0), -1.0);
            ga.finish(A, OP_ASSIGN, 0.0);
            grid_laplacian_pow(A, -1, -g3c);     // :555 and :568-570 in one sweep
        }
    } else if (order > 2) {
        phi3->space = MFB_KSPACE;
        // 3a  :522-526
        conv3(c, H(phi, 0, 0), H(phi, 1, 1), H(phi, 2, 2), phi3, OP_ASSIGN, 0.0);
        conv3(c, H(phi, 0, 1), H(phi, 0, 2), H(phi, 1, 2), phi3, OP_MULADD, 2.0);
        conv3(c, H(phi, 1, 2), H(phi, 1, 2), H(phi, 0, 0), phi3, OP_SUB, 0.0);
        conv3(c, H(phi, 0, 2), H(phi, 0, 2), H(phi, 1, 1), phi3, OP_SUB, 0.0);
        conv3(c, H(phi, 0, 1), H(phi, 0, 1), H(phi, 2, 2), phi3, OP_SUB, 0.0);
        // 3b  :533-538
        conv2(c, H(phi, 0, 0), HS(phi2, 1, 1, 2, 2, +1.0), phi3, OP_MULADD, -5.0 / 7.0);
        conv2(c, H(phi, 1, 1), HS(phi2, 2, 2, 0, 0, +1.0), phi3, OP_MULADD, -5.0 / 7.0);
        conv2(c, H(phi, 2, 2), HS(phi2, 0, 0, 1, 1, +1.0), phi3, OP_MULADD, -5.0 / 7.0);
        conv2(c, H(phi, 0, 1), H(phi2, 0, 1), phi3, OP_MULADD, 10.0 / 7.0);
        conv2(c, H(phi, 0, 2), H(phi2, 0, 2), phi3, OP_MULADD, 10.0 / 7.0);
        conv2(c, H(phi, 1, 2), H(phi2, 1, 2), phi3, OP_MULADD, 10.0 / 7.0);
        grid_laplacian_pow(phi3, -1, -1.0);  // :539
        // transverse term  :545-556
        for (int d = 0; d < 3; ++d) {
            const int p = (d + 1) % 3, q = (d + 2) % 3;
            mfb_grid* A = A3[d];
            A->space = MFB_KSPACE;
            conv2(c, H(phi2, d, p), H(phi, d, q), A, OP_ASSIGN, 0.0);
            conv2(c, H(phi2, d, q), H(phi, d, p), A, OP_SUB, 0.0);
            conv2(c, H(phi, p, q), HS(phi2, p, p, q, q, -1.0), A, OP_ADD, 0.0);
            conv2(c, H(phi2, p, q), HS(phi, p, p, q, q, -1.0), A, OP_SUB, 0.0);
            grid_laplacian_pow(A, -1, -1.0);
        }
    }
    c->cache_on = false;
    conv_cache_clear(c);
    // :561-574
    grid_scale(phi, g1);
    if (order > 1) grid_scale(phi2, g2);
    if (order > 2 && !grouped) {
        grid_scale(phi3, g3);
        for (int d = 0; d < 3; ++d) grid_scale(A3[d], g3c);
    }
}

// ------------------------------------------------------------------------------------------
// particles
// ------------------------------------------------------------------------------------------
static const double LATTICE_SHIFTS[4][8][3] = {
    {{0, 0, 0}},
    {{0, 0, 0}, {0.5, 0.5, 0.5}},
    {{0, 0, 0}, {0, 0.5, 0.5}, {0.5, 0, 0.5}, {0.5, 0.5, 0}},
    {{0, 0, 0}, {0, 0, 0.5}, {0, 0.5, 0}, {0, 0.5, 0.5}, {0.5, 0, 0}, {0.5, 0, 0.5}, {0.5, 0.5, 0}, {0.5, 0.5, 0.5}},
};
static const double SECOND_LATTICE_SHIFT[4][3] = {{0.5, 0.5, 0.5}, {0.5, 0.5, 0}, {0.5, 0.5, 0.5}, {0.25, 0.25, 0.25}};

static mfb::LatticeMask make_mask(const mfb_particles* p) {
    mfb::LatticeMask M;
    for (int s = 0; s < 8; ++s) M.sel[s] = p->mask[s] == p->mask_index ? 1 : 0;
    for (int ab = 0; ab < 4; ++ab) {
        M.pair[ab] = M.sel[2 * ab] + M.sel[2 * ab + 1];
        M.row[ab] = long(p->n[2] / 2) * M.pair[ab];
    }
    for (int a = 0; a < 2; ++a) M.plane[a] = long(p->n[1] / 2) * (M.row[2 * a] + M.row[2 * a + 1]);
    return M;
}
// kind 0: positions, 1: velocities, 2: masses
static void particles_readout(mfb_particles* p, int kind, int lattice_index, int idim, double unit, mfb_grid* f) {
    require(f->n[0] == p->n[0] && f->n[1] == p->n[1] && f->n[2] == p->n[2], "field/particle grid size mismatch");
    if (p->lattice_type == MFB_LATTICE_MASKED || p->lattice_type == MFB_LATTICE_GLASS) {
        const bool masked = p->lattice_type == MFB_LATTICE_MASKED;
        if (masked) require(lattice_index == p->mask_index, "For masked lattice type, lattice index must be 0 or 1.");
        require(masked || kind != 2, "cannot have individual particle masses for glasses");
        require(f->nloc0 == p->nloc0 && f->off0 == p->off0, "field/particle slab mismatch");
        const int at = kind == 2 ? MFB_ATTR_MASS : kind;
        if (p->copy_pending[at]) {
            cuda_check(cudaStreamWaitEvent(g_ctx.stream, p->copied[at], 0), "cudaStreamWaitEvent");
            p->copy_pending[at] = false;
        }
        to_rspace(f);
        void* dst = (kind == 0) ? p->pos : ((kind == 1) ? p->vel : p->mass);
        require(dst != nullptr, "particle buffer not allocated");
        if (masked) {
            const mfb::LatticeMask M = make_mask(p);
            const double u = (kind == 2) ? unit / double(p->np_global) : unit;   // pmeanmass = munit / global_num_particles_ (:363)
            const unsigned nb = stream_blocks(size_t(p->nloc0) * p->n[1] * p->n[2]);
            if (p->real64) {
                MFB_LAUNCH_ROLE(ROLE_PARTICLES, masked_readout_kernel<double>, nb, EW_THREADS, (double*)dst, f->d, M, f->nloc0, f->n[0], f->n[1],
                           f->n[2], f->pitch, f->off0, kind, idim, u);
            } else {
                MFB_LAUNCH_ROLE(ROLE_PARTICLES, masked_readout_kernel<float>, nb, EW_THREADS, (float*)dst, f->d, M, f->nloc0, f->n[0], f->n[1],
                           f->n[2], f->pitch, f->off0, kind, idim, u);
            }
            check_launch("masked_readout");
        } else {
            const unsigned nb = stream_blocks(p->np);
            if (p->real64) {
                MFB_LAUNCH_ROLE(ROLE_PARTICLES, glass_readout_kernel<double>, nb, EW_THREADS, (double*)dst, f->d, p->glass_pos, long(p->np), f->n[0],
                           f->n[1], f->n[2], f->pitch, kind == 0 ? 1 : 0, idim, unit);
            } else {
                MFB_LAUNCH_ROLE(ROLE_PARTICLES, glass_readout_kernel<float>, nb, EW_THREADS, (float*)dst, f->d, p->glass_pos, long(p->np), f->n[0],
                           f->n[1], f->n[2], f->pitch, kind == 0 ? 1 : 0, idim, unit);
            }
            check_launch("glass_readout");
        }
        return;
    }
    require(lattice_index == 0 || lattice_index == 1, "For Bravais lattice type, lattice index must be 0 or 1.");
    const int lt = p->lattice_type;
    const int nshift = 1 << lt;
    require(f->nloc0 == p->nloc0 && f->off0 == p->off0, "field/particle slab mismatch");
    const long ncell = long(p->nloc0) * p->n[1] * p->n[2];          // local cells
    const double pmean = unit / (double(p->n[0]) * p->n[1] * p->n[2] * nshift);
    if (p->copy_pending[kind == 2 ? MFB_ATTR_MASS : kind]) {  // a streamed D2H copy still reads this buffer
        const int at = kind == 2 ? MFB_ATTR_MASS : kind;
        cuda_check(cudaStreamWaitEvent(g_ctx.stream, p->copied[at], 0), "cudaStreamWaitEvent");
        p->copy_pending[at] = false;
    }
    if (f->space == MFB_KSPACE && nshift == 1 && lattice_index == 0 && g_ctx.fused_readout) {
        // single-shift load handed a k-space field: the read-out is the epilogue of the c2r pass and the real field is never
        // stored (the reference's set_* calls leave the field "invalidated" as well, include/particle_generator.hh:290)
        ZReadout ro;
        memset(&ro, 0, sizeof(ro));
        ro.dst = (kind == 0) ? p->pos : ((kind == 1) ? p->vel : p->mass);
        require(ro.dst != nullptr, "particle buffer not allocated");
        ro.real64 = p->real64;
        ro.stride = (kind == 2) ? 1 : 3;
        ro.comp = (kind == 2) ? 0 : idim;
        ro.slot0 = 0;
        ro.add_lattice = (kind == 0) ? 1 : 0;
        ro.d = idim;
        ro.shift = (kind == 0) ? LATTICE_SHIFTS[lt][0][idim] : 0.0;
        ro.lunit = unit;
        ro.mul = (kind == 2) ? pmean : 1.0;
        ro.n0 = f->n[0]; ro.n1 = f->n[1]; ro.n2 = f->n[2]; ro.i0 = f->off0;
        grid_fft_backward(f, &ro);
        f->space = MFB_RSPACE;   // contents unspecified from here on
        return;
    }
    to_rspace(f);
    for (int is = 0; is < nshift; ++is) {
        if (is == 0 && lattice_index == 1) grid_shift_field(f, SECOND_LATTICE_SHIFT[lt], true);
        if (is > 0) {
            double rel[3];
            for (int d = 0; d < 3; ++d) rel[d] = LATTICE_SHIFTS[lt][is][d] - LATTICE_SHIFTS[lt][is - 1][d];
            grid_shift_field(f, rel, true);
        }
        const unsigned nb = stream_blocks(size_t(ncell));
        const long slot0 = long(is) * ncell;
        double shift = 0.0;
        if (kind == 0) shift = LATTICE_SHIFTS[lt][is][idim] + (lattice_index == 1 ? SECOND_LATTICE_SHIFT[lt][idim] : 0.0);
        void* dst = (kind == 0) ? p->pos : ((kind == 1) ? p->vel : p->mass);
        require(dst != nullptr, "particle buffer not allocated");
        const int stride = (kind == 2) ? 1 : 3;
        const int comp = (kind == 2) ? 0 : idim;
        const int addl = (kind == 0) ? 1 : 0;
        const double mul = (kind == 2) ? pmean : 1.0;
        if (p->real64) {
            MFB_LAUNCH_ROLE(ROLE_PARTICLES, lattice_readout_kernel<double>, nb, EW_THREADS, (double*)dst, f->d, f->nloc0, f->n[0], f->n[1],
                       f->n[2], f->pitch, f->off0, slot0, stride, comp, addl, idim, shift, unit, mul);
        } else {
            MFB_LAUNCH_ROLE(ROLE_PARTICLES, lattice_readout_kernel<float>, nb, EW_THREADS, (float*)dst, f->d, f->nloc0, f->n[0], f->n[1],
                       f->n[2], f->pitch, f->off0, slot0, stride, comp, addl, idim, shift, unit, mul);
        }
        check_launch("lattice_readout");
    }
}

}  // namespace mfb

using namespace mfb;

// ==========================================================================================
// extern "C"
// ==========================================================================================
extern "C" {

const char* mfb_last_error(void) { return g_err.c_str(); }
const char* mfb_build_info(void) {
#ifdef MFB_EMU
    return "cpu-emulation (TEST ONLY)";
#else
    return "cuda sm_100a";
#endif
}

/* CUDA devices visible to this process (0 without a GPU or on error): a host with one rank per GPU picks rank % count */
int mfb_device_count(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ndev;
}
/* free / total bytes of the selected device (after mfb_init): what a host sizes its exchange block with */
int mfb_device_mem_info(size_t* free_bytes, size_t* total_bytes) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    size_t f = 0, t = 0;
    cuda_check(cudaMemGetInfo(&f, &t), "cudaMemGetInfo");
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    MFB_CATCH
}
int mfb_init(int device) {
    MFB_TRY
    if (g_ctx.initialised) return 0;
    int ndev = 0;
    cuda_check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount");
    require(ndev > 0, "libmfb200: no CUDA device visible - there is no CPU fallback");
    require(device >= 0 && device < ndev, "libmfb200: invalid device index");
    cuda_check(cudaSetDevice(device), "cudaSetDevice");
    cuda_check(cudaStreamCreate(&g_ctx.stream), "cudaStreamCreate");
#ifdef MFB_EMU
    g_ctx.num_sms = 3;   // small and odd: the persistent kernels run several rounds with a ragged tail
#else
    cuda_check(cudaDeviceGetAttribute(&g_ctx.num_sms, cudaDevAttrMultiProcessorCount, device), "cudaDeviceGetAttribute");
#endif
    g_ctx.device = device;
    g_ctx.initialised = true;
    MFB_CATCH
}
int mfb_finalize(void) {
    MFB_TRY
    if (!g_ctx.initialised) return 0;
    cudaStreamSynchronize(g_ctx.stream);
    for (auto& kv : g_ctx.tw) cudaFree(kv.second);
    g_ctx.tw.clear();
    for (auto& kv : g_ctx.tw2) cudaFree(kv.second);
    g_ctx.tw2.clear();
    if (g_ctx.partial) cudaFree(g_ctx.partial);
    g_ctx.partial = nullptr; g_ctx.partial_cap = 0;
    for (int t = 0; t < 2; ++t) {
        if (g_ctx.table[t]) cudaFree(g_ctx.table[t]);
        g_ctx.table[t] = nullptr; g_ctx.table_cap[t] = 0;
    }
    cudaStreamDestroy(g_ctx.stream);
    if (g_ctx.copy_stream) { cudaStreamDestroy(g_ctx.copy_stream); g_ctx.copy_stream = nullptr; }
    if (g_ctx.in_stream) { cudaStreamDestroy(g_ctx.in_stream); g_ctx.in_stream = nullptr; }
    g_ctx.initialised = false;
    MFB_CATCH
}
// slab mode: a device barrier that timed out (a peer died) leaves garbage behind - every call that hands results to the host
// reports it instead of relying on the caller to remember mfb_dist_check
static void dist_require_alive() {
    DistState& d = D();
    if (!d.ready || d.staged || !d.err) return;
    int e = 0;
    cuda_check(cudaMemcpyAsync(&e, d.err, sizeof(int), cudaMemcpyDeviceToHost, g_ctx.stream), "read barrier status");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    require(e == 0, "slab exchange barrier timed out: a peer rank is not responding");
}
int mfb_synchronize(void) {
    MFB_TRY
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "cudaStreamSynchronize");
    dist_require_alive();
    MFB_CATCH
}
void* mfb_stream(void) { return (void*)g_ctx.stream; }
void* mfb_copy_stream(void) {   // created on first use; NULL if that fails
    if (!g_ctx.copy_stream && g_ctx.initialised && cudaStreamCreate(&g_ctx.copy_stream) != cudaSuccess) g_ctx.copy_stream = nullptr;
    return (void*)g_ctx.copy_stream;
}
long long mfb_kernel_launches(void) { return g_ctx.launches; }
int mfb_supported_fft_size(int n) { return supported_size(n) ? 1 : 0; }
/* switches for tests: "strided_v2" = 0: every strided pass through the general kernel (strided_pass_kernel) instead of the
   compile-time pad / truncate specialisations; "sources_v2" = 0: per-product source stage; "conv_grouping", "conv_own_accumulator",
   "sources_max_slots", "sources_chunk": see mfb_internal.h */
int mfb_set_option(const char* name, int value) {
    MFB_TRY
    if (std::string(name) == "strided_v2") g_ctx.strided_v2 = value != 0;
    else if (std::string(name) == "conv_grouping") g_ctx.conv_grouping = value != 0;
    else if (std::string(name) == "conv_own_accumulator") g_ctx.conv_own_accumulator = value != 0;  // 0: as if HBM were tight
    else if (std::string(name) == "sources_v2") g_ctx.sources_v2 = value != 0;           // 0: per-product source stage (ConvGroup)
    else if (std::string(name) == "fused_readout") g_ctx.fused_readout = value != 0;     // 0: c2r and lattice read-out as two passes
    else if (std::string(name) == "exchange_blocked") g_ctx.exchange_blocked = value != 0; // 0: plain [..][position][kz] exchange chunks
    else if (std::string(name) == "y_pad_tile4") g_ctx.y_pad_tile4 = value != 0;         // 0: 8-line tiles for the padded y passes at n = 1536
    else if (std::string(name) == "zm_prefetch") g_ctx.zm_prefetch = value != 0;         // 0: direct-load multi-product z pass
    else if (std::string(name) == "sources_shared") g_ctx.sources_shared = value != 0;   // 0: multi-product z pass in every stage
    else if (std::string(name) == "sources_alias") g_ctx.sources_alias = value != 0;     // 0: shared last stage takes a slot per target
    else if (std::string(name) == "sources_max_slots") g_ctx.sources_max_slots = value;  // cap the slot workspace (tests)
    else if (std::string(name) == "sources_chunk") g_ctx.sources_chunk = value;          // cap the x planes per chunk (tests)
    else throw std::runtime_error(std::string("unknown option ") + name);
    MFB_CATCH
}

// ---------------------------------------------------------------- profiling
int mfb_profile_enable(int on) {
    MFB_TRY
    prof_collect();
    g_ctx.profiling = on != 0;
    MFB_CATCH
}
int mfb_profile_reset(void) {
    MFB_TRY
    prof_collect();
    for (int r = 0; r < ROLE_COUNT; ++r) { g_ctx.role_ms[r] = 0; g_ctx.role_bytes[r] = 0; g_ctx.role_count[r] = 0; }
    MFB_CATCH
}
int mfb_profile_num_roles(void) { return ROLE_COUNT; }
const char* mfb_profile_role_name(int role) {
    static const char* names[ROLE_COUNT] = {"conv_x_bwd", "conv_y_bwd", "conv_z_fused", "conv_y_fwd", "conv_x_fwd", "fft_x",
                                            "fft_y", "fft_z_r2c", "fft_z_c2r", "elementwise", "reduce", "particles", "barrier",
                                            "fft_z_c2c"};
    return (role >= 0 && role < ROLE_COUNT) ? names[role] : "?";
}
int mfb_profile_get(int role, double* ms, long long* count, double* bytes) {
    MFB_TRY
    require(role >= 0 && role < ROLE_COUNT, "bad role");
    prof_collect();
    *ms = g_ctx.role_ms[role]; *count = g_ctx.role_count[role]; *bytes = g_ctx.role_bytes[role];
    MFB_CATCH
}


// ---------------------------------------------------------------- slab decomposition over GPUs (one process per GPU)
size_t mfb_dist_xbuf_bytes(const int n[3], int with_conv, int nranks) {
    const size_t P = size_t(nranks > 0 ? nranks : 1);
    size_t plain = size_t(n[0]) * n[1] * size_t((n[2] / 2 + 1 + 7) & ~7) / P;   // z pitch of the FFT exchange chunks: dist_zpitch
    size_t conv = with_conv ? size_t(3 * n[0] / 2) * n[1] * (n[2] / 2) / P : 0;
    size_t e = plain > conv ? plain : conv;
    if (e < 64) e = 64;
    return ((e * sizeof(double2) + 255) / 256) * 256;
}
static void dist_layout_block(size_t xbuf_bytes) {
    DistState& d = D();
    char* b = (char*)d.block;
    d.xbuf_elems = xbuf_bytes / sizeof(double2);
    for (int k = 0; k < d.nxbuf; ++k) d.xbuf[k] = (double2*)(b + size_t(k) * xbuf_bytes);
    d.flags = (unsigned long long*)(b + size_t(d.nxbuf) * xbuf_bytes);
    d.err = (int*)(b + size_t(d.nxbuf) * xbuf_bytes + 256);
}
// nslots: exchange buffers in the peer-mapped block.  2 = the ping-pong pair of the 3-D FFTs (all the per-product
// convolver needs); more (up to 2 + MFB_CONV_MAX_SLOTS) let the v2 source stage keep x-transformed spectra resident.
int mfb_dist_init_slots(int rank, int nranks, size_t xbuf_bytes, int nslots, void* ipc_handle_out) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    require(!D().ready, "mfb_dist_init: already initialised");
    require(nranks >= 1 && nranks <= MFB_MAX_RANKS && rank >= 0 && rank < nranks, "mfb_dist_init: bad rank / world size (max 8)");
    require(xbuf_bytes % 256 == 0, "mfb_dist_init: xbuf_bytes must come from mfb_dist_xbuf_bytes");
    require(nslots >= 2 && nslots <= MFB_MAX_XBUF, "mfb_dist_init: 2..18 exchange buffers");
#ifdef MFB_EMU
    (void)ipc_handle_out;
    throw std::runtime_error("peer-mapped exchange is not available in the emulation build; use mfb_dist_init_staged");
#else
    DistState& d = D();
    d.rank = rank; d.nranks = nranks; d.staged = false; d.nxbuf = nslots;
    const size_t total = size_t(nslots) * xbuf_bytes + 4096;
    d.block = dev_alloc(total);
    cuda_check(cudaMemset((char*)d.block + size_t(nslots) * xbuf_bytes, 0, 4096), "memset exchange flags");
    dist_layout_block(xbuf_bytes);
    cudaIpcMemHandle_t h;
    cuda_check(cudaIpcGetMemHandle(&h, d.block), "cudaIpcGetMemHandle");
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &h, sizeof(h));
#endif
    MFB_CATCH
}
int mfb_dist_init(int rank, int nranks, size_t xbuf_bytes, void* ipc_handle_out) {
    return mfb_dist_init_slots(rank, nranks, xbuf_bytes, 2, ipc_handle_out);
}
int mfb_dist_connect(const void* all_handles) {
    MFB_TRY
#ifdef MFB_EMU
    (void)all_handles;
    throw std::runtime_error("peer-mapped exchange is not available in the emulation build");
#else
    DistState& d = D();
    require(d.block != nullptr && !d.ready, "mfb_dist_connect: call mfb_dist_init first");
    const size_t xb = d.xbuf_elems * sizeof(double2);
    for (int s = 0; s < d.nranks; ++s) {
        void* base = d.block;
        if (s != d.rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)all_handles + 64 * s, 64);
            cuda_check(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (peer exchange buffer)");
            d.peer_block[s] = base;
        }
        for (int k = 0; k < d.nxbuf; ++k) d.peer_xbuf[k][s] = (double2*)((char*)base + size_t(k) * xb);
        d.peer_flags[s] = (unsigned long long*)((char*)base + size_t(d.nxbuf) * xb);
    }
    d.ready = true;
#endif
    MFB_CATCH
}
int mfb_dist_init_staged_slots(int rank, int nranks, size_t xbuf_bytes, int nslots, mfb_exchange_fn cb, void* user) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    require(!D().ready, "mfb_dist_init_staged: already initialised");
    require(nranks >= 1 && nranks <= MFB_MAX_RANKS && rank >= 0 && rank < nranks, "mfb_dist_init_staged: bad rank / world size (max 8)");
    require(cb != nullptr, "mfb_dist_init_staged: exchange callback required");
    require(nslots >= 2 && nslots <= MFB_MAX_XBUF, "mfb_dist_init_staged: 2..18 exchange buffers");
    DistState& d = D();
    d.rank = rank; d.nranks = nranks; d.staged = true; d.cb = cb; d.cb_user = user; d.nxbuf = nslots;
    const size_t total = size_t(nslots) * xbuf_bytes + 4096;
    d.block = dev_alloc(total);
    cuda_check(cudaMemset((char*)d.block + size_t(nslots) * xbuf_bytes, 0, 4096), "memset exchange flags");
    dist_layout_block(xbuf_bytes);
    d.sendbuf = (double2*)dev_alloc(size_t(nslots) * xbuf_bytes);   // one staging buffer per exchange buffer
    d.ready = true;
    MFB_CATCH
}
int mfb_dist_init_staged(int rank, int nranks, size_t xbuf_bytes, mfb_exchange_fn cb, void* user) {
    return mfb_dist_init_staged_slots(rank, nranks, xbuf_bytes, 2, cb, user);
}
int mfb_dist_nslots(void) { return D().ready ? D().nxbuf : 0; }
int mfb_dist_shutdown(void) {
    MFB_TRY
    DistState& d = D();
    if (g_ctx.initialised) cudaStreamSynchronize(g_ctx.stream);
#ifndef MFB_EMU
    for (int s = 0; s < MFB_MAX_RANKS; ++s)
        if (d.peer_block[s]) cudaIpcCloseMemHandle(d.peer_block[s]);
#endif
    if (d.sendbuf) cudaFree(d.sendbuf);
    if (d.block) cudaFree(d.block);
    d = DistState();
    MFB_CATCH
}
int mfb_dist_rank(void) { return D().rank; }
int mfb_dist_nranks(void) { return D().ready ? D().nranks : 1; }
long long mfb_dist_exchanges(void) { return D().nexch; }
/* non-zero if a device barrier timed out (a peer died): results are invalid */
int mfb_dist_check(void) {
    MFB_TRY
    DistState& d = D();
    if (d.ready && !d.staged && d.err) {
        int e = 0;
        cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
        cuda_check(cudaMemcpy(&e, d.err, sizeof(int), cudaMemcpyDeviceToHost), "read barrier status");
        require(e == 0, "slab exchange barrier timed out: a peer rank is not responding");
    }
    MFB_CATCH
}
int mfb_grid_local(const mfb_grid* g, int out[4]) {
    out[0] = g->off0; out[1] = g->nloc0; out[2] = g->off1; out[3] = g->nloc1;
    return g->dist;
}

// ---------------------------------------------------------------- grids
int mfb_grid_create(mfb_grid** out, const int n[3], const double L[3], int space) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    for (int d = 0; d < 3; ++d) {
        require(supported_size(n[d]), "unsupported grid size (supported: 32..2048 powers of two and 3*2^k up to 3072)");
        require(L[d] > 0.0, "box length must be positive");
    }
    mfb_grid* g = new mfb_grid();
    for (int d = 0; d < 3; ++d) { g->n[d] = n[d]; g->L[d] = L[d]; }
    g->nzc = n[2] / 2 + 1;
    g->pitch = n[2] + 2;
    g->ntot = size_t(n[0]) * n[1] * (n[2] + 2);
    g->dist = 0; g->nloc0 = n[0]; g->off0 = 0; g->nloc1 = n[1]; g->off1 = 0;
    if (dist_on()) {
        const int P = D().nranks, r = D().rank;
        if (n[0] % P || n[1] % P) { delete g; throw std::runtime_error("slab decomposition needs grid sizes divisible by the number of GPUs"); }
        g->dist = 1;
        g->nloc0 = n[0] / P; g->off0 = r * g->nloc0;
        g->nloc1 = n[1] / P; g->off1 = r * g->nloc1;
        const size_t rows = std::max(size_t(g->nloc0) * n[1], size_t(g->nloc1) * n[0]);
        g->ntot = rows * (n[2] + 2);
    }
    g->space = space;
    try { g->d = (double*)dev_alloc(g->ntot * sizeof(double)); }
    catch (...) { delete g; throw; }   // out of HBM: no half-built object is leaked
    *out = g;
    MFB_CATCH
}
int mfb_grid_destroy(mfb_grid* g) {
    MFB_TRY
    if (g) {
        if (g->upload_pending && g_ctx.in_stream) cudaStreamSynchronize(g_ctx.in_stream);
        if (g->uploaded) cudaEventDestroy(g->uploaded);
        if (g->d) cudaFree(g->d);
        delete g;
    }
    MFB_CATCH
}
int mfb_grid_space(const mfb_grid* g) { return g->space; }
size_t mfb_grid_ntot(const mfb_grid* g) { return g->ntot; }
void* mfb_grid_device_ptr(mfb_grid* g) { return g->d; }
int mfb_grid_upload(mfb_grid* g, const double* host) {
    MFB_TRY
    cuda_check(cudaMemcpyAsync(g->d, host, g->ntot * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream), "upload");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    MFB_CATCH
}
// Streamed ingest - the host -> device counterpart of mfb_particles_stream_out.  The RNG plugin's next field (pinned host
// memory, the raw layout of mfb_grid_upload) is copied into `g` on an ingest stream of its own, ordered after everything
// queued on the compute stream so far (the last reader of the old contents), while the compute stream goes on with work
// that no longer touches `g` - in a series of runs the noise of run k+1 travels under the source stage of run k.
// mfb_grid_stream_in_wait makes the compute stream (not the host) wait for the copy.
int mfb_grid_stream_in(mfb_grid* g, const double* host) {
    MFB_TRY
    require(host != nullptr, "null host pointer");
    if (!g_ctx.in_stream) cuda_check(cudaStreamCreate(&g_ctx.in_stream), "cudaStreamCreate(ingest)");
    if (!g->uploaded) cuda_check(cudaEventCreateWithFlags(&g->uploaded, cudaEventDisableTiming), "cudaEventCreate");
    cuda_check(cudaEventRecord(g->uploaded, g_ctx.stream), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(g_ctx.in_stream, g->uploaded, 0), "cudaStreamWaitEvent");
    cuda_check(cudaMemcpyAsync(g->d, host, g->ntot * sizeof(double), cudaMemcpyHostToDevice, g_ctx.in_stream), "H2D stream_in");
    cuda_check(cudaEventRecord(g->uploaded, g_ctx.in_stream), "cudaEventRecord");
    g->upload_pending = true;
    MFB_CATCH
}
int mfb_grid_stream_in_wait(mfb_grid* g) {
    MFB_TRY
    if (g->upload_pending) {
        cuda_check(cudaStreamWaitEvent(g_ctx.stream, g->uploaded, 0), "cudaStreamWaitEvent");
        g->upload_pending = false;
    }
    MFB_CATCH
}
int mfb_grid_download(const mfb_grid* g, double* host) {
    MFB_TRY
    cuda_check(cudaMemcpyAsync(host, g->d, g->ntot * sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream), "download");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    if (g->dist) dist_require_alive();
    MFB_CATCH
}
int mfb_fft_forward(mfb_grid* g, int do_transform) {
    MFB_TRY
    if (g->space != MFB_KSPACE) {
        if (do_transform) grid_fft_forward(g);
        g->space = MFB_KSPACE;
    }
    MFB_CATCH
}
int mfb_fft_backward(mfb_grid* g, int do_transform) {
    MFB_TRY
    if (g->space != MFB_RSPACE) {
        if (do_transform) grid_fft_backward(g);
        g->space = MFB_RSPACE;
    }
    MFB_CATCH
}
int mfb_grid_zero(mfb_grid* g) {
    MFB_TRY
    cuda_check(cudaMemsetAsync(g->d, 0, g->ntot * sizeof(double), g_ctx.stream), "zero");
    MFB_CATCH
}
int mfb_grid_copy(mfb_grid* dst, const mfb_grid* src) {
    MFB_TRY
    require(dst->n[0] == src->n[0] && dst->n[1] == src->n[1] && dst->n[2] == src->n[2], "copy_from: grid size mismatch");
    dst->space = src->space;  // copy_from flips the flag without transforming (grid_fft.hh:160-166)
    cuda_check(cudaMemcpyAsync(dst->d, src->d, src->ntot * sizeof(double), cudaMemcpyDeviceToDevice, g_ctx.stream), "copy");
    MFB_CATCH
}
int mfb_grid_scale(mfb_grid* g, double a) {
    MFB_TRY
    grid_scale(g, a);
    MFB_CATCH
}
int mfb_grid_axpy(mfb_grid* y, double a, const mfb_grid* x, double b) {
    MFB_TRY
    require(y->ntot == x->ntot && y->space == x->space, "axpy: grids must have equal size and space");
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, axpy_kernel, stream_blocks(y->ntot), EW_THREADS, y->d, x->d, y->ntot, a, b);
    check_launch("axpy");
    MFB_CATCH
}
int mfb_grid_affine(mfb_grid* g, double a, double b) {
    MFB_TRY
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, affine_kernel, stream_blocks(g->ntot), EW_THREADS, g->d, g->ntot, a, b);
    check_launch("affine");
    MFB_CATCH
}
int mfb_grid_zero_dc(mfb_grid* g) {
    MFB_TRY
    grid_zero_dc(g);
    MFB_CATCH
}
int mfb_grid_inverse_laplacian(mfb_grid* g) {
    MFB_TRY
    grid_laplacian_pow(g, -1, -1.0);
    MFB_CATCH
}
int mfb_grid_laplacian(mfb_grid* g, int negative) {
    MFB_TRY
    grid_laplacian_pow(g, +1, negative ? 1.0 : -1.0);
    MFB_CATCH
}
int mfb_grid_shift_field(mfb_grid* g, const double s[3], int transform_back) {
    MFB_TRY
    grid_shift_field(g, s, transform_back != 0);
    MFB_CATCH
}
int mfb_grid_mean(mfb_grid* g, double* out) {
    MFB_TRY
    require(g->space == MFB_RSPACE, "mean(): grid must be in real space");
    double s[2];
    grid_sums(g, s);
    *out = s[0] / (double(g->n[0]) * g->n[1] * g->n[2]);
    MFB_CATCH
}
int mfb_grid_std(mfb_grid* g, double* out) {
    MFB_TRY
    double s[2];
    grid_sums(g, s);
    const double cnt = (g->space == MFB_RSPACE) ? double(g->n[0]) * g->n[1] * g->n[2] : double(g->n[0]) * g->n[1] * g->nzc;
    const double m1 = s[0] / cnt, m2 = s[1] / cnt;
    *out = std::sqrt(m2 - m1 * m1);
    MFB_CATCH
}
int mfb_grid_2norm(mfb_grid* g, double* out) {
    MFB_TRY
    require(g->space == MFB_RSPACE, "compute_2norm(): grid must be in real space");
    double s[2];
    grid_sums(g, s);
    *out = s[1] / (double(g->n[0]) * g->n[1] * g->n[2]);
    MFB_CATCH
}
int mfb_grid_max_neighbour_diff(mfb_grid* g, double* out) {
    MFB_TRY
    require(g->space == MFB_RSPACE, "max_neighbour_diff: grid must be in real space");
    const int nb = 148 * 4;
    double* part = partial_buffer(2 * nb + 2);
    // slabs: like the reference's loop (i < size(0) - 1 on the LOCAL planes, then MPI_Allreduce(MAX), src/ic_generator.cc:715-733)
    // every rank leaves out the x neighbour of its last plane
    MFB_LAUNCH_ROLE(ROLE_REDUCE, max_neighbour_diff_kernel, nb, RED_THREADS, g->d, g->dist ? g->nloc0 : g->n[0], g->n[1], g->n[2], g->pitch, part);
    check_launch("max_neighbour_diff");
    MFB_LAUNCH_ROLE(ROLE_REDUCE, reduce_final_kernel, 1, RED_THREADS, part, nb, 1, part + 2 * nb);
    check_launch("reduce_final");
    if (g->dist) dist_allreduce_max1(part + 2 * nb);
    cuda_check(cudaMemcpyAsync(out, part + 2 * nb, sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream), "D2H");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    MFB_CATCH
}
int mfb_fourier_resample(mfb_grid* src, mfb_grid* dst) {
    MFB_TRY
    require(!src->dist && !dst->dist, "FourierInterpolateCopyTo is not available for slab-decomposed grids");
    to_kspace(src);
    dst->space = MFB_KSPACE;
    if (src->n[0] == dst->n[0] && src->n[1] == dst->n[1] && src->n[2] == dst->n[2]) {
        cuda_check(cudaMemcpyAsync(dst->d, src->d, src->ntot * sizeof(double), cudaMemcpyDeviceToDevice, g_ctx.stream), "copy");
    } else {
        KGeom kd = kgeom(dst), ks = kgeom(src);
        MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, resample_kernel, kgrid_blocks(kd), EW_THREADS, reinterpret_cast<double2*>(dst->d), kd,
                   reinterpret_cast<const double2*>(src->d), ks);
        check_launch("resample");
    }
    MFB_CATCH
}

// ---------------------------------------------------------------- pipeline operators
int mfb_noise_condition(mfb_grid* w, double volfac, int fix, int invert, int remove_corners) {
    MFB_TRY
    require(w->space == MFB_KSPACE, "noise_condition: grid must be in k-space");
    KGeom k = kgeom(w);
    const double kny = k.kf0 * w->n[0] / 2;
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, noise_condition_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(w->d), k,
               (invert ? -1.0 : 1.0) / volfac, fix, remove_corners, kny);
    check_launch("noise_condition");
    MFB_CATCH
}
static const double* upload_table(const double* table, size_t ntable, size_t need, int slot) {
    require(table != nullptr && ntable >= need, "radial table too short: need 3*(N/2)^2+1 entries");
    if (g_ctx.table_cap[slot] < ntable) {
        if (g_ctx.table[slot]) cudaFree(g_ctx.table[slot]);
        g_ctx.table[slot] = (double*)dev_alloc(ntable * sizeof(double));
        g_ctx.table_cap[slot] = ntable;
    }
    cuda_check(cudaMemcpyAsync(g_ctx.table[slot], table, ntable * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream), "table upload");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    return g_ctx.table[slot];
}
int mfb_apply_radial_table(mfb_grid* out, const mfb_grid* in, const double* table, size_t ntable, int inv_k2) {
    MFB_TRY
    require(in->space == MFB_KSPACE, "apply_radial_table: input must be in k-space");
    require(in->n[0] == in->n[1] && in->n[1] == in->n[2] && in->L[0] == in->L[1] && in->L[1] == in->L[2],
            "apply_radial_table: cubic grid required");
    require(out->ntot == in->ntot, "apply_radial_table: grid size mismatch");
    const int h = in->n[0] / 2;
    const double* dt = upload_table(table, ntable, size_t(3) * h * h + 1, 0);
    out->space = MFB_KSPACE;
    KGeom k = kgeom(in);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, radial_table_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(out->d),
               reinterpret_cast<const double2*>(in->d), dt, k, inv_k2 ? 1 : 0);
    check_launch("radial_table");
    MFB_CATCH
}
int mfb_add_tides(mfb_grid* phi2, const mfb_grid* phi, const double lam[3]) {
    MFB_TRY
    require(phi2->space == MFB_KSPACE && phi->space == MFB_KSPACE, "add_tides: k-space grids required");
    KGeom k = kgeom(phi2);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, tides_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(phi2->d),
               reinterpret_cast<const double2*>(phi->d), k, lam[0], lam[1], lam[2]);
    check_launch("tides");
    MFB_CATCH
}
int mfb_gradient_create_plt(mfb_gradient** out, int ng, int ngmapto, double boxlen, const double* const tables[4]) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    require(ng >= 2 && ng % 2 == 0 && ngmapto >= 2 && boxlen > 0.0, "lattice_gradient: bad helper mesh / grid size / box length");
    require(!dist_on(), "the PLT gradient (Grid_FFT::get_cic_kspace) is only available on a single slab");
    mfb_gradient* lg = new mfb_gradient();
    lg->ng = ng; lg->ngmapto = ngmapto; lg->boxlen = boxlen;
    for (int t = 0; t < 4; ++t) lg->tab[t] = nullptr;
    const size_t bytes = size_t(ng) * ng * (ng / 2 + 1) * sizeof(double);
    try {
        for (int t = 0; t < 4; ++t) {
            require(tables[t] != nullptr, "lattice_gradient: null table");
            lg->tab[t] = (double*)dev_alloc(bytes);
            cuda_check(cudaMemcpyAsync(lg->tab[t], tables[t], bytes, cudaMemcpyHostToDevice, g_ctx.stream), "PLT table upload");
        }
        cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    } catch (...) {
        for (int t = 0; t < 4; ++t) if (lg->tab[t]) cudaFree(lg->tab[t]);
        delete lg;
        throw;
    }
    *out = lg;
    MFB_CATCH
}
int mfb_gradient_destroy(mfb_gradient* lg) {
    MFB_TRY
    if (lg) {
        cudaStreamSynchronize(g_ctx.stream);
        for (int t = 0; t < 4; ++t) if (lg->tab[t]) cudaFree(lg->tab[t]);
        delete lg;
    }
    MFB_CATCH
}
int mfb_lpt_assemble(mfb_grid* out, int d, const mfb_grid* phi, const mfb_grid* phi2, const mfb_grid* phi3,
                     mfb_grid* const A3[3], double w1, double w2, double w3, double wc, double scale, const mfb_grid* wnoise,
                     const double* vbc_table, size_t ntable, double cvbc) {
    return mfb_lpt_assemble_lg(out, d, phi, phi2, phi3, A3, w1, w2, w3, wc, scale, wnoise, vbc_table, ntable, cvbc, nullptr, 0, 0);
}
int mfb_lpt_assemble_lg(mfb_grid* out, int d, const mfb_grid* phi, const mfb_grid* phi2, const mfb_grid* phi3,
                        mfb_grid* const A3[3], double w1, double w2, double w3, double wc, double scale, const mfb_grid* wnoise,
                        const double* vbc_table, size_t ntable, double cvbc, const mfb_gradient* lg, int velocity,
                        int cic_compensation) {
    MFB_TRY
    require(d >= 0 && d < 3, "assemble: bad dimension");
    require(phi->space == MFB_KSPACE, "assemble: phi must be in k-space");
    require(!phi2 || phi2->space == MFB_KSPACE, "assemble: phi2 must be in k-space");
    require(!phi3 || phi3->space == MFB_KSPACE, "assemble: phi3 must be in k-space");
    for (const mfb_grid* g : {phi2, phi3, wnoise, (const mfb_grid*)out})
        require(!g || (g->ntot == phi->ntot && g->n[0] == phi->n[0] && g->n[1] == phi->n[1] && g->n[2] == phi->n[2]),
                "assemble: grid size mismatch");
    if (A3 && A3[0] && A3[1] && A3[2])
        for (int k = 0; k < 3; ++k) require(A3[k]->ntot == phi->ntot, "assemble: grid size mismatch");
    AssembleArgs a;
    memset(&a, 0, sizeof(a));
    a.d = d; a.p = (d + 1) % 3; a.q = (d + 2) % 3;
    a.phi = reinterpret_cast<const double2*>(phi->d);
    a.phi2 = phi2 ? reinterpret_cast<const double2*>(phi2->d) : nullptr;
    a.phi3 = phi3 ? reinterpret_cast<const double2*>(phi3->d) : nullptr;
    if (A3 && A3[0] && A3[1] && A3[2]) {
        require(A3[a.p]->space == MFB_KSPACE && A3[a.q]->space == MFB_KSPACE, "assemble: A3 must be in k-space");
        a.Ap = reinterpret_cast<const double2*>(A3[a.p]->d);
        a.Aq = reinterpret_cast<const double2*>(A3[a.q]->d);
    }
    a.w1 = w1; a.w2 = w2; a.w3 = w3; a.wc = wc; a.scale = scale;
    if (vbc_table) {
        require(wnoise && wnoise->space == MFB_KSPACE, "assemble: wnoise (k-space) needed for the vbc term");
        require(phi->n[0] == phi->n[1] && phi->n[1] == phi->n[2] && phi->L[0] == phi->L[1] && phi->L[1] == phi->L[2],
                "assemble: the vbc term needs a cubic grid (radial table on integer |n|^2)");
        const int h = phi->n[0] / 2;
        a.vbc_table = upload_table(vbc_table, ntable, size_t(3) * h * h + 1, 1);
        a.wnoise = reinterpret_cast<const double2*>(wnoise->d);
        a.cvbc = cvbc;
    }
    if (lg) {
        require(!phi->dist, "the PLT gradient is only available on a single slab");
        require(phi->n[0] == lg->ngmapto && phi->n[1] == lg->ngmapto && phi->n[2] == lg->ngmapto,
                "lattice_gradient was set up for another GridRes");
        for (int t = 0; t < 4; ++t) a.lg.tab[t] = lg->tab[t];
        a.lg.ng = lg->ng;
        a.lg.mapratio = double(lg->ng) / double(lg->ngmapto);
        a.lg.boxlen = lg->boxlen;
        a.vfac_corr = velocity ? 1 : 0;
    }
    a.cic_comp = cic_compensation ? 1 : 0;
    out->space = MFB_KSPACE;
    KGeom k = kgeom(phi);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, assemble_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(out->d), a, k);
    check_launch("assemble");
    MFB_CATCH
}
int mfb_noise_fill_ngenic(mfb_grid* g, unsigned long long seed) {
    MFB_TRY
    require(g->n[0] == g->n[1] && g->n[1] == g->n[2], "the NGenIC layout needs a cubic grid");
    g->space = MFB_KSPACE;   // Fill_Grid declares the grid k-space without transforming (:86-87)
    KGeom k = kgeom(g);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ngenic_fill_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(g->d), k, seed);
    check_launch("ngenic_fill");
    MFB_CATCH
}
/* the same layout from the reference's own generator (RANLUX, gsl_rng_ranlxd1): the seed table of RNG_ngenic's constructor
   (src/plugins/random_ngenic.cc:42-75) is drawn on the host - N^2 numbers of one stream, in the plugin's ring order - and the
   column streams run on the device */
int mfb_noise_fill_ngenic_ranlxd(mfb_grid* g, unsigned long long seed) {
    MFB_TRY
    require(g->n[0] == g->n[1] && g->n[1] == g->n[2], "the NGenIC layout needs a cubic grid");
    const size_t n = size_t(g->n[0]);
    std::vector<unsigned> table(n * n, 0u);
    Ranlxd1 r;
    r.seed((unsigned long)seed);
    auto draw = [&]() -> unsigned { return unsigned(0x7fffffff * r.uniform()); };
    for (size_t i = 0; i < n / 2; ++i) {
        for (size_t j = 0; j < i; ++j) table[i * n + j] = draw();
        for (size_t j = 0; j < i + 1; ++j) table[j * n + i] = draw();
        for (size_t j = 0; j < i; ++j) table[(n - 1 - i) * n + j] = draw();
        for (size_t j = 0; j < i + 1; ++j) table[(n - 1 - j) * n + i] = draw();
        for (size_t j = 0; j < i; ++j) table[i * n + (n - 1 - j)] = draw();
        for (size_t j = 0; j < i + 1; ++j) table[j * n + (n - 1 - i)] = draw();
        for (size_t j = 0; j < i; ++j) table[(n - 1 - i) * n + (n - 1 - j)] = draw();
        for (size_t j = 0; j < i + 1; ++j) table[(n - 1 - j) * n + (n - 1 - i)] = draw();
    }
    unsigned* dtab = (unsigned*)dev_alloc(table.size() * sizeof(unsigned));
    cuda_check(cudaMemcpyAsync(dtab, table.data(), table.size() * sizeof(unsigned), cudaMemcpyHostToDevice, g_ctx.stream), "seed table upload");
    g->space = MFB_KSPACE;
    KGeom k = kgeom(g);
    const long cols = long(k.s0) * k.s1;
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ngenic_ranlux_fill_kernel, unsigned((cols + 127) / 128), 128, reinterpret_cast<double2*>(g->d), k, dtab);
    check_launch("ngenic_ranlux_fill");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    cudaFree(dtab);
    MFB_CATCH
}
int mfb_density_from_phi(mfb_grid* out, const mfb_grid* phi) {
    MFB_TRY
    require(phi->space == MFB_KSPACE, "density_from_phi: phi must be in k-space");
    out->space = MFB_KSPACE;
    KGeom k = kgeom(phi);
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, laplacian_pow_kernel, kgrid_blocks(k), EW_THREADS, reinterpret_cast<double2*>(out->d),
               reinterpret_cast<const double2*>(phi->d), k, 1, 1.0);
    check_launch("density");
    MFB_CATCH
}

// ---------------------------------------------------------------- complex grids + PPT / Eulerian branch
int mfb_cgrid_create(mfb_cgrid** out, const int n[3], const double L[3]) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    for (int d = 0; d < 3; ++d) require(supported_size(n[d]), "unsupported grid size (no FFT plan)");
    mfb_cgrid* g = new mfb_cgrid();
    for (int d = 0; d < 3; ++d) { g->n[d] = n[d]; g->L[d] = L[d]; }
    g->space = MFB_RSPACE;
    g->nelem = size_t(n[0]) * n[1] * n[2];
    g->dist = 0; g->nloc0 = n[0]; g->off0 = 0; g->nloc1 = n[1]; g->off1 = 0;
    if (dist_on()) {
        const int P = D().nranks, r = D().rank;
        if (n[0] % P || n[1] % P || n[2] % 2) { delete g; throw std::runtime_error("slab decomposition needs grid sizes divisible by the number of GPUs"); }
        g->dist = 1;
        g->nloc0 = n[0] / P; g->off0 = r * g->nloc0;
        g->nloc1 = n[1] / P; g->off1 = r * g->nloc1;
        g->nelem = std::max(size_t(g->nloc0) * n[1], size_t(g->nloc1) * n[0]) * n[2];
    }
    try { g->d = (double2*)dev_alloc(g->nelem * sizeof(double2)); }
    catch (...) { delete g; throw; }
    cuda_check(cudaMemsetAsync(g->d, 0, g->nelem * sizeof(double2), g_ctx.stream), "memset");
    *out = g;
    MFB_CATCH
}
int mfb_cgrid_destroy(mfb_cgrid* g) {
    MFB_TRY
    if (g) { cudaFree(g->d); delete g; }
    MFB_CATCH
}
int mfb_cgrid_space(const mfb_cgrid* g) { return g->space; }
int mfb_cgrid_local(const mfb_cgrid* g, int out[4]) {
    out[0] = g->off0; out[1] = g->nloc0; out[2] = g->off1; out[3] = g->nloc1;
    return g->dist;
}
size_t mfb_cgrid_nelem(const mfb_cgrid* g) { return g->nelem; }
int mfb_cgrid_upload(mfb_cgrid* g, const double* host_interleaved, int space) {
    MFB_TRY
    cuda_check(cudaMemcpyAsync(g->d, host_interleaved, g->nelem * sizeof(double2), cudaMemcpyHostToDevice, g_ctx.stream), "H2D");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    g->space = space;
    MFB_CATCH
}
int mfb_cgrid_download(const mfb_cgrid* g, double* host_interleaved) {
    MFB_TRY
    cuda_check(cudaMemcpyAsync(host_interleaved, g->d, g->nelem * sizeof(double2), cudaMemcpyDeviceToHost, g_ctx.stream), "D2H");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    MFB_CATCH
}
int mfb_cgrid_copy(mfb_cgrid* dst, const mfb_cgrid* src) {
    MFB_TRY
    require(dst->nelem == src->nelem && dst->dist == src->dist, "copy_from: size mismatch");
    cuda_check(cudaMemcpyAsync(dst->d, src->d, src->nelem * sizeof(double2), cudaMemcpyDeviceToDevice, g_ctx.stream), "copy");
    dst->space = src->space;
    MFB_CATCH
}
int mfb_cgrid_fft_forward(mfb_cgrid* g, int do_transform) {
    MFB_TRY
    if (g->space != MFB_KSPACE) { if (do_transform) cgrid_fft(g, 0); g->space = MFB_KSPACE; }
    MFB_CATCH
}
int mfb_cgrid_fft_backward(mfb_cgrid* g, int do_transform) {
    MFB_TRY
    if (g->space != MFB_RSPACE) { if (do_transform) cgrid_fft(g, 1); g->space = MFB_RSPACE; }
    MFB_CATCH
}
int mfb_cgrid_gradient(mfb_cgrid* g, int idim) {
    MFB_TRY
    require(idim >= 0 && idim < 3, "bad dimension");
    require(g->space == MFB_KSPACE, "gradient: complex grid must be in k-space");
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, cgrid_gradient_kernel, stream_blocks(cgrid_local_elems(g)), EW_THREADS, g->d, cgeom(g), idim);
    check_launch("cgrid_gradient");
    MFB_CATCH
}
int mfb_ppt_sqrt_density(mfb_grid* rho, double C_species) {
    MFB_TRY
    require(rho->space == MFB_RSPACE, "PPT: rho must be a real-space grid");
    CGeom c;
    c.n0 = rho->n[0]; c.n1 = rho->n[1]; c.n2 = rho->n[2]; c.kf0 = c.kf1 = c.kf2 = 0.0; c.rpitch = rho->pitch;
    c.s0 = rho->dist ? rho->nloc0 : rho->n[0]; c.s1 = rho->n[1]; c.tr = 0; c.off = 0;
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ppt_sqrt1p_kernel, stream_blocks(size_t(c.s0) * c.n1 * c.n2), EW_THREADS, rho->d, c, C_species);
    check_launch("ppt_sqrt1p");
    MFB_CATCH
}
int mfb_ppt_init_psi(mfb_cgrid* psi, const mfb_grid* rho, const mfb_grid* phi, const mfb_grid* phi2, double hbar, double Dplus0) {
    MFB_TRY
    check_ppt_real(psi, rho, "rho"); check_ppt_real(psi, phi, "phi");
    if (phi2) check_ppt_real(psi, phi2, "phi2");
    psi->space = MFB_RSPACE;
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ppt_init_psi_kernel, stream_blocks(cgrid_local_elems(psi)), EW_THREADS, psi->d, rho->d, phi->d,
                    phi2 ? phi2->d : (const double*)nullptr, cgeom(psi), 1.0 / hbar, Dplus0);
    check_launch("ppt_init_psi");
    MFB_CATCH
}
int mfb_ppt_drift(mfb_cgrid* psi, double hbar, double Dplus0) {
    MFB_TRY
    require(psi->space == MFB_KSPACE, "PPT drift: psi must be in k-space");
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ppt_drift_kernel, stream_blocks(cgrid_local_elems(psi)), EW_THREADS, psi->d, cgeom(psi), hbar, Dplus0);
    check_launch("ppt_drift");
    MFB_CATCH
}
int mfb_ppt_kick(mfb_cgrid* psi, const mfb_grid* phi2, double hbar, double Dplus0) {
    MFB_TRY
    require(psi->space == MFB_RSPACE, "PPT kick: psi must be in real space");
    check_ppt_real(psi, phi2, "phi2");
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ppt_kick_kernel, stream_blocks(cgrid_local_elems(psi)), EW_THREADS, psi->d, phi2->d, cgeom(psi), 1.0 / hbar, Dplus0);
    check_launch("ppt_kick");
    MFB_CATCH
}
int mfb_ppt_density(mfb_grid* rho, const mfb_cgrid* psi) {
    MFB_TRY
    require(psi->space == MFB_RSPACE, "PPT density: psi must be in real space");
    rho->space = MFB_RSPACE;
    check_ppt_real(psi, rho, "rho");
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ppt_density_kernel, stream_blocks(cgrid_local_elems(psi)), EW_THREADS, rho->d, psi->d, cgeom(psi));
    check_launch("ppt_density");
    MFB_CATCH
}
int mfb_ppt_velocity(mfb_grid* out, const mfb_cgrid* psi, const mfb_cgrid* grad_psi, const mfb_grid* rho, double hbar, double vunit) {
    MFB_TRY
    require(psi->space == MFB_RSPACE && grad_psi->space == MFB_RSPACE, "PPT velocity: psi and grad_psi must be in real space");
    require(psi->nelem == grad_psi->nelem, "PPT velocity: size mismatch");
    out->space = MFB_RSPACE;
    check_ppt_real(psi, out, "out"); check_ppt_real(psi, rho, "rho");
    MFB_LAUNCH_ROLE(ROLE_ELEMENTWISE, ppt_velocity_kernel, stream_blocks(cgrid_local_elems(psi)), EW_THREADS, out->d, psi->d, grad_psi->d, rho->d,
                    cgeom(psi), hbar, vunit);
    check_launch("ppt_velocity");
    MFB_CATCH
}

// ---------------------------------------------------------------- convolver
int mfb_conv_create(mfb_conv** out, const int n[3], const double L[3]) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    mfb_conv* c = new mfb_conv();
    for (int d = 0; d < 3; ++d) {
        require(n[d] % 2 == 0, "Orszag convolver for LPTorder>1 requires even grid size!");  // src/ic_generator.cc:118-121
        c->n[d] = n[d]; c->L[d] = L[d]; c->np[d] = 3 * n[d] / 2;
        require(supported_size(n[d]) && supported_size(c->np[d]), "unsupported grid size for the Orszag convolver");
    }
    const size_t nzin = n[2] / 2;
    c->dist = 0; c->p0loc = c->np[0]; c->p0off = 0;
    if (dist_on()) {
        const int P = D().nranks;
        require(n[0] % P == 0 && n[1] % P == 0 && c->np[0] % P == 0, "slab decomposition needs N and 3N/2 divisible by the number of GPUs");
        c->dist = 1; c->p0loc = c->np[0] / P; c->p0off = D().rank * c->p0loc;
    }
    c->w1_elems = c->dist ? 0 : size_t(c->np[0]) * n[1] * nzin;   // slab mode: the exchange buffers play W1's role
    c->w2_elems = size_t(c->p0loc) * c->np[1] * nzin;
    c->naive = 0; c->nb1 = c->nb2 = nullptr;
    c->Wacc = nullptr; c->acc_borrowed = false; c->lender = nullptr;
    c->mode = 0; c->nxs = 0; c->chunk = nullptr; c->chunk_elems = 0; c->chunk_owned = 0;
    c->W1 = c->W2a = c->W2b = nullptr; c->fbuf = nullptr; c->legacy_ready = 0;
    // no workspace yet: it is carved at the first use, for the path that use takes - the slot workspace of the v2 source
    // stage (mfb_sources2.cuh) or the per-product buffers W1, W2a, W2b, fbuf[, Wacc] (conv_legacy_ensure)
    *out = c;
    MFB_CATCH
}
int mfb_conv_create_naive(mfb_conv** out, const int n[3], const double L[3]) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    // slabs: the products run through the exchange buffers of the 3-D FFTs (naive_backward_operand / naive_forward_output)
    mfb_conv* c = new mfb_conv();
    for (int d = 0; d < 3; ++d) {
        c->n[d] = n[d]; c->L[d] = L[d]; c->np[d] = n[d];
        require(supported_size(n[d]), "unsupported grid size for the convolver");
    }
    c->naive = 1;
    c->dist = dist_on() ? 1 : 0;
    if (mfb_grid_create(&c->nb1, n, L, MFB_KSPACE) || mfb_grid_create(&c->nb2, n, L, MFB_KSPACE)) throw std::runtime_error(g_err);
    *out = c;
    MFB_CATCH
}
int mfb_conv_destroy(mfb_conv* c) {
    MFB_TRY
    if (c) {
        if (c->lender) c->lender->borrower = nullptr;
        if (g_ctx.initialised) cudaStreamSynchronize(g_ctx.stream);
        conv_free_slots(c);
        if (c->W1) cudaFree(c->W1);
        if (c->W2a) cudaFree(c->W2a);
        if (c->W2b) cudaFree(c->W2b);
        if (c->Wacc && !c->acc_borrowed) cudaFree(c->Wacc);
        if (c->fbuf) mfb_grid_destroy(c->fbuf);
        if (c->nb1) mfb_grid_destroy(c->nb1);
        if (c->nb2) mfb_grid_destroy(c->nb2);
        delete c;
    }
    MFB_CATCH
}
size_t mfb_conv_workspace_bytes(const mfb_conv* c) {
    if (c->naive) return (c->nb1->ntot + c->nb2->ntot) * sizeof(double);
    if (c->mode == 1) {   // slot workspace of the v2 source stage: own slots (incl. W1) + own chunk workspace
        size_t b = 0;
        for (int i = 0; i < c->nxs; ++i) if (c->xs[i].owned || c->xs[i].p == c->W1) b += conv_slot_elems(c) * sizeof(double2);
        return b + (c->chunk_owned ? c->chunk_elems * sizeof(double2) : 0);
    }
    return ((c->W1 ? c->w1_elems : 0) + ((c->W2a ? 1 : 0) + (c->W2b ? 1 : 0) + ((c->Wacc && !c->acc_borrowed) ? 1 : 0)) * c->w2_elems) * sizeof(double2) +
           (c->fbuf ? c->fbuf->ntot * sizeof(double) : 0);
}
int mfb_conv_grouped(const mfb_conv* cc) {
    mfb_conv* c = const_cast<mfb_conv*>(cc);
    if (!c->naive && c->mode == 0 && !c->legacy_ready) {   // answer for the per-product workspace: carve it
        try { conv_legacy_ensure(c); } catch (const std::exception& e) { set_error(e.what()); return 0; }
    }
    return can_group(c) ? 1 : 0;
}
/* Lend the (idle) position/velocity arrays of a particle container to the convolver: the per-product path uses them as the
   accumulator of product groups when it has none of its own, the v2 source stage as slots / chunk workspace.  p == NULL
   ends the loan. */
int mfb_conv_borrow_accumulator(mfb_conv* c, mfb_particles* p) {
    MFB_TRY
    require(!c->naive, "the NaiveConvolver has no accumulator");
    if (p == c->lender) return 0;
    // whatever the workspace looks like now, it may sit in the old lender's memory: drop the loan first
    if (c->lender) {
        cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
        if (c->mode == 1) { conv_free_slots(c); c->mode = 0; }
        if (c->acc_borrowed) { c->Wacc = nullptr; c->acc_borrowed = false; }
        c->lender->borrower = nullptr;
        c->lender = nullptr;
    }
    if (p == nullptr) return 0;
    require(p->borrower == nullptr, "these particle arrays are already lent to another convolver");
    if (c->mode == 1) { cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync"); conv_free_slots(c); c->mode = 0; }   // re-carved with the loan at the next use
    c->lender = p;
    p->borrower = c;
    if (c->mode == 0 && c->Wacc == nullptr && c->W2a != nullptr && p->posvel_bytes >= c->w2_elems * sizeof(double2)) {
        c->Wacc = (double2*)p->pos;
        c->acc_borrowed = true;
    }
    MFB_CATCH
}
/* which path the last mfb_lpt_sources call took: 1 = per-product (ConvGroup), 2 = shared x passes + multi-product z pass;
   info[0..5] = slots, x-plane loops, x passes, backward y passes, z launches, x planes per chunk (0 if not applicable) */
int mfb_conv_sources_info(const mfb_conv* c, long long info[6]) {
    info[0] = c->last_slots; info[1] = c->last_loops; info[2] = c->last_xpass; info[3] = c->last_ypass; info[4] = c->last_zlaunch;
    info[5] = 0;
    return c->last_path;
}
static void check_conv_grids(mfb_conv* c, mfb_grid* a, mfb_grid* b, mfb_grid* o) {
    for (mfb_grid* g : {a, b, o})
        require(g && g->n[0] == c->n[0] && g->n[1] == c->n[1] && g->n[2] == c->n[2], "convolver: grid size mismatch");
}
static void check_dims(const int* d, int cnt) {
    for (int i = 0; i < cnt; ++i) require(d[i] >= 0 && d[i] < 3, "convolver: axis index must be 0,1,2");
}
int mfb_conv_hessians(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* r, const int d2r[2], mfb_grid* out, int op,
                      double alpha) {
    MFB_TRY
    check_conv_grids(c, l, r, out); check_dims(d2l, 2); check_dims(d2r, 2);
    conv2(c, hessian_operand(l, d2l), hessian_operand(r, d2r), out, op, alpha);
    MFB_CATCH
}
int mfb_conv_hessians3(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* m, const int d2m[2], mfb_grid* r,
                       const int d2r[2], mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, l, m, out); check_conv_grids(c, r, r, out);
    check_dims(d2l, 2); check_dims(d2m, 2); check_dims(d2r, 2);
    conv3(c, hessian_operand(l, d2l), hessian_operand(m, d2m), hessian_operand(r, d2r), out, op, alpha);
    MFB_CATCH
}
int mfb_conv_sum_hessians(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* r, const int d2r1[2], const int d2r2[2],
                          mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, l, r, out); check_dims(d2l, 2); check_dims(d2r1, 2); check_dims(d2r2, 2);
    conv2(c, hessian_operand(l, d2l), Operand{r, d2r1[0], d2r1[1], d2r2[0], d2r2[1], -1.0, -1.0, 0}, out, op, alpha);
    MFB_CATCH
}
int mfb_conv_diff_hessians(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* r, const int d2r1[2], const int d2r2[2],
                           mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, l, r, out); check_dims(d2l, 2); check_dims(d2r1, 2); check_dims(d2r2, 2);
    conv2(c, hessian_operand(l, d2l), Operand{r, d2r1[0], d2r1[1], d2r2[0], d2r2[1], -1.0, +1.0, 0}, out, op, alpha);
    MFB_CATCH
}
int mfb_conv_gradients(mfb_conv* c, mfb_grid* l, int d1l, mfb_grid* r, int d1r, mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, l, r, out); check_dims(&d1l, 1); check_dims(&d1r, 1);
    conv2(c, Operand{l, d1l, -1, -1, -1, 1.0, 0.0, 1}, Operand{r, d1r, -1, -1, -1, 1.0, 0.0, 1}, out, op, alpha);
    MFB_CATCH
}
int mfb_conv_gradient_hessian(mfb_conv* c, mfb_grid* l, int d1l, mfb_grid* r, const int d2r[2], mfb_grid* out, int op,
                              double alpha) {
    MFB_TRY
    check_conv_grids(c, l, r, out); check_dims(&d1l, 1); check_dims(d2r, 2);
    // (-i k_a) f  and  -k_a k_b g   (include/convolution.hh:116-124); Nyquist never enters the padded grid
    // (the un-padded NaiveConvolver does see it: these two factors come from get_k, Nyquist = +N/2)
    conv2(c, Operand{l, d1l, -1, -1, -1, -1.0, 0.0, 1, 1}, Operand{r, d2r[0], d2r[1], -1, -1, -1.0, 0.0, 0, 1}, out, op, alpha);
    MFB_CATCH
}
int mfb_conv_multiply(mfb_conv* c, mfb_grid* l, mfb_grid* r, mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, l, r, out);
    conv2(c, Operand{l, -1, -1, -1, -1, 1.0, 0.0, 0}, Operand{r, -1, -1, -1, -1, 1.0, 0.0, 0}, out, op, alpha);
    MFB_CATCH
}
/* convolve2 / convolve3 with caller-formed spectra (include/convolution.hh:512-549): the k-space operands are device grids the
   caller has filled (the host mirror evaluates the reference's kfunc lambdas into them); no derivative factor is applied */
int mfb_conv_generic(mfb_conv* c, mfb_grid* specA, mfb_grid* specB, mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, specA, specB, out);
    require(specA->space == MFB_KSPACE && specB->space == MFB_KSPACE, "mfb_conv_generic: the operand spectra must be in k-space");
    conv2(c, Operand{specA, -1, -1, -1, -1, 1.0, 0.0, 0}, Operand{specB, -1, -1, -1, -1, 1.0, 0.0, 0}, out, op, alpha);
    MFB_CATCH
}
int mfb_conv_generic3(mfb_conv* c, mfb_grid* specA, mfb_grid* specB, mfb_grid* specC, mfb_grid* out, int op, double alpha) {
    MFB_TRY
    check_conv_grids(c, specA, specB, out); check_conv_grids(c, specC, specC, out);
    require(specA->space == MFB_KSPACE && specB->space == MFB_KSPACE && specC->space == MFB_KSPACE,
            "mfb_conv_generic3: the operand spectra must be in k-space");
    conv3(c, Operand{specA, -1, -1, -1, -1, 1.0, 0.0, 0}, Operand{specB, -1, -1, -1, -1, 1.0, 0.0, 0},
          Operand{specC, -1, -1, -1, -1, 1.0, 0.0, 0}, out, op, alpha);
    MFB_CATCH
}
int mfb_lpt_sources(mfb_conv* c, int order, double Dplus0, mfb_grid* phi, mfb_grid* phi2, mfb_grid* phi3, mfb_grid* const A3[3],
                    const double* tides_lambda) {
    MFB_TRY
    lpt_sources(c, order, Dplus0, phi, phi2, phi3, A3, tides_lambda);
    MFB_CATCH
}

// ---------------------------------------------------------------- particles
int mfb_particles_create(mfb_particles** out, const int n[3], int lattice_type, int real64, int ids64, int with_masses) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    require(lattice_type >= 0 && lattice_type <= 3, "only Bravais lattices (sc,bcc,fcc,rsc) are supported on the device path");
    mfb_particles* p = new mfb_particles();
    for (int d = 0; d < 3; ++d) p->n[d] = n[d];
    p->lattice_type = lattice_type; p->real64 = real64; p->ids64 = ids64;
    const size_t ov = size_t(1) << lattice_type;
    p->nloc0 = n[0]; p->off0 = 0;
    if (dist_on()) {
        require(n[0] % D().nranks == 0, "slab decomposition needs grid sizes divisible by the number of GPUs");
        p->nloc0 = n[0] / D().nranks; p->off0 = D().rank * p->nloc0;
    }
    p->np = ov * size_t(p->nloc0) * n[1] * n[2];   // get_local_num_particles (particle_container.hh:150-152)
    const size_t rs = real64 ? 8 : 4, is = ids64 ? 8 : 4;
    p->posvel_bytes = 6 * p->np * rs;
    p->pos = p->ids = p->mass = nullptr;
    try {
        p->pos = dev_alloc(p->posvel_bytes);             // one block: it can be lent to the convolver as its accumulator
        p->vel = (char*)p->pos + 3 * p->np * rs;
        p->ids = dev_alloc(p->np * is);
        p->mass = with_masses ? dev_alloc(p->np * rs) : nullptr;
    } catch (...) {
        if (p->pos) cudaFree(p->pos);
        if (p->ids) cudaFree(p->ids);
        delete p;
        throw;
    }
    *out = p;
    MFB_CATCH
}
static void particles_alloc(mfb_particles* p, int with_masses) {
    const size_t rs = p->real64 ? 8 : 4, is = p->ids64 ? 8 : 4;
    p->posvel_bytes = 6 * p->np * rs;
    p->pos = p->ids = p->mass = nullptr;
    try {
        p->pos = dev_alloc(std::max<size_t>(p->posvel_bytes, 16));
        p->vel = (char*)p->pos + 3 * p->np * rs;
        p->ids = dev_alloc(std::max<size_t>(p->np * is, 16));
        p->mass = with_masses ? dev_alloc(std::max<size_t>(p->np * rs, 16)) : nullptr;
    } catch (...) {
        if (p->pos) cudaFree(p->pos);
        if (p->ids) cudaFree(p->ids);
        throw;
    }
}
int mfb_particles_create_masked(mfb_particles** out, const int n[3], int lattice_index, const int mask[8], int real64, int ids64,
                                int with_masses) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    require((size_t(n[0]) * n[1] * n[2]) % 8 == 0 && n[0] % 2 == 0 && n[1] % 2 == 0 && n[2] % 2 == 0,
            "For masked lattice type, linear field resolution must be a multiple of two.");
    require(lattice_index == 0 || lattice_index == 1, "For masked lattice type, lattice index must be 0 or 1.");
    mfb_particles* p = new mfb_particles();
    for (int d = 0; d < 3; ++d) p->n[d] = n[d];
    p->lattice_type = MFB_LATTICE_MASKED; p->real64 = real64; p->ids64 = ids64;
    p->mask_index = lattice_index;
    for (int s = 0; s < 8; ++s) p->mask[s] = mask[s];
    p->nloc0 = n[0]; p->off0 = 0;
    if (dist_on()) {
        if (n[0] % (2 * D().nranks)) { delete p; throw std::runtime_error("masked lattice on slabs needs an even number of planes per GPU"); }
        p->nloc0 = n[0] / D().nranks; p->off0 = D().rank * p->nloc0;
    }
    const mfb::LatticeMask M = make_mask(p);
    p->np = size_t(p->nloc0 / 2) * size_t(M.plane[0] + M.plane[1]);
    p->np_global = size_t(n[0] / 2) * size_t(M.plane[0] + M.plane[1]);
    if (p->np == 0) { delete p; throw std::runtime_error("masked lattice: no cell carries this species"); }
    try { particles_alloc(p, with_masses); } catch (...) { delete p; throw; }
    *out = p;
    MFB_CATCH
}
int mfb_particles_create_glass(mfb_particles** out, const int n[3], const double* posr, size_t np, unsigned long long glass_offset,
                               int real64, int ids64) {
    MFB_TRY
    require(g_ctx.initialised, "mfb_init has not been called");
    require(!dist_on(), "glass loads are only available on a single slab");
    require(posr != nullptr && np > 0, "glass load: no particle positions");
    for (size_t i = 0; i < 3 * np; ++i)
        require(posr[i] >= 0.0 && posr[i] < double(n[i % 3]), "glass load: position outside the grid (grid units expected)");
    mfb_particles* p = new mfb_particles();
    for (int d = 0; d < 3; ++d) p->n[d] = n[d];
    p->lattice_type = MFB_LATTICE_GLASS; p->real64 = real64; p->ids64 = ids64;
    p->nloc0 = n[0]; p->off0 = 0;
    p->np = p->np_global = np;
    p->glass_off = glass_offset;
    try {
        particles_alloc(p, 0);
        try {
            p->glass_pos = (double*)dev_alloc(3 * np * sizeof(double));
            cuda_check(cudaMemcpyAsync(p->glass_pos, posr, 3 * np * sizeof(double), cudaMemcpyHostToDevice, g_ctx.stream), "glass positions");
            cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
        } catch (...) {
            cudaFree(p->pos); cudaFree(p->ids);
            if (p->glass_pos) cudaFree(p->glass_pos);
            throw;
        }
    } catch (...) { delete p; throw; }
    *out = p;
    MFB_CATCH
}
int mfb_particles_destroy(mfb_particles* p) {
    MFB_TRY
    if (p) {
        if (p->borrower) {   // the loan ends with the lender: the convolver falls back to what it owns
            cudaStreamSynchronize(g_ctx.stream);
            mfb_conv* b = p->borrower;
            if (b->mode == 1) { conv_free_slots(b); b->mode = 0; }
            if (b->acc_borrowed) { b->Wacc = nullptr; b->acc_borrowed = false; }
            b->lender = nullptr;
        }
        if (g_ctx.copy_stream) cudaStreamSynchronize(g_ctx.copy_stream);
        for (int a = 0; a < 4; ++a) if (p->copied[a]) cudaEventDestroy(p->copied[a]);
        cudaFree(p->pos); cudaFree(p->ids);
        if (p->mass) cudaFree(p->mass);
        if (p->glass_pos) cudaFree(p->glass_pos);
        delete p;
    }
    MFB_CATCH
}
size_t mfb_particles_count(const mfb_particles* p) { return p->np; }
int mfb_particles_set_ids(mfb_particles* p, unsigned long long id_offset_species) {
    MFB_TRY
    if (p->lattice_type == MFB_LATTICE_MASKED) {
        const mfb::LatticeMask M = make_mask(p);
        const unsigned long long off = id_offset_species * (unsigned long long)p->n[0] * (unsigned long long)p->n[1] * (unsigned long long)p->n[2];
        const unsigned nb = stream_blocks(size_t(p->nloc0) * p->n[1] * p->n[2]);
        if (p->ids64) {
            MFB_LAUNCH_ROLE(ROLE_PARTICLES, masked_ids_kernel<unsigned long long>, nb, EW_THREADS, (unsigned long long*)p->ids, M, p->nloc0, p->n[1], p->n[2], p->off0, off);
        } else {
            MFB_LAUNCH_ROLE(ROLE_PARTICLES, masked_ids_kernel<unsigned int>, nb, EW_THREADS, (unsigned int*)p->ids, M, p->nloc0, p->n[1], p->n[2], p->off0, off);
        }
        check_launch("masked_ids");
        return 0;
    }
    if (p->lattice_type == MFB_LATTICE_GLASS) {
        const unsigned nb = stream_blocks(p->np);
        const unsigned long long off = id_offset_species + p->glass_off;   // :277-286: IDoffset + i + glass_ptr_->offset()
        if (p->ids64) {
            MFB_LAUNCH_ROLE(ROLE_PARTICLES, glass_ids_kernel<unsigned long long>, nb, EW_THREADS, (unsigned long long*)p->ids, long(p->np), off);
        } else {
            MFB_LAUNCH_ROLE(ROLE_PARTICLES, glass_ids_kernel<unsigned int>, nb, EW_THREADS, (unsigned int*)p->ids, long(p->np), off);
        }
        check_launch("glass_ids");
        return 0;
    }
    const long ncell = long(p->n[0]) * p->n[1] * p->n[2];            // global
    const long nloc = long(p->nloc0) * p->n[1] * p->n[2];            // this slab
    const long cell0 = long(p->off0) * p->n[1] * p->n[2];
    const int ov = 1 << p->lattice_type;
    const unsigned long long off = id_offset_species * (unsigned long long)ov * (unsigned long long)ncell;
    const unsigned nb = stream_blocks(p->np);
    if (p->ids64) {
        MFB_LAUNCH_ROLE(ROLE_PARTICLES, particle_ids_kernel<unsigned long long>, nb, EW_THREADS, (unsigned long long*)p->ids, nloc, ov, off, cell0);
    } else {
        MFB_LAUNCH_ROLE(ROLE_PARTICLES, particle_ids_kernel<unsigned int>, nb, EW_THREADS, (unsigned int*)p->ids, nloc, ov, off, cell0);
    }
    check_launch("particle_ids");
    MFB_CATCH
}
int mfb_particles_set_positions(mfb_particles* p, int lattice_index, int idim, double lunit, mfb_grid* field) {
    MFB_TRY
    require(idim >= 0 && idim < 3, "bad dimension");
    particles_readout(p, 0, lattice_index, idim, lunit, field);
    MFB_CATCH
}
int mfb_particles_set_velocities(mfb_particles* p, int lattice_index, int idim, mfb_grid* field) {
    MFB_TRY
    require(idim >= 0 && idim < 3, "bad dimension");
    particles_readout(p, 1, lattice_index, idim, 1.0, field);
    MFB_CATCH
}
int mfb_particles_set_masses(mfb_particles* p, int lattice_index, double munit, mfb_grid* field) {
    MFB_TRY
    require(p->mass != nullptr, "particles were created without masses");
    particles_readout(p, 2, lattice_index, 0, munit, field);
    MFB_CATCH
}
int mfb_particles_download(const mfb_particles* p, void* pos, void* vel, void* ids, void* mass) {
    MFB_TRY
    const size_t rs = p->real64 ? 8 : 4, is = p->ids64 ? 8 : 4;
    if (pos) cuda_check(cudaMemcpyAsync(pos, p->pos, 3 * p->np * rs, cudaMemcpyDeviceToHost, g_ctx.stream), "D2H pos");
    if (vel) cuda_check(cudaMemcpyAsync(vel, p->vel, 3 * p->np * rs, cudaMemcpyDeviceToHost, g_ctx.stream), "D2H vel");
    if (ids) cuda_check(cudaMemcpyAsync(ids, p->ids, p->np * is, cudaMemcpyDeviceToHost, g_ctx.stream), "D2H ids");
    if (mass && p->mass) cuda_check(cudaMemcpyAsync(mass, p->mass, p->np * rs, cudaMemcpyDeviceToHost, g_ctx.stream), "D2H mass");
    cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync");
    dist_require_alive();
    MFB_CATCH
}
int mfb_particles_stream_out(mfb_particles* p, int attr, void* host) {
    MFB_TRY
    require(attr >= 0 && attr < 4 && host != nullptr, "bad particle attribute / host pointer");
    const size_t rs = p->real64 ? 8 : 4, is = p->ids64 ? 8 : 4;
    const void* src = attr == MFB_ATTR_POS ? p->pos : attr == MFB_ATTR_VEL ? p->vel : attr == MFB_ATTR_IDS ? p->ids : p->mass;
    require(src != nullptr, "particle buffer not allocated");
    const size_t bytes = attr == MFB_ATTR_IDS ? p->np * is : (attr == MFB_ATTR_MASS ? p->np * rs : 3 * p->np * rs);
    if (!g_ctx.copy_stream) cuda_check(cudaStreamCreate(&g_ctx.copy_stream), "cudaStreamCreate(copy)");
    if (!p->copied[attr]) cuda_check(cudaEventCreateWithFlags(&p->copied[attr], cudaEventDisableTiming), "cudaEventCreate");
    // order the copy after the producer kernels: reuse the attribute's event as the compute-side marker first
    cuda_check(cudaEventRecord(p->copied[attr], g_ctx.stream), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(g_ctx.copy_stream, p->copied[attr], 0), "cudaStreamWaitEvent");
    cuda_check(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, g_ctx.copy_stream), "D2H stream_out");
    cuda_check(cudaEventRecord(p->copied[attr], g_ctx.copy_stream), "cudaEventRecord");
    p->copy_pending[attr] = true;
    MFB_CATCH
}
/* the same for particles [first, first + count) of one attribute: a host that cannot hold the whole container (3LPT 2048^3:
   275 GB of particles on a 196 GB host) hands chunks to its consumer through a small ring of page-locked buffers */
int mfb_particles_stream_out_range(mfb_particles* p, int attr, size_t first, size_t count, void* host) {
    MFB_TRY
    require(attr >= 0 && attr < 4 && host != nullptr, "bad particle attribute / host pointer");
    require(first <= p->np && count <= p->np - first, "particle range outside the container");
    const size_t rs = p->real64 ? 8 : 4, is = p->ids64 ? 8 : 4;
    const char* src = static_cast<const char*>(attr == MFB_ATTR_POS ? p->pos : attr == MFB_ATTR_VEL ? p->vel : attr == MFB_ATTR_IDS ? p->ids : p->mass);
    require(src != nullptr, "particle buffer not allocated");
    const size_t elem = attr == MFB_ATTR_IDS ? is : (attr == MFB_ATTR_MASS ? rs : 3 * rs);
    if (count == 0) return 0;
    if (!g_ctx.copy_stream) cuda_check(cudaStreamCreate(&g_ctx.copy_stream), "cudaStreamCreate(copy)");
    if (!p->copied[attr]) cuda_check(cudaEventCreateWithFlags(&p->copied[attr], cudaEventDisableTiming), "cudaEventCreate");
    cuda_check(cudaEventRecord(p->copied[attr], g_ctx.stream), "cudaEventRecord");
    cuda_check(cudaStreamWaitEvent(g_ctx.copy_stream, p->copied[attr], 0), "cudaStreamWaitEvent");
    cuda_check(cudaMemcpyAsync(host, src + first * elem, count * elem, cudaMemcpyDeviceToHost, g_ctx.copy_stream), "D2H stream_out_range");
    cuda_check(cudaEventRecord(p->copied[attr], g_ctx.copy_stream), "cudaEventRecord");
    p->copy_pending[attr] = true;
    MFB_CATCH
}
int mfb_particles_stream_wait(mfb_particles* p) {
    MFB_TRY
    (void)p;
    if (g_ctx.copy_stream) cuda_check(cudaStreamSynchronize(g_ctx.copy_stream), "sync(copy)");
    MFB_CATCH
}

}  // extern "C"
