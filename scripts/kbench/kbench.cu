// kbench.cu - micro-benchmark of FFT pass kernel variants on synthetic data (development tool, not product).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -I../../monofonic_b200/csrc kbench.cu -o kbench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include "kbench_legacy_kernels.cuh"
using namespace mfb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static std::vector<double2> make_tw(int n) {
    std::vector<double2> h(n);
    for (int m = 0; m < n; ++m) { long double a = -2.0L * 3.14159265358979323846264338327950288L * m / n; h[m].x = (double)cosl(a); h[m].y = (double)sinl(a); }
    return h;
}
template <class P> static double2* make_tw2() {
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
    double2* d; CK(cudaMalloc(&d, h.size() * 16)); CK(cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice));
    return d;
}
__global__ void fill_kernel(double2* p, size_t n, unsigned seed) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        unsigned long long x = (i + 1) * 6364136223846793005ull + seed; x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29;
        p[i] = make_double2(double(int(x & 0xffff) - 32768) / 32768.0, double(int((x >> 20) & 0xffff) - 32768) / 32768.0);
    }
}
__global__ void maxdiff_kernel(const double2* a, const double2* b, size_t n, double* out) {
    double m = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
        m = fmax(m, fmax(fabs(a[i].x - b[i].x), fabs(a[i].y - b[i].y)));
    __shared__ double s[256];
    s[threadIdx.x] = m; __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + k]); __syncthreads(); }
    if (threadIdx.x == 0) atomicMax((unsigned long long*)out, __double_as_longlong(s[0]));  // non-negative doubles order like integers
}

struct Bench {
    int rmw = 0;
    int n, n_small, nz, na, mode;  // mode: SM_PAD (n_small -> n) or SM_TRUNC (n -> n_small)
    double2 *in, *out, *ref; double2* tw; double* dmax;
    size_t in_elems, out_elems;
    StridedArgs args() const {
        StridedArgs a; memset(&a, 0, sizeof(a));
        const int nin = (mode == SM_PAD) ? n_small : n, nout = (mode == SM_TRUNC) ? n_small : n;
        a.in = in; a.out = out;
        a.in_ps = nz; a.in_as = ptrdiff_t(nin) * nz;
        a.out_ps = nz; a.out_as = ptrdiff_t(nout) * nz;
        a.na = na; a.nz = nz; a.nz_fft = nz; a.n_in = nin; a.n_out = nout;
        a.inverse = (mode == SM_PAD); a.scale = 1.0; a.op = rmw ? OP_ADD : OP_ASSIGN; a.a_skip = -1;
        if (mode == SM_PAD) {  // an operand factor like the conv x pass (cost is part of the real kernel)
            a.f.enabled = 1; a.f.a1 = 1; a.f.b1 = 1; a.f.a2 = -1; a.f.b2 = -1; a.f.c1 = -1.0; a.f.c2 = 0.0; a.f.imag = 0;
            for (int d = 0; d < 3; ++d) { a.f.n[d] = n_small; a.f.kf[d] = 0.0628; }
            a.pos_axis = 1; a.a_axis = 0;
        }
        return a;
    }
    double bytes() const { return 16.0 * (double(in_elems) + double(out_elems)); }
};

template <class K> static float time_kernel(K launch, int reps = 8) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / reps;
}
static void report(Bench& B, const char* label, float ms, bool is_ref, int regs, int occ) {
    double md = 0;
    if (is_ref) CK(cudaMemcpy(B.ref, B.out, B.out_elems * 16, cudaMemcpyDeviceToDevice));
    else {
        CK(cudaMemset(B.dmax, 0, 8));
        maxdiff_kernel<<<1184, 256>>>(B.out, B.ref, B.out_elems, B.dmax);
        CK(cudaMemcpy(&md, B.dmax, 8, cudaMemcpyDeviceToHost));
    }
    printf("%-46s %8.3f ms %8.1f GB/s  regs %3d  CTA/SM %d  maxdiff %.2e\n", label, ms, B.bytes() / ms / 1e6, regs, occ, md);
    fflush(stdout);
}
template <class P> static void run_v1(Bench& B, const char* label, bool is_ref) {
    auto k = strided_pass_kernel<P>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM_BYTES));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::NTHREADS, P::SMEM_BYTES));
    StridedArgs a = B.args();
    const unsigned grid = unsigned(B.na) * unsigned((B.nz + P::T - 1) / P::T);
    CK(cudaMemset(B.out, 0xff, B.out_elems * 16));
    float ms = time_kernel([&] { k<<<grid, P::NTHREADS, P::SMEM_BYTES>>>(a, B.tw); });
    report(B, label, ms, is_ref, fa.numRegs, occ);
}
template <class P, int MODE, int MINB, bool BATCH0 = false> static void run_v2(Bench& B, const char* label) {
    if (MODE != B.mode) return;
    auto k = strided2_kernel<P, MODE, MINB, BATCH0>;
    double2* tw2 = make_tw2<P>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM_BYTES));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::NTHREADS, P::SMEM_BYTES));
    StridedArgs a = B.args();
    const unsigned grid = unsigned(B.na) * unsigned((B.nz + P::T - 1) / P::T);
    CK(cudaMemset(B.out, 0xff, B.out_elems * 16));
    float ms = time_kernel([&] { k<<<grid, P::NTHREADS, P::SMEM_BYTES>>>(a, tw2); });
    report(B, label, ms, false, fa.numRegs, occ);
    cudaFree(tw2);
}

template <class P, int MODE> static void run_v3(Bench& B, const char* label, int ctas = 148) {
    if (MODE != B.mode) return;
    auto k = strided3_kernel<P, MODE>;
    double2* tw2 = make_tw2<P>();
    constexpr size_t smem = 2 * P::SMEM_BYTES;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::NTHREADS, smem));
    StridedArgs a = B.args();
    CK(cudaMemset(B.out, 0xff, B.out_elems * 16));
    float ms = time_kernel([&] { k<<<ctas, P::NTHREADS, smem>>>(a, tw2); });
    report(B, label, ms, false, fa.numRegs, occ);
    cudaFree(tw2);
}
#define V3(MODE, ...) run_v3<Plan<__VA_ARGS__>, MODE>(B, "v3 double-buffered <" #__VA_ARGS__ ">")
// (run_v4 / strided4_kernel, the TMA bulk double-buffered strided pass, is gone: see kbench_legacy_kernels.cuh)

static Bench make_bench(int n, int mode, int na) {
    Bench B; B.n = n; B.n_small = 2 * n / 3; B.nz = B.n_small / 2; B.na = na; B.mode = mode;
    const size_t small = size_t(na) * B.n_small * B.nz, big = size_t(na) * n * B.nz;
    B.in_elems = (mode == SM_PAD) ? small : big; B.out_elems = (mode == SM_PAD) ? big : small;
    CK(cudaMalloc(&B.in, B.in_elems * 16)); CK(cudaMalloc(&B.out, B.out_elems * 16)); CK(cudaMalloc(&B.ref, B.out_elems * 16));
    CK(cudaMalloc(&B.dmax, 8));
    auto h = make_tw(n); CK(cudaMalloc(&B.tw, n * 16)); CK(cudaMemcpy(B.tw, h.data(), n * 16, cudaMemcpyHostToDevice));
    fill_kernel<<<1184, 256>>>(B.in, B.in_elems, 1234u + n);
    CK(cudaDeviceSynchronize());
    printf("== n=%d %s  na=%d nz=%d  (%.2f GB in, %.2f GB out)\n", n, mode == SM_PAD ? "PAD y-bwd" : "TRUNC y-fwd", na, B.nz, B.in_elems * 16e-9, B.out_elems * 16e-9);
    return B;
}
static void free_bench(Bench& B) { cudaFree(B.in); cudaFree(B.out); cudaFree(B.ref); cudaFree(B.tw); cudaFree(B.dmax); }


// ------------------------------------------------------------------ fused z pass
struct ZBench { int n, nz; long nrows; double2 *A, *B, *out, *ref, *tw; double* dmax; };
static ZBench make_zbench(int n, long nrows) {
    ZBench Z; Z.n = n; Z.nz = n / 3; Z.nrows = nrows;
    const size_t e = size_t(nrows) * Z.nz;
    CK(cudaMalloc(&Z.A, e * 16)); CK(cudaMalloc(&Z.B, e * 16)); CK(cudaMalloc(&Z.out, e * 16)); CK(cudaMalloc(&Z.ref, e * 16)); CK(cudaMalloc(&Z.dmax, 8));
    auto h = make_tw(n); CK(cudaMalloc(&Z.tw, n * 16)); CK(cudaMemcpy(Z.tw, h.data(), n * 16, cudaMemcpyHostToDevice));
    fill_kernel<<<1184, 256>>>(Z.A, e, 77u); fill_kernel<<<1184, 256>>>(Z.B, e, 99u);
    CK(cudaDeviceSynchronize());
    printf("== z_conv n=%d nrows=%ld nz=%d (%.2f GB per operand)\n", n, nrows, Z.nz, e * 16e-9);
    return Z;
}
static void zreport(ZBench& Z, const char* label, float ms, bool is_ref, int regs, int occ, size_t smem) {
    const size_t e = size_t(Z.nrows) * Z.nz;
    double md = 0;
    if (is_ref) CK(cudaMemcpy(Z.ref, Z.out, e * 16, cudaMemcpyDeviceToDevice));
    else { CK(cudaMemset(Z.dmax, 0, 8)); maxdiff_kernel<<<1184, 256>>>(Z.out, Z.ref, e, Z.dmax); CK(cudaMemcpy(&md, Z.dmax, 8, cudaMemcpyDeviceToHost)); }
    printf("%-50s %8.3f ms %8.1f GB/s  regs %3d  CTA/SM %d smem %3zu KB  maxdiff %.2e\n", label, ms, 3.0 * e * 16 / ms / 1e6, regs, occ, smem / 1024, md);
    fflush(stdout);
}
template <class P> static void zrun_v1(ZBench& Z, const char* label) {
    auto k = z_conv_kernel<P>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM_BYTES));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::NTHREADS, P::SMEM_BYTES));
    const unsigned grid = unsigned((Z.nrows + P::T - 1) / P::T);
    float ms = time_kernel([&] { k<<<grid, P::NTHREADS, P::SMEM_BYTES>>>(Z.A, Z.B, Z.out, Z.nz, Z.nz, Z.nrows, Z.nz, 0.37, Z.tw, 0); });
    zreport(Z, label, ms, true, fa.numRegs, occ, P::SMEM_BYTES);
}
template <class P, int RPC, int MINB> static void zrun_v2(ZBench& Z, const char* label) {
    auto k = z_conv2_kernel<P, RPC, MINB>;
    double2* tw2 = make_tw2<P>();
    constexpr size_t smem = z_conv2_smem<P, RPC>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::TPL * RPC, smem));
    const unsigned grid = unsigned((Z.nrows + RPC - 1) / RPC);
    CK(cudaMemset(Z.out, 0xff, size_t(Z.nrows) * Z.nz * 16));
    float ms = time_kernel([&] { k<<<grid, P::TPL * RPC, smem>>>(Z.A, Z.B, Z.out, Z.nz, Z.nz, Z.nrows, 0.37, tw2, 0); });
    zreport(Z, label, ms, false, fa.numRegs, occ, smem);
    cudaFree(tw2);
}
#define Z2(RPC, MINB, ...) zrun_v2<Plan<__VA_ARGS__>, RPC, MINB>(Z, "z2 <" #__VA_ARGS__ "> rpc " #RPC " minb " #MINB)
static void zbench_all(int which) {
    if (which == 0 || which == 768) {
        ZBench Z = make_zbench(768, 768L * 192);
        zrun_v1<Plan<768, 32, 8, 12, 8, 8>>(Z, "v1 <768,32,8,12,8,8>");
        Z2(2, 1, 768, 64, 2, 12, 8, 8); Z2(2, 4, 768, 64, 2, 12, 8, 8); Z2(2, 5, 768, 64, 2, 12, 8, 8);
        Z2(2, 4, 768, 32, 2, 12, 8, 8); Z2(2, 8, 768, 32, 2, 12, 8, 8);
        Z2(4, 2, 768, 32, 2, 12, 8, 8); Z2(4, 4, 768, 32, 2, 12, 8, 8);
        Z2(4, 2, 768, 64, 2, 12, 8, 8); Z2(2, 3, 768, 96, 2, 12, 8, 8); Z2(2, 2, 768, 96, 2, 12, 8, 8);
        Z2(2, 4, 768, 64, 2, 6, 4, 4, 8); Z2(2, 4, 768, 64, 2, 6, 16, 8); Z2(2, 5, 768, 64, 2, 12, 4, 4, 4);
        Z2(8, 1, 768, 32, 2, 12, 8, 8); Z2(8, 1, 768, 64, 2, 12, 8, 8);
    }
    if (which == 0 || which == 1536) {
        ZBench Z = make_zbench(1536, 1536L * 48);
        zrun_v1<Plan<1536, 32, 8, 16, 12, 8>>(Z, "v1 <1536,32,8,16,12,8>");
        Z2(2, 1, 1536, 64, 2, 12, 16, 8); Z2(2, 2, 1536, 64, 2, 12, 16, 8); Z2(2, 1, 1536, 128, 2, 12, 16, 8); Z2(2, 2, 1536, 128, 2, 12, 16, 8);
        Z2(2, 2, 1536, 32, 2, 12, 16, 8); Z2(4, 1, 1536, 64, 2, 12, 16, 8); Z2(4, 1, 1536, 128, 2, 12, 16, 8);
        Z2(2, 2, 1536, 64, 2, 6, 8, 4, 8); Z2(2, 2, 1536, 128, 2, 6, 8, 4, 8); Z2(2, 2, 1536, 192, 2, 6, 8, 4, 8);
        Z2(2, 2, 1536, 64, 2, 12, 4, 4, 8); Z2(2, 2, 1536, 128, 2, 12, 4, 4, 8);
    }
}

template <class P, int RPC, int MINB> static void zrun_v3(ZBench& Z, const char* label) {
    auto k = z_conv3_kernel<P, RPC, MINB>;
    double2* tw2 = make_tw2<P>();
    constexpr size_t smem = z_conv2_smem<P, RPC>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::TPL * RPC, smem));
    const unsigned grid = unsigned((Z.nrows + RPC - 1) / RPC);
    CK(cudaMemset(Z.out, 0xff, size_t(Z.nrows) * Z.nz * 16));
    float ms = time_kernel([&] { k<<<grid, P::TPL * RPC, smem>>>(Z.A, Z.B, Z.out, Z.nz, Z.nz, Z.nrows, 0.37, tw2, 0); });
    zreport(Z, label, ms, false, fa.numRegs, occ, smem);
    cudaFree(tw2);
}
#define Z3(RPC, MINB, ...) zrun_v3<Plan<__VA_ARGS__>, RPC, MINB>(Z, "z3 <" #__VA_ARGS__ "> rpc " #RPC " minb " #MINB)
template <class P, int RPC, int MINB> static void zrun_v3d(ZBench& Z, const char* label) {
    auto k = z_conv3_kernel<P, RPC, MINB, true>;
    double2* tw2 = make_tw2<P>();
    constexpr size_t smem = z_conv3_direct_smem<P, RPC>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::TPL * RPC, smem));
    const unsigned grid = unsigned((Z.nrows + RPC - 1) / RPC);
    CK(cudaMemset(Z.out, 0xff, size_t(Z.nrows) * Z.nz * 16));
    float ms = time_kernel([&] { k<<<grid, P::TPL * RPC, smem>>>(Z.A, Z.B, Z.out, Z.nz, Z.nz, Z.nrows, 0.37, tw2, 0); });
    zreport(Z, label, ms, false, fa.numRegs, occ, smem);
    cudaFree(tw2);
}
#define Z3D(RPC, MINB, ...) zrun_v3d<Plan<__VA_ARGS__>, RPC, MINB>(Z, "z3 direct <" #__VA_ARGS__ "> rpc " #RPC " minb " #MINB)
// accumulate launches (product groups): how the old sum rows are read (z_conv3_kernel EARLY_ACC 0 / 1 / 2)
template <class P, int RPC, int MINB, int EARLY> static void zrun_acc(ZBench& Z, const char* label) {
    auto k = z_conv3_kernel<P, RPC, MINB, true, EARLY>;
    double2* tw2 = make_tw2<P>();
    constexpr size_t smem = z_conv3_direct_smem<P, RPC>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, P::TPL * RPC, smem));
    const unsigned grid = unsigned((Z.nrows + RPC - 1) / RPC);
    const size_t e = size_t(Z.nrows) * Z.nz;
    CK(cudaMemset(Z.out, 0, e * 16));
    k<<<grid, P::TPL * RPC, smem>>>(Z.A, Z.B, Z.out, Z.nz, Z.nz, Z.nrows, 0.37, tw2, 1);   // 0 + product == product, bit for bit
    double md = 0;
    CK(cudaMemset(Z.dmax, 0, 8)); maxdiff_kernel<<<1184, 256>>>(Z.out, Z.ref, e, Z.dmax); CK(cudaMemcpy(&md, Z.dmax, 8, cudaMemcpyDeviceToHost));
    float ms = time_kernel([&] { k<<<grid, P::TPL * RPC, smem>>>(Z.A, Z.B, Z.out, Z.nz, Z.nz, Z.nrows, 0.37, tw2, 1); });
    printf("%-58s %8.3f ms %8.1f GB/s  regs %3d  CTA/SM %d  maxdiff(0+product vs product) %.2e\n", label, ms, 4.0 * e * 16 / ms / 1e6, fa.numRegs, occ, md);
    fflush(stdout);
    cudaFree(tw2);
}
#define ZACC(EARLY, RPC, MINB, ...) zrun_acc<Plan<__VA_ARGS__>, RPC, MINB, EARLY>(Z, "z3 direct accumulate early=" #EARLY " <" #__VA_ARGS__ "> minb " #MINB)
static void zbench_acc() {
    {
        ZBench Z = make_zbench(1536, 1536L * 192);
        zrun_v3d<Plan<1536, 64, 2, 12, 16, 8>, 2, 3>(Z, "z3 direct (no accumulate) <1536,64,2,12,16,8> minb 3");
        CK(cudaMemcpy(Z.ref, Z.out, size_t(Z.nrows) * Z.nz * 16, cudaMemcpyDeviceToDevice));
        ZACC(0, 2, 3, 1536, 64, 2, 12, 16, 8); ZACC(1, 2, 3, 1536, 64, 2, 12, 16, 8); ZACC(2, 2, 3, 1536, 64, 2, 12, 16, 8);
    }
    {
        ZBench Z = make_zbench(768, 768L * 384);
        zrun_v3d<Plan<768, 64, 2, 12, 8, 8>, 2, 5>(Z, "z3 direct (no accumulate) <768,64,2,12,8,8> minb 5");
        CK(cudaMemcpy(Z.ref, Z.out, size_t(Z.nrows) * Z.nz * 16, cudaMemcpyDeviceToDevice));
        ZACC(0, 2, 5, 768, 64, 2, 12, 8, 8); ZACC(1, 2, 5, 768, 64, 2, 12, 8, 8); ZACC(2, 2, 5, 768, 64, 2, 12, 8, 8);
    }
    {
        ZBench Z = make_zbench(3072, 3072L * 48);
        zrun_v3d<Plan<3072, 128, 2, 12, 16, 16>, 2, 2>(Z, "z3 direct (no accumulate) <3072,128,2,12,16,16> minb 2");
        CK(cudaMemcpy(Z.ref, Z.out, size_t(Z.nrows) * Z.nz * 16, cudaMemcpyDeviceToDevice));
        ZACC(0, 2, 2, 3072, 128, 2, 12, 16, 16); ZACC(2, 2, 2, 3072, 128, 2, 12, 16, 16);
    }
}
static void zbench4(int which) {
    if (which == 0 || which == 768) {
        ZBench Z = make_zbench(768, 768L * 768);
        zrun_v1<Plan<768, 32, 8, 12, 8, 8>>(Z, "v1 <768,32,8,12,8,8>");
        Z3(2, 5, 768, 64, 2, 12, 8, 8);
        Z3D(2, 5, 768, 64, 2, 12, 8, 8); Z3D(2, 6, 768, 64, 2, 12, 8, 8); Z3D(2, 8, 768, 64, 2, 12, 8, 8);
        Z3D(2, 8, 768, 32, 2, 12, 8, 8); Z3D(2, 6, 768, 32, 2, 12, 8, 8); Z3D(2, 4, 768, 32, 2, 12, 8, 8);
    }
    if (which == 0 || which == 1536) {
        ZBench Z = make_zbench(1536, 1536L * 384);
        zrun_v1<Plan<1536, 32, 8, 16, 12, 8>>(Z, "v1 <1536,32,8,16,12,8>");
        Z3(2, 2, 1536, 128, 2, 12, 16, 8);
        Z3D(2, 2, 1536, 128, 2, 12, 16, 8); Z3D(2, 4, 1536, 64, 2, 12, 16, 8); Z3D(2, 3, 1536, 64, 2, 12, 16, 8);
        Z3D(2, 4, 1536, 64, 2, 12, 4, 4, 8); Z3D(2, 4, 1536, 32, 2, 12, 16, 8); Z3D(2, 4, 1536, 32, 2, 12, 4, 4, 8);
        Z3D(2, 3, 1536, 96, 2, 12, 16, 8); Z3D(2, 2, 1536, 64, 2, 12, 16, 8); Z3D(2, 4, 1536, 64, 2, 12, 8, 2, 8);
    }
    if (which == 0 || which == 3072) {
        ZBench Z = make_zbench(3072, 3072L * 96);
        zrun_v1<Plan<3072, 64, 4, 16, 16, 12>>(Z, "v1 <3072,64,4,16,16,12>");
        Z3D(2, 2, 3072, 128, 2, 6, 8, 8, 8); Z3D(2, 1, 3072, 128, 2, 6, 8, 8, 8); Z3D(2, 2, 3072, 128, 2, 12, 16, 16); Z3D(2, 2, 3072, 256, 2, 12, 16, 16);
        Z3D(2, 1, 3072, 128, 2, 12, 16, 16); Z3D(2, 2, 3072, 64, 2, 12, 16, 16); Z3D(2, 2, 3072, 128, 2, 12, 4, 8, 8); Z3D(2, 2, 3072, 256, 2, 12, 4, 8, 8);
        Z3D(2, 2, 3072, 192, 2, 12, 16, 16); Z3D(2, 2, 3072, 256, 2, 6, 8, 8, 8);
    }
}
static void zbench3(int which) {
    if (which == 0 || which == 768) {
        ZBench Z = make_zbench(768, 768L * 192);
        zrun_v1<Plan<768, 32, 8, 12, 8, 8>>(Z, "v1 <768,32,8,12,8,8>");
        Z2(2, 4, 768, 32, 2, 12, 8, 8);
        Z3(2, 4, 768, 32, 2, 12, 8, 8); Z3(2, 5, 768, 32, 2, 12, 8, 8); Z3(2, 3, 768, 32, 2, 12, 8, 8);
        Z3(2, 4, 768, 64, 2, 12, 8, 8); Z3(2, 3, 768, 64, 2, 12, 8, 8); Z3(2, 5, 768, 64, 2, 12, 8, 8);
        Z3(4, 2, 768, 32, 2, 12, 8, 8); Z3(4, 2, 768, 64, 2, 12, 8, 8);
        Z3(2, 4, 768, 32, 2, 6, 16, 8); Z3(2, 4, 768, 64, 2, 6, 16, 8); Z3(2, 4, 768, 48, 2, 12, 8, 8);
    }
    if (which == 0 || which == 1536) {
        ZBench Z = make_zbench(1536, 1536L * 48);
        zrun_v1<Plan<1536, 32, 8, 16, 12, 8>>(Z, "v1 <1536,32,8,16,12,8>");
        Z2(2, 2, 1536, 64, 2, 12, 4, 4, 8);
        Z3(2, 2, 1536, 64, 2, 12, 4, 4, 8); Z3(2, 2, 1536, 32, 2, 12, 4, 4, 8); Z3(2, 2, 1536, 128, 2, 12, 4, 4, 8);
        Z3(2, 2, 1536, 64, 2, 12, 16, 8); Z3(2, 2, 1536, 128, 2, 12, 16, 8); Z3(2, 1, 1536, 64, 2, 12, 16, 8);
        Z3(2, 2, 1536, 64, 2, 6, 8, 4, 8); Z3(2, 2, 1536, 64, 2, 12, 8, 2, 8); Z3(2, 2, 1536, 96, 2, 12, 16, 8);
    }
}

#define V2(MODE, MINB, ...) run_v2<Plan<__VA_ARGS__>, MODE, MINB>(B, "v2 <" #__VA_ARGS__ "> minb " #MINB)
#define V2B(MODE, MINB, ...) run_v2<Plan<__VA_ARGS__>, MODE, MINB, true>(B, "v2 batch0 <" #__VA_ARGS__ "> minb " #MINB)

static void sbench2(int which) {
    if (which == 0 || which == 768) {
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(768, mode, 256);
            run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 <768,32,8,12,8,8>", true);
            V2(SM_PAD, 2, 768, 32, 8, 12, 8, 8); V2B(SM_PAD, 2, 768, 32, 8, 12, 8, 8); V2B(SM_PAD, 1, 768, 32, 8, 12, 8, 8);
            V2B(SM_PAD, 2, 768, 32, 8, 6, 16, 8); V2B(SM_PAD, 4, 768, 32, 4, 12, 8, 8); V2B(SM_PAD, 2, 768, 16, 8, 12, 8, 8);
            V2(SM_TRUNC, 2, 768, 32, 8, 8, 8, 12); V2B(SM_TRUNC, 2, 768, 32, 8, 8, 8, 12); V2B(SM_TRUNC, 2, 768, 48, 8, 8, 8, 12);
            if (mode == SM_TRUNC) { B.rmw = 1; printf("  -- RMW (op ADD)\n"); run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 rmw", true); V2(SM_TRUNC, 2, 768, 32, 8, 8, 8, 12); V2B(SM_TRUNC, 2, 768, 32, 8, 8, 8, 12); }
            free_bench(B);
        }
    }
    if (which == 0 || which == 1536) {
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(1536, mode, 96);
            run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 <1536,32,8,16,12,8>", true);
            V2(SM_PAD, 1, 1536, 64, 8, 12, 16, 8); V2B(SM_PAD, 1, 1536, 64, 8, 12, 16, 8); V2B(SM_PAD, 2, 1536, 64, 4, 12, 16, 8);
            V2B(SM_PAD, 1, 1536, 32, 8, 12, 16, 8); V2B(SM_PAD, 2, 1536, 32, 4, 12, 16, 8); V2B(SM_PAD, 1, 1536, 64, 8, 12, 8, 16);
            V2(SM_TRUNC, 2, 1536, 64, 4, 8, 16, 12); V2B(SM_TRUNC, 2, 1536, 64, 4, 8, 16, 12); V2B(SM_TRUNC, 1, 1536, 64, 8, 8, 16, 12);
            V2B(SM_TRUNC, 1, 1536, 96, 8, 8, 16, 12); V2B(SM_TRUNC, 2, 1536, 96, 4, 8, 16, 12); V2B(SM_TRUNC, 1, 1536, 96, 8, 16, 8, 12);
            if (mode == SM_TRUNC) { B.rmw = 1; printf("  -- RMW (op ADD)\n"); run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 rmw", true); V2(SM_TRUNC, 2, 1536, 64, 4, 8, 16, 12); V2B(SM_TRUNC, 1, 1536, 64, 8, 8, 16, 12); }
            free_bench(B);
        }
    }
}
// ---------------------------------------------------------------------------------------------------------
// "seg": what HBM delivers for the ACCESS PATTERN of a strided pass, without any FFT: a CTA copies one tile =
// n positions x SEG contiguous complex values (SEG*16 bytes) at a position stride of nz*16 bytes, tiles ordered
// (a, z-tile) like the pass kernels.  UNR loads per thread are issued before the stores.
// XLIKE: the x-pass layout [position][a][z] (position stride na*nz*16 bytes = 8.4 MB at 1024^3) instead of the y-pass
// layout [a][position][z] (position stride nz*16 bytes = 8 KB)
template <int SEG, int UNR, bool XLIKE = false>
__global__ void __launch_bounds__(256) seg_copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, int n, int nz, int na) {
    const int tiles_per_a = nz / SEG;
    const int a = blockIdx.x / tiles_per_a, z0 = (blockIdx.x - a * tiles_per_a) * SEG;
    const int lane = threadIdx.x % SEG, q = threadIdx.x / SEG;
    constexpr int PPI = 256 / SEG;   // positions per iteration
    const ptrdiff_t ps = XLIKE ? ptrdiff_t(na) * nz : nz, as = XLIKE ? nz : ptrdiff_t(n) * nz;
    const double2* ip = in + ptrdiff_t(a) * as + z0 + lane;
    double2* op = out + ptrdiff_t(a) * as + z0 + lane;
    for (int p0 = q; p0 < n; p0 += PPI * UNR) {
        double2 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) { const int p = p0 + u * PPI; if (p < n) v[u] = mfb_ld_stream(ip + ptrdiff_t(p) * ps); }
#pragma unroll
        for (int u = 0; u < UNR; ++u) { const int p = p0 + u * PPI; if (p < n) mfb_st_stream(op + ptrdiff_t(p) * ps, v[u]); }
    }
}
template <int SEG, int UNR> static void segx_one(int n, int nz, int na, const double2* in, double2* out) {
    const unsigned grid = unsigned(na) * unsigned(nz / SEG);
    auto launch = [&] { seg_copy_kernel<SEG, UNR, true><<<grid, 256>>>(in, out, n, nz, na); };
    const float ms = time_kernel(launch);
    const double gb = 32.0 * double(n) * nz * na / 1e9;
    printf("x-like seg %4d B x unroll %2d  n=%d nz=%d na=%d  (position stride %.1f MB)  %.3f ms  %8.1f GB/s\n", SEG * 16, UNR, n, nz, na,
           double(na) * nz * 16 / 1e6, ms, gb / (ms * 1e-3));
}
template <int SEG, int UNR> static void seg_one(int n, int nz, int na, const double2* in, double2* out) {
    const unsigned grid = unsigned(na) * unsigned(nz / SEG);
    auto launch = [&] { seg_copy_kernel<SEG, UNR><<<grid, 256>>>(in, out, n, nz, na); };
    const float ms = time_kernel(launch);
    const double gb = 32.0 * double(n) * nz * na / 1e9;
    printf("seg %4d B x unroll %2d  n=%d nz=%d na=%d   %.3f ms  %8.1f GB/s\n", SEG * 16, UNR, n, nz, na, ms, gb / (ms * 1e-3));
}
static void segbench(int n) {
    const int nz = 512, na = (n >= 1536) ? 96 : 256;
    const size_t elems = size_t(n) * nz * na;
    double2 *in, *out;
    CK(cudaMalloc(&in, elems * 16)); CK(cudaMalloc(&out, elems * 16));
    fill_kernel<<<1184, 256>>>(in, elems, 1u);
    CK(cudaDeviceSynchronize());
    printf("== segment-copy pattern test (%.2f GB each way)\n", elems * 16 / 1e9);
    seg_one<2, 8>(n, nz, na, in, out);  seg_one<4, 8>(n, nz, na, in, out);   seg_one<8, 8>(n, nz, na, in, out);
    seg_one<16, 8>(n, nz, na, in, out); seg_one<32, 8>(n, nz, na, in, out);  seg_one<64, 8>(n, nz, na, in, out);
    seg_one<8, 4>(n, nz, na, in, out);  seg_one<8, 16>(n, nz, na, in, out);  seg_one<16, 16>(n, nz, na, in, out);
    seg_one<4, 16>(n, nz, na, in, out); seg_one<32, 16>(n, nz, na, in, out);
    CK(cudaFree(in)); CK(cudaFree(out));
    {   // the x pass of a 1024^3 product: 1536 positions x 1024 a x 512 z  (12.9 GB each way)
        const int xn = 1536, xna = 1024;
        const size_t xe = size_t(xn) * nz * xna;
        CK(cudaMalloc(&in, xe * 16)); CK(cudaMalloc(&out, xe * 16));
        fill_kernel<<<1184, 256>>>(in, xe, 1u);
        CK(cudaDeviceSynchronize());
        printf("== same test at the full size of one 1024^3 operand (%.2f GB each way): y-like, then x-like layout\n", xe * 16 / 1e9);
        seg_one<8, 8>(xn, nz, xna, in, out); seg_one<4, 8>(xn, nz, xna, in, out);
        segx_one<8, 8>(xn, nz, xna, in, out); segx_one<4, 8>(xn, nz, xna, in, out); segx_one<8, 16>(xn, nz, xna, in, out); segx_one<16, 8>(xn, nz, xna, in, out);
        CK(cudaFree(in)); CK(cudaFree(out));
    }
}

int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    const char* what = argc > 2 ? argv[2] : "all";
    if (!strcmp(what, "seg")) { segbench(which ? which : 1536); return 0; }
    if (!strcmp(what, "z4")) { zbench4(which); return 0; }
    if (!strcmp(what, "zacc")) { zbench_acc(); return 0; }
    if (!strcmp(what, "t2")) {   // narrow tiles for the y-like passes: 2-line tiles (32-byte segments), 4 CTAs per SM
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(1536, mode, 384);
            run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 <1536,32,8,16,12,8>", true);
            V2(SM_PAD, 1, 1536, 64, 8, 12, 16, 8); V2(SM_PAD, 2, 1536, 64, 4, 12, 16, 8);
            V2(SM_PAD, 4, 1536, 64, 2, 12, 16, 8); V2(SM_PAD, 4, 1536, 32, 2, 12, 16, 8); V2(SM_PAD, 3, 1536, 64, 2, 12, 16, 8); V2(SM_PAD, 4, 1536, 128, 2, 12, 16, 8);
            V2(SM_PAD, 3, 1536, 32, 4, 12, 16, 8);
            V2(SM_TRUNC, 2, 1536, 64, 4, 8, 16, 12);
            V2(SM_TRUNC, 4, 1536, 64, 2, 8, 16, 12); V2(SM_TRUNC, 4, 1536, 32, 2, 8, 16, 12); V2(SM_TRUNC, 3, 1536, 64, 2, 8, 16, 12); V2(SM_TRUNC, 4, 1536, 128, 2, 8, 16, 12);
            free_bench(B);
        }
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(768, mode, 768);
            run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 <768,32,8,12,8,8>", true);
            V2(SM_PAD, 2, 768, 32, 8, 12, 8, 8); V2(SM_PAD, 4, 768, 32, 4, 12, 8, 8); V2(SM_PAD, 8, 768, 32, 2, 12, 8, 8); V2(SM_PAD, 6, 768, 64, 2, 12, 8, 8);
            V2(SM_TRUNC, 2, 768, 32, 8, 8, 8, 12); V2(SM_TRUNC, 4, 768, 32, 4, 8, 8, 12); V2(SM_TRUNC, 8, 768, 32, 2, 8, 8, 12); V2(SM_TRUNC, 6, 768, 64, 2, 8, 8, 12);
            free_bench(B);
        }
        return 0;
    }
    if (!strcmp(what, "s4")) {   // round 2b: TMA (cp.async.bulk) staging of the next tile against the LDGSTS version and v2
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(1536, mode, 384);
            run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 <1536,32,8,16,12,8>", true);
            run_v2<Plan<1536, 64, 8, 12, 16, 8>, SM_PAD, 1, true>(B, "v2 <1536,64,8,12,16,8> minb 1 batch0");
            run_v2<Plan<1536, 64, 4, 12, 16, 8>, SM_PAD, 2>(B, "v2 <1536,64,4,12,16,8> minb 2");
            run_v2<Plan<1536, 64, 4, 8, 16, 12>, SM_TRUNC, 2>(B, "v2 <1536,64,4,8,16,12> minb 2");
            V3(SM_PAD, 1536, 64, 4, 12, 16, 8); V3(SM_PAD, 1536, 128, 4, 12, 16, 8);
            V3(SM_TRUNC, 1536, 64, 4, 8, 16, 12);
            free_bench(B);
        }
        for (int mode : {SM_PAD, SM_TRUNC}) {   // n = 768: two 8-line tiles fit
            Bench B = make_bench(768, mode, 768);
            run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 <768,32,8,12,8,8>", true);
            run_v2<Plan<768, 32, 8, 12, 8, 8>, SM_PAD, 2>(B, "v2 <768,32,8,12,8,8> minb 2");
            run_v2<Plan<768, 32, 8, 8, 8, 12>, SM_TRUNC, 2>(B, "v2 <768,32,8,8,8,12> minb 2");
            free_bench(B);
        }
        return 0;
    }
    if (!strcmp(what, "s3")) {
        if (which == 0 || which == 768) for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(768, mode, 768);
            run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 <768,32,8,12,8,8>", true);
            run_v2<Plan<768, 32, 8, 12, 8, 8>, SM_PAD, 2>(B, "v2 <768,32,8,12,8,8> minb 2");
            run_v2<Plan<768, 32, 8, 8, 8, 12>, SM_TRUNC, 2>(B, "v2 <768,32,8,8,8,12> minb 2");
            V3(SM_PAD, 768, 32, 8, 12, 8, 8); V3(SM_PAD, 768, 64, 8, 12, 8, 8); V3(SM_PAD, 768, 64, 4, 12, 8, 8); V3(SM_PAD, 768, 128, 4, 12, 8, 8);
            V3(SM_PAD, 768, 64, 8, 6, 16, 8);
            V3(SM_TRUNC, 768, 32, 8, 8, 8, 12); V3(SM_TRUNC, 768, 64, 8, 8, 8, 12); V3(SM_TRUNC, 768, 64, 4, 8, 8, 12); V3(SM_TRUNC, 768, 128, 4, 8, 8, 12);
            if (mode == SM_TRUNC) { B.rmw = 1; printf("  -- RMW (op ADD)\n"); run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 rmw", true);
                run_v2<Plan<768, 32, 8, 8, 8, 12>, SM_TRUNC, 2>(B, "v2 <768,32,8,8,8,12> minb 2"); V3(SM_TRUNC, 768, 32, 8, 8, 8, 12); V3(SM_TRUNC, 768, 64, 8, 8, 8, 12); }
            free_bench(B);
        }
        if (which == 0 || which == 1536) for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(1536, mode, 384);
            run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 <1536,32,8,16,12,8>", true);
            run_v2<Plan<1536, 64, 8, 12, 16, 8>, SM_PAD, 1>(B, "v2 <1536,64,8,12,16,8> minb 1");
            run_v2<Plan<1536, 64, 4, 8, 16, 12>, SM_TRUNC, 2>(B, "v2 <1536,64,4,8,16,12> minb 2");
            V3(SM_PAD, 1536, 64, 4, 12, 16, 8); V3(SM_PAD, 1536, 128, 4, 12, 16, 8); V3(SM_PAD, 1536, 32, 4, 12, 16, 8); V3(SM_PAD, 1536, 128, 2, 12, 16, 8);
            V3(SM_PAD, 1536, 128, 4, 12, 8, 16); V3(SM_PAD, 1536, 256, 2, 12, 16, 8); V3(SM_PAD, 1536, 128, 4, 12, 8, 4, 4); V3(SM_PAD, 1536, 96, 4, 12, 16, 8);
            V3(SM_TRUNC, 1536, 64, 4, 8, 16, 12); V3(SM_TRUNC, 1536, 128, 4, 8, 16, 12); V3(SM_TRUNC, 1536, 32, 4, 8, 16, 12); V3(SM_TRUNC, 1536, 128, 4, 16, 8, 12);
            V3(SM_TRUNC, 1536, 128, 2, 8, 16, 12); V3(SM_TRUNC, 1536, 96, 4, 8, 16, 12);
            if (mode == SM_TRUNC) { B.rmw = 1; printf("  -- RMW (op ADD)\n"); run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 rmw", true);
                run_v2<Plan<1536, 64, 8, 8, 16, 12>, SM_TRUNC, 1>(B, "v2 <1536,64,8,8,16,12> minb 1"); V3(SM_TRUNC, 1536, 64, 4, 8, 16, 12); V3(SM_TRUNC, 1536, 128, 4, 8, 16, 12); }
            free_bench(B);
        }
        return 0;
    }
    if (!strcmp(what, "z") || !strcmp(what, "all")) zbench_all(which);
    if (!strcmp(what, "z")) return 0;
    if (!strcmp(what, "z3")) { zbench3(which); return 0; }
    if (!strcmp(what, "s2")) { sbench2(which); return 0; }
    if (!strcmp(what, "zs")) { zbench3(which); sbench2(which); return 0; }
    if (which == 0 || which == 768) {
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(768, mode, 256);
            run_v1<Plan<768, 32, 8, 12, 8, 8>>(B, "v1 <768,32,8,12,8,8>", true);
            V2(SM_PAD, 1, 768, 32, 8, 12, 8, 8);   V2(SM_PAD, 2, 768, 32, 8, 12, 8, 8);
            V2(SM_PAD, 2, 768, 64, 8, 12, 8, 8);   V2(SM_PAD, 1, 768, 64, 8, 12, 8, 8);
            V2(SM_PAD, 2, 768, 32, 4, 12, 8, 8);   V2(SM_PAD, 4, 768, 32, 4, 12, 8, 8);
            V2(SM_PAD, 2, 768, 64, 4, 12, 8, 8);   V2(SM_PAD, 3, 768, 64, 4, 12, 8, 8);
            V2(SM_PAD, 2, 768, 32, 8, 6, 8, 16);   V2(SM_PAD, 2, 768, 64, 8, 6, 16, 8);
            V2(SM_PAD, 2, 768, 32, 8, 12, 16, 4);  V2(SM_PAD, 2, 768, 64, 8, 6, 8, 4, 4);
            V2(SM_TRUNC, 1, 768, 32, 8, 8, 8, 12); V2(SM_TRUNC, 2, 768, 32, 8, 8, 8, 12);
            V2(SM_TRUNC, 2, 768, 64, 8, 8, 8, 12); V2(SM_TRUNC, 1, 768, 64, 8, 8, 8, 12);
            V2(SM_TRUNC, 2, 768, 32, 4, 8, 8, 12); V2(SM_TRUNC, 4, 768, 32, 4, 8, 8, 12);
            V2(SM_TRUNC, 2, 768, 64, 4, 8, 8, 12); V2(SM_TRUNC, 3, 768, 64, 4, 8, 8, 12);
            V2(SM_TRUNC, 2, 768, 32, 8, 16, 8, 6); V2(SM_TRUNC, 2, 768, 64, 8, 8, 16, 6);
            V2(SM_TRUNC, 2, 768, 64, 8, 4, 4, 8, 6);
            free_bench(B);
        }
    }
    if (which == 0 || which == 1536) {
        for (int mode : {SM_PAD, SM_TRUNC}) {
            Bench B = make_bench(1536, mode, 96);
            run_v1<Plan<1536, 32, 8, 16, 12, 8>>(B, "v1 <1536,32,8,16,12,8>", true);
            V2(SM_PAD, 1, 1536, 32, 8, 12, 16, 8);  V2(SM_PAD, 1, 1536, 64, 8, 12, 16, 8);
            V2(SM_PAD, 1, 1536, 32, 4, 12, 16, 8);  V2(SM_PAD, 2, 1536, 32, 4, 12, 16, 8);
            V2(SM_PAD, 1, 1536, 64, 4, 12, 16, 8);  V2(SM_PAD, 2, 1536, 64, 4, 12, 16, 8);
            V2(SM_PAD, 2, 1536, 128, 4, 12, 16, 8); V2(SM_PAD, 1, 1536, 128, 4, 12, 16, 8);
            V2(SM_PAD, 2, 1536, 64, 4, 12, 8, 16);  V2(SM_PAD, 2, 1536, 64, 4, 6, 16, 16);
            V2(SM_PAD, 2, 1536, 64, 4, 12, 8, 4, 4); V2(SM_PAD, 2, 1536, 128, 4, 12, 8, 4, 4);
            V2(SM_PAD, 2, 1536, 64, 4, 6, 8, 8, 4);  V2(SM_PAD, 1, 1536, 128, 8, 6, 8, 8, 4);
            V2(SM_TRUNC, 1, 1536, 32, 8, 8, 16, 12);  V2(SM_TRUNC, 1, 1536, 64, 8, 8, 16, 12);
            V2(SM_TRUNC, 1, 1536, 32, 4, 8, 16, 12);  V2(SM_TRUNC, 2, 1536, 32, 4, 8, 16, 12);
            V2(SM_TRUNC, 1, 1536, 64, 4, 8, 16, 12);  V2(SM_TRUNC, 2, 1536, 64, 4, 8, 16, 12);
            V2(SM_TRUNC, 2, 1536, 128, 4, 8, 16, 12); V2(SM_TRUNC, 2, 1536, 64, 4, 16, 8, 12);
            V2(SM_TRUNC, 2, 1536, 64, 4, 16, 16, 6);  V2(SM_TRUNC, 2, 1536, 64, 4, 4, 4, 8, 12);
            V2(SM_TRUNC, 2, 1536, 128, 4, 4, 4, 8, 12); V2(SM_TRUNC, 2, 1536, 64, 4, 4, 8, 8, 6);
            V2(SM_TRUNC, 1, 1536, 128, 8, 4, 8, 8, 6);
            free_bench(B);
        }
    }
    return 0;
}
