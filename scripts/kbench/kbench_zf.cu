// kbench_zf.cu - micro-benchmark of the shared-field z pass (z_fields_kernel, fft_kernels_zf.cuh) against the multi-product
// launches it replaces (z_multi2_kernel, fft_kernels_zm.cuh) on synthetic rows, n = 1536 (3LPT 1024^3).  Development tool, not product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -I../../monofonic_b200/csrc kbench_zf.cu -o kbench_zf
//   kbench_zf [x-planes of 1536 rows, default 96] [variant filter: 0 = all, 1 = product variants only]
// Every variant is checked against the multi-product result (relative to the largest element of the target).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <type_traits>
#include <vector>
#include "fft_kernels_zf.cuh"
using namespace mfb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

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
__global__ void maxdiff_kernel(const double2* a, const double2* b, size_t n, double* out) {   // out[0] = max|a-b|, out[1] = max|b|
    double m = 0, r = 0;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        m = fmax(m, fmax(fabs(a[i].x - b[i].x), fabs(a[i].y - b[i].y)));
        r = fmax(r, fmax(fabs(b[i].x), fabs(b[i].y)));
    }
    __shared__ double s[256], t[256];
    s[threadIdx.x] = m; t[threadIdx.x] = r; __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) { s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + k]); t[threadIdx.x] = fmax(t[threadIdx.x], t[threadIdx.x + k]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) {   // non-negative doubles order like integers
        atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(s[0]));
        atomicMax((unsigned long long*)(out + 1), (unsigned long long)__double_as_longlong(t[0]));
    }
}
template <class K> static float time_kernel(K launch, int reps = 6) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms / reps;
}

constexpr int N = 1536, NZ = N / 3;
using PlanZ = Plan<N, 128, 2, 12, 16, 8>;

struct Setup {
    long nrows;
    double2* f[12];       // base rows of the 12 fields: xx yy zz xy xz yz of phi, then of phi2
    double2* out[5];      // targets of the variant under test
    double2* ref[5];      // targets of the multi-product pass
    double* dmax;
    double2* tw;
    int e[12] = {0, 0, 2, 0, 1, 1, 0, 0, 2, 0, 1, 1};                       // k_z power the z pass applies on load
    double c[12] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};          // (i k_a)(i k_b)
    double kfz = 0.0628;
};

// ---- the launches the shared-field pass replaces: one multi-product launch per target
static ZmOperand op1(const Setup& S, int i) {
    ZmOperand o; memset(&o, 0, sizeof(o));
    o.n = 1; o.t[0].p = S.f[i]; o.t[0].c = S.c[i]; o.t[0].e = S.e[i]; o.t[1] = o.t[0];
    return o;
}
static ZmOperand op2(const Setup& S, int i, int j, double sj) {   // field i + sj * field j
    ZmOperand o = op1(S, i);
    o.n = 2; o.t[1].p = S.f[j]; o.t[1].c = sj * S.c[j]; o.t[1].e = S.e[j];
    return o;
}
static void add(ZmArgs& a, const ZmOperand& x, const ZmOperand& y, double coeff, bool square = false) {
    ZmProduct& p = a.prod[a.nprod++];
    p.a = x; p.b = y; p.coeff = coeff; p.square = square ? 1 : 0;
}
enum { xx, yy, zz, xy, xz, yz, XX, YY, ZZ, XY, XZ, YZ };
static std::vector<ZmArgs> multi_targets(const Setup& S, int prog) {
    std::vector<ZmArgs> v;
    auto fresh = [&](int it) { ZmArgs a; memset(&a, 0, sizeof(a)); a.out = S.ref[it]; a.in_pitch = a.out_pitch = NZ; a.nrows = S.nrows; a.kfz = S.kfz; return a; };
    if (prog == ZF_PROG_A) {   // src/ic_generator.cc:482-486, 522-526
        ZmArgs a = fresh(0);
        add(a, op1(S, xx), op2(S, yy, zz, +1.0), 1.0); add(a, op1(S, yy), op1(S, zz), 1.0);
        add(a, op1(S, xy), op1(S, xy), -1.0, true); add(a, op1(S, xz), op1(S, xz), -1.0, true); add(a, op1(S, yz), op1(S, yz), -1.0, true);
        v.push_back(a);
        a = fresh(1); add(a, op1(S, xx), op1(S, yy), 1.0); add(a, op1(S, xy), op1(S, xy), -1.0, true); v.push_back(a);
        a = fresh(2); add(a, op1(S, xy), op1(S, xz), 1.0); v.push_back(a);
        a = fresh(3); add(a, op1(S, yz), op1(S, yz), 1.0, true); v.push_back(a);
        a = fresh(4); add(a, op1(S, xz), op1(S, xz), 1.0, true); v.push_back(a);
    } else {                   // :533-538, :545-556
        ZmArgs a = fresh(0);
        add(a, op1(S, xx), op2(S, YY, ZZ, +1.0), -5.0 / 7.0); add(a, op1(S, yy), op2(S, ZZ, XX, +1.0), -5.0 / 7.0);
        add(a, op1(S, zz), op2(S, XX, YY, +1.0), -5.0 / 7.0);
        add(a, op1(S, xy), op1(S, XY), 10.0 / 7.0); add(a, op1(S, xz), op1(S, XZ), 10.0 / 7.0); add(a, op1(S, yz), op1(S, YZ), 10.0 / 7.0);
        v.push_back(a);
        const int dd[3][6] = {{XY, xz, XZ, xy, yz, YY}, {YZ, xy, XY, yz, xz, ZZ}, {XZ, yz, YZ, xz, xy, XX}};
        const int pq[3][4] = {{YY, ZZ, yy, zz}, {ZZ, XX, zz, xx}, {XX, YY, xx, yy}};
        const int PQ[3] = {YZ, XZ, XY};
        for (int d = 0; d < 3; ++d) {
            a = fresh(1 + d);
            add(a, op1(S, dd[d][0]), op1(S, dd[d][1]), 1.0); add(a, op1(S, dd[d][2]), op1(S, dd[d][3]), -1.0);
            add(a, op1(S, dd[d][4]), op2(S, pq[d][0], pq[d][1], -1.0), 1.0); add(a, op1(S, PQ[d]), op2(S, pq[d][2], pq[d][3], -1.0), -1.0);
            v.push_back(a);
        }
    }
    return v;
}
static float run_multi(const Setup& S, int prog, double* cffts) {
    auto k = z_multi2_kernel<PlanZ, 128, 3, true>;
    constexpr size_t smem = z_multi2_smem<PlanZ, true>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<ZmArgs> T = multi_targets(S, prog);
    double n = 0;
    for (const ZmArgs& a : T) { for (int i = 0; i < a.nprod; ++i) n += a.prod[i].square ? 0.5 : 1.0; n += 0.5; }
    *cffts = n;
    const unsigned grid = unsigned((S.nrows + 1) / 2);
    return time_kernel([&] { for (const ZmArgs& a : T) k<<<grid, 128, smem>>>(a, S.tw); });
}

template <int NTG, int G, int PROG, bool PF, bool COOP, class PL = PlanZ> static void run_fields(const Setup& S, const char* label, float multi_ms) {
    using PR = ZfProg<PROG>;
    auto k = z_fields_kernel<PL, NTG, G, PROG, PF, COOP>;
    constexpr size_t smem = z_fields_smem<PL, G, PROG, PF>();
    double2* tw = std::is_same<PL, PlanZ>::value ? S.tw : make_tw2<PL>();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NTG * G, smem));
    int dev, sms; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ZfArgs a; memset(&a, 0, sizeof(a));
    const int off = 0;
    for (int i = 0; i < PR::NF; ++i) { a.f[i].p = S.f[off + i]; a.f[i].c = S.c[i]; a.f[i].e = S.e[i]; }
    for (int t = 0; t < PR::NT; ++t) { a.out[t] = S.out[t]; CK(cudaMemset(S.out[t], 0xff, size_t(S.nrows) * NZ * 16)); }
    a.in_pitch = a.out_pitch = NZ; a.nrows = S.nrows; a.kfz = S.kfz;
    a.coef[0] = -5.0 / 7.0; a.coef[1] = 10.0 / 7.0;
    const unsigned grid = unsigned(std::min<long>(S.nrows, long(sms) * occ));
    const float ms = time_kernel([&] { k<<<grid, NTG * G, smem>>>(a, tw); });
    double worst = 0;
    for (int t = 0; t < PR::NT; ++t) {
        double h[2];
        CK(cudaMemset(S.dmax, 0, 16));
        maxdiff_kernel<<<1184, 256>>>(S.out[t], S.ref[t], size_t(S.nrows) * NZ, S.dmax);
        CK(cudaMemcpy(h, S.dmax, 16, cudaMemcpyDeviceToHost));
        worst = fmax(worst, h[0] / h[1]);
    }
    const double bytes = double(S.nrows) * NZ * 16.0 * (PR::NF + PR::NT);
    const double cffts = PR::NF / 2 + (PR::NT + 1) / 2;
    printf("%-40s regs %3d  smem %6zu  CTAs/SM %d (%2d warps) grid %4u  %8.3f ms  %7.1f GB/s alg  %6.2f ms per cFFT sweep of 1536^2 rows  x%.2f vs multi  rel diff %.1e\n",
           label, fa.numRegs, smem, occ, occ * NTG * G / 32, grid, ms, bytes / ms / 1e6, ms / cffts * (1536.0 * 1536.0 / S.nrows), multi_ms / ms, worst);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const int planes = argc > 1 ? atoi(argv[1]) : 96;
    const int filter = argc > 2 ? atoi(argv[2]) : 0;
    Setup S;
    S.nrows = long(planes) * N;
    const size_t elems = size_t(S.nrows) * NZ;
    for (int i = 0; i < 12; ++i) { CK(cudaMalloc(&S.f[i], elems * 16)); fill_kernel<<<1184, 256>>>(S.f[i], elems, 77u + i); }
    for (int t = 0; t < 5; ++t) { CK(cudaMalloc(&S.out[t], elems * 16)); CK(cudaMalloc(&S.ref[t], elems * 16)); }
    CK(cudaMalloc(&S.dmax, 16));
    S.tw = make_tw2<PlanZ>();
    CK(cudaDeviceSynchronize());
    double cffts;
    {
        const float ms = run_multi(S, ZF_PROG_C, &cffts);
        const double bytes = double(S.nrows) * NZ * 16.0 * (33 + 4);
        printf("== stage C, %ld rows: multi-product pass, 4 launches  %8.3f ms  %7.1f GB/s alg  %.1f cFFTs per row, %6.2f ms per cFFT sweep of 1536^2 rows\n",
               S.nrows, ms, bytes / ms / 1e6, cffts, ms / cffts * (1536.0 * 1536.0 / S.nrows));
        run_fields<64, 6, ZF_PROG_C, false, false>(S, "fields C: 6 groups x 64, direct loads", ms);
        if (filter == 0) {
            run_fields<128, 3, ZF_PROG_C, true, false>(S, "fields C: 3 groups x 128, TMA rows", ms);
            run_fields<128, 3, ZF_PROG_C, false, false>(S, "fields C: 3 groups x 128, direct loads", ms);
            run_fields<128, 2, ZF_PROG_C, true, false>(S, "fields C: 2 groups x 128, TMA rows", ms);
            run_fields<128, 3, ZF_PROG_C, true, true>(S, "fields C: 3 x 128, TMA, forward teams", ms);
            run_fields<64, 6, ZF_PROG_C, false, true>(S, "fields C: 6 x 64, direct, forward teams", ms);
            run_fields<64, 6, ZF_PROG_C, false, false, Plan<N, 128, 2, 12, 8, 16>>(S, "fields C: 6 x 64, direct, radices 12 8 16", ms);
            run_fields<64, 6, ZF_PROG_C, false, false, Plan<N, 128, 2, 6, 16, 16>>(S, "fields C: 6 x 64, direct, radices 6 16 16", ms);
            run_fields<64, 6, ZF_PROG_C, false, false, Plan<N, 128, 2, 12, 8, 4, 4>>(S, "fields C: 6 x 64, direct, radices 12 8 4 4", ms);
        }
    }
    {
        const float ms = run_multi(S, ZF_PROG_A, &cffts);
        const double bytes = double(S.nrows) * NZ * 16.0 * (16 + 5);
        printf("== stage A, %ld rows: multi-product pass, 5 launches  %8.3f ms  %7.1f GB/s alg  %.1f cFFTs per row, %6.2f ms per cFFT sweep of 1536^2 rows\n",
               S.nrows, ms, bytes / ms / 1e6, cffts, ms / cffts * (1536.0 * 1536.0 / S.nrows));
        run_fields<128, 3, ZF_PROG_A, true, false>(S, "fields A: 3 groups x 128, TMA rows", ms);
        if (filter == 0) {
            run_fields<128, 3, ZF_PROG_A, false, false>(S, "fields A: 3 groups x 128, direct loads", ms);
            run_fields<64, 3, ZF_PROG_A, true, false>(S, "fields A: 3 groups x 64, TMA rows", ms);
            run_fields<64, 3, ZF_PROG_A, false, false>(S, "fields A: 3 groups x 64, direct loads", ms);
        }
    }
    return 0;
}
