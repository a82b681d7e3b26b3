// cuda_emu.h - TEST INFRASTRUCTURE ONLY (see monofonic_b200/csrc/mfb_platform.h).
//
// A tiny CUDA-on-CPU shim: each CUDA thread of a block is an OS thread, __syncthreads() is a
// real barrier, blocks run one after the other.  Just enough of the runtime API to compile
// monofonic_b200/csrc/*.cu with g++ -DMFB_EMU -x c++ so that kernel index logic can be
// validated against the oracle without a GPU.  Slow by design; only used at 32^3-ish sizes.
#pragma once
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(x) alignas(x)
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_emu { unsigned x, y, z; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
using std::min;
using std::max;
using std::abs;
inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

namespace mfb_emu {
// Every block runs on one OS worker thread; its CUDA threads are ucontext fibers scheduled
// round-robin, __syncthreads() yields to the scheduler (=> a full round == one barrier).
// Blocks are distributed over the workers; threadIdx/blockIdx/shared memory are per worker.
extern thread_local uint3_emu t_threadIdx;
extern thread_local uint3_emu t_blockIdx;
extern thread_local unsigned char* t_dyn_smem;
extern dim3 g_blockDim, g_gridDim;
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void sync_threads();
void bar_sync(int id, int count);

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F body) {
    run_grid(grid, block, smem, std::function<void()>(body));
}
}  // namespace mfb_emu

#define threadIdx (mfb_emu::t_threadIdx)
#define blockIdx (mfb_emu::t_blockIdx)
#define blockDim (mfb_emu::g_blockDim)
#define gridDim (mfb_emu::g_gridDim)
inline void __syncthreads() { mfb_emu::sync_threads(); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline void sincospi(double x, double* s, double* c) { *s = std::sin(M_PI * x); *c = std::cos(M_PI * x); }
#ifndef __GLIBC__
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
#endif
inline double atomicAdd(double* p, double v) {
    static std::mutex m; std::lock_guard<std::mutex> lk(m); double o = *p; *p += v; return o;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    static std::mutex m; std::lock_guard<std::mutex> lk(m); auto o = *p; *p += v; return o;
}
inline double atomicMaxDouble_emu(double* p, double v) {
    static std::mutex m; std::lock_guard<std::mutex> lk(m); double o = *p; if (v > o) *p = v; return o;
}

#define MFB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(mfb_emu::t_dyn_smem)
#define MFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    mfb_emu::launch(dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })
#define MFB_KERNEL_SMEM_OPTIN(kernel, bytes) (cudaSuccess)

// ---- minimal runtime API -------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = size_t(64) << 30; return 0; }
