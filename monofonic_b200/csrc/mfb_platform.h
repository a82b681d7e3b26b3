// mfb_platform.h - build-mode switch for the monofonIC-B200 kernels.
//
// Product build: nvcc, sm_100a.  All kernels are plain CUDA.
//
// Test-only build (-DMFB_EMU, g++): the very same kernel sources are compiled against
// tests/emu/cuda_emu.h, which runs every CUDA thread of a block as an OS thread with a
// real barrier for __syncthreads().  It exists so that the index arithmetic of the kernels
// can be checked against the oracle in the GPU-less build container (`pytest -m "not gpu"`).
// It is NOT a fallback: the Python package and the C++ host shims only ever load
// libmfb200.so (the nvcc build) and fail loudly when it is missing.
#pragma once

#ifdef MFB_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define MFB_DYN_SMEM(type, name)                                     \
    extern __shared__ __align__(16) unsigned char name##_raw_[];     \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define MFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MFB_KERNEL_SMEM_OPTIN(kernel, bytes) \
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif

#include <cstddef>
#include <cstdint>
#include <cstring>

#define MFB_HD __host__ __device__ __forceinline__
#define MFB_D __device__ __forceinline__

// streaming (read-once / write-once) global accesses: keep them out of L1 so that the
// twiddle table stays cached there.
#if !defined(MFB_EMU)
MFB_D double2 mfb_ld_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
MFB_D void mfb_st_stream(double2* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
MFB_D double mfb_ld_stream1(const double* p) {
    double v;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// 16-byte asynchronous global -> shared copy (LDGSTS): no register staging, the thread is not stalled
MFB_D void mfb_cp_async16(double2* smem_dst, const double2* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
// ---- TMA bulk copies (cp.async.bulk, SASS: UBLKCP) global -> shared with mbarrier completion: one thread issues, nobody's
// registers or LSU slots are spent on the data, the copy runs under whatever the CTA computes meanwhile
MFB_D void mfb_mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
MFB_D void mfb_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
MFB_D void mfb_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
MFB_D void mfb_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
MFB_D void mfb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
MFB_D void mfb_cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
MFB_D void mfb_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
MFB_D void mfb_cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }   // all but the newest group
#else
// emulation: the "asynchronous" copy is done by the issuing thread on the spot; the kernels put a barrier between issue and use
inline void mfb_mbar_init(unsigned long long*, unsigned) {}
inline void mfb_mbar_expect_tx(unsigned long long*, unsigned) {}
inline void mfb_bulk_g2s(void* d, const void* s, unsigned bytes, unsigned long long*) { memcpy(d, s, bytes); }
inline void mfb_mbar_wait(unsigned long long*, unsigned) {}
inline void mfb_fence_proxy_async() {}
inline void mfb_cp_async_commit() {}
inline void mfb_cp_async_wait_but_one() {}
inline void mfb_cp_async16(double2* d, const double2* s) { *d = *s; }
inline void mfb_cp_async_wait_all() {}
inline double2 mfb_ld_stream(const double2* p) { return *p; }
inline void mfb_st_stream(double2* p, double2 v) { *p = v; }
inline double mfb_ld_stream1(const double* p) { return *p; }
#endif
