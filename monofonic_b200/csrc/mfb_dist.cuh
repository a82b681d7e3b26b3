// mfb_dist.cuh - slab-decomposed (multi-GPU) layer of libmfb200: one process per GPU.
//
// Decomposition = the reference's FFTW-MPI slabs (src/grid_fft.cc:120-128): real space split along axis 0
// (local_0_start_/size_), k-space along axis 1 (local_1_start_/size_) in the transposed layout
// [ky_loc][kx][kz].  The ONE exchange per 3-D FFT (FFTW's internal all-to-all) is fused into the
// FFT pass that precedes it: the strided pass writes every output 128-byte segment straight into the owning
// rank's receive buffer (peer memory mapped through CUDA IPC, stores travel over NVLink/NVSwitch), so the
// transposed data never exists as a separate send buffer and the transfer overlaps the butterflies of the
// other tiles.  A device-side flag barrier (release/acquire at system scope through the same peer mapping)
// closes the exchange.  Inside a dealiased product the exchange sits between the x and y passes and moves the
// pruned 1.5 R(N)/P intermediate, not R(3N/2)/P.
//
// When peers are not mapped ("staged" mode: CPU emulation tests over gloo, or an NCCL all-to-all baseline) the
// same kernels write into a local staging buffer laid out by destination and a host collective delivers it.
#pragma once
#include "mfb_internal.h"

namespace mfb {

struct Exchange {
    double2* recv;                   // where this rank finds the received data: [src rank][chunk]
    double2* base[MFB_MAX_RANKS];    // where to write the chunk destined to rank s
    double2* send;                   // staged mode: the staging buffer the chunks were written to
    size_t chunk;                    // complex elements per (src,dst) pair
};

#ifndef MFB_EMU
struct PeerFlags { unsigned long long* p[MFB_MAX_RANKS]; };
// thread t publishes this rank's epoch to rank t, then waits until rank t has published the same epoch here.
__global__ void dist_barrier_kernel(PeerFlags peers, unsigned long long* mine, int rank, int nranks,
                                    unsigned long long epoch, int* err) {
    const int t = threadIdx.x;
    if (t >= nranks) return;
    __threadfence_system();  // the preceding kernel's peer stores are performed before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.p[t] + rank), "l"(epoch) : "memory");
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + t) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 60000000000ull) { *err = 1; break; }  // 60 s: a peer died; do not hang the GPU
        __nanosleep(200);
    }
    __threadfence_system();
}
#endif

inline DistState& D() { return g_ctx.dist; }
inline bool dist_on() { return g_ctx.dist.ready && g_ctx.dist.nranks > 1; }

inline Exchange begin_exchange(size_t chunk_elems) {
    DistState& d = D();
    require(d.ready, "distributed layer not initialised (mfb_dist_init)");
    require(chunk_elems * size_t(d.nranks) <= d.xbuf_elems, "exchange buffer too small for this grid (mfb_dist_xbuf_bytes)");
    Exchange x;
    const int e = int(d.nexch & 1);
    x.recv = d.xbuf[e];   // buffers 0 and 1 ping-pong
    x.chunk = chunk_elems;
    x.send = d.staged ? d.sendbuf + size_t(e) * d.xbuf_elems : nullptr;
    for (int s = 0; s < MFB_MAX_RANKS; ++s) x.base[s] = nullptr;
    for (int s = 0; s < d.nranks; ++s)
        x.base[s] = d.staged ? x.send + size_t(s) * chunk_elems : d.peer_xbuf[e][s] + size_t(d.rank) * chunk_elems;
    return x;
}

// Exchange into a NAMED buffer of the block (k >= 2: the slots the v2 source stage keeps x-transformed spectra and
// target intermediates in).  The caller orders reuse: between the last read of a slot and the next peer write into it at
// least one barrier must lie in program order (mfb_sources2.cuh does that with one barrier per x-plane loop).
inline Exchange slot_exchange(int k, size_t chunk_elems) {
    DistState& d = D();
    require(d.ready, "distributed layer not initialised (mfb_dist_init)");
    require(k >= 0 && k < d.nxbuf, "internal: exchange slot out of range");
    require(chunk_elems * size_t(d.nranks) <= d.xbuf_elems, "exchange buffer too small for this grid (mfb_dist_xbuf_bytes)");
    Exchange x;
    x.recv = d.xbuf[k];
    x.chunk = chunk_elems;
    x.send = d.staged ? d.sendbuf + size_t(k) * d.xbuf_elems : nullptr;   // staged mode: one staging buffer per slot
    for (int s = 0; s < MFB_MAX_RANKS; ++s) x.base[s] = nullptr;
    for (int s = 0; s < d.nranks; ++s)
        x.base[s] = d.staged ? x.send + size_t(s) * chunk_elems : d.peer_xbuf[k][s] + size_t(d.rank) * chunk_elems;
    return x;
}

// Closes an exchange: afterwards x.recv holds [src][chunk] on every rank.  Ping-pong between the two receive
// buffers makes ONE barrier per exchange sufficient as long as the consumer of exchange k is enqueued before the
// producer of exchange k+1 (stream order), which holds for every caller in this library.
inline void end_exchange(const Exchange& x) {
    DistState& d = D();
    if (d.staged) {
        cuda_check(cudaStreamSynchronize(g_ctx.stream), "sync before host exchange");
        d.cb(x.send, x.recv, x.chunk * sizeof(double2), d.cb_user);
    } else {
#ifndef MFB_EMU
        ++d.epoch;
        PeerFlags pf;
        for (int s = 0; s < MFB_MAX_RANKS; ++s) pf.p[s] = d.peer_flags[s];
        RoleScope rs(ROLE_BARRIER, 0.0);
        dist_barrier_kernel<<<1, 32, 0, g_ctx.stream>>>(pf, d.flags, d.rank, d.nranks, d.epoch, d.err);
        check_launch("dist_barrier_kernel");
#else
        throw std::runtime_error("peer-mapped exchange is not available in the emulation build");
#endif
    }
    ++d.nexch;
}

// sum of (a, b) over all ranks, identical bits on every rank (fixed summation order)
__global__ void dist_put2_kernel(Exchange x, int nranks, const double* __restrict__ src) {
    const int t = threadIdx.x;
    if (t < nranks) x.base[t][0] = make_double2(src[0], src[1]);
}
__global__ void dist_sum2_kernel(const double2* __restrict__ recv, int nranks, double* __restrict__ out) {
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int r = 0; r < nranks; ++r) { a += recv[r].x; b += recv[r].y; }
        out[0] = a; out[1] = b;
    }
}
inline void dist_allreduce_sum2(double* dev_pair) {
    DistState& d = D();
    Exchange x = begin_exchange(1);
    MFB_LAUNCH(dist_put2_kernel, 1, 32, 0, g_ctx.stream, x, d.nranks, dev_pair);
    check_launch("dist_put2");
    end_exchange(x);
    MFB_LAUNCH(dist_sum2_kernel, 1, 32, 0, g_ctx.stream, x.recv, d.nranks, dev_pair);
    check_launch("dist_sum2");
}

// max of a scalar over all ranks (MPI_Allreduce(MAX) of src/ic_generator.cc:730-733)
__global__ void dist_max1_kernel(const double2* __restrict__ recv, int nranks, double* __restrict__ out) {
    if (threadIdx.x == 0) {
        double a = recv[0].x;
        for (int r = 1; r < nranks; ++r) a = fmax(a, recv[r].x);
        out[0] = a;
    }
}
inline void dist_allreduce_max1(double* dev_pair) {   // dev_pair[0] in/out, dev_pair[1] scratch
    DistState& d = D();
    Exchange x = begin_exchange(1);
    MFB_LAUNCH(dist_put2_kernel, 1, 32, 0, g_ctx.stream, x, d.nranks, dev_pair);
    check_launch("dist_put2");
    end_exchange(x);
    MFB_LAUNCH(dist_max1_kernel, 1, 32, 0, g_ctx.stream, x.recv, d.nranks, dev_pair);
    check_launch("dist_max1");
}

}  // namespace mfb
