// CUDA-on-CPU test shim: fiber scheduler (TEST INFRASTRUCTURE ONLY, see cuda_emu.h)
#include "cuda_emu.h"

#include <ucontext.h>

#include <memory>

namespace mfb_emu {
thread_local uint3_emu t_threadIdx;
thread_local uint3_emu t_blockIdx;
thread_local unsigned char* t_dyn_smem = nullptr;
dim3 g_blockDim, g_gridDim;

namespace {
constexpr size_t STACK = 256 * 1024;
struct Worker {
    ucontext_t main_ctx;
    std::vector<ucontext_t> ctx;
    std::vector<std::unique_ptr<char[]>> stacks;
    std::vector<char> done;
    const std::function<void()>* body = nullptr;
    int cur = 0;
    unsigned alive = 0;
    // barriers of the running block: 0 is __syncthreads() (every live thread), 1..15 are PTX named barriers with an explicit
    // thread count (bar.sync id, n).  The last arrival opens the barrier, waiting fibers yield until the generation changes.
    struct Bar { unsigned arrived = 0, gen = 0; } bars[16];
};
thread_local Worker* t_worker = nullptr;

void trampoline() {
    Worker* w = t_worker;
    (*w->body)();
    w->done[w->cur] = 1;
    swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

void run_block(Worker& w, dim3 block, const std::function<void()>& body) {
    const unsigned nthr = block.x * block.y * block.z;
    if (w.ctx.size() < nthr) {
        w.ctx.resize(nthr);
        while (w.stacks.size() < nthr) w.stacks.emplace_back(new char[STACK]);
    }
    w.done.assign(nthr, 0);
    w.body = &body;
    w.alive = nthr;
    for (auto& b : w.bars) b = Worker::Bar();
    for (unsigned t = 0; t < nthr; ++t) {
        getcontext(&w.ctx[t]);
        w.ctx[t].uc_stack.ss_sp = w.stacks[t].get();
        w.ctx[t].uc_stack.ss_size = STACK;
        w.ctx[t].uc_link = &w.main_ctx;
        makecontext(&w.ctx[t], trampoline, 0);
    }
    unsigned remaining = nthr;
    while (remaining) {
        for (unsigned t = 0; t < nthr; ++t) {
            if (w.done[t]) continue;
            w.cur = int(t);
            t_threadIdx.x = t % block.x;
            t_threadIdx.y = (t / block.x) % block.y;
            t_threadIdx.z = t / (block.x * block.y);
            swapcontext(&w.main_ctx, &w.ctx[t]);
            if (w.done[t]) {
                --remaining;
                --w.alive;
                // threads that have left the kernel no longer take part in __syncthreads()
                Worker::Bar& b = w.bars[0];
                if (b.arrived > 0 && b.arrived >= w.alive) { b.arrived = 0; ++b.gen; }
            }
        }
    }
}
}  // namespace

void sync_threads() {
    Worker* w = t_worker;
    Worker::Bar& b = w->bars[0];
    const unsigned g = b.gen;
    if (++b.arrived >= w->alive) { b.arrived = 0; ++b.gen; return; }
    while (b.gen == g) swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

void bar_sync(int id, int count) {
    Worker* w = t_worker;
    Worker::Bar& b = w->bars[id & 15];
    const unsigned g = b.gen;
    if (++b.arrived == unsigned(count)) { b.arrived = 0; ++b.gen; return; }
    while (b.gen == g) swapcontext(&w->ctx[w->cur], &w->main_ctx);
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    g_blockDim = block;
    g_gridDim = grid;
    const size_t nblocks = size_t(grid.x) * grid.y * grid.z;
    unsigned nw = std::thread::hardware_concurrency();
    if (nw == 0) nw = 4;
    if (nw > 16) nw = 16;
    if (nw > nblocks) nw = unsigned(nblocks ? nblocks : 1);
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Worker w;
        t_worker = &w;
        std::vector<unsigned char> dyn(smem + 128);
        t_dyn_smem = dyn.data() + (128 - (reinterpret_cast<uintptr_t>(dyn.data()) & 127)) % 128;
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            t_blockIdx.x = unsigned(b % grid.x);
            t_blockIdx.y = unsigned((b / grid.x) % grid.y);
            t_blockIdx.z = unsigned(b / (size_t(grid.x) * grid.y));
            run_block(w, block, body);
        }
        t_worker = nullptr;
        t_dyn_smem = nullptr;
    };
    std::vector<std::thread> pool;
    for (unsigned i = 1; i < nw; ++i) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}
}  // namespace mfb_emu
