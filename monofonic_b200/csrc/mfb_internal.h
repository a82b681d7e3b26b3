// mfb_internal.h - internal types of libmfb200 (not part of the C ABI)
#pragma once
#include <map>
#include <string>
#include <vector>

#include "mfb_platform.h"
#include "../../include/mfb200.h"

struct mfb_grid {
    int n[3];
    double L[3];
    int space;        // MFB_RSPACE / MFB_KSPACE
    double* d;        // device buffer, ntot doubles: real [n0][n1][n2+2] == complex [n0][n1][n2/2+1]
    size_t ntot;
    int nzc;          // n2/2+1
    ptrdiff_t pitch;  // n2+2 doubles per z-row
    // slab decomposition (src/grid_fft.cc:120-128): real space split along axis 0, k-space along axis 1 with the
    // transposed layout [ky_loc][kx][kz].  dist == 0: one slab holding everything, k-layout [kx][ky][kz].
    int dist;
    int nloc0, off0;  // local_0_size_, local_0_start_
    int nloc1, off1;  // local_1_size_, local_1_start_
    // streamed ingest (mfb_grid_stream_in): event recorded on the ingest stream after the H2D copy of the next field
    cudaEvent_t uploaded = nullptr;
    bool upload_pending = false;
};

// Grid_FFT<ccomplex_t> (PPT / Eulerian branch): dense [n0][n1][n2] complex, c2c transforms.  Slabs like mfb_grid: real space
// [x_loc][n1][n2], k-space transposed [ky_loc][n0][n2]
struct mfb_cgrid {
    int n[3];
    double L[3];
    int space;
    double2* d;
    size_t nelem;     // local elements
    int dist = 0;
    int nloc0 = 0, off0 = 0, nloc1 = 0, off1 = 0;
};

// identity of a backward-transformed operand held in one of the convolver's operand buffers
struct mfb_operand_key {
    const void* g; int a1, b1, a2, b2; double c1, c2; int imag, keep_nyq; bool valid;
};

// one U-sized buffer ([P0loc][N1][N2/2] complex) of the slot workspace of the v2 source stage (mfb_sources2.cuh): holds the
// x-transformed spectrum k_x^mx f^ of a source potential ("X buffer", cached by (grid, mx)) or the y-forward output of a target
struct mfb_conv_slot {
    double2* p;
    int owned;          // allocated by the convolver (else: lent particle memory, W1, or a peer-mapped exchange slot)
    int xslot;          // slab mode: index of the exchange slot behind it (-1: none)
    const void* g; int mx; int valid;   // cached X buffer
    int pinned;
    long stamp;
    int lent;           // carved from the lender's particle arrays (a streamed hand-off copy may still be reading it)
    int zb;             // slab mode: z-block height of the exchange layout the buffer was written with (0: plain)
};
constexpr int MFB_CONV_MAX_SLOTS = 16;

struct mfb_particles;
struct mfb_conv {
    int n[3], np[3];
    double L[3];
    double2 *W1, *W2a, *W2b;  // workspaces: [P0][N1][N2/2], [P0][P1][N2/2] x2  
    mfb_operand_key slot[2];  // what W2a / W2b hold (only trusted inside one mfb_lpt_sources call: cache_on)
    bool cache_on;
    bool acc_borrowed;        // Wacc points into the particle arrays of `lender` (mfb_conv_borrow_accumulator)
    mfb_particles* lender;
    double2* Wacc;            // optional [P0][P1][N2/2]: running sum of a product group in z-Fourier space (NULL: no grouping)
    size_t w1_elems, w2_elems;
    int dist, p0loc, p0off;   // slab of the padded grid along axis 0 (real space)
    mfb_grid* fbuf;           // N-grid for the intermediate of 3-field products (convolution.hh:469)
    int naive;                // NaiveConvolver (convolution.hh:296-426): no padding, products on the N-grid itself
    mfb_grid *nb1, *nb2;      // its fbuf1_, fbuf2_
    long nconv;
    // ---- slot workspace of the v2 source stage (mode 1).  mode 0: the buffers above.
    int mode, legacy_ready;
    mfb_conv_slot xs[MFB_CONV_MAX_SLOTS];
    int nxs;
    double2* chunk; size_t chunk_elems; int chunk_owned;   // operand / target chunks of the x-plane loop
    int chunk_lent;                                        // ... carved from the lender's particle arrays
    long stamp;
    int last_path, last_slots, last_loops;                 // what the last mfb_lpt_sources call did (bench / tests)
    long last_xpass, last_ypass, last_zlaunch;
};

struct mfb_particles {
    int n[3];
    int lattice_type, real64, ids64;
    size_t np;        // local number of particles
    int nloc0, off0;  // slab of lattice planes held by this rank
    void *pos, *vel, *ids, *mass;  // device AoS (particle_container.hh:82-108); pos and vel are one allocation (vel = pos + 3 np)
    size_t posvel_bytes;
    struct mfb_conv* borrower = nullptr;  // convolver that currently uses pos/vel as its accumulator (loan ends with either side)
    // streamed hand-off: per attribute, event recorded on the copy stream after its D2H copy (null: none pending)
    cudaEvent_t copied[4] = {nullptr, nullptr, nullptr, nullptr};
    bool copy_pending[4] = {false, false, false, false};
    // masked sc lattice (lattice_type -2): the species this container holds and the 2x2x2 type mask
    int mask_index = 0;
    int mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t np_global = 0;
    // glass load (lattice_type -1): positions in grid units on the device, offset of this rank's particles
    double* glass_pos = nullptr;
    unsigned long long glass_off = 0;
};
// the gradient operator of src/ic_generator.cc:377-381 when it is particle::lattice_gradient (PLT): four tables on the device
struct mfb_gradient {
    int ng, ngmapto;
    double boxlen;
    double* tab[4];
};

namespace mfb {
constexpr int MFB_MAX_XBUF = 2 + MFB_CONV_MAX_SLOTS;
// profiling roles: one per kernel use in the pipeline (bench.py reports time and algorithmic bytes per role)
enum Role {
    ROLE_CONV_X_BWD = 0, ROLE_CONV_Y_BWD, ROLE_CONV_Z, ROLE_CONV_Y_FWD, ROLE_CONV_X_FWD,
    ROLE_FFT_X, ROLE_FFT_Y, ROLE_FFT_Z_R2C, ROLE_FFT_Z_C2R, ROLE_ELEMENTWISE, ROLE_REDUCE, ROLE_PARTICLES,
    ROLE_BARRIER, ROLE_FFT_Z_C2C, ROLE_COUNT
};
// host collective used when peers are not mapped (CPU emulation tests, NCCL baseline): deliver chunk s of
// `send` (P chunks of chunk_bytes) to rank s, receive rank r's chunk into chunk r of `recv`.
struct DistState {
    int rank = 0, nranks = 1;
    bool ready = false, staged = false;
    void* block = nullptr;               // [xbuf0][xbuf1][flags]: one IPC-exported allocation
    size_t xbuf_elems = 0;               // complex elements per exchange buffer
    int nxbuf = 2;                       // exchange buffers in the block: 0,1 ping-pong for the 3-D FFTs, >= 2: slots of the convolver
    double2* xbuf[MFB_MAX_XBUF] = {};
    double2* sendbuf = nullptr;          // staged mode only
    double2* peer_xbuf[MFB_MAX_XBUF][8] = {};
    unsigned long long* flags = nullptr; // flags[r]: last barrier epoch rank r has reached
    unsigned long long* peer_flags[8] = {};
    int* err = nullptr;                  // device flag set when a barrier times out
    unsigned long long epoch = 0;
    long nexch = 0;
    mfb_exchange_fn cb = nullptr;
    void* cb_user = nullptr;
    void* peer_block[8] = {};
};
struct ProfEvent { cudaEvent_t a, b; int role; double bytes; };
struct Context {
    bool initialised = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H hand-off of particle attributes, overlaps the compute stream
    cudaStream_t in_stream = nullptr;    // H2D ingest of the next noise field (mfb_grid_stream_in), the other PCIe direction
    std::map<int, double2*> tw;  // twiddle tables exp(-2 pi i m/n)
    std::map<std::string, double2*> tw2;  // stage-major twiddle tables per plan (Plan::tw2_off)
    double* partial = nullptr;   // reduction scratch
    size_t partial_cap = 0;
    double* table[2] = {nullptr, nullptr};  // radial tables (amplitude, vbc)
    size_t table_cap[2] = {0, 0};
    long long launches = 0;
    int num_sms = 148;
    bool strided_v2 = true;  // compile-time pad / truncate specialisations of the strided pass (fft_kernels2.cuh); 0: general kernel
    bool conv_own_accumulator = true;  // tests: 0 makes mfb_conv_create behave as under memory pressure
    bool conv_grouping = true;  // sum the products of one target before the forward passes (needs mfb_conv::Wacc)
    bool exchange_blocked = true;  // source stage on slabs: z-blocked exchange chunks (contiguous peer stores per tile)
    bool y_pad_tile4 = false;   // padded y passes at n = 1536: two 4-line tiles per SM instead of one 8-line tile (no gain in the pipeline)
    bool fused_readout = true;  // sc loads handed a k-space field: lattice read-out as the epilogue of the c2r pass
    bool zm_prefetch = true;  // multi-product z pass: operand rows staged by TMA bulk copies (0: direct loads)
    bool sources_v2 = true;  // mfb_lpt_sources: shared x passes + multi-product z pass on the slot workspace (mfb_sources2.cuh)
    bool sources_shared = true;  // 3LPT: stages A and C as shared-field loops (z_fields_kernel: every Hessian component transformed once per row)
    bool sources_alias = true;   // ... whose last stage lets the targets overwrite consumed X planes on a single slab (0: needs 10 slots)
    int sources_max_slots = 0;   // tests: cap the slot workspace (0: what HBM allows)
    int sources_chunk = 0;       // tests: cap the x planes per chunk (0: what the chunk workspace allows)
    // optional per-role device timing (cudaEvents on `stream` around every launch)
    bool profiling = false;
    int cur_role = ROLE_ELEMENTWISE;
    double cur_bytes = 0.0;
    std::vector<ProfEvent> pending;
    std::vector<cudaEvent_t> event_pool;
    double role_ms[ROLE_COUNT] = {0};
    double role_bytes[ROLE_COUNT] = {0};
    long long role_count[ROLE_COUNT] = {0};
    DistState dist;
};
// RAII: tags the launches issued in its scope with a role and their algorithmic HBM bytes
struct RoleScope {
    int prev_role; double prev_bytes; bool armed;
    RoleScope(int role, double bytes);
    ~RoleScope();
};
extern Context g_ctx;
void set_error(const std::string& s);
void cuda_check(cudaError_t e, const char* what);
void require(bool c, const char* msg);
void check_launch(const char* what);
void* dev_alloc(size_t bytes);
}  // namespace mfb
