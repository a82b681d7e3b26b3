// mfb_internal.h - internal types of libmfb200 (not part of the C ABI)
#pragma once
#include <map>
#include <string>

#include "mfb_platform.h"

struct mfb_grid {
    int n[3];
    double L[3];
    int space;        // MFB_RSPACE / MFB_KSPACE
    double* d;        // device buffer, ntot doubles: real [n0][n1][n2+2] == complex [n0][n1][n2/2+1]
    size_t ntot;
    int nzc;          // n2/2+1
    ptrdiff_t pitch;  // n2+2 doubles per z-row
};

struct mfb_conv {
    int n[3], np[3];
    double L[3];
    double2 *W1, *W2a, *W2b;  // workspaces: [P0][N1][N2/2], [P0][P1][N2/2] x2  
    size_t w1_elems, w2_elems;
    mfb_grid* fbuf;           // N-grid for the intermediate of 3-field products (convolution.hh:469)
    long nconv;
};

struct mfb_particles {
    int n[3];
    int lattice_type, real64, ids64;
    size_t np;
    void *pos, *vel, *ids, *mass;  // device AoS (particle_container.hh:82-108)
};

namespace mfb {
struct Context {
    bool initialised = false;
    int device = -1;
    cudaStream_t stream = nullptr;
    std::map<int, double2*> tw;  // twiddle tables exp(-2 pi i m/n)
    double* partial = nullptr;   // reduction scratch
    size_t partial_cap = 0;
    double* table[2] = {nullptr, nullptr};  // radial tables (amplitude, vbc)
    size_t table_cap[2] = {0, 0};
    long long launches = 0;
};
extern Context g_ctx;
void set_error(const std::string& s);
void cuda_check(cudaError_t e, const char* what);
void require(bool c, const char* msg);
void check_launch(const char* what);
void* dev_alloc(size_t bytes);
}  // namespace mfb
