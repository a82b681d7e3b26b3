/* mfb200.h - C ABI of libmfb200: the B200-native LPT field pipeline of monofonIC (MUSIC2).
 *
 * This is the drop-in boundary (SURVEY.md section 8b / DESIGN.md section 2).  The reference has no
 * FFI around this path (header templates in one executable), so every entry point names the
 * reference member function / loop it replaces (paths relative to the reference repo).  The host
 * classes in include/mfb/ (C++) and monofonic_b200/ (Python, ctypes) are thin mirrors of the
 * reference's Grid_FFT / OrszagConvolver / op:: / particle::lattice_generator on top of this ABI.
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; mfb_last_error() gives the message
 *     (the reference throws std::runtime_error / aborts; the C++ shim re-throws).
 *   - FP64 only (CODE_PRECISION=DOUBLE, the reference default, CMakeLists.txt:89-96).
 *   - a grid is stored on the device exactly like the reference's non-distributed Grid_FFT
 *     (src/grid_fft.cc:45-54,73-74): real view [n0][n1][n2+2], complex view [n0][n1][n2/2+1],
 *     in place.  Multi-GPU slabs (mfb_dist_*) use the reference's MPI layout: real slab along
 *     axis 0, k-space slab along axis 1, transposed [ky_loc][kx][kz] (src/grid_fft.cc:120-128).
 *   - FFTs carry the symmetric 1/sqrt(n0 n1 n2) normalisation (src/grid_fft.cc:199-205).
 *   - no CPU fallback: without a CUDA device mfb_init fails.
 */
#ifndef MFB200_H
#define MFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mfb_grid mfb_grid; /* opaque: replaces Grid_FFT<real_t> (include/grid_fft.hh:40-914) */
typedef struct mfb_conv mfb_conv; /* opaque: replaces OrszagConvolver<real_t> (include/convolution.hh:429-669) */
typedef struct mfb_cgrid mfb_cgrid;   /* opaque: replaces Grid_FFT<ccomplex_t> (PPT branch, src/ic_generator.cc:690,778) */
typedef struct mfb_particles mfb_particles; /* opaque: replaces particle::container (include/particle_container.hh:30-189) */

enum { MFB_RSPACE = 1, MFB_KSPACE = 0 };                           /* space_t, include/grid_fft.hh:29 */
enum { MFB_OP_ASSIGN = 0, MFB_OP_ADD = 1, MFB_OP_SUB = 2, MFB_OP_MULADD = 3 }; /* include/operators.hh:25-35 */
enum { MFB_LATTICE_SC = 0, MFB_LATTICE_BCC = 1, MFB_LATTICE_FCC = 2, MFB_LATTICE_RSC = 3,
       MFB_LATTICE_GLASS = -1, MFB_LATTICE_MASKED = -2 }; /* particle_generator.hh:31-39 */
typedef struct mfb_gradient mfb_gradient;

/* ---- library / device ------------------------------------------------------------------ */
int mfb_device_count(void);          /* CUDA devices visible to this process (0: none) */
int mfb_init(int device);            /* selects the CUDA device, creates the stream; fails without a GPU */
int mfb_finalize(void);
int mfb_device_mem_info(size_t* free_bytes, size_t* total_bytes);   /* of the selected device; sizes the exchange block (mfb_dist_init_slots) */
const char* mfb_last_error(void);
const char* mfb_build_info(void);    /* "cuda sm_100a" for the product build */
int mfb_synchronize(void);
void* mfb_stream(void);              /* the cudaStream_t all work is issued on */
void* mfb_copy_stream(void);         /* the cudaStream_t of the streamed particle hand-off (events of a consumer go here) */
long long mfb_kernel_launches(void); /* number of kernels launched so far (bench.py gpu_launches) */
int mfb_supported_fft_size(int n);   /* 1 if an FFT plan exists for line length n */
int mfb_set_option(const char* name, int value); /* tuning switches, e.g. "strided_v2" = 0/1 (kernel generation) */

/* per-role device timing (CUDA events on the library stream around every kernel launch); the reference's
   analogue are the "took : ...s" stage timers (src/ic_generator.cc:470,500,528,540,557) */
int mfb_profile_enable(int on);
int mfb_profile_reset(void);
int mfb_profile_num_roles(void);
const char* mfb_profile_role_name(int role);
int mfb_profile_get(int role, double* ms, long long* count, double* algorithmic_bytes);

/* ---- slab decomposition over the GPUs of one node (one process per GPU) ---------------------------
   Replaces the FFTW-MPI slab machinery (src/grid_fft.cc:120-140 fftw_mpi_local_size_3d_transposed + the
   TRANSPOSED_OUT/IN plans) and the slice-wise Isend/Recv pad/unpad (src/grid_fft.cc:401-485).  After
   mfb_dist_init + mfb_dist_connect every grid / convolver / particle container created by this process is a
   slab: real space split along axis 0, k-space along axis 1 in the transposed layout [ky_loc][kx][kz].
   The exchange of each 3-D FFT is fused into the preceding FFT pass (stores into the peer's buffer, mapped via
   CUDA IPC, over NVLink); the host only has to carry the 64-byte IPC handles between the processes once
   (any transport: torch.distributed, MPI_Allgather, ...). */
size_t mfb_dist_xbuf_bytes(const int n[3], int with_conv, int nranks); /* exchange-buffer size for grids up to n */
int mfb_dist_init(int rank, int nranks, size_t xbuf_bytes, void* ipc_handle_out /* 64 bytes */);
/* the same with `nslots` exchange buffers in the peer-mapped block (2..18; mfb_dist_init == 2, the ping-pong pair of the 3-D
   FFTs).  With >= 2 + 7 (3LPT; 2 + 4 at 2LPT) mfb_lpt_sources keeps the x-transformed spectra of its source potentials
   resident in the extra buffers: 10 exchanges per 3LPT step instead of 49 + 9. */
int mfb_dist_init_slots(int rank, int nranks, size_t xbuf_bytes, int nslots, void* ipc_handle_out /* 64 bytes */);
int mfb_dist_connect(const void* all_handles /* nranks x 64 bytes, rank order */);
/* the same slab algorithm with the exchange done by a host collective instead of peer stores (CPU tests over
   gloo; NCCL all-to-all baseline).  cb(send, recv, chunk_bytes, user): chunk s of send -> rank s; chunk r of
   recv <- rank r. */
typedef void (*mfb_exchange_fn)(const void* send, void* recv, size_t chunk_bytes, void* user);
int mfb_dist_init_staged(int rank, int nranks, size_t xbuf_bytes, mfb_exchange_fn cb, void* user);
int mfb_dist_init_staged_slots(int rank, int nranks, size_t xbuf_bytes, int nslots, mfb_exchange_fn cb, void* user);
int mfb_dist_nslots(void);                /* exchange buffers in the block (0: no slab decomposition) */
int mfb_dist_shutdown(void);
int mfb_dist_rank(void);
int mfb_dist_nranks(void);
long long mfb_dist_exchanges(void);       /* number of slab exchanges performed so far */
int mfb_dist_check(void);                 /* fails if a device barrier timed out (dead peer) */
/* out = {local_0_start_, local_0_size_, local_1_start_, local_1_size_} (include/grid_fft.hh:60-63); returns 1 for a slab */
int mfb_grid_local(const mfb_grid* g, int out[4]);

/* ---- Grid_FFT ---------------------------------------------------------------------------- */
/* Grid_FFT(N, L, allocate=true, initialspace) + allocate()        include/grid_fft.hh:77-83, src/grid_fft.cc:40-197 */
int mfb_grid_create(mfb_grid** g, const int n[3], const double L[3], int space);
int mfb_grid_destroy(mfb_grid* g);                                 /* reset(), include/grid_fft.hh:95-101 */
int mfb_grid_space(const mfb_grid* g);
size_t mfb_grid_ntot(const mfb_grid* g);                           /* ntot_ in doubles, src/grid_fft.cc:45 */
void* mfb_grid_device_ptr(mfb_grid* g);
/* host<->device copies of the whole buffer (ntot doubles) - the coherence primitive under relem/kelem/data_ */
int mfb_grid_upload(mfb_grid* g, const double* host);
/* Streamed ingest (the host -> device side of the hand-off): asynchronous copy of the RNG plugin's next
   field (pinned host memory, raw layout) into `g` on an ingest stream, ordered after all work queued so far; the compute
   stream continues with work that does not touch `g`.  mfb_grid_stream_in_wait makes the compute stream - not the host -
   wait for the copy before `g` is used again.  Replaces the blocking Fill_Grid -> device copy of a series of runs. */
/* Device noise fill in the NGenIC layout (RNG_ngenic::Fill_Grid, src/plugins/random_ngenic.cc:84-154): `g`
   becomes white noise in k-space, unit variance per complex mode, DC and every index-N/2 mode zero, Hermitian k_z = 0 plane
   by the plugin's mirror rule.  Generator: Philox4x32-10 keyed with `seed`, counter = (column, k_z) with the column taken
   from the plugin's ring-ordered seed table, so a larger grid draws the same low-k columns.  NOT gsl ranlxd1: bit-exact
   NGenIC phases stay with the host plugin (GSL is not available offline); this entry point is for runs that only need
   reproducible Gaussian noise, and removes the host fill + H2D copy from them.  Slabs: every rank fills its own k-slab. */
int mfb_noise_fill_ngenic(mfb_grid* g, unsigned long long seed);
/* ... and with the reference's own generator - RANLUX, GSL's gsl_rng_ranlxd1 (:44-45) - one stream per (k_x, k_y) column seeded
   from the plugin's seed table: the NGenIC white noise of the reference itself for this [random] seed (checked against its golden
   files), without the host fill and the H2D copy */
int mfb_noise_fill_ngenic_ranlxd(mfb_grid* g, unsigned long long seed);
int mfb_grid_stream_in(mfb_grid* g, const double* host);
int mfb_grid_stream_in_wait(mfb_grid* g);
int mfb_grid_download(const mfb_grid* g, double* host);
/* FourierTransformForward/Backward(do_transform)                  src/grid_fft.cc:208-270 */
int mfb_fft_forward(mfb_grid* g, int do_transform);
int mfb_fft_backward(mfb_grid* g, int do_transform);
int mfb_grid_zero(mfb_grid* g);                                    /* zero(), include/grid_fft.hh:150-155 */
int mfb_grid_copy(mfb_grid* dst, const mfb_grid* src);             /* copy_from(), :157-177 */
int mfb_grid_scale(mfb_grid* g, double a);                         /* operator*=, :392-403 */
int mfb_grid_axpy(mfb_grid* y, double a, const mfb_grid* x, double b); /* y = a*y + b*x (same space) */
int mfb_grid_affine(mfb_grid* g, double a, double b);              /* x <- a + b*x in real space (src/ic_generator.cc:673-675) */
int mfb_grid_zero_dc(mfb_grid* g);                                 /* zero_DC_mode(), :844-887 (both spaces) */
int mfb_grid_inverse_laplacian(mfb_grid* g);                       /* apply_InverseLaplacian(), :440-449 */
int mfb_grid_laplacian(mfb_grid* g, int negative);                 /* apply_Laplacian / apply_negative_Laplacian, :418-438 */
int mfb_grid_shift_field(mfb_grid* g, const double s[3], int transform_back); /* shift_field(), :832-842 */
int mfb_grid_mean(mfb_grid* g, double* out);                       /* mean(), :555-595 */
int mfb_grid_std(mfb_grid* g, double* out);                        /* std(), :507-553 */
int mfb_grid_2norm(mfb_grid* g, double* out);                      /* compute_2norm(), :485-505 (real space) */
int mfb_grid_max_neighbour_diff(mfb_grid* g, double* out);         /* src/ic_generator.cc:713-735 */
int mfb_fourier_resample(mfb_grid* src, mfb_grid* dst);            /* FourierInterpolateCopyTo(), src/grid_fft.cc:273-492 */

/* ---- pipeline k-space operators (src/ic_generator.cc) ------------------------------------- */
/* wnoise <- (+-1) wnoise/volfac [, w/|w| fixing][, corner-mode cut]            :341-358 */
int mfb_noise_condition(mfb_grid* w, double volfac, int fix, int invert, int remove_corners);
/* out = in * A(|k|) * (inv_k2 ? -1/k^2 : 1), DC=0.  table[m] = A(kfac*sqrt(m)), m = |n|^2 = 0..3(N/2)^2:
   the transfer plugin is evaluated on the host once per distinct |k|           :461-468, :667-670 */
int mfb_apply_radial_table(mfb_grid* out, const mfb_grid* in, const double* table, size_t ntable, int inv_k2);
/* phi2 -= (sum(lambda) k^2 + 4/3 sum lambda_d k_d^2) phi                        :492-496 */
int mfb_add_tides(mfb_grid* phi2, const mfb_grid* phi, const double lambda[3]);
/* out = G_d (w1 phi + w2 phi2 + w3 phi3) + wc (G_p A_q - G_q A_p), * scale, DC=0 (k-space)
   displacement: w=(1,1,1,1), scale=lunit/L  :841-870;  velocity: w=vfac*(1,2,3,3), scale=vunit/L  :896-941.
   phi2/phi3/A3 may be NULL (lower LPT order).  vbc_* add  - cvbc*T(|k|)*W*G_d/k^2  (:918-921) when non-NULL. */
/* The gradient operator `lg` of src/ic_generator.cc:377-381 when the reference is built with ENABLE_PLT:
   particle::lattice_gradient (include/particle_plt.hh:50-587).  The host computes the eigen-decomposition of the lattice's
   dynamical matrix on an ng^3 helper mesh (GSL, once per run - stays host code); the device needs the four k-space tables
   its gradient() / vfac_corr() read: tables[0..3] = Re of D_xx_, D_yy_, D_zz_ (radial / theta / phi eigenvalue) and D_xy_
   (growth-rate correction), each [ng][ng][ng/2+1] doubles.  ngmapto = GridRes.  Single slab, like the reference
   (Grid_FFT::get_cic_kspace "doesn't work with MPI"). */
int mfb_gradient_create_plt(mfb_gradient** lg, int ng, int ngmapto, double boxlen, const double* const tables[4]);
int mfb_gradient_destroy(mfb_gradient* lg);
/* mfb_lpt_assemble with the gradient operator made explicit.  lg == NULL: op::fourier_gradient.  velocity != 0: the mode is
   multiplied by lg.vfac_corr (src/ic_generator.cc:929; 1 for the Fourier gradient).  cic_compensation != 0: glass loads,
   grid_interpolate::compensation_kernel (:861, :924; include/grid_interpolate.hh:195-208). */
int mfb_lpt_assemble_lg(mfb_grid* out, int d, const mfb_grid* phi, const mfb_grid* phi2, const mfb_grid* phi3,
                        mfb_grid* const A3[3], double w1, double w2, double w3, double wc, double scale, const mfb_grid* wnoise,
                        const double* vbc_table, size_t ntable, double cvbc, const mfb_gradient* lg, int velocity,
                        int cic_compensation);
int mfb_lpt_assemble(mfb_grid* out, int d, const mfb_grid* phi, const mfb_grid* phi2, const mfb_grid* phi3,
                     mfb_grid* const A3[3], double w1, double w2, double w3, double wc, double scale,
                     const mfb_grid* wnoise, const double* vbc_table, size_t ntable, double cvbc);
/* delta = k^2 phi  (first-order SPT density, :962-973) */
int mfb_density_from_phi(mfb_grid* out, const mfb_grid* phi);

/* ---- Grid_FFT<ccomplex_t> and the PPT / Eulerian branch (src/ic_generator.cc:685-803) ------------------------
   Dense [n0][n1][n2] complex grids with c2c transforms and the same symmetric norm (src/grid_fft.cc:199-270). Single slab. */
int mfb_cgrid_create(mfb_cgrid** g, const int n[3], const double L[3]);
int mfb_cgrid_destroy(mfb_cgrid* g);
int mfb_cgrid_space(const mfb_cgrid* g);
size_t mfb_cgrid_nelem(const mfb_cgrid* g);     /* local elements of the buffer (slabs: the larger of the two layouts) */
/* slab decomposition of a complex grid, as mfb_grid_local: out = {local_0_start, local_0_size, local_1_start, local_1_size};
   real space [x_loc][n1][n2], k-space transposed [ky_loc][n0][n2] (src/grid_fft.cc:130-133).  Returns 1 when distributed. */
int mfb_cgrid_local(const mfb_cgrid* g, int out[4]);
int mfb_cgrid_upload(mfb_cgrid* g, const double* host_interleaved, int space);
int mfb_cgrid_download(const mfb_cgrid* g, double* host_interleaved);
int mfb_cgrid_copy(mfb_cgrid* dst, const mfb_cgrid* src);                    /* copy_from */
int mfb_cgrid_fft_forward(mfb_cgrid* g, int do_transform);                   /* FourierTransformForward (c2c) */
int mfb_cgrid_fft_backward(mfb_cgrid* g, int do_transform);
int mfb_cgrid_gradient(mfb_cgrid* g, int idim);                              /* x *= i k_d, k from get_k (:790-792) */
int mfb_ppt_sqrt_density(mfb_grid* rho, double C_species);                   /* rho <- sqrt(1 + C rho) (:703-705) */
/* psi <- rho * exp(i (phi [+ phi2]) / (hbar D)), all real-space (:741-751); phi2 may be NULL (LPTorder 1) */
int mfb_ppt_init_psi(mfb_cgrid* psi, const mfb_grid* rho, const mfb_grid* phi, const mfb_grid* phi2, double hbar, double Dplus0);
int mfb_ppt_drift(mfb_cgrid* psi, double hbar, double Dplus0);               /* k-space: psi *= exp(-i hbar k^2 D / 2) (:757-760) */
int mfb_ppt_kick(mfb_cgrid* psi, const mfb_grid* phi2, double hbar, double Dplus0); /* psi *= exp(i phi2/(hbar D)) (:763-767) */
int mfb_ppt_density(mfb_grid* rho, const mfb_cgrid* psi);                    /* rho <- |psi|^2 - 1 (:772-775) */
/* out <- vunit * hbar * Im(conj(psi) grad_psi) / (1 + rho)  (:796-799) */
int mfb_ppt_velocity(mfb_grid* out, const mfb_cgrid* psi, const mfb_cgrid* grad_psi, const mfb_grid* rho, double hbar, double vunit);

/* ---- OrszagConvolver ------------------------------------------------------------------------ */
int mfb_conv_create(mfb_conv** c, const int n[3], const double L[3]);      /* include/convolution.hh:463-495 */
/* NaiveConvolver (include/convolution.hh:296-426; the reference selects it at compile time with CONVOLVER_TYPE=NAIVE):
   same named products below on the same handle type, but without padding - aliased products on the N-grid itself,
   Nyquist planes included, three-field products multiplied in one go.  Single slab only. */
int mfb_conv_create_naive(mfb_conv** c, const int n[3], const double L[3]);
int mfb_conv_destroy(mfb_conv* c);
size_t mfb_conv_workspace_bytes(const mfb_conv* c);
/* 1 when the optional accumulator workspace ([P0][P1][N2/2] complex, taken at creation only if HBM is plentiful) exists:
   mfb_lpt_sources then sums the products of one target in z-Fourier space before ONE forward transform */
int mfb_conv_grouped(const mfb_conv* c);
/* When HBM was too tight for an own accumulator (3LPT 2048^3 on 8 GPUs), lend the convolver the position/velocity arrays of
   a particle container as accumulator: they hold nothing while the sources are computed (src/ic_generator.cc:475-574 runs
   before the species loop :832-955).  p == NULL returns the loan.  The caller keeps `p` alive while the loan lasts. */
int mfb_conv_borrow_accumulator(mfb_conv* c, mfb_particles* p);
/* Which path the last mfb_lpt_sources call took.  1: one padded transform pair per product (OrszagConvolver::convolve2,
   include/convolution.hh:512-535, grouped per target when an accumulator exists).  2: shared x passes + x-plane chunks +
   multi-product z pass (needs >= 7 [2LPT: 4] slots of 1.5 R(N) each, own HBM plus lent particle arrays).
   info = {slots, x-plane loops, backward x passes, backward y passes, z launches, 0}. */
int mfb_conv_sources_info(const mfb_conv* c, long long info[6]);
/* All products: inputs are k-space N-grids, `out` is a k-space N-grid updated with `op` (alpha for MULADD).
   dims are axis pairs {a,b}: the operand is (i k_a)(i k_b) f^ with the gradient zeroed at Nyquist. */
int mfb_conv_hessians(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* r, const int d2r[2],
                      mfb_grid* out, int op, double alpha);                 /* convolve_Hessians 2-field :137-160 */
int mfb_conv_hessians3(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* m, const int d2m[2], mfb_grid* r,
                       const int d2r[2], mfb_grid* out, int op, double alpha); /* 3-field :172-203 (= two convolve2, :542-549) */
int mfb_conv_sum_hessians(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* r, const int d2r1[2],
                          const int d2r2[2], mfb_grid* out, int op, double alpha); /* convolve_SumOfHessians :214-239 */
int mfb_conv_diff_hessians(mfb_conv* c, mfb_grid* l, const int d2l[2], mfb_grid* r, const int d2r1[2],
                           const int d2r2[2], mfb_grid* out, int op, double alpha); /* convolve_DifferenceOfHessians :250-274 */
int mfb_conv_gradients(mfb_conv* c, mfb_grid* l, int d1l, mfb_grid* r, int d1r, mfb_grid* out, int op,
                       double alpha);                                       /* convolve_Gradients :75-96 */
int mfb_conv_gradient_hessian(mfb_conv* c, mfb_grid* l, int d1l, mfb_grid* r, const int d2r[2], mfb_grid* out,
                              int op, double alpha);                        /* convolve_Gradient_and_Hessian :106-127 */
int mfb_conv_multiply(mfb_conv* c, mfb_grid* l, mfb_grid* r, mfb_grid* out, int op, double alpha); /* multiply_field :282-292 */
/* OrszagConvolver::convolve2 / convolve3 with caller-formed spectra (:512-535, :542-549; NaiveConvolver :312-401): the reference
   takes host lambdas kfunc(i,j,k); here the operands are k-space device grids holding kfunc's values (the C++ mirror
   evaluates the lambdas into temporaries, include/mfb/convolution.hh).  convolve3 truncates between the two products. */
int mfb_conv_generic(mfb_conv* c, mfb_grid* specA, mfb_grid* specB, mfb_grid* out, int op, double alpha);
int mfb_conv_generic3(mfb_conv* c, mfb_grid* specA, mfb_grid* specB, mfb_grid* specC, mfb_grid* out, int op, double alpha);
/* The whole source stage of ic_generator::run (:475-574): phi2, phi3, A3 from phi (un-scaled, k-space), then the
   growth-factor scaling g1,g2,g3,g3c.  phi2/phi3/A3 may be NULL below the requested order. */
int mfb_lpt_sources(mfb_conv* c, int order, double Dplus0, mfb_grid* phi, mfb_grid* phi2, mfb_grid* phi3,
                    mfb_grid* const A3[3], const double* tides_lambda);

/* ---- particle::lattice_generator -------------------------------------------------------------- */
/* container::allocate (particle_container.hh:60-80): device AoS buffers pos[3*np], vel[3*np], ids[np], mass[np] */
int mfb_particles_create(mfb_particles** p, const int n[3], int lattice_type, int real64, int ids64, int with_masses);
int mfb_particles_destroy(mfb_particles* p);
size_t mfb_particles_count(const mfb_particles* p);
/* lattice_generator ctor, Bravais branch (particle_generator.hh:129-170) */
int mfb_particles_set_ids(mfb_particles* p, unsigned long long id_offset_species);
/* set_positions / set_velocities / set_masses, Bravais branch (:422-467, :520-562, :292-350).
   `field` is a real-space grid; it is phase-shifted in place exactly like the reference. */
int mfb_particles_set_positions(mfb_particles* p, int lattice_index, int idim, double lunit, mfb_grid* field);
int mfb_particles_set_velocities(mfb_particles* p, int lattice_index, int idim, mfb_grid* field);
int mfb_particles_set_masses(mfb_particles* p, int lattice_index, double munit, mfb_grid* field);
/* Masked sc lattice (lattice_type -2, particle_generator.hh:171-260): mask[4a+2b+c] is the species (0, 1 or -1 = nobody)
   of the cells with parity (i%2, j%2, k%2) = (a,b,c), as set up from ParticleMaskType (:184-205); the container holds the
   cells of species `lattice_index` in (i,j,k) order.  The set_* calls above then take the masked branches (:351-398 masses
   with get_mean_mask_value, :479-497 positions, :557-571 velocities); set_ids writes IDoffset*N^3 + cell index (:236-243).
   Grid sizes must be even (slabs: even local sizes). */
int mfb_particles_create_masked(mfb_particles** p, const int n[3], int lattice_index, const int mask[8], int real64, int ids64,
                                int with_masses);
/* Glass load (lattice_type -1, :262-290): `posr` = glass_loader::glass_posr, np positions in grid units [np][3] (host), what
   the tiled glass file gives; IDs are IDoffset + i + glass_offset; positions pos/n*lunit + CIC(field, pos), velocities
   CIC(field, pos) (:498-515, :573-590; grid_interpolate::get_cic_at, include/grid_interpolate.hh:103-148).  The component
   fields of a glass load carry the CIC compensation kernel (mfb_lpt_assemble_lg).  Single slab. */
int mfb_particles_create_glass(mfb_particles** p, const int n[3], const double* posr, size_t np, unsigned long long glass_offset,
                               int real64, int ids64);
/* device -> host copies into the container vectors the output plugins read (get_*_ptr, particle_container.hh:110-148) */
int mfb_particles_download(const mfb_particles* p, void* pos, void* vel, void* ids, void* mass);
/* Streamed hand-off to the output plugin (replaces the host-side fill of the container that write_particle_data then
   reads, src/ic_generator.cc:832-955): enqueue the device -> host copy of ONE attribute (MFB_ATTR_*) on a separate copy
   stream, ordered after everything queued so far on the compute stream, so that it overlaps the kernels of the next
   component.  `host` should be page-locked.  A later set_* on the same attribute waits for the copy on the device.
   mfb_particles_stream_wait blocks the host until all streamed copies of `p` have landed. */
#define MFB_ATTR_POS 0
#define MFB_ATTR_VEL 1
#define MFB_ATTR_IDS 2
#define MFB_ATTR_MASS 3
int mfb_particles_stream_out(mfb_particles* p, int attr, void* host);
/* particles [first, first + count) of one attribute: for hosts that hand chunks to their consumer through a ring of page-locked
   buffers instead of holding the whole container (3LPT 2048^3: 275 GB of particles) */
int mfb_particles_stream_out_range(mfb_particles* p, int attr, size_t first, size_t count, void* host);
int mfb_particles_stream_wait(mfb_particles* p);

#ifdef __cplusplus
}
#endif
#endif /* MFB200_H */
