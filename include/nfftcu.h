/*
 * nfftcu.h -- the drop-in boundary: C ABI of the B200 (sm_100a) NFFT engine, libnfftcu.so.
 *
 * Plain C, plain pointers and sizes, no CUDA or torch types in any signature.  This is the
 * layer a maintainer of NFFT3 binds in place of the bodies of the plan functions in
 * kernel/nfft/nfft.c; the host layer shipped here (nfft_b200/csrc/nfft3_host.c ->
 * libnfft3_b200.so) does exactly that and re-exports the reference's own symbols
 * nfft_init_guru / nfft_precompute_one_psi / nfft_trafo / nfft_adjoint / nfft_finalize (and the
 * nfftf_ twins, include/nfft3.h:163-187 of the reference) on top of the entry points below.
 * INTEGRATION.md shows the binding.
 *
 * Every function returns 0 on success and a negative NFFTCU_E* code on failure;
 * nfftcu_last_error() returns the message (thread-local).  The reference has no error codes:
 * its convention is nfft_die(msg) -> die_hook -> exit (kernel/util/malloc.c), which the host
 * layer reproduces by passing nfftcu_last_error() to nfft_die.
 *
 * Conventions (identical to the reference):
 *   real  R = double (NFFTCU_DOUBLE) | float (NFFTCU_FLOAT);  complex = interleaved (re,im)
 *   x[j*d+t] in [-1/2,1/2), j<M;   f[j], j<M;   f_hat[k], row-major over k_t+N_t/2;
 *   oversampled grid g[l], row-major over l_t in [0,n_t).
 * There is no CPU fallback anywhere behind this header: without a CUDA device every compute
 * entry point fails with NFFTCU_ENODEV.
 */
#ifndef NFFTCU_H
#define NFFTCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFFTCU_DOUBLE 0
#define NFFTCU_FLOAT 1

#define NFFTCU_OK 0
#define NFFTCU_EINVAL (-1)   /* bad argument */
#define NFFTCU_ENODEV (-2)   /* no usable CUDA device */
#define NFFTCU_ECUDA (-3)    /* CUDA runtime error, see nfftcu_last_error() */
#define NFFTCU_ENOMEM (-4)
#define NFFTCU_ESTATE (-5)   /* call order violated (e.g. transform before set_nodes) */

#define NFFTCU_MAX_D 8

typedef struct nfftcu_ctx_s nfftcu_ctx;

/* options for nfftcu_set_option */
#define NFFTCU_OPT_TIMING 1        /* 1: record CUDA-event stage times (D,F,B) per transform */
#define NFFTCU_OPT_PSI_TABLE 2     /* 1: keep a per-node window table in sorted order (PRE_PSI) */
#define NFFTCU_OPT_B_KERNEL 3      /* 0 auto | 1 generic gather/scatter | 2 register pencils | 3 DMMA */
#define NFFTCU_OPT_B_FLUSH 5       /* DMMA spreading: 0 auto | 1 RED.ADD from registers | 2 staged TMA bulk reductions */
#define NFFTCU_OPT_FFT_PRUNE 6     /* 1 (default): band-pruned FFT passes and D without zero padding inside trafo/adjoint | 0: full passes */
#define NFFTCU_OPT_FFT_KERNEL 7    /* 0 auto (register-resident Stockham for 2^k lengths 64..2048) | 1 shared-memory Stockham only */
#define NFFTCU_OPT_WINDOW_IMAGES 8 /* 3-D tensor kernels: per-batch window images built with the node set and fed by TMA: 0 auto (when they fit) | 1 off | 2 on */
#define NFFTCU_OPT_NODE_ORDER 4    /* 0 auto | 1 reference row-major key | 2 tile-binned */

const char *nfftcu_last_error(void);
int nfftcu_device_count(void);

/* ---- plan life cycle --------------------------------------------------------------------
 * replaces init_help (kernel/nfft/nfft.c:5950-6046): copies d,N,n,m, computes
 * b_t = pi(2-1/sigma_t) (include/infft.h:216-222) and c_t[k] = 1/phi_hat_t(k)
 * (precompute_phi_hut, nfft.c:5754-5770) and allocates the oversampled grid (the g1/g2 of
 * FFTW_INIT, nfft.c:6012-6017) and the FFT twiddles (the fftw_plan_dft pair, nfft.c:6030-6031)
 * on `device`.  `flags` are the reference's plan flags (include/nfft3.h:195-208). */
int nfftcu_create(nfftcu_ctx **out, int precision, int d, const int64_t *N, const int64_t *n,
                  int64_t m, int64_t M, unsigned flags, int device);
/* replaces the device side of nfft_finalize (nfft.c:6209-6270) */
int nfftcu_destroy(nfftcu_ctx *ctx);

/* c_phi_inv[t] (N_t reals, host) for plans with PRE_PHI_HUT: plan member c_phi_inv */
int nfftcu_get_c_phi_inv(nfftcu_ctx *ctx, int t, void *out_host);
/* b[t], sigma[t] as the reference stores them in the plan (R-typed, host) */
int nfftcu_get_window_params(nfftcu_ctx *ctx, void *b_host, void *sigma_host);

/* ---- nodes ------------------------------------------------------------------------------
 * replaces sort0/sort (nfft.c:75-123) + nfft_sort_node_indices_radix_lsdf
 * (kernel/util/sort.c:91-167) and the node-dependent part of nfft_precompute_one_psi /
 * precompute_psi (nfft.c:5819-5844, 5938-5948): uploads x (M*d reals), builds the
 * reference's sort key floor(n_t*x_jt - m) mod n_t, sorts stably, stores nodes in sorted
 * order and (optionally) the per-node window table. */
int nfftcu_set_nodes(nfftcu_ctx *ctx, const void *x_host);
int nfftcu_set_nodes_dev(nfftcu_ctx *ctx, const void *x_dev);
/* Plans without a node-bound psi flag get no notification when the caller changes x; the
 * reference simply re-sorts on every transform (nfft.c:4889, 5351).  nfftcu_set_nodes compares
 * the uploaded nodes with the resident ones on the device and redoes the sort only when they
 * differ; this counter increments whenever it did (0 = no nodes yet). */
int64_t nfftcu_nodes_version(nfftcu_ctx *ctx);
/* the NFFT_SORT_NODES witness: index_x[2k] = key, index_x[2k+1] = original node index
 * (2*M int64, host) -- bit-exact against the reference permutation */
int nfftcu_get_index_x(nfftcu_ctx *ctx, int64_t *index_x_host);

/* ---- transforms, host buffers (what nfft_trafo / nfft_adjoint bind) -----------------------
 * nfft_trafo (nfft.c:5655-5701): f = B F D f_hat;  nfft_adjoint (5703-5749): f_hat = D^T F^H B^T f.
 * Copies in, runs on the plan's stream, copies out, returns after completion.  Falls back to
 * the exact NDFT kernels when any N_t <= m or n_t <= 2m+2, as the reference does (5658-5664). */
int nfftcu_trafo(nfftcu_ctx *ctx, const void *f_hat_host, void *f_host);
int nfftcu_adjoint(nfftcu_ctx *ctx, const void *f_host, void *f_hat_host);
/* The same with an unannounced node refresh, for plans without a psi flag, where the reference re-reads
 * x on every call (nfft.c:4889, 5351): x_host is uploaded and compared with the resident nodes on a side
 * stream while the transform already runs with the resident nodes; only when they differ are the nodes
 * re-sorted and the transform repeated.  *changed (may be NULL) reports whether that happened, i.e.
 * whether index_x has to be fetched again. */
int nfftcu_trafo_refresh(nfftcu_ctx *ctx, const void *x_host, const void *f_hat_host, void *f_host, int *changed);
int nfftcu_adjoint_refresh(nfftcu_ctx *ctx, const void *x_host, const void *f_host, void *f_hat_host, int *changed);
/* nfft_trafo_direct / nfft_adjoint_direct (nfft.c:145-297): exact NDFT */
int nfftcu_trafo_direct(nfftcu_ctx *ctx, const void *f_hat_host, void *f_host);
int nfftcu_adjoint_direct(nfftcu_ctx *ctx, const void *f_host, void *f_hat_host);

/* ---- transforms, device buffers (benchmark / multi-GPU driver) ----------------------------
 * Same operations on device pointers of the plan's device; asynchronous on the plan's stream. */
int nfftcu_trafo_dev(nfftcu_ctx *ctx, const void *f_hat_dev, void *f_dev);
int nfftcu_adjoint_dev(nfftcu_ctx *ctx, const void *f_dev, void *f_hat_dev);
int nfftcu_trafo_direct_dev(nfftcu_ctx *ctx, const void *f_hat_dev, void *f_dev);
int nfftcu_adjoint_direct_dev(nfftcu_ctx *ctx, const void *f_dev, void *f_hat_dev);

/* ---- single stages on the plan's internal grid (kernel-level parity tests, profiling) ------
 * D   (nfft.c:5415-5513): grid := zero-padded, fftshifted f_hat * c
 * F   (nfft.c:5516/5557): grid := DFT(grid), sign -1 forward / +1 backward, unnormalised
 * B   (nfft.c:4687-4914): f := interpolate(grid)       BT (5126-5384): grid := spread(f)
 * DT  (nfft.c:5560-5650): f_hat := grid(corners) * c */
int nfftcu_stage_D(nfftcu_ctx *ctx, const void *f_hat_dev);
int nfftcu_stage_F(nfftcu_ctx *ctx, int sign);
int nfftcu_stage_B(nfftcu_ctx *ctx, void *f_dev);
int nfftcu_stage_BT(nfftcu_ctx *ctx, const void *f_dev);
int nfftcu_stage_DT(nfftcu_ctx *ctx, void *f_hat_dev);
void *nfftcu_grid_ptr(nfftcu_ctx *ctx);           /* device pointer, n_total complex */

/* ---- plumbing ----------------------------------------------------------------------------- */
int nfftcu_set_option(nfftcu_ctx *ctx, int option, int64_t value);
int nfftcu_set_stream(nfftcu_ctx *ctx, void *cuda_stream);   /* cudaStream_t; NULL = own stream */
void *nfftcu_get_stream(nfftcu_ctx *ctx);
int nfftcu_sync(nfftcu_ctx *ctx);
/* milliseconds of the last transform's D, F, B(^T) stages -> plan member MEASURE_TIME_t
 * (include/nfft3.h:139); valid when NFFTCU_OPT_TIMING is on */
int nfftcu_stage_times(nfftcu_ctx *ctx, float ms[3]);
/* duration of the main interpolation / spreading kernel launch of the last transform (CUDA events on the
 * plan's stream; 0 when NFFTCU_OPT_TIMING is off or the generic kernels ran): the roofline numerator of bench.py */
int nfftcu_b_kernel_time(nfftcu_ctx *ctx, float *ms);
/* number of kernels this context has launched since creation (bench.py "gpu_launches") */
int64_t nfftcu_launch_count(nfftcu_ctx *ctx);

/* device / pinned memory for C callers that do not link the CUDA runtime themselves */
int nfftcu_malloc_device(void **ptr, size_t bytes, int device);
int nfftcu_free_device(void *ptr);
int nfftcu_malloc_pinned(void **ptr, size_t bytes);
int nfftcu_free_pinned(void *ptr);
int nfftcu_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int nfftcu_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);

/* ---- device-resident inverse-NFFT iterations --------------------------------------------------------------
 * Replaces the host loops of kernel/solver/solver.c (solver_before_loop_complex 81-125, solver_loop_one_step_complex
 * 347-360 with LANDWEBER 128-174, STEEPEST_DESCENT 177-229, CGNR 232-292, CGNE 295-344) and the vector kernels of
 * kernel/util/vector1.c-vector3.c they call, for an mv plan that is an NFFT plan of this library: all vectors stay
 * in HBM, alpha / beta are computed on the device, one synchronisation per step.  `flags` are the reference's solver
 * flags (include/nfft3.h:823-829).  The reference-facing wrappers solver_*_complex / solverf_*_complex of
 * libnfft3_b200.so sit on top of these. */
typedef struct nfftcu_solver_s nfftcu_solver;
#define NFFTCU_SOLVER_Y 0            /* M complex: right-hand side                      (solver_plan_complex.y) */
#define NFFTCU_SOLVER_W 1            /* M reals, PRECOMPUTE_WEIGHT                      (.w) */
#define NFFTCU_SOLVER_W_HAT 2        /* N_total reals, PRECOMPUTE_DAMP                  (.w_hat) */
#define NFFTCU_SOLVER_F_HAT_ITER 3   /* N_total complex: iterate                        (.f_hat_iter) */
#define NFFTCU_SOLVER_R_ITER 4       /* M complex: residual                             (.r_iter) */
#define NFFTCU_SOLVER_Z_HAT_ITER 5   /* N_total complex; aliases P_HAT_ITER unless CGNR (.z_hat_iter) */
#define NFFTCU_SOLVER_P_HAT_ITER 6   /* N_total complex: search direction               (.p_hat_iter) */
#define NFFTCU_SOLVER_V_ITER 7       /* M complex, CGNR / STEEPEST_DESCENT              (.v_iter) */
/* scal[8] = alpha_iter, beta_iter, dot_r_iter, dot_r_iter_old, dot_z_hat_iter, dot_z_hat_iter_old, dot_p_hat_iter,
 * dot_v_iter (the scalar members of solver_plan_complex, include/nfft3.h:772-779), as doubles */
int nfftcu_solver_create(nfftcu_solver **out, nfftcu_ctx *plan, unsigned flags);
int nfftcu_solver_destroy(nfftcu_solver *s);
int nfftcu_solver_upload(nfftcu_solver *s, int which, const void *host);      /* host -> device vector */
int nfftcu_solver_download(nfftcu_solver *s, int which, void *host);          /* device vector -> host */
void *nfftcu_solver_vector(nfftcu_solver *s, int which);                      /* device pointer */
/* r = y - A f_hat_iter, z_hat = A^H (w r), initial norms.  f_hat_iter_host / r_iter_host (may be NULL) receive the
 * host mirrors of the iterate and the residual; scal receives the scalars. */
int nfftcu_solver_before_loop(nfftcu_solver *s, void *f_hat_iter_host, void *r_iter_host, double scal[8]);
/* one iteration.  LANDWEBER reads scal[0] (alpha_iter is set by the caller in the reference too). */
int nfftcu_solver_step(nfftcu_solver *s, void *f_hat_iter_host, void *r_iter_host, double scal[8]);

#ifdef __cplusplus
}
#endif
#endif /* NFFTCU_H */
