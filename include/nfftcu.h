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
/* window family: the reference fixes it at configure time (--with-window, include/infft.h:146-222); here it is a
 * create-time flag OR-ed into `flags` (bits the reference's plan flags do not use) */
#define NFFTCU_WINDOW_KAISER_BESSEL 0
#define NFFTCU_WINDOW_GAUSSIAN 1
#define NFFTCU_FLAG_GAUSSIAN (1u << 30)   /* Gaussian window (include/infft.h:154-173) instead of Kaiser-Bessel */
#define NFFTCU_MAX_PEERS 8            /* GPUs of one NVSwitch domain that can share an adjoint reduction */
#define NFFTCU_PEER_HANDLE_BYTES 192  /* opaque per-rank blob of nfftcu_peer_export (three CUDA IPC handles) */

typedef struct nfftcu_ctx_s nfftcu_ctx;

/* options for nfftcu_set_option */
#define NFFTCU_OPT_TIMING 1        /* 1: record CUDA-event stage times (D,F,B) per transform */
#define NFFTCU_OPT_PSI_TABLE 2     /* 1: keep a per-node window table in sorted order (PRE_PSI) */
#define NFFTCU_OPT_B_KERNEL 3      /* 0 auto | 1 generic gather/scatter | 2 register pencils | 3 DMMA */
#define NFFTCU_OPT_B_FLUSH 5       /* DMMA spreading: 0 auto | 1 RED.ADD from registers | 2 staged TMA bulk reductions */
#define NFFTCU_OPT_FFT_PRUNE 6     /* 1 (default): band-pruned FFT passes and D without zero padding inside trafo/adjoint | 0: full passes */
#define NFFTCU_OPT_FFT_KERNEL 7    /* 0 auto (register-resident Stockham for 2^k lengths 64..2048) | 1 shared-memory Stockham only */
#define NFFTCU_OPT_WINDOW_IMAGES 8 /* 3-D tensor kernels: per-batch window images built with the node set and fed by TMA: 0 auto (when they fit) | 1 off | 2 on */
#define NFFTCU_OPT_SLAB_FFT 9      /* 0 auto: when the nodes occupy a slab of the first axis (a rank of a node-sharded run), the pruned FFT passes and the B^T memset touch only the slab's planes | 1 off */
#define NFFTCU_OPT_TC5 10          /* fp32 plans, d = 3, m <= 6: B / B^T on tcgen05.mma kind::tf32 with the grid window and the accumulators in tensor memory (tc5.cu): 0 auto (both) | 1 off (mma.sync TF32 kernels) | 2 B only | 3 B and B^T */
#define NFFTCU_OPT_NODE_ORDER 4    /* 0 auto | 1 reference row-major key | 2 tile-binned */

const char *nfftcu_last_error(void);

/* debugging aid of the tcgen05 kernels (tc5.cu): 32 cycle totals of the timing probes (env NFFT_B200_TC5_DBG & 8), cleared on read */
int nfftcu_tc5_debug(unsigned long long *out32);
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

/* Units of the internal grid (single-stage calls only; nfftcu_trafo / nfftcu_adjoint are unaffected): the device
 * window values of dimension t carry an exact power-of-two factor s_t (1 for fp64 plans; for fp32 plans
 * 2^-round(log2 phi_hat_t(0)), which keeps the fp32 grid from overflowing where the reference's does) and c_t carries
 * 1/s_t.  With S = prod s_t: stage_D writes g_ref / S, stage_B returns S * B g, stage_BT writes S * B^T f, stage_DT
 * returns D^T g / S.  scale_host receives s_t, t < d, as doubles. */
int nfftcu_get_window_scale(nfftcu_ctx *ctx, double *scale_host);
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
 * x on every call (nfft.c:4889, 5351): the transform starts at once with the resident nodes while host threads
 * compute a 64-bit fingerprint of x_host; only when it differs from the fingerprint of the array the resident
 * nodes were uploaded from are the nodes uploaded, re-sorted and the transform repeated (no PCIe traffic for
 * unchanged nodes).  NFFT_B200_EXACT_NODE_CHECK=1, or resident nodes that came from nfftcu_set_nodes_dev, select the
 * exact variant: x_host is uploaded on a side stream and compared word for word on the device.  *changed (may be
 * NULL) reports whether the nodes were replaced, i.e. whether index_x has to be fetched again. */
int nfftcu_trafo_refresh(nfftcu_ctx *ctx, const void *x_host, const void *f_hat_host, void *f_host, int *changed);
int nfftcu_adjoint_refresh(nfftcu_ctx *ctx, const void *x_host, const void *f_host, void *f_hat_host, int *changed);
/* Split-phase variants (no reference counterpart): *_begin enqueues the H2D copy, the transform and the D2H copy on the
 * plan's stream and returns at once; nfftcu_end waits for them.  With page-locked host buffers (nfft_malloc) two plans --
 * e.g. the trafo plan and the adjoint plan of an iteration on independent data -- overlap their copies with each other's
 * kernels; bench.py reports that as e2e.overlapped.  The resident nodes are used as they are; the host buffers must not
 * be touched between begin and end. */
int nfftcu_trafo_begin(nfftcu_ctx *ctx, const void *f_hat_host, void *f_host);
int nfftcu_adjoint_begin(nfftcu_ctx *ctx, const void *f_host, void *f_hat_host);
int nfftcu_end(nfftcu_ctx *ctx);
/* nfft_trafo_direct / nfft_adjoint_direct (nfft.c:145-297): exact NDFT */
int nfftcu_trafo_direct(nfftcu_ctx *ctx, const void *f_hat_host, void *f_host);
int nfftcu_adjoint_direct(nfftcu_ctx *ctx, const void *f_host, void *f_hat_host);

/* ---- transforms, device buffers (benchmark / multi-GPU driver) ----------------------------
 * Same operations on device pointers of the plan's device; asynchronous on the plan's stream. */
int nfftcu_trafo_dev(nfftcu_ctx *ctx, const void *f_hat_dev, void *f_dev);
int nfftcu_adjoint_dev(nfftcu_ctx *ctx, const void *f_dev, void *f_hat_dev);
int nfftcu_trafo_direct_dev(nfftcu_ctx *ctx, const void *f_hat_dev, void *f_dev);
int nfftcu_adjoint_direct_dev(nfftcu_ctx *ctx, const void *f_dev, void *f_hat_dev);

/* ---- batched transforms: K right-hand sides on one node set --------------------------------
 * No reference API exists for this; the consumer is the per-coil loop of applications/mri/mri2d/
 * reconstruct_data_2d.c:52-139 (every coil: the same trajectory, its own data).  Layout: f_hat [K][N_total],
 * f [K][M], interleaved complex.  The node-dependent state (sort, tile binning, window images) is built once per
 * node set and shared by all K; D, the FFT passes, the 2-D tile kernels and D^T process the K right-hand sides in
 * one launch each (right-hand side = one more grid dimension). */
int nfftcu_trafo_batch(nfftcu_ctx *ctx, int K, const void *f_hat_host, void *f_host);
int nfftcu_adjoint_batch(nfftcu_ctx *ctx, int K, const void *f_host, void *f_hat_host);
int nfftcu_trafo_batch_dev(nfftcu_ctx *ctx, int K, const void *f_hat_dev, void *f_dev);
int nfftcu_adjoint_batch_dev(nfftcu_ctx *ctx, int K, const void *f_dev, void *f_hat_dev);

/* f_dst = A_dst ( b .* (A_src^H f_src) ) with the intermediate coefficients kept in HBM: the far field of
 * applications/fastsum/fastsum.c:1196-1220 (nfft_adjoint(&mv1); mv2.f_hat[k] = b[k] * mv1.f_hat[k]; nfft_trafo(&mv2)),
 * which on the host-pointer API moves f_hat across PCIe three times.  src and dst: two plans with the same N on
 * the same device (source nodes x, target nodes y); b_host: N_total complex (NULL = no multiply). */
int nfftcu_adjoint_mul_trafo(nfftcu_ctx *src, nfftcu_ctx *dst, const void *f_src_host, const void *b_host,
                             void *f_dst_host);

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
int nfftcu_set_stream(nfftcu_ctx *ctx, void *cuda_stream);   /* cudaStream_t; NULL = own (non-blocking) stream; the legacy default stream is cudaStreamLegacy = (void*) 1 */
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

/* Roofline denominators measured on `device` right now (peaks.cu): FP64 tensor-core rate (mma.sync m8n8k4 f64, the
 * pipe the fp64 B / B^T kernels are bound by), legacy mma.sync m16n8k8 TF32 rate (fp32 plans), and a 1 GiB device copy
 * (read + write bytes).  Any pointer may be NULL.  bench.py calls this in-process and records the values it used. */
int nfftcu_measure_peaks(int device, double *fp64_tensor_tflops, double *tf32_mma_sync_tflops, double *copy_gbs);

/* device / pinned memory for C callers that do not link the CUDA runtime themselves */
int nfftcu_malloc_device(void **ptr, size_t bytes, int device);
int nfftcu_free_device(void *ptr);
int nfftcu_malloc_pinned(void **ptr, size_t bytes);
int nfftcu_free_pinned(void *ptr);
/* backing store of nfft_malloc / nfft_free (kernel/util/malloc.c): buffers of NFFT_B200_PINNED_MALLOC_MIN bytes
 * (default 256 KiB; 0 = never) or more are page-locked and portable across devices, smaller ones (and every request
 * when no CUDA device exists) come from the C heap; nfftcu_host_free accepts both kinds.  Returns NULL when out of memory. */
void *nfftcu_host_alloc(size_t bytes);
void nfftcu_host_free(void *ptr);
/* device and page-locked buffers freed by finalized plans are cached per process (mempool.cu, NFFT_B200_POOL_MB);
 * this returns the cache to the driver, for processes that share the GPU with other CUDA users */
void nfftcu_pool_trim(void);
/* the 64-bit fingerprint of a host array that decides whether nodes have changed (host threads; independent of the
 * thread count: 1 MiB chunks hashed separately and combined in order) */
uint64_t nfftcu_fingerprint(const void *data, size_t bytes);
int nfftcu_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes);
int nfftcu_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes);

/* ---- multi-GPU, node-sharded (north_star configs[3]) ------------------------------------------------
 * The reference is a single-node CPU code; these entry points are what a multi-GPU build of nfft_trafo /
 * nfft_adjoint sits on.  Nodes are sorted once by the reference key (nfft.c:75-109) and cut into equal-count
 * slabs of the sorted order, one per GPU, so that every GPU works on a compact slab of the grid.
 *
 * (1) nfftcu_get_sorted_slab: positions [begin, end) of the reference-sorted node list of a plan -- the nodes
 *     (x_out_dev, (end-begin)*d reals) and their original indices (perm_out_dev) -- copied into device buffers of the
 *     plan's device.  A process-per-GPU driver calls it on a plan holding the whole node set to obtain its slab.
 * (2) fused D^T + cross-GPU reduction over NVLink peer memory (peer.cu): nfftcu_peer_export writes an opaque blob of
 *     NFFTCU_PEER_HANDLE_BYTES describing this rank's grid / exchange buffer / flags; the caller gathers the blobs of
 *     all ranks (rank order) and passes them to nfftcu_peer_attach.  nfftcu_adjoint_dev_peer then computes
 *     f_hat = sum_r D^T F^H B_r^T f_r with the sum taken INSIDE the D^T kernel through peer loads; every rank
 *     receives the complete f_hat.  All ranks must call it the same number of times (device-side flag barriers).
 * (3) nfftcu_group_*: ONE process driving several GPUs behind the host-pointer interface of nfft_trafo /
 *     nfft_adjoint (what libnfft3_b200.so uses when NFFT_B200_DEVICES lists more than one device): x is sorted on the
 *     first device and distributed as slabs; trafo replicates f_hat, runs D+F redundantly and B per slab; adjoint
 *     runs B^T+F per slab and the fused D^T+reduce, every device delivering its slice of f_hat.  The permutation
 *     between caller order and slab order of f is an all-to-all over peer memory, so that every device moves only
 *     M/P samples over its own host link. */
int nfftcu_get_sorted_slab(nfftcu_ctx *ctx, int64_t begin, int64_t end, void *x_out_dev, uint32_t *perm_out_dev);
int nfftcu_peer_export(nfftcu_ctx *ctx, void *handles);
int nfftcu_peer_attach(nfftcu_ctx *ctx, int rank, int world, const void *all_handles);
int nfftcu_peer_detach(nfftcu_ctx *ctx);
int nfftcu_adjoint_dev_peer(nfftcu_ctx *ctx, const void *f_dev, void *f_hat_dev);
int nfftcu_peer_reduce_only(nfftcu_ctx *ctx, void *f_hat_dev);   /* the fused D^T + reduce alone (profiling) */
int nfftcu_peer_error(nfftcu_ctx *ctx);   /* 1: a flag barrier timed out waiting for a peer, results are invalid */

typedef struct nfftcu_group_s nfftcu_group;
int nfftcu_group_create(nfftcu_group **out, int precision, int d, const int64_t *N, const int64_t *n, int64_t m,
                        int64_t M, unsigned flags, const int *devices, int ndevices);
int nfftcu_group_destroy(nfftcu_group *g);
int nfftcu_group_set_nodes(nfftcu_group *g, const void *x_host);          /* returns after the slabs are resident */
int64_t nfftcu_group_nodes_version(nfftcu_group *g);
int nfftcu_group_get_index_x(nfftcu_group *g, int64_t *index_x_host);    /* the reference's index_x, 2*M entries */
int nfftcu_group_trafo(nfftcu_group *g, const void *f_hat_host, void *f_host);
int nfftcu_group_adjoint(nfftcu_group *g, const void *f_host, void *f_hat_host);
/* the same with the unannounced node refresh of nfftcu_trafo_refresh (fingerprint of x_host) */
int nfftcu_group_trafo_refresh(nfftcu_group *g, const void *x_host, const void *f_hat_host, void *f_host, int *changed);
int nfftcu_group_adjoint_refresh(nfftcu_group *g, const void *x_host, const void *f_host, void *f_hat_host, int *changed);
int nfftcu_group_direct(nfftcu_group *g, int adjoint, const void *in_host, void *out_host);   /* exact NDFT, device 0 */
int nfftcu_group_size(nfftcu_group *g);
nfftcu_ctx *nfftcu_group_ctx(nfftcu_group *g, int rank);                 /* the per-device plan (options, timing) */
/* milliseconds of the last group transform: [0] host->device copies, [1] device compute incl. the peer exchange,
 * [2] device->host copies (wall clock of the slowest device each) */
int nfftcu_group_times(nfftcu_group *g, float ms[3]);

/* ---- field-inhomogeneity transforms kept on the device ----------------------------------------
 * Replace the host loops of kernel/mri/mri.c: mri_inh_2d1d_trafo 57-103 / _adjoint 105-150 (N3 + 1 two-dimensional
 * NFFTs with cexp / PHI_HUT scaling of f_hat and PHI-weighted accumulation of f between them) on a 2-D plan, and
 * mri_inh_3d_trafo 197-228 / _adjoint 230-255 (one 3-D NFFT, the window applied along the third frequency axis
 * before and 1/PHI_HUT(N3, N3 x_j2) after) on a 3-D plan with N[2] = N3.  Double-precision plans only, window =
 * the reference's window_funct_plan (Kaiser-Bessel, n = N3, b = pi (2 - 1/sigma3), m = the plan's m).
 * w: N_total (2d1d) resp. N[0]*N[1] (3d) doubles; t: M doubles; x_host: the plan's nodes (M x 3); adjoint = 0: in =
 * f_hat, out = f; adjoint = 1: in = f, out = f_hat.  f_scaled_host (3d adjoint, may be NULL) receives f / PHI_HUT,
 * which the reference leaves in that->f. */
int nfftcu_mri_inh_2d1d(nfftcu_ctx *ctx, int adjoint, int N3, double sigma3, const double *w_host,
                        const double *t_host, const void *in_host, void *out_host);
int nfftcu_mri_inh_3d(nfftcu_ctx *ctx, int adjoint, int N3, double sigma3, const double *w_host,
                      const double *x_host, const void *in_host, void *out_host, void *f_scaled_host);

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
/* K right-hand sides iterated in lock-step on ONE plan (multi-coil reconstruction: the coils of
 * applications/mri/mri2d/reconstruct_data_2d.c:52-139 share the trajectory): vectors are [K][M] / [K][N_total], the
 * weights w / w_hat are shared, scal is [K][8]; the transforms of a step are nfftcu_*_batch_dev.  Every right-hand side
 * follows exactly the iteration of a K = 1 solver (own alpha / beta). */
int nfftcu_solver_create_batch(nfftcu_solver **out, nfftcu_ctx *plan, unsigned flags, int K);
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
