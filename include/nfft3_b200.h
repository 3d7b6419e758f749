/*
 * nfft3_b200.h -- the NFFT3 plan API as exported by libnfft3_b200.so (double: nfft_*, single:
 * nfftf_*; nfftl_ long double is not provided).
 *
 * This header is layout- and signature-compatible with the reference's include/nfft3.h
 * (MACRO_MV_PLAN at nfft3.h:54-60, the plan body at 109-161, prototypes at 163-187, flags at
 * 195-208, the malloc/die hooks at 69-82).  Code written against the reference's own
 * <nfft3.h> (kernel/solver/solver.c, kernel/mri/mri.c, applications/fastsum, ...) links
 * against libnfft3_b200.so unchanged; this header exists so that the host layer and new users
 * can be compiled without FFTW's header.  tests/test_abi.py checks every member offset against
 * the reference header when it is available.
 *
 * Members documented "device" are not host arrays in this implementation:
 *   my_fftw_plan1  holds the nfftcu_ctx* of the plan (include/nfftcu.h), my_fftw_plan2 is NULL
 *   g, g_hat, g1, g2, psi, psi_index_g, psi_index_f, spline_coeffs are NULL (the oversampled
 *   grid and the window table live in HBM)
 * c_phi_inv, b, sigma, N, n, index_x and the MALLOC_* buffers are host arrays as in the reference.
 */
#ifndef NFFT3_B200_H
#define NFFT3_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
typedef double nfft_b200_cdouble[2];
typedef float nfft_b200_cfloat[2];
#else
#include <complex.h>
typedef double _Complex nfft_b200_cdouble;
typedef float _Complex nfft_b200_cfloat;
#endif

typedef ptrdiff_t NFFT_INT;

#define NFFT_B200_DEFINE_API(X, R, C)                                                            \
  typedef void *(*X(malloc_type_function))(size_t n);                                            \
  typedef void (*X(free_type_function))(void *p);                                                \
  typedef void (*X(die_type_function))(const char *errString);                                   \
  extern X(malloc_type_function) X(malloc_hook);                                                 \
  extern X(free_type_function) X(free_hook);                                                     \
  extern X(die_type_function) X(die_hook);                                                       \
  void *X(malloc)(size_t n);                                                                     \
  void X(free)(void *p);                                                                         \
  void X(die)(const char *s);                                                                    \
                                                                                                 \
  typedef struct {                                                                               \
    NFFT_INT N_total;          /* number of Fourier coefficients */                              \
    NFFT_INT M_total;          /* number of nodes */                                             \
    C *f_hat;                  /* Fourier coefficients, host */                                  \
    C *f;                      /* samples, host */                                               \
    void (*mv_trafo)(void *);                                                                    \
    void (*mv_adjoint)(void *);                                                                  \
  } X(mv_plan_complex);                                                                          \
                                                                                                 \
  typedef struct {                                                                               \
    NFFT_INT N_total;                                                                            \
    NFFT_INT M_total;                                                                            \
    C *f_hat;                                                                                    \
    C *f;                                                                                        \
    void (*mv_trafo)(void *);                                                                    \
    void (*mv_adjoint)(void *);                                                                  \
    NFFT_INT d;                /* rank */                                                        \
    NFFT_INT *N;               /* bandwidths */                                                  \
    R *sigma;                  /* oversampling factors n/N */                                    \
    NFFT_INT *n;               /* FFT lengths */                                                 \
    NFFT_INT n_total;                                                                            \
    NFFT_INT m;                /* window cut-off */                                              \
    R *b;                      /* window shape parameters */                                     \
    NFFT_INT K;                /* PRE_LIN_PSI table size (accepted, unused on device) */         \
    unsigned flags;                                                                              \
    unsigned fftw_flags;                                                                         \
    R *x;                      /* nodes, host, x[j*d+t] in [-1/2,1/2) */                         \
    R MEASURE_TIME_t[3];       /* seconds of the last transform's D, F, B stages */              \
    void *my_fftw_plan1;       /* nfftcu_ctx* */                                                 \
    void *my_fftw_plan2;       /* NULL */                                                        \
    R **c_phi_inv;             /* host, per dimension N_t reals (PRE_PHI_HUT) */                 \
    R *psi;                    /* NULL: window table is device-side */                           \
    NFFT_INT *psi_index_g;     /* NULL */                                                        \
    NFFT_INT *psi_index_f;     /* NULL */                                                        \
    C *g;                      /* NULL: grid is device-side */                                   \
    C *g_hat;                  /* NULL */                                                        \
    C *g1;                     /* NULL */                                                        \
    C *g2;                     /* NULL */                                                        \
    R *spline_coeffs;          /* NULL */                                                        \
    NFFT_INT *index_x;         /* host, 2*M (key, node) pairs when NFFT_SORT_NODES */            \
  } X(plan);                                                                                     \
                                                                                                 \
  void X(trafo_direct)(const X(plan) *ths);                                                      \
  void X(adjoint_direct)(const X(plan) *ths);                                                    \
  void X(trafo)(X(plan) *ths);                                                                   \
  void X(trafo_1d)(X(plan) *ths);                                                                \
  void X(trafo_2d)(X(plan) *ths);                                                                \
  void X(trafo_3d)(X(plan) *ths);                                                                \
  void X(adjoint)(X(plan) *ths);                                                                 \
  void X(adjoint_1d)(X(plan) *ths);                                                              \
  void X(adjoint_2d)(X(plan) *ths);                                                              \
  void X(adjoint_3d)(X(plan) *ths);                                                              \
  void X(init_1d)(X(plan) *ths, int N1, int M);                                                  \
  void X(init_2d)(X(plan) *ths, int N1, int N2, int M);                                          \
  void X(init_3d)(X(plan) *ths, int N1, int N2, int N3, int M);                                  \
  void X(init)(X(plan) *ths, int d, int *N, int M);                                              \
  void X(init_guru)(X(plan) *ths, int d, int *N, int M, int *n, int m, unsigned flags,           \
                    unsigned fftw_flags);                                                        \
  void X(init_lin)(X(plan) *ths, int d, int *N, int M, int *n, int m, int K, unsigned flags,     \
                   unsigned fftw_flags);                                                         \
  void X(precompute_one_psi)(X(plan) *ths);                                                      \
  void X(precompute_psi)(X(plan) *ths);                                                          \
  void X(precompute_full_psi)(X(plan) *ths);                                                     \
  void X(precompute_fg_psi)(X(plan) *ths);                                                       \
  void X(precompute_lin_psi)(X(plan) *ths);                                                      \
  const char *X(check)(X(plan) *ths);                                                            \
  void X(finalize)(X(plan) *ths);

#define NFFT_B200_MANGLE_DOUBLE(name) nfft_##name
#define NFFT_B200_MANGLE_FLOAT(name) nfftf_##name

NFFT_B200_DEFINE_API(NFFT_B200_MANGLE_DOUBLE, double, nfft_b200_cdouble)
NFFT_B200_DEFINE_API(NFFT_B200_MANGLE_FLOAT, float, nfft_b200_cfloat)

/* ---- inverse-NFFT solver (layout of include/nfft3.h:758-786, SOLVER_DEFINE_API, complex variant) ------------
 * solver_*_complex / solverf_*_complex of libnfft3_b200.so run the iteration on the device (include/nfftcu.h,
 * nfftcu_solver_*) when mv is an NFFT plan of this library; any other mv plan is refused loudly (link the
 * reference's kernel/solver/solver.c for those -- it runs unmodified on top of nfft_trafo / nfft_adjoint).
 * Host members: y, w, w_hat and the initial f_hat_iter are read at solver_before_loop; f_hat_iter, r_iter and all
 * scalar members are current after every call; z_hat_iter, p_hat_iter and v_iter are allocated as in the
 * reference but only refreshed when the environment variable NFFT_B200_SOLVER_MIRROR_ALL is set. */
#define NFFT_B200_DEFINE_SOLVER_API(X, Y, R, C)                                                  \
  typedef struct {                                                                               \
    Y(mv_plan_complex) *mv;                                                                      \
    unsigned flags;                                                                              \
    R *w;                                                                                        \
    R *w_hat;                                                                                    \
    C *y;                                                                                        \
    C *f_hat_iter;                                                                               \
    C *r_iter;                                                                                   \
    C *z_hat_iter;                                                                               \
    C *p_hat_iter;                                                                               \
    C *v_iter;                                                                                   \
    R alpha_iter;                                                                                \
    R beta_iter;                                                                                 \
    R dot_r_iter;                                                                                \
    R dot_r_iter_old;                                                                            \
    R dot_z_hat_iter;                                                                            \
    R dot_z_hat_iter_old;                                                                        \
    R dot_p_hat_iter;                                                                            \
    R dot_v_iter;                                                                                \
  } X(plan_complex);                                                                             \
  void X(init_advanced_complex)(X(plan_complex) *ths, Y(mv_plan_complex) *mv, unsigned flags);   \
  void X(init_complex)(X(plan_complex) *ths, Y(mv_plan_complex) *mv);                            \
  void X(before_loop_complex)(X(plan_complex) *ths);                                             \
  void X(loop_one_step_complex)(X(plan_complex) *ths);                                           \
  void X(finalize_complex)(X(plan_complex) *ths);

#define NFFT_B200_SOLVER_MANGLE_DOUBLE(name) solver_##name
#define NFFT_B200_SOLVER_MANGLE_FLOAT(name) solverf_##name
NFFT_B200_DEFINE_SOLVER_API(NFFT_B200_SOLVER_MANGLE_DOUBLE, NFFT_B200_MANGLE_DOUBLE, double, nfft_b200_cdouble)
NFFT_B200_DEFINE_SOLVER_API(NFFT_B200_SOLVER_MANGLE_FLOAT, NFFT_B200_MANGLE_FLOAT, float, nfft_b200_cfloat)

/* ---- field-inhomogeneity transforms (layout and names of include/nfft3.h:510-541, MRI_DEFINE_API; double only,
 * like kernel/mri/mri.c) -----------------------------------------------------------------------------------------
 * mri_inh_2d1d_* / mri_inh_3d_* of libnfft3_b200.so keep the N3 + 1 NFFTs and the PHI / PHI_HUT / cexp scaling
 * between them on the device (nfft_b200/csrc/mri_host.c -> mri.cu); host-visible side effects of the reference
 * (buffer replacement in the 2d1d transforms, in-place scaling of f in the 3d adjoint) are preserved. */
typedef struct {
  NFFT_INT N_total;
  NFFT_INT M_total;
  nfft_b200_cdouble *f_hat;
  nfft_b200_cdouble *f;
  void (*mv_trafo)(void *);
  void (*mv_adjoint)(void *);
  nfft_plan plan;
  int N3;
  double sigma3;
  double *t;
  double *w;
} mri_inh_2d1d_plan;
typedef mri_inh_2d1d_plan mri_inh_3d_plan;   /* identical member lists in the reference */
void mri_inh_2d1d_trafo(mri_inh_2d1d_plan *ths);
void mri_inh_2d1d_adjoint(mri_inh_2d1d_plan *ths);
void mri_inh_2d1d_init_guru(mri_inh_2d1d_plan *ths, int *N, int M, int *n, int m, double sigma, unsigned nfft_flags,
    unsigned fftw_flags);
void mri_inh_2d1d_finalize(mri_inh_2d1d_plan *ths);
void mri_inh_3d_trafo(mri_inh_3d_plan *ths);
void mri_inh_3d_adjoint(mri_inh_3d_plan *ths);
void mri_inh_3d_init_guru(mri_inh_3d_plan *ths, int *N, int M, int *n, int m, double sigma, unsigned nfft_flags,
    unsigned fftw_flags);
void mri_inh_3d_finalize(mri_inh_3d_plan *ths);

/* solver flags (values of include/nfft3.h:823-829) */
#ifndef LANDWEBER
#define LANDWEBER (1U << 0)
#define STEEPEST_DESCENT (1U << 1)
#define CGNR (1U << 2)
#define CGNE (1U << 3)
#define NORMS_FOR_LANDWEBER (1U << 4)
#define PRECOMPUTE_WEIGHT (1U << 5)
#define PRECOMPUTE_DAMP (1U << 6)
#endif

/* plan flags (values of include/nfft3.h:195-208) */
#ifndef PRE_PHI_HUT
#define PRE_PHI_HUT (1U << 0)
#define FG_PSI (1U << 1)
#define PRE_LIN_PSI (1U << 2)
#define PRE_FG_PSI (1U << 3)
#define PRE_PSI (1U << 4)
#define PRE_FULL_PSI (1U << 5)
#define MALLOC_X (1U << 6)
#define MALLOC_F_HAT (1U << 7)
#define MALLOC_F (1U << 8)
#define FFT_OUT_OF_PLACE (1U << 9)
#define FFTW_INIT (1U << 10)
#define NFFT_SORT_NODES (1U << 11)
#define NFFT_OMP_BLOCKWISE_ADJOINT (1U << 12)
#define PRE_ONE_PSI (PRE_LIN_PSI | PRE_FG_PSI | PRE_PSI | PRE_FULL_PSI)
#endif

/* split-phase extensions (no reference counterpart): begin enqueues copy-in + transform + copy-out on the plan's stream
 * and returns, wait returns when they are done; two plans overlap their copies with each other's kernels.  Buffers must
 * be nfft_malloc'ed (page-locked) and untouched until the wait; the resident nodes are used as they are. */
void nfft_b200_trafo_begin(nfft_plan *ths);
void nfft_b200_adjoint_begin(nfft_plan *ths);
void nfft_b200_wait(nfft_plan *ths);
void nfftf_b200_trafo_begin(nfftf_plan *ths);
void nfftf_b200_adjoint_begin(nfftf_plan *ths);
void nfftf_b200_wait(nfftf_plan *ths);

/* util entry points of include/nfft3.h:839-890 that the plan initialisers depend on */
NFFT_INT nfft_next_power_of_2(const NFFT_INT N);
NFFT_INT nfftf_next_power_of_2(const NFFT_INT N);
NFFT_INT nfft_get_default_window_cut_off(void);
NFFT_INT nfftf_get_default_window_cut_off(void);
const char *nfft_get_window_name(void);
const char *nfftf_get_window_name(void);

#ifdef __cplusplus
}
#endif
#endif /* NFFT3_B200_H */
